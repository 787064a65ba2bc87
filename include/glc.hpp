/*
 * glc.hpp -- C++17 host-side mirror of the reference crate's public `codec` / `flac` module API
 * (ajcm474/gapless-lossy-codec v0.5.0, src/lib.rs:1-5) over the C ABI of glc.h.
 *
 * The reference is compiled (Rust) code and this image has no Rust toolchain, so the host side above
 * the C ABI is mirrored in C++: same type names, same method names, same argument meaning and the
 * same error behaviour, so that tests/cpp/reference_suite.cpp reads like the reference test files (tests/ of the crate).
 * (The Rust shim a maintainer would add is INTEGRATION.md; the Python mirror used by the pytest
 * parity suite is gapless_lossy_codec_b200/codec.py.)
 *
 *   Rust (reference)                                   C++ (this header)
 *   ------------------------------------------------   ---------------------------------------------
 *   Encoder::new(sr)                  codec.rs:406      glc::codec::Encoder enc(sr);
 *   enc.encode(&samples, ch)?         codec.rs:421      glc::codec::EncodedAudio e = enc.encode(samples, ch);
 *   Decoder::new(ch, sr)              codec.rs:581      glc::codec::Decoder dec(ch, sr);
 *   dec.decode(&e, None)?             codec.rs:744      std::vector<float> pcm = dec.decode(e);
 *   dec.decode_streaming(Arc(e), tx)  codec.rs:595      glc::codec::ChunkReceiver rx = dec.decode_streaming(e, tx);
 *   save_encoded / load_encoded       codec.rs:774,781  glc::codec::save_encoded / load_encoded
 *   flac::encode_flac_with_level      flac.rs:947       glc::flac::encode_flac_with_level
 *   flac::encode_flac                 flac.rs:1055      glc::flac::encode_flac            (level 5)
 *   flac::export_to_flac[_with_level] flac.rs:1065,1080 glc::flac::export_to_flac[_with_level]
 *
 * Result<T> / anyhow::Error  ->  glc::Error (std::runtime_error carrying the glc_status);
 * the reference's panic on <= 512 samples per channel -> glc::Error(GLC_ERR_TOO_SHORT).
 * There is no CPU fallback: without a B200 the first call throws glc::Error(GLC_ERR_NO_DEVICE).
 * Header-only; link with -lglc_b200.
 */
#ifndef GLC_HPP
#define GLC_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "glc.h"

namespace glc
{

/* anyhow::Error of the reference: message + the ABI status it came from. */
class Error : public std::runtime_error
{
  public:
    Error(glc_status st, const std::string &what) : std::runtime_error(what), status_(st) {}
    glc_status status() const noexcept { return status_; }

  private:
    glc_status status_;
};

inline void check(glc_status st, const char *what)
{
    if (st != GLC_OK)
        throw Error(st, std::string(what) + ": " + glc_last_error());
}

/* One CUDA device (glc_ctx).  The reference has no such object (its tables live inside Encoder /
 * Decoder, codec.rs:396-402, 571-577); here they are built once per device and shared. */
class Context
{
  public:
    explicit Context(int device = 0, glc_mode mode = GLC_MODE_EXACT)
    {
        check(glc_ctx_create(device, mode, &h_), "glc_ctx_create");
    }
    ~Context()
    {
        if (h_)
            glc_ctx_destroy(h_);
    }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    glc_ctx *handle() const noexcept { return h_; }
    /* process-wide default on device 0, EXACT (bit-exact) mode */
    static Context &global()
    {
        static Context ctx(0, GLC_MODE_EXACT);
        return ctx;
    }

  private:
    glc_ctx *h_ = nullptr;
};

namespace codec
{

constexpr std::size_t FRAME_SIZE = GLC_FRAME_SIZE;             /* codec.rs:15 */
constexpr std::size_t HOP_SIZE = GLC_HOP_SIZE;                 /* codec.rs:16 */
constexpr std::size_t FRAMES_PER_CHUNK = GLC_FRAMES_PER_CHUNK; /* codec.rs:18 */

struct AudioHeader /* codec.rs:39-45 */
{
    std::uint32_t sample_rate = 0;
    std::uint16_t channels = 0;
    std::uint64_t total_samples = 0;
};

struct GaplessInfo /* codec.rs:47-53 */
{
    std::uint32_t encoder_delay = 0;
    std::uint32_t padding = 0;
    std::uint64_t original_length = 0;
};

struct EncodedFrame /* codec.rs:55-69 */
{
    std::vector<std::vector<std::pair<std::uint16_t, std::int16_t>>> sparse_coeffs_per_channel;
    std::vector<float> scale_factors;
    std::optional<std::vector<std::int16_t>> raw_pcm;
};

struct EncodedAudio /* codec.rs:31-37 */
{
    AudioHeader header;
    std::vector<EncodedFrame> frames;
    GaplessInfo gapless_info;
};

struct Progress /* enum Progress, codec.rs:71-79 */
{
    enum Kind
    {
        Encoding,
        Decoding,
        Exporting,
        Complete,
        Error,
        Status
    } kind;
    float value;      /* Encoding / Decoding / Exporting */
    std::string text; /* Complete / Error / Status */
};
/* Option<Sender<Progress>>: an empty std::function is None. */
using ProgressSender = std::function<void(const Progress &)>;

struct AudioChunk /* codec.rs:81-85 */
{
    std::vector<float> samples;
    bool is_last = false;
};

namespace detail
{

/* Owning flat image (struct glc_encoded) of a nested EncodedAudio: what crosses the ABI. */
struct Flat
{
    glc_encoded e{};
    std::vector<std::uint8_t> frame_is_raw;
    std::vector<std::uint32_t> nnz;
    std::vector<std::uint64_t> pair_offset, raw_offset;
    std::vector<glc_pair> pairs;
    std::vector<float> scales;
    std::vector<std::int16_t> raw;

    explicit Flat(const EncodedAudio &a)
    {
        const std::size_t ch = a.header.channels, nf = a.frames.size();
        frame_is_raw.assign(nf, 0);
        nnz.assign(nf * ch, 0);
        scales.assign(nf * ch, 0.0f);
        pair_offset.assign(nf * ch + 1, 0);
        raw_offset.assign(nf + 1, 0);
        for (std::size_t f = 0; f < nf; ++f)
        {
            const EncodedFrame &fr = a.frames[f];
            if (fr.raw_pcm)
            {
                frame_is_raw[f] = 1;
                raw.insert(raw.end(), fr.raw_pcm->begin(), fr.raw_pcm->end());
            }
            else
            {
                /* the reference indexes [ch] unchecked and would panic (codec.rs:651-652) */
                if (fr.sparse_coeffs_per_channel.size() < ch || fr.scale_factors.size() < ch)
                    throw glc::Error(GLC_ERR_CORRUPT, "frame " + std::to_string(f) + " has fewer channels than the header");
                for (std::size_t c = 0; c < ch; ++c)
                {
                    const auto &v = fr.sparse_coeffs_per_channel[c];
                    nnz[f * ch + c] = (std::uint32_t)v.size();
                    scales[f * ch + c] = fr.scale_factors[c];
                    for (const auto &p : v)
                        pairs.push_back(glc_pair{p.first, p.second});
                }
            }
            for (std::size_t c = 0; c < ch; ++c)
                pair_offset[f * ch + c + 1] = pair_offset[f * ch + c] + nnz[f * ch + c];
            raw_offset[f + 1] = raw.size();
        }
        e.sample_rate = a.header.sample_rate;
        e.channels = a.header.channels;
        e.total_samples = a.header.total_samples;
        e.encoder_delay = a.gapless_info.encoder_delay;
        e.padding = a.gapless_info.padding;
        e.original_length = a.gapless_info.original_length;
        e.n_frames = nf;
        e.frame_is_raw = frame_is_raw.data();
        e.nnz = nnz.data();
        e.pair_offset = pair_offset.data();
        e.pairs = pairs.data();
        e.scales = scales.data();
        e.raw_offset = raw_offset.data();
        e.raw = raw.data();
    }
    Flat(const Flat &) = delete;
    Flat &operator=(const Flat &) = delete;
};

/* flat -> nested, the shapes Encoder::encode builds at codec.rs:517-540 */
inline EncodedAudio nest(const glc_encoded &e)
{
    EncodedAudio a;
    a.header = AudioHeader{e.sample_rate, e.channels, e.total_samples};
    a.gapless_info = GaplessInfo{e.encoder_delay, e.padding, e.original_length};
    const std::size_t ch = e.channels;
    a.frames.resize(e.n_frames);
    for (std::size_t f = 0; f < e.n_frames; ++f)
    {
        EncodedFrame &fr = a.frames[f];
        if (e.frame_is_raw[f])
            fr.raw_pcm.emplace(e.raw + e.raw_offset[f], e.raw + e.raw_offset[f + 1]);
        else
        {
            fr.sparse_coeffs_per_channel.resize(ch);
            fr.scale_factors.assign(e.scales + f * ch, e.scales + (f + 1) * ch);
            for (std::size_t c = 0; c < ch; ++c)
            {
                auto &v = fr.sparse_coeffs_per_channel[c];
                const std::uint64_t a0 = e.pair_offset[f * ch + c], a1 = e.pair_offset[f * ch + c + 1];
                v.reserve(a1 - a0);
                for (std::uint64_t i = a0; i < a1; ++i)
                    v.emplace_back(e.pairs[i].idx, e.pairs[i].q);
            }
        }
    }
    return a;
}

} // namespace detail

/* codec::Encoder, codec.rs:396-566.  Reusable across files (tests/test_codec.rs:150-153). */
class Encoder
{
  public:
    explicit Encoder(std::uint32_t sample_rate, Context &ctx = Context::global()) : ctx_(ctx)
    {
        check(glc_encoder_new(ctx.handle(), sample_rate, &h_), "Encoder::new");
    }
    ~Encoder()
    {
        if (h_)
            glc_encoder_free(h_);
    }
    Encoder(const Encoder &) = delete;
    Encoder &operator=(const Encoder &) = delete;

    /* Encoder::encode(&mut self, samples: &[f32], channels: u16) -> Result<EncodedAudio> */
    EncodedAudio encode(const float *samples, std::size_t n, std::uint16_t channels)
    {
        glc_encoded *out = nullptr;
        check(glc_encode(h_, samples, n, channels, &out), "Encoder::encode");
        std::unique_ptr<glc_encoded, std::function<void(glc_encoded *)>> guard(
            out, [this](glc_encoded *p) { glc_encoded_free(ctx_.handle(), p); });
        return detail::nest(*out);
    }
    EncodedAudio encode(const std::vector<float> &samples, std::uint16_t channels)
    {
        return encode(samples.data(), samples.size(), channels);
    }
    /* many encode() calls in one device pass: the data-parallel entry point */
    std::vector<EncodedAudio> encode_batch(const std::vector<std::vector<float>> &files, const std::vector<std::uint16_t> &channels)
    {
        const std::uint32_t n = (std::uint32_t)files.size();
        std::vector<const float *> ptrs(n);
        std::vector<std::uint64_t> ns(n);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            ptrs[i] = files[i].data();
            ns[i] = files[i].size();
        }
        std::vector<glc_encoded *> outs(n, nullptr);
        check(glc_encode_batch(h_, n, ptrs.data(), ns.data(), channels.data(), outs.data()), "Encoder::encode_batch");
        std::vector<EncodedAudio> res;
        res.reserve(n);
        for (std::uint32_t i = 0; i < n; ++i)
        {
            res.push_back(detail::nest(*outs[i]));
            glc_encoded_free(ctx_.handle(), outs[i]);
        }
        return res;
    }

  private:
    Context &ctx_;
    glc_encoder *h_ = nullptr;
};

/* Receiver<AudioChunk> of decode_streaming (codec.rs:595, 612).  The reference feeds a bounded(5)
 * channel from a worker thread; here recv() pulls the next chunk from the device on demand, which
 * delivers the same sequence of chunks and Progress events in the same order. */
class ChunkReceiver
{
  public:
    ChunkReceiver(glc_decoder *dec, std::shared_ptr<const EncodedAudio> encoded, ProgressSender progress)
        : encoded_(std::move(encoded)), flat_(new detail::Flat(*encoded_)), progress_(std::move(progress))
    {
        if (progress_)
            progress_(Progress{Progress::Status, 0.0f,
                               "Starting streaming decode of " + std::to_string(encoded_->frames.size()) + " frames"});
        check(glc_decode_stream_open(dec, &flat_->e, &s_), "Decoder::decode_streaming");
    }
    ~ChunkReceiver()
    {
        if (s_)
            glc_decode_stream_close(s_);
    }
    ChunkReceiver(ChunkReceiver &&o) noexcept
        : encoded_(std::move(o.encoded_)), flat_(std::move(o.flat_)), progress_(std::move(o.progress_)), s_(o.s_), done_(o.done_)
    {
        o.s_ = nullptr;
    }
    ChunkReceiver(const ChunkReceiver &) = delete;
    ChunkReceiver &operator=(const ChunkReceiver &) = delete;

    /* rx.recv(): the next chunk, or nullopt once the sender has hung up (after is_last). */
    std::optional<AudioChunk> recv()
    {
        if (done_)
            return std::nullopt;
        const float *p = nullptr;
        std::uint64_t n = 0;
        int last = 0;
        float pct = 0.0f;
        check(glc_decode_stream_next(s_, &p, &n, &last, &pct), "decode_streaming recv");
        if (progress_ && !last)
            progress_(Progress{Progress::Decoding, pct, ""}); /* codec.rs:712 */
        AudioChunk c;
        c.samples.assign(p, p + n);
        c.is_last = last != 0;
        if (last)
        {
            done_ = true;
            if (progress_)
                progress_(Progress{Progress::Complete, 0.0f, "Decoded " + std::to_string(encoded_->frames.size()) + " frames"});
        }
        return c;
    }

  private:
    std::shared_ptr<const EncodedAudio> encoded_;
    std::unique_ptr<detail::Flat> flat_;
    ProgressSender progress_;
    glc_stream *s_ = nullptr;
    bool done_ = false;
};

/* codec::Decoder, codec.rs:571-769.  `channels` / `sample_rate` are informational, as in the
 * reference (fields never read; the stream header wins, codec.rs:598). */
class Decoder
{
  public:
    Decoder(std::size_t channels, std::uint32_t sample_rate, Context &ctx = Context::global()) : ctx_(ctx)
    {
        check(glc_decoder_new(ctx.handle(), (std::uint32_t)channels, sample_rate, &h_), "Decoder::new");
    }
    ~Decoder()
    {
        if (h_)
            glc_decoder_free(h_);
    }
    Decoder(const Decoder &) = delete;
    Decoder &operator=(const Decoder &) = delete;

    /* Decoder::decode(&mut self, &EncodedAudio, Option<Sender<Progress>>) -> Result<Vec<f32>>:
     * every chunk concatenated, then the gapless trim (codec.rs:744-768). */
    std::vector<float> decode(const EncodedAudio &encoded, const ProgressSender &progress = {})
    {
        detail::Flat flat(encoded);
        if (progress)
            progress(Progress{Progress::Status, 0.0f, "Starting streaming decode of " + std::to_string(encoded.frames.size()) + " frames"});
        float *p = nullptr;
        std::uint64_t n = 0;
        check(glc_decode(h_, &flat.e, &p, &n), "Decoder::decode");
        std::vector<float> out(p, p + n);
        glc_free(ctx_.handle(), p);
        if (progress)
            progress(Progress{Progress::Complete, 0.0f, "Decoded " + std::to_string(encoded.frames.size()) + " frames"});
        return out;
    }

    /* Decoder::decode_streaming(&mut self, Arc<EncodedAudio>, Option<Sender<Progress>>) -> Receiver<AudioChunk> */
    ChunkReceiver decode_streaming(std::shared_ptr<const EncodedAudio> encoded, ProgressSender progress = {})
    {
        return ChunkReceiver(h_, std::move(encoded), std::move(progress));
    }

    /* the samples `glc -d x.glc` writes to its WAV (main.rs:95-105 -> audio::convert_f32_to_i16, audio.rs:11-16),
     * converted on the device */
    std::vector<std::int16_t> decode_pcm16(const EncodedAudio &encoded)
    {
        detail::Flat flat(encoded);
        std::int16_t *p = nullptr;
        std::uint64_t n = 0;
        check(glc_decode_i16(h_, &flat.e, &p, &n), "Decoder::decode_pcm16");
        std::vector<std::int16_t> out(p, p + n);
        glc_free(ctx_.handle(), p);
        return out;
    }

    /* `glc -d x.glc --flac-level N` (main.rs:55-113) with the decoded PCM kept on the device */
    std::vector<std::uint8_t> decode_to_flac(const EncodedAudio &encoded, std::uint8_t level = 5)
    {
        detail::Flat flat(encoded);
        std::uint8_t *b = nullptr;
        std::uint64_t n = 0;
        check(glc_decode_to_flac(h_, &flat.e, level, &b, &n), "Decoder::decode_to_flac");
        std::vector<std::uint8_t> out(b, b + n);
        glc_free(ctx_.handle(), b);
        return out;
    }

  private:
    Context &ctx_;
    glc_decoder *h_ = nullptr;
};

/* bincode image of save_encoded / load_encoded (codec.rs:774-786) */
inline std::vector<std::uint8_t> encoded_to_bytes(const EncodedAudio &encoded, Context &ctx = Context::global())
{
    detail::Flat flat(encoded);
    std::uint8_t *b = nullptr;
    std::uint64_t n = 0;
    check(glc_encoded_to_bincode(ctx.handle(), &flat.e, &b, &n), "bincode::serialize");
    std::vector<std::uint8_t> out(b, b + n);
    glc_free(ctx.handle(), b);
    return out;
}

inline EncodedAudio encoded_from_bytes(const std::uint8_t *bytes, std::size_t len, Context &ctx = Context::global())
{
    glc_encoded *e = nullptr;
    check(glc_encoded_from_bincode(ctx.handle(), bytes, len, &e), "bincode::deserialize");
    EncodedAudio a = detail::nest(*e);
    glc_encoded_free(ctx.handle(), e);
    return a;
}

namespace detail
{
inline void write_file(const std::string &path, const std::uint8_t *p, std::size_t n)
{
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f)
        throw glc::Error(GLC_ERR_INVALID_ARG, "cannot create " + path);
    const std::size_t w = n ? std::fwrite(p, 1, n, f) : 0;
    std::fclose(f);
    if (w != n)
        throw glc::Error(GLC_ERR_INVALID_ARG, "short write to " + path);
}
inline std::vector<std::uint8_t> read_file(const std::string &path)
{
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f)
        throw glc::Error(GLC_ERR_INVALID_ARG, "cannot open " + path);
    std::vector<std::uint8_t> buf;
    std::uint8_t tmp[65536];
    std::size_t r;
    while ((r = std::fread(tmp, 1, sizeof tmp, f)) > 0)
        buf.insert(buf.end(), tmp, tmp + r);
    std::fclose(f);
    return buf;
}
} // namespace detail

/* pub fn save_encoded(encoded: &EncodedAudio, path: &Path) -> Result<()>            codec.rs:774 */
inline void save_encoded(const EncodedAudio &encoded, const std::string &path)
{
    const std::vector<std::uint8_t> b = encoded_to_bytes(encoded);
    detail::write_file(path, b.data(), b.size());
}

/* pub fn load_encoded(path: &Path) -> Result<EncodedAudio>                         codec.rs:781 */
inline EncodedAudio load_encoded(const std::string &path)
{
    const std::vector<std::uint8_t> b = detail::read_file(path);
    return encoded_from_bytes(b.data(), b.size());
}

} // namespace codec

namespace flac
{

/* pub fn encode_flac_with_level(samples, sample_rate, channels, compression_level) -> Result<Vec<u8>>
 *                                                                                   flac.rs:947 */
inline std::vector<std::uint8_t> encode_flac_with_level(const float *samples, std::size_t n, std::uint32_t sample_rate,
                                                        std::uint16_t channels, std::uint8_t compression_level,
                                                        Context &ctx = Context::global())
{
    std::uint8_t *b = nullptr;
    std::uint64_t len = 0;
    check(glc_flac_encode(ctx.handle(), samples, n, sample_rate, channels, compression_level, &b, &len), "encode_flac_with_level");
    std::vector<std::uint8_t> out(b, b + len);
    glc_free(ctx.handle(), b);
    return out;
}
inline std::vector<std::uint8_t> encode_flac_with_level(const std::vector<float> &samples, std::uint32_t sample_rate,
                                                        std::uint16_t channels, std::uint8_t compression_level)
{
    return encode_flac_with_level(samples.data(), samples.size(), sample_rate, channels, compression_level);
}

/* pub fn encode_flac(samples, sample_rate, channels): level 5                       flac.rs:1055 */
inline std::vector<std::uint8_t> encode_flac(const std::vector<float> &samples, std::uint32_t sample_rate, std::uint16_t channels)
{
    return encode_flac_with_level(samples, sample_rate, channels, 5);
}

/* pub fn export_to_flac_with_level(path, samples, sample_rate, channels, level)     flac.rs:1065 */
inline void export_to_flac_with_level(const std::string &path, const std::vector<float> &samples, std::uint32_t sample_rate,
                                      std::uint16_t channels, std::uint8_t compression_level)
{
    const std::vector<std::uint8_t> b = encode_flac_with_level(samples, sample_rate, channels, compression_level);
    codec::detail::write_file(path, b.data(), b.size());
}

/* pub fn export_to_flac(path, samples, sample_rate, channels)                       flac.rs:1080 */
inline void export_to_flac(const std::string &path, const std::vector<float> &samples, std::uint32_t sample_rate,
                           std::uint16_t channels)
{
    export_to_flac_with_level(path, samples, sample_rate, channels, 5);
}

} // namespace flac
} // namespace glc

#endif /* GLC_HPP */
