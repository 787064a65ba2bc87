// glc_flac_api.cu -- host side of the FLAC path: flac::encode_flac_with_level (reference
// src/flac.rs:947-1052) over the two device passes of glc_flac_kernels.cu.
//
//   host:   argument checks in the reference's order (:963-978), block size by level (:983-995),
//           "fLaC" + STREAMINFO (:1001, :908-944), MD5 of the i16 little-endian samples (:1004,
//           :305-318).  MD5 is a serial chain per file, so it runs on host threads (one per file,
//           converting f32 -> i16 on the fly) concurrently with the GPU passes.
//   device: pass A (sizes + Rice parameters) -> scan -> pass B (bitstream + CRCs) -> D2H straight
//           into each file's output buffer behind its 42-byte header.
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include <math.h>
#include <stdlib.h>

#include "glc_internal.cuh"

using namespace glc;

namespace
{

#define FL_CUDA(expr)                                                                                        \
    do                                                                                                       \
    {                                                                                                        \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return set_error(GLC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                                      \
    } while (0)

// ---- MD5 (RFC 1321), streaming; fed the samples' little-endian i16 bytes ----
struct Md5
{
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    uint64_t total = 0;
    uint8_t buf[64];
    uint32_t fill = 0;
    static inline uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
    void block(const uint8_t *p)
    {
        static const uint32_t K[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
            0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
            0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
            0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
            0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
        uint32_t x[16];
        memcpy(x, p, 64); // little-endian host
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
        for (int i = 0; i < 64; ++i)
        {
            uint32_t f;
            int g;
            const int r = i >> 4;
            if (r == 0)
            {
                f = (b & c) | (~b & d);
                g = i;
            }
            else if (r == 1)
            {
                f = (d & b) | (~d & c);
                g = (5 * i + 1) & 15;
            }
            else if (r == 2)
            {
                f = b ^ c ^ d;
                g = (3 * i + 5) & 15;
            }
            else
            {
                f = c ^ (b | ~d);
                g = (7 * i) & 15;
            }
            const uint32_t t = d;
            d = c;
            c = b;
            b = b + rol(a + f + K[i] + x[g], S[r][i & 3]);
            a = t;
        }
        st[0] += a;
        st[1] += b;
        st[2] += c;
        st[3] += d;
    }
    void update(const uint8_t *p, size_t n)
    {
        total += n;
        if (fill)
        {
            const size_t take = std::min<size_t>(64 - fill, n);
            memcpy(buf + fill, p, take);
            fill += (uint32_t)take;
            p += take;
            n -= take;
            if (fill == 64)
            {
                block(buf);
                fill = 0;
            }
        }
        while (n >= 64)
        {
            block(p);
            p += 64;
            n -= 64;
        }
        if (n)
        {
            memcpy(buf, p, n);
            fill = (uint32_t)n;
        }
    }
    void finish(uint8_t out[16])
    {
        const uint64_t bits = total * 8;
        uint8_t pad[72] = {0x80};
        const size_t padlen = (fill < 56) ? (56 - fill) : (120 - fill);
        update(pad, padlen);
        uint8_t lenb[8];
        for (int i = 0; i < 8; ++i)
            lenb[i] = (uint8_t)(bits >> (8 * i));
        update(lenb, 8);
        memcpy(out, st, 16);
    }
};

// MD5 over ((s*32767).clamp(-32768,32767) as i16) little-endian, src/flac.rs:955-958, 305-318
void md5_of_pcm(const float *pcm, uint64_t n, uint8_t out[16])
{
    Md5 m;
    int16_t chunk[4096];
    uint64_t i = 0;
    while (i < n)
    {
        const uint64_t take = std::min<uint64_t>(4096, n - i);
        for (uint64_t j = 0; j < take; ++j)
        {
            const float v = pcm[i + j] * 32767.0f;
            int16_t q;
            if (v != v)
                q = 0;
            else if (v <= -32768.0f)
                q = -32768;
            else if (v >= 32767.0f)
                q = 32767;
            else
                q = (int16_t)(int32_t)v;
            chunk[j] = q;
        }
        m.update(reinterpret_cast<const uint8_t *>(chunk), take * 2);
        i += take;
    }
    m.finish(out);
}

struct BitOut
{
    uint8_t *p;
    uint32_t nbits = 0;
    void put(uint64_t v, int n)
    {
        for (int b = n - 1; b >= 0; --b, ++nbits)
            if ((v >> b) & 1u)
                p[nbits >> 3] |= (uint8_t)(0x80u >> (nbits & 7));
    }
};

} // namespace

extern "C" glc_status glc_flac_encode_batch(glc_ctx *ctx, uint32_t n_files, const float *const *pcm,
                                            const uint64_t *n_samples, const uint32_t *sample_rate,
                                            const uint16_t *channels, uint8_t level, uint8_t **bytes, uint64_t *len)
{
    if (!ctx || !pcm || !n_samples || !sample_rate || !channels || !bytes || !len || n_files == 0)
        return set_error(GLC_ERR_INVALID_ARG, "null/empty argument");
    // ---- checks, in the reference's order: length first (:963), then level (:972) ----
    for (uint32_t i = 0; i < n_files; ++i)
    {
        if (!pcm[i] && n_samples[i])
            return set_error(GLC_ERR_INVALID_ARG, "file %u: pcm is null", i);
        if (channels[i] == 0)
            return set_error(GLC_ERR_INVALID_ARG, "file %u: channels is 0", i);
        if (channels[i] > 8)
            return set_error(GLC_ERR_INVALID_ARG, "file %u: FLAC carries at most 8 channels", i);
        const uint64_t total = n_samples[i] / channels[i];
        if (total < 16)
            return set_error(GLC_ERR_FLAC_TOO_SHORT, "FLAC requires at least 16 samples per channel, got %llu",
                             (unsigned long long)total);
    }
    if (level > 8)
        return set_error(GLC_ERR_FLAC_LEVEL, "Invalid compression level %u, must be 0-8", (unsigned)level);

    FL_CUDA(cudaSetDevice(ctx_device(ctx)));
    cudaStream_t cs = ctx_compute_stream(ctx);
    cudaDeviceProp prop;
    FL_CUDA(cudaGetDeviceProperties(&prop, ctx_device(ctx)));

    std::vector<FlacFileDesc> files(n_files);
    uint64_t tot_pcm = 0, tot_blocks = 0;
    uint32_t max_ch = 1, max_bs = 16;
    for (uint32_t i = 0; i < n_files; ++i)
    {
        FlacFileDesc &f = files[i];
        const uint64_t total = n_samples[i] / channels[i];
        uint64_t bs = level <= 2 ? 1152 : 4096; // src/flac.rs:983-995
        bs = std::max<uint64_t>(std::min<uint64_t>(bs, total), 16);
        f.pcm_off = tot_pcm;
        f.i16_off = tot_pcm;
        f.n_samples = n_samples[i];
        f.first_block = tot_blocks;
        f.block_size = (uint32_t)bs;
        f.channels = channels[i];
        f.sample_rate = sample_rate[i];
        f.n_blocks = (uint32_t)((total + bs - 1) / bs); // loop of :1024-1049
        tot_blocks += f.n_blocks;
        tot_pcm += (n_samples[i] + 3) & ~(uint64_t)3;
        max_ch = std::max<uint32_t>(max_ch, channels[i]);
        max_bs = std::max<uint32_t>(max_bs, (uint32_t)bs);
    }

    // ---- MD5 on host threads, overlapped with everything below ----
    std::vector<std::vector<uint8_t>> md5(n_files, std::vector<uint8_t>(16));
    std::vector<std::thread> workers;
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned nthreads = std::min<unsigned>(hw, n_files);
        for (unsigned t = 0; t < nthreads; ++t)
            workers.emplace_back([&, t]() {
                for (uint32_t i = t; i < n_files; i += nthreads)
                    md5_of_pcm(pcm[i], n_samples[i], md5[i].data());
            });
    }
    auto join_all = [&]() {
        for (auto &w : workers)
            if (w.joinable())
                w.join();
    };

    float *d_pcm = nullptr;
    int16_t *d_i16 = nullptr;
    FlacFileDesc *d_files = nullptr;
    uint8_t *d_k = nullptr, *d_out = nullptr;
    uint32_t *d_fbytes = nullptr, *d_scratch = nullptr;
    uint64_t *d_foff = nullptr;
    glc_status st = GLC_OK;
    uint32_t *h_fbytes = nullptr;
    std::vector<uint8_t *> outs(n_files, nullptr);
    do
    {
#define FL_STEP(expr)                                                                             \
    {                                                                                             \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
        {                                                                                         \
            st = set_error(GLC_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));          \
            break;                                                                                \
        }                                                                                         \
    }
        FL_STEP(dev_alloc(ctx, (void **)&d_pcm, std::max<uint64_t>(tot_pcm, 1) * 4, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_i16, std::max<uint64_t>(tot_pcm, 1) * 2, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_files, sizeof(FlacFileDesc) * n_files, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_k, std::max<uint64_t>(tot_blocks, 1) * max_ch * 64, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_fbytes, std::max<uint64_t>(tot_blocks, 1) * 4, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_foff, (tot_blocks + 1) * 8, cs));
        FL_STEP(cudaMemcpyAsync(d_files, files.data(), sizeof(FlacFileDesc) * n_files, cudaMemcpyHostToDevice, cs));
        bool bad = false;
        for (uint32_t i = 0; i < n_files && !bad; ++i)
            if (cudaMemcpyAsync(d_pcm + files[i].pcm_off, pcm[i], n_samples[i] * 4, cudaMemcpyHostToDevice, cs) !=
                cudaSuccess)
                bad = true;
        if (bad)
        {
            st = set_error(GLC_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        ctx_count_bytes(ctx, tot_pcm * 4, 0);

        FlacLaunch fl{};
        fl.pcm_arena = d_pcm;
        fl.files = d_files;
        fl.n_files = n_files;
        fl.n_blocks_total = tot_blocks;
        fl.level = level;
        fl.frame_bytes = d_fbytes;
        fl.i16_arena = d_i16;
        void *tok = nullptr;
        ctx_count_launch(ctx, GLC_K_FLAC_BLOCK, 1);
        ctx_time_begin(ctx, GLC_K_FLAC_BLOCK, &tok);
        FL_STEP(launch_flac_measure(fl, d_k, max_ch, max_bs, cs));
        ctx_time_end(ctx, tok);
        ctx_count_launch(ctx, GLC_K_SCAN, 1);
        ctx_time_begin(ctx, GLC_K_SCAN, &tok);
        FL_STEP(launch_scan_u32_u64(d_fbytes, d_foff, tot_blocks, cs));
        ctx_time_end(ctx, tok);

        // frame sizes come back to the host: total size, largest frame, per-file extents
        h_fbytes = (uint32_t *)pinned_alloc(ctx, std::max<uint64_t>(tot_blocks, 1) * 4);
        if (!h_fbytes)
        {
            st = set_error(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
            break;
        }
        FL_STEP(cudaMemcpyAsync(h_fbytes, d_fbytes, tot_blocks * 4, cudaMemcpyDeviceToHost, cs));
        FL_STEP(cudaStreamSynchronize(cs));
        uint64_t tot_bytes = 0;
        uint32_t max_frame = 16;
        std::vector<uint64_t> file_bytes(n_files, 0), file_off(n_files, 0);
        for (uint32_t i = 0; i < n_files; ++i)
        {
            file_off[i] = tot_bytes;
            for (uint32_t b = 0; b < files[i].n_blocks; ++b)
            {
                const uint32_t fb = h_fbytes[files[i].first_block + b];
                file_bytes[i] += fb;
                max_frame = std::max(max_frame, fb);
            }
            tot_bytes += file_bytes[i];
        }
        FL_STEP(dev_alloc(ctx, (void **)&d_out, std::max<uint64_t>(tot_bytes, 1), cs));
        if (const size_t sw = flac_emit_scratch_words(tot_blocks, max_ch, max_bs, max_frame, prop.multiProcessorCount))
            FL_STEP(dev_alloc(ctx, (void **)&d_scratch, sw * 4, cs));
        ctx_count_launch(ctx, GLC_K_FLAC_GATHER, 1);
        ctx_time_begin(ctx, GLC_K_FLAC_GATHER, &tok);
        FL_STEP(launch_flac_emit(fl, d_k, max_ch, max_bs, max_frame, d_foff, d_out, d_scratch,
                                 prop.multiProcessorCount, cs));
        ctx_time_end(ctx, tok);

        // outputs: 4 + 38 header bytes, then the file's frames copied straight from the device
        for (uint32_t i = 0; i < n_files; ++i)
        {
            outs[i] = (uint8_t *)pinned_alloc(ctx, 42 + file_bytes[i]);
            if (!outs[i])
            {
                st = set_error(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
                break;
            }
            if (file_bytes[i])
                if (cudaMemcpyAsync(outs[i] + 42, d_out + file_off[i], file_bytes[i], cudaMemcpyDeviceToHost, cs) !=
                    cudaSuccess)
                {
                    st = set_error(GLC_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
                    break;
                }
            len[i] = 42 + file_bytes[i];
        }
        if (st != GLC_OK)
            break;
        ctx_count_bytes(ctx, 0, tot_bytes + tot_blocks * 4);
        FL_STEP(cudaStreamSynchronize(cs));
    } while (0);

    join_all();
    if (st == GLC_OK)
    {
        for (uint32_t i = 0; i < n_files; ++i)
        {
            uint8_t *h = outs[i];
            memset(h, 0, 42);
            memcpy(h, "fLaC", 4); // src/flac.rs:9,1001
            BitOut w{h + 4};
            const uint64_t total = n_samples[i] / channels[i];
            w.put(1, 1);  // last metadata block        src/flac.rs:923
            w.put(0, 7);  // STREAMINFO                  :925
            w.put(34, 24);
            w.put(files[i].block_size & 0xFFFFu, 16); // min = max block size (:1009-1010)
            w.put(files[i].block_size & 0xFFFFu, 16);
            w.put(0, 24); // unknown min/max frame size (:1011-1012)
            w.put(0, 24);
            w.put(sample_rate[i], 20);
            w.put((uint64_t)(channels[i] - 1), 3);
            w.put(15, 5); // bits per sample - 1
            w.put(total, 36);
            memcpy(h + 4 + 4 + 18, md5[i].data(), 16);
            bytes[i] = h;
        }
    }
    else
    {
        for (auto p : outs)
            if (p)
                pinned_release(ctx, p);
    }
    if (h_fbytes)
        pinned_release(ctx, h_fbytes);
    dev_free(ctx, d_pcm, cs);
    dev_free(ctx, d_i16, cs);
    dev_free(ctx, d_files, cs);
    dev_free(ctx, d_k, cs);
    dev_free(ctx, d_fbytes, cs);
    dev_free(ctx, d_foff, cs);
    dev_free(ctx, d_out, cs);
    dev_free(ctx, d_scratch, cs);
    return st;
}

extern "C" glc_status glc_flac_encode(glc_ctx *ctx, const float *pcm, uint64_t n_samples, uint32_t sample_rate,
                                      uint16_t channels, uint8_t level, uint8_t **bytes, uint64_t *len)
{
    if (!bytes || !len)
        return set_error(GLC_ERR_INVALID_ARG, "null argument");
    const float *files[1] = {pcm};
    return glc_flac_encode_batch(ctx, 1, files, &n_samples, &sample_rate, &channels, level, bytes, len);
}
