/*
 * glc.h -- C ABI of libglc_b200.so: the B200-native (sm_100a) encode/decode hot path of
 * gapless-lossy-codec, a drop-in for the reference crate's `codec` and `flac` modules.
 *
 * The reference (ajcm474/gapless-lossy-codec v0.5.0) has no FFI layer: its boundary is the Rust
 * module API of src/lib.rs:1-5.  Each entry point below names the reference item it replaces
 * (file:line relative to /root/reference).  INTEGRATION.md shows the Rust shim (`extern "C"`
 * declarations + replacement `codec.rs`/`flac.rs`) a maintainer would add, and include/glc.hpp is
 * a C++ mirror of the same API used by tests/cpp.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns a glc_status (0 = OK); glc_last_error() gives the message of the last
 *     failure on the calling thread; nothing unwinds across the ABI;
 *   - there is NO CPU fallback: without a usable CUDA device glc_ctx_create fails with
 *     GLC_ERR_NO_DEVICE and nothing else can be called;
 *   - outputs are allocated by the library (pinned host memory from a per-context pool) and are
 *     released with glc_encoded_free / glc_free;
 *   - PCM is interleaved f32 in [-1, 1], exactly what the reference's functions take.
 *   - input buffers may live in ordinary pageable memory (the reference's entry points borrow slices);
 *     they are then staged through a ring of pinned chunks filled by a few host threads
 *     (GLC_COPY_THREADS, default 4).  Buffers from glc_host_alloc, or pinned by the caller, go to the DMA
 *     engine directly.
 *   - a context and the objects made from it may be used by one thread at a time.
 */
#ifndef GLC_H
#define GLC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLC_FRAME_SIZE 2048u      /* src/codec.rs:15 */
#define GLC_HOP_SIZE 1024u        /* src/codec.rs:16 */
#define GLC_FRAMES_PER_CHUNK 500u /* src/codec.rs:18 */
#define GLC_ABI_VERSION 7u

typedef enum glc_status
{
    GLC_OK = 0,
    GLC_ERR_INVALID_ARG = 1,
    GLC_ERR_TOO_SHORT = 2,   /* codec input with <= 512 samples per channel: the reference panics
                                (src/codec.rs:449-452 then :474) */
    GLC_ERR_NO_MEMORY = 3,
    GLC_ERR_FLAC_TOO_SHORT = 4, /* "FLAC requires at least 16 samples per channel" src/flac.rs:963-969 */
    GLC_ERR_FLAC_LEVEL = 5,     /* "Invalid compression level" src/flac.rs:972-978 */
    GLC_ERR_NO_DEVICE = 6,      /* no CUDA device / driver: there is no CPU fallback */
    GLC_ERR_CUDA = 7,
    GLC_ERR_CORRUPT = 8,        /* malformed encoded stream / container */
    GLC_ERR_UNSUPPORTED = 9
} glc_status;

/* Transform arithmetic.  EXACT reproduces the reference's table-driven direct-form MDCT/IMDCT
 * operation by operation (src/codec.rs:359-390) and is the parity-gated default.  FAST is the
 * FFT-based true MDCT/IMDCT named in BASELINE.json, fused with the quantiser: same stream layout,
 * frame counts, gapless metadata and sample counts, but a TOLERANCE parity class (the reference's
 * f32 table is not the true MDCT basis).  FLAC and the container are
 * integer paths and identical in both modes. */
typedef enum glc_mode
{
    GLC_MODE_EXACT = 0,
    GLC_MODE_FAST = 1
} glc_mode;

/* (index, value) pair of EncodedFrame::sparse_coeffs_per_channel, src/codec.rs:62 */
typedef struct glc_pair
{
    uint16_t idx;
    int16_t q;
} glc_pair;

/* Flat image of EncodedAudio (src/codec.rs:31-69).  Frame f, channel c is "frame-channel"
 * fc = f*channels + c.  A frame is either sparse (frame_is_raw[f] == 0: nnz/pairs/scales valid for
 * its channels) or raw (frame_is_raw[f] == 1: raw_pcm = Some(..), 2048*channels i16 stored planar
 * [channel][2048] exactly as src/codec.rs:469-502 pushes them; nnz = 0 and scales = 0). */
typedef struct glc_encoded
{
    uint32_t sample_rate;    /* AudioHeader.sample_rate        src/codec.rs:42 */
    uint16_t channels;       /* AudioHeader.channels           :43 */
    uint16_t reserved0;
    uint64_t total_samples;  /* AudioHeader.total_samples      :44 (interleaved count) */
    uint32_t encoder_delay;  /* GaplessInfo.encoder_delay      :50 */
    uint32_t padding;        /* GaplessInfo.padding            :51 */
    uint64_t original_length;/* GaplessInfo.original_length    :52 */
    uint64_t n_frames;       /* frames.len() */
    uint8_t *frame_is_raw;   /* [n_frames] */
    uint32_t *nnz;           /* [n_frames*channels] */
    uint64_t *pair_offset;   /* [n_frames*channels + 1] exclusive scan of nnz */
    glc_pair *pairs;         /* [pair_offset[n_frames*channels]] ascending idx within a frame-channel */
    float *scales;           /* [n_frames*channels] */
    uint64_t *raw_offset;    /* [n_frames + 1] exclusive scan, i16 units */
    int16_t *raw;            /* [raw_offset[n_frames]] */
} glc_encoded;

typedef struct glc_ctx glc_ctx;         /* one CUDA device: streams, pinned pools, device tables */
typedef struct glc_encoder glc_encoder; /* codec::Encoder  src/codec.rs:396-402 */
typedef struct glc_decoder glc_decoder; /* codec::Decoder  src/codec.rs:571-577 */
typedef struct glc_stream glc_stream;   /* the Receiver<AudioChunk> of decode_streaming */

/* ------------------------------------------------------------------ context */

uint32_t glc_abi_version(void);
/* Message of the last error raised on this thread ("" if none). */
const char *glc_last_error(void);
/* Number of CUDA devices visible; 0 (with GLC_ERR_NO_DEVICE) when there is no driver/GPU. */
glc_status glc_device_count(int *count);
/* Create a context on CUDA device `device`.  Builds the reference's tables on the host with the
 * host libm (MdctTables::new, src/codec.rs:326-356) and uploads them. */
glc_status glc_ctx_create(int device, glc_mode mode, glc_ctx **out);
void glc_ctx_destroy(glc_ctx *ctx);
/* Tuning knobs: `reserved` is ignored (it once selected kernel variants at run time; variants are
 * compile-time choices now); wave_rows = frame-channel rows per encode
 * pipeline wave, 0 = automatic (multiples of 37 row tiles). */
glc_status glc_ctx_set_tuning(glc_ctx *ctx, int reserved, uint64_t wave_rows);
/* Pinned host memory for callers that want zero-staging transfers (optional). */
glc_status glc_host_alloc(glc_ctx *ctx, size_t bytes, void **out);
void glc_host_free(glc_ctx *ctx, void *p);
/* Release any pointer returned through a `**out` byte/float buffer of this API. */
void glc_free(glc_ctx *ctx, void *p);

/* -------------------------------------------------------------------- codec */

/* Encoder::new(sample_rate)                                       src/codec.rs:406-418 */
glc_status glc_encoder_new(glc_ctx *ctx, uint32_t sample_rate, glc_encoder **out);
void glc_encoder_free(glc_encoder *enc);
/* Encoder::encode(&mut self, samples, channels) -> EncodedAudio     src/codec.rs:421-565 */
glc_status glc_encode(glc_encoder *enc, const float *pcm, uint64_t n_samples, uint16_t channels,
                      glc_encoded **out);
/* Many Encoder::encode calls in one device pass (files batched into one launch sequence);
 * out[i] corresponds to pcm[i].  The data-parallel entry point used for sharding by file. */
glc_status glc_encode_batch(glc_encoder *enc, uint32_t n_files, const float *const *pcm,
                            const uint64_t *n_samples, const uint16_t *channels,
                            glc_encoded **out /* [n_files] */);
/* Integer-PCM ingest ("next" row of SURVEY.md 8f): Encoder::encode over the samples that
 * audio::load_wav / load_flac would have produced from integer sources, `s as f32 / 2^(bits-1)`
 * (src/audio.rs:51-59, 76-80).  The integers cross PCIe (half the bytes for 16-bit sources) and the
 * conversion runs on the device; the result is bit-identical to glc_encode over the converted f32. */
glc_status glc_encode_i16(glc_encoder *enc, const int16_t *pcm, uint64_t n_samples, uint16_t channels,
                          glc_encoded **out);
glc_status glc_encode_batch_i16(glc_encoder *enc, uint32_t n_files, const int16_t *const *pcm,
                                const uint64_t *n_samples, const uint16_t *channels,
                                glc_encoded **out /* [n_files] */);
/* 8..32-bit samples in an int32 container (24-bit WAV/FLAC as hound/claxon deliver them). */
glc_status glc_encode_i32(glc_encoder *enc, const int32_t *pcm, uint64_t n_samples, uint16_t channels,
                          uint32_t bits_per_sample, glc_encoded **out);
void glc_encoded_free(glc_ctx *ctx, glc_encoded *e);

/* Decoder::new(channels, sample_rate)                              src/codec.rs:581-592
 * (both arguments are informational in the reference too: the stream header wins, :598) */
glc_status glc_decoder_new(glc_ctx *ctx, uint32_t channels, uint32_t sample_rate,
                           glc_decoder **out);
void glc_decoder_free(glc_decoder *dec);
/* Decoder::decode(&mut self, &EncodedAudio, None) -> Vec<f32>      src/codec.rs:744-768
 * (gapless trim applied: drops encoder_delay interleaved VALUES, truncates to original_length) */
glc_status glc_decode(glc_decoder *dec, const glc_encoded *enc, float **pcm, uint64_t *n_samples);
/* Concatenation of every AudioChunk of decode_streaming (no trim)   src/codec.rs:595-741 */
glc_status glc_decode_untrimmed(glc_decoder *dec, const glc_encoded *enc, float **pcm,
                                uint64_t *n_samples);
glc_status glc_decode_batch(glc_decoder *dec, uint32_t n_files, const glc_encoded *const *enc,
                            float **pcm /* [n_files] */, uint64_t *n_samples /* [n_files] */);

/* 16-bit PCM output ("next" row of SURVEY.md 8f, WAV side): what `glc -d file.glc` writes by default
 * (src/main.rs:95-105: Decoder::decode, then audio::export_to_wav), i.e. the decoded samples after
 * convert_f32_to_i16, `(s * 32767.0).clamp(-32768.0, 32767.0) as i16` (src/audio.rs:11-16).  The
 * conversion runs on the device; the result equals that conversion applied to glc_decode's output and
 * half the bytes cross PCIe.  Release with glc_free. */
glc_status glc_decode_i16(glc_decoder *dec, const glc_encoded *enc, int16_t **pcm, uint64_t *n_samples);
glc_status glc_decode_batch_i16(glc_decoder *dec, uint32_t n_files, const glc_encoded *const *enc,
                                int16_t **pcm /* [n_files] */, uint64_t *n_samples /* [n_files] */);

/* Decoder::decode_streaming: pull API.  Each _next yields one AudioChunk: exactly
 * FRAMES_PER_CHUNK*HOP_SIZE*channels values while more frames remain, then the tail with
 * is_last = 1 (src/codec.rs:708-717, :723-732).  `progress_percent` is the value the reference
 * sends as Progress::Decoding before that chunk (:712), or -1 for the last chunk.  The chunk
 * memory belongs to the stream and is valid until the next _next/_close.  Each _next decodes only
 * its own frames on the device (incremental); `enc` is borrowed and must outlive the stream. */
glc_status glc_decode_stream_open(glc_decoder *dec, const glc_encoded *enc, glc_stream **out);
glc_status glc_decode_stream_next(glc_stream *s, const float **samples, uint64_t *n_samples,
                                  int *is_last, float *progress_percent);
void glc_decode_stream_close(glc_stream *s);

/* --------------------------------------------------------------------- flac */

/* flac::encode_flac_with_level(samples, sample_rate, channels, level) -> Vec<u8>
 *                                                                   src/flac.rs:947-1052
 * flac::encode_flac is level 5 (:1055-1062); export_to_flac* write these bytes to a file. */
glc_status glc_flac_encode(glc_ctx *ctx, const float *pcm, uint64_t n_samples, uint32_t sample_rate,
                           uint16_t channels, uint8_t level, uint8_t **bytes, uint64_t *len);
glc_status glc_flac_encode_batch(glc_ctx *ctx, uint32_t n_files, const float *const *pcm,
                                 const uint64_t *n_samples, const uint32_t *sample_rate,
                                 const uint16_t *channels, uint8_t level, uint8_t **bytes,
                                 uint64_t *len);

/* Fused decode -> FLAC ("next" row of SURVEY.md 8f): what `glc -d file.glc --flac-level N` does
 * (src/main.rs:55-113: Decoder::decode, then flac::export_to_flac_with_level) with the decoded PCM
 * kept in HBM between the two steps.  Bytes are identical to glc_flac_encode over glc_decode's output. */
glc_status glc_decode_to_flac(glc_decoder *dec, const glc_encoded *enc, uint8_t level, uint8_t **bytes,
                              uint64_t *len);
glc_status glc_decode_to_flac_batch(glc_decoder *dec, uint32_t n_files, const glc_encoded *const *enc,
                                    uint8_t level, uint8_t **bytes /* [n_files] */, uint64_t *len /* [n_files] */);

/* ------------------------------------------------- one call, several devices */

/* The path shards by file with no exchange step: the reference loops over files (src/main.rs:546-583)
 * and parallelises inside a file (rayon, src/codec.rs:462, 620).  These calls split the files of ONE
 * batch over n_shards encoders / decoders / contexts (normally one per GPU of the box, each made from its
 * own glc_ctx), run the ordinary batch call of every shard on its own host thread and hand the outputs
 * back in input order -- a host-side gather, no collective.  shard_of[i] (required) receives the shard
 * that processed file i: its context owns out[i] / pcm[i] / bytes[i] and must be passed to
 * glc_encoded_free / glc_free.  Every file is validated before any work is planned; on failure nothing is
 * returned (outputs that other shards had already produced are released by the call). */
glc_status glc_plan_shards(uint32_t n_files, const uint64_t *weights, uint32_t n_shards, uint32_t *shard_of);
glc_status glc_encode_batch_sharded(glc_encoder *const *encs, uint32_t n_shards, uint32_t n_files,
                                    const float *const *pcm, const uint64_t *n_samples, const uint16_t *channels,
                                    glc_encoded **out /* [n_files] */, uint32_t *shard_of /* [n_files] */);
glc_status glc_decode_batch_sharded(glc_decoder *const *decs, uint32_t n_shards, uint32_t n_files,
                                    const glc_encoded *const *enc, float **pcm /* [n_files] */,
                                    uint64_t *n_samples /* [n_files] */, uint32_t *shard_of /* [n_files] */);
glc_status glc_flac_encode_batch_sharded(glc_ctx *const *ctxs, uint32_t n_shards, uint32_t n_files,
                                         const float *const *pcm, const uint64_t *n_samples,
                                         const uint32_t *sample_rate, const uint16_t *channels, uint8_t level,
                                         uint8_t **bytes /* [n_files] */, uint64_t *len /* [n_files] */,
                                         uint32_t *shard_of /* [n_files] */);

/* ---------------------------------------------------------------- container */

/* save_encoded / load_encoded byte image (bincode 1.3)            src/codec.rs:774-786 */
glc_status glc_encoded_to_bincode(glc_ctx *ctx, const glc_encoded *e, uint8_t **bytes,
                                  uint64_t *len);
glc_status glc_encoded_from_bincode(glc_ctx *ctx, const uint8_t *bytes, uint64_t len,
                                    glc_encoded **out);

/* -------------------------------------------- device-resident + measurement */

/* Kernel ids for glc_stats. */
enum
{
    GLC_K_MDCT_EXACT = 0,   /* direct-form MDCT contraction (dominant encode kernel) */
    GLC_K_QUANT_PACK = 1,   /* scale, masking thresholds, quantize, ordered compaction, raw decision */
    GLC_K_SCAN = 2,
    GLC_K_GATHER = 3,       /* stream compaction + raw-PCM frames */
    GLC_K_DEQUANT = 4,      /* row selection, per-tile step masks and stage list, sparse -> A stages (4 launches) */
    GLC_K_IMDCT_EXACT = 5,  /* direct IMDCT (sparse per 2-row warp) + synthesis window (dominant decode kernel) */
    GLC_K_OLA = 6,          /* raw frames, overlap-add, interleave */
    GLC_K_FLAC_BLOCK = 7,   /* one CTA per FLAC block */
    GLC_K_FLAC_GATHER = 8,
    GLC_K_MISC = 9,
    GLC_K_WINDOW_TILE = 10, /* padding + window -> tiled A operand of the MDCT */
    GLC_K_FAST_ENCODE = 11, /* FAST mode: fused window + FFT MDCT + quantise + pack */
    GLC_K_FAST_DECODE = 12, /* FAST mode: fused dequantise + FFT IMDCT + synthesis window */
    GLC_K_COUNT = 13
};

typedef struct glc_stats
{
    uint64_t launches[GLC_K_COUNT]; /* kernels launched since the last reset */
    double kernel_ms[GLC_K_COUNT];  /* CUDA-event time per kernel id; filled only while timing is on */
    uint64_t h2d_bytes, d2h_bytes;  /* bytes moved by cudaMemcpyAsync since the last reset */
    uint64_t staged_bytes;          /* of h2d_bytes: came from pageable memory through the pinned staging ring */
    uint64_t pinned_allocs, pinned_alloc_bytes; /* cudaHostAlloc calls made by the pinned pool (growth, not reuse) */
    uint64_t dev_allocs, dev_alloc_bytes;       /* cudaMalloc calls made by the device pool */
} glc_stats;

void glc_stats_reset(glc_ctx *ctx);
void glc_stats_get(glc_ctx *ctx, glc_stats *out);
/* When on, every kernel launch is bracketed by CUDA events on its stream (adds a sync at
 * glc_stats_get); off by default. */
void glc_stats_enable_kernel_timing(glc_ctx *ctx, int on);

/* Device-resident codec pass used for the kernel-only figure of bench.py: the PCM already lives
 * in HBM (glc_dev_upload), the encoded stream stays in HBM, and so does the decoded PCM.
 * glc_dev_* never touch host buffers inside the call. */
typedef struct glc_dev_pcm glc_dev_pcm;
typedef struct glc_dev_encoded glc_dev_encoded;
glc_status glc_dev_upload(glc_ctx *ctx, const float *pcm, uint64_t n_samples, uint16_t channels,
                          glc_dev_pcm **out);
void glc_dev_pcm_free(glc_dev_pcm *p);
glc_status glc_dev_encode(glc_encoder *enc, const glc_dev_pcm *pcm, glc_dev_encoded **out);
glc_status glc_dev_decode(glc_decoder *dec, const glc_dev_encoded *enc, glc_dev_pcm **out);
glc_status glc_dev_encoded_download(const glc_dev_encoded *enc, glc_encoded **out);
glc_status glc_dev_pcm_download(const glc_dev_pcm *p, float **pcm, uint64_t *n_samples);
void glc_dev_encoded_free(glc_dev_encoded *e);
/* CUDA-event stopwatch on the context's compute stream (begin records, end records + syncs). */
glc_status glc_timer_begin(glc_ctx *ctx);
glc_status glc_timer_end(glc_ctx *ctx, float *elapsed_ms);
glc_status glc_ctx_sync(glc_ctx *ctx);
/* Writes `bytes` of HBM with a fill kernel (L2 flush helper for benchmarks, > 126 MB). */
glc_status glc_flush_l2(glc_ctx *ctx);
/* Pure-DMA probe: moves h2d_bytes host -> device and d2h_bytes device -> host between pinned host memory and
 * HBM, both directions at once (on two streams) when `concurrent` is non-zero, one after the other otherwise,
 * with no kernel in between; *elapsed_ms is measured with CUDA events on the copy streams (from the common start
 * to the later of the two ends).  The floor any end-to-end
 * figure of this path can reach on this box: bench.py runs it on every rank at the same time. */
glc_status glc_dma_probe(glc_ctx *ctx, uint64_t h2d_bytes, uint64_t d2h_bytes, int concurrent, float *elapsed_ms);
/* FP32 multiply-then-add micro-benchmark (independent chains on every SM; packed = 0: scalar FMUL + FADD,
 * packed != 0: the same chains as f32x2 instructions, what the EXACT contractions issue): returns achieved
 * 1e12 lane-operations per second; the measured roof for the EXACT transform kernels. */
glc_status glc_measure_fp32_issue(glc_ctx *ctx, int packed, double *tera_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GLC_H */
