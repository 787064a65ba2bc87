// glc_internal.cuh -- internal declarations shared by the translation units of libglc_b200.so.
// Nothing here is part of the ABI (include/glc.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <vector>

#include "glc.h"

namespace glc
{

constexpr int kFrame = 2048; // FRAME_SIZE, reference src/codec.rs:15
constexpr int kHop = 1024;   // HOP_SIZE,   reference src/codec.rs:16
constexpr int kMaxBands = 64;

// ---- tiling of the EXACT transform kernels ----
constexpr int kBM = 128;     // frame-channels (rows) per CTA tile
constexpr int kBN = 128;     // outputs per CTA tile (coefficients for MDCT, samples for IMDCT)
constexpr int kKC = 32;      // reduction steps per pipeline stage
constexpr int kGemmThreads = 256;
// IMDCT (warp-sparse variant): tiles of kImdctBM compacted rows; a warp owns kImdctRowsPerWarp rows of
// the tile and all kImdctBN outputs of the CTA; a pipeline stage is kImdctKC consecutive coefficient
// indices and one A stage carries, after its [kImdctKC][kImdctBM] values, one step mask per warp.
// The shape is a compile-time choice (tools/imdct_sweep.sh builds and times variants on the GPU box).
// Measured on the hour-long bench signal: 2 rows x 256 outputs per warp, 32-index
// stages, 2-slot ring, 3 CTAs/SM = 13.85 ms; 1 row x 512 outputs (no union of index sets, 16 accumulators
// per thread, 1 CTA of 32 warps) = 15.6 ms; 2 rows x 512 outputs = 15.2 ms -- a thread's table operand is
// used by `rows per warp` rows only, and below two rows per warp the kernel is bound by shared-memory
// wavefronts (17 per 32 arithmetic instructions) instead of by issue slots.
#ifndef GLC_IMDCT_BN
#define GLC_IMDCT_BN 256
#endif
#ifndef GLC_IMDCT_KC
#define GLC_IMDCT_KC 32
#endif
#ifndef GLC_IMDCT_RW
#define GLC_IMDCT_RW 2
#endif
constexpr int kImdctBM = 32;
constexpr int kImdctBN = GLC_IMDCT_BN;
constexpr int kImdctKC = GLC_IMDCT_KC;
constexpr int kImdctStages = kHop / kImdctKC;                          // 32
constexpr int kImdctRowsPerWarp = GLC_IMDCT_RW;
constexpr int kImdctWarps = kImdctBM / kImdctRowsPerWarp;              // 16
constexpr int kImdctThreads = kImdctWarps * 32;                        // 512
// GLC_IMDCT_ROWMASK: one step mask per ROW instead of one per warp; the warp still walks the union of its rows'
// index sets (one table load per step) but executes a row's multiply-adds only at the steps that row holds
// (warp-uniform branches)
// GLC_MDCT_X2 / GLC_IMDCT_X2: the multiply-adds of the contraction as packed f32x2 instructions (two lanes per
// issue slot, each lane rounded exactly like the scalar FMUL / FADD pair; see mul_then_add_x2 in glc_exact_gemm.cu)
#ifndef GLC_MDCT_X2
#define GLC_MDCT_X2 1
#endif
#ifndef GLC_IMDCT_X2
#define GLC_IMDCT_X2 1
#endif
#ifndef GLC_IMDCT_ROWMASK
#define GLC_IMDCT_ROWMASK 0
#endif
constexpr int kImdctMasks = GLC_IMDCT_ROWMASK ? kImdctBM : kImdctWarps;
constexpr int kImdctAStageFloats = kImdctKC * kImdctBM + kImdctMasks;  // 1024 values + 16 masks = 4 160 B
constexpr size_t kImdctATileFloats = (size_t)kImdctStages * kImdctAStageFloats;

// One input file inside a batched encode (device copy lives in FileTable::d_files).
struct FileDesc
{
    uint64_t pcm_off;   // offset (floats) of the file's first interleaved sample in the PCM arena
    uint64_t len;       // samples per channel (L)
    uint64_t first_row; // first frame-channel row of this file in the batch
    uint64_t first_frame; // first frame index of this file in the batch-wide frame numbering
    uint32_t channels;
    uint32_t n_frames;
};

// Host-built constant tables (reference: MdctTables::new src/codec.rs:326-356 and
// PerceptualWeights::new :102-183).  Built with the host libm, never on the device.
struct HostTables
{
    float *cos_tab; // [1024][2048] reference layout tab[k*2048 + i]
    float window[kFrame];
    float norm;
    float noise_floor_factor; // powf(10, -48/20)
};

struct HostPerceptual
{
    float inv_w[kHop];        // 1 / max(w[k], 0.1)            src/codec.rs:228
    int32_t band_edges[kMaxBands];
    int32_t n_edges;
    float band_pf[kMaxBands]; // 1 / max(avg_weight, 0.1)      src/codec.rs:218-222
    float band_cnt[kMaxBands];
    float cf;                 // max(1 - 0.7, 0.01)            src/codec.rs:221
};

void build_host_tables(HostTables *t);
void free_host_tables(HostTables *t);
void build_host_perceptual(uint32_t sample_rate, HostPerceptual *p);
// Re-tile the cosine table for the two transform kernels: [n_block][stage][kKC][kBN]
void tile_table_for_mdct(const float *cos_tab, float *out);  // rows = i (2048), cols = k (1024)
void tile_table_for_imdct(const float *cos_tab, float *out); // [n_block][k][kImdctBN]

// Device-side perceptual model, passed by value-pointer to the quantize/pack kernel.
struct DevPerceptual
{
    float inv_w[kHop];
    int32_t band_edges[kMaxBands];
    float band_pf[kMaxBands];
    float band_cnt[kMaxBands];
    uint8_t band_of[kHop];    // band index of every bin
    int32_t n_edges;
    float cf;
    float noise_floor_factor;
    float pad;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: one process may drive several
// devices (glc_*_batch_sharded), so "configured" is remembered per device, not per process.
// `mask` is a static std::atomic<uint64_t> of the call site.
#define GLC_SET_MAX_DYN_SMEM_ONCE(kernel, bytes)                                                              \
    do                                                                                                        \
    {                                                                                                         \
        static std::atomic<unsigned long long> configured_mask_{0};                                           \
        int dev_ = 0;                                                                                         \
        cudaGetDevice(&dev_);                                                                                 \
        if (!((configured_mask_.load(std::memory_order_acquire) >> (dev_ & 63)) & 1ull))                      \
        {                                                                                                     \
            cudaError_t e_ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
            if (e_ != cudaSuccess)                                                                            \
                return e_;                                                                                    \
            configured_mask_.fetch_or(1ull << (dev_ & 63), std::memory_order_release);                        \
        }                                                                                                     \
    } while (0)

// ---- kernel launchers (each returns the launch error) ----
struct MdctLaunch
{
    const float *pcm_arena;
    const FileDesc *files;
    uint32_t n_files;
    uint64_t row_begin, row_end; // rows (frame-channels) of the batch handled by this launch
    const float *tab_tiled;      // [n_block][stage][kKC][kBN]
    const float *window;
    float norm;
    float *coefs;   // [n_rows][1024], indexed by absolute row
    float *a_tiles; // scratch for the windowed A operand: mdct_a_tile_floats(row_end - row_begin) floats
};
size_t mdct_a_tile_floats(uint64_t n_rows);
cudaError_t launch_window_tiles(const MdctLaunch &p, cudaStream_t s); // PCM -> a_tiles (padding + window)
cudaError_t launch_mdct_exact(const MdctLaunch &p, cudaStream_t s);   // a_tiles x table -> coefs

struct ImdctLaunch
{
    const float *a_tiles;      // [tile][stage]{[kImdctKC][kImdctBM] dequantised coefficients, [kImdctWarps] step masks}
    const uint8_t *stage_list; // [tile][kImdctStages] ascending indices of the stages that hold a coefficient
    const uint32_t *n_stages;  // [tile] entries of stage_list
    const uint32_t *n_tiles;   // device-side number of live tiles
    uint64_t max_slots;        // worst-case number of compacted rows (sizes the grid and `blocks`)
    const float *tab;          // tile_table_for_imdct layout
    const float *window;
    float norm;
    float *blocks; // [slot][2048] windowed IMDCT output
};
cudaError_t launch_imdct_exact(const ImdctLaunch &p, cudaStream_t s);

struct QuantPackLaunch
{
    const float *coefs; // [n_rows][1024], indexed by absolute row
    const FileDesc *files;
    uint32_t n_files;
    const uint64_t *first_group;     // [n_files] first frame group of each file (quant_groups_for, scanned)
    uint64_t group_begin, group_end; // groups handled by this launch
    const DevPerceptual *perc;
    glc_pair *slots;    // [n_rows][1024]
    uint32_t *nnz;      // [n_rows]
    float *scales;      // [n_rows]
    uint8_t *is_raw;    // [n_frames_total]
    uint32_t *raw_len;  // [n_frames_total] 0 or 2048*ch
};
uint64_t quant_groups_for(uint32_t n_frames, uint32_t channels);
uint32_t quant_frames_per_group(uint32_t channels);
cudaError_t launch_quant_pack(const QuantPackLaunch &p, cudaStream_t s);

// exclusive scans: u32 -> u64, n+1 outputs
// (`base`, nullable device pointer: running total before in[0]; may alias out[0])
cudaError_t launch_scan_u32_u64(const uint32_t *in, uint64_t *out, uint64_t n, cudaStream_t s,
                                const uint64_t *base = nullptr);

struct GatherLaunch
{
    const glc_pair *slots;
    const uint32_t *nnz;
    const uint64_t *pair_off;
    glc_pair *pairs;
    const uint8_t *is_raw;
    const uint64_t *raw_off;
    int16_t *raw;
    const float *pcm_arena;
    const FileDesc *files;
    uint32_t n_files;
    const float *window;
    uint64_t row_begin, row_end;     // rows whose pairs are compacted by this launch
    uint64_t frame_begin, frame_end; // frames whose raw bodies are produced by this launch
    // Host-bound encodes keep only two waves of compact output on the device (a ring): `pairs` / `raw` then point
    // at this wave's half and the wave's first offsets (read on the device: pair_off[row_begin],
    // raw_off[frame_begin]) are subtracted from every destination.  Null = absolute offsets.
    const uint64_t *pair_bias;
    const uint64_t *raw_bias;
};
cudaError_t launch_gather(const GatherLaunch &p, cudaStream_t s);

// Decode front end: which rows go through the IMDCT at all (sparse frames with at least one pair),
// their compaction into tiles of kImdctBM rows, and per tile the step masks / stage list / A stages of the IMDCT.
struct DequantLaunch
{
    const glc_pair *pairs;
    const uint64_t *pair_off; // [n_rows+1]
    const float *scales;
    const uint8_t *is_raw;    // [n_frames_total]
    const struct DecFileDesc *files;
    uint32_t n_files;
    uint64_t row_begin, row_end; // rows of this wave (absolute batch rows)
    uint64_t slot_base;       // first slot of this wave in the batch-wide `blocks` array
    int32_t *row_slot;        // [batch rows] absolute slot of a row or -1 (indexed by absolute row)
    // wave-local scratch, indexed from 0 = row_begin / tile 0 of the wave:
    uint32_t *flags;          // [n]   1 = row is transformed
    uint64_t *slot_off;       // [n+1] exclusive scan of flags
    uint32_t *active_rows;    // [n]   local slot -> absolute row
    uint32_t *n_tiles;        // [1]
    uint8_t *stage_list;      // [max_tiles][kImdctStages]
    uint32_t *n_stages;       // [max_tiles]
    float *a_tiles;           // [max_tiles][kImdctATileFloats]
};
cudaError_t launch_dequant(const DequantLaunch &p, cudaStream_t s); // flags -> scan -> scatter -> tiles

// One encoded stream inside a batched decode.
struct DecFileDesc
{
    uint64_t first_row;   // first frame-channel row in the batch
    uint64_t first_frame; // first frame in the batch-wide numbering
    uint64_t out_off;     // offset (floats) of this file's untrimmed PCM in the output arena
    uint64_t n_frames;
    uint32_t channels;
    uint32_t pad;
};

struct OlaLaunch
{
    const float *blocks;      // [slot][2048] windowed IMDCT output
    const int32_t *row_slot;  // [n_rows] slot of a row in `blocks`, -1 = all-zero block
    const uint8_t *is_raw;    // [n_frames_total]
    const uint64_t *raw_off;  // [n_frames_total+1]
    const int16_t *raw;
    const DecFileDesc *files;
    uint32_t n_files;
    uint64_t hop_begin, hop_end; // batch-wide hop ids (frame index + file index) produced by this launch
    float *out;               // per file interleaved, (n_frames+1)*1024*ch values at out_off
    uint32_t tile_channels;   // largest channel count in 3..8 of the batch (0 = none): sizes the shared-memory tile
};
cudaError_t launch_ola(const OlaLaunch &p, cudaStream_t s);

// ---- FAST transform mode (FFT-based, tolerance class; glc_fast_kernels.cu) ----
struct FastEncodeLaunch
{
    const float *pcm_arena;
    const FileDesc *files;
    uint32_t n_files;
    const uint64_t *first_group; // [n_files] first CTA group of each file (fast_groups_for per file, scanned)
    uint64_t group_begin, group_end;
    const float *window;
    const float2 *twiddles; // fast_twiddle_table
    float norm;
    const DevPerceptual *perc;
    // fast_encode_kernel: per-row / per-frame results + every group's compact block in its slot
    uint32_t *nnz;       // [n_rows]
    float *scales;       // [n_rows]
    uint8_t *is_raw;     // [n_frames_total]
    glc_pair *slots;     // [n_rows][1024]: a group's rows are its slot (pairs to the front, raw bodies to the back)
    uint32_t *grp_pairs; // [n_groups] pairs of the group's sparse frames
    uint32_t *grp_raw;   // [n_groups] frame-channels of the group's raw frames (units of 2048 i16)
    unsigned int *ticket; // group counter of THIS launch, zero at launch
    // fast_place_kernel: exclusive scans of the two totals -> final positions
    const uint64_t *grp_pair_off; // [n_groups + 1]
    const uint64_t *grp_raw_off;  // [n_groups + 1]
    uint64_t *pair_off;  // [n_rows + 1]
    uint64_t *raw_off;   // [n_frames_total + 1]
    glc_pair *pairs;
    int16_t *raw;
    // ring of two waves for host-bound encodes (see GatherLaunch): the placement subtracts
    // grp_pair_off[group_begin] / grp_raw_off[group_begin] from its destinations when set
    bool ring;
};
cudaError_t launch_fast_place(const FastEncodeLaunch &p, cudaStream_t s);
constexpr int kFastTwiddleFloats = 2 * (32 * 16 + 16);
void fast_twiddle_table(float norm, float *out); // host, double precision
uint64_t fast_groups_for(uint32_t n_frames, uint32_t channels);
cudaError_t launch_fast_encode(const FastEncodeLaunch &p, cudaStream_t s);

struct FastDecodeLaunch
{
    const glc_pair *pairs;
    const uint64_t *pair_off;
    const float *scales;
    const uint8_t *is_raw;
    const DecFileDesc *files;
    uint32_t n_files;
    uint64_t row_begin, row_end;
    const float *window;
    const float2 *twiddles;
    float norm;
    uint64_t slot_base; // first slot of this wave in `blocks` (a ring of two waves)
    int32_t *row_slot;  // [batch rows] slot_base + (row - row_begin) when transformed, -1 otherwise
    float *blocks;      // [2 waves of slots][2048]
};
cudaError_t launch_fast_decode(const FastDecodeLaunch &p, cudaStream_t s);

cudaError_t launch_pcm_convert(const void *stage, int elem_bytes, uint64_t off, uint64_t n, float inv_max, float *arena,
                               cudaStream_t s);
cudaError_t launch_pcm_to_i16(const float *in, int16_t *out, uint64_t n, cudaStream_t s); // WAV export conversion
cudaError_t launch_fill(float *p, uint64_t n, float v, cudaStream_t s);
cudaError_t launch_fp32_issue_bench(int packed, int iters, float *sink, int blocks, cudaStream_t s);

// ---- FLAC ----
struct FlacFileDesc
{
    uint64_t pcm_off;     // floats
    uint64_t n_samples;   // interleaved count
    uint64_t first_block; // index of this file's first block in the batch
    uint64_t i16_off;     // offset into the i16 arena
    uint32_t n_blocks;
    uint32_t block_size;
    uint32_t channels;
    uint32_t sample_rate;
};
struct FlacLaunch
{
    const float *pcm_arena;
    const FlacFileDesc *files;
    uint32_t n_files;
    uint64_t n_blocks_total;
    int level;
    uint32_t *frame_bytes; // [n_blocks_total] exact size of every frame
    int16_t *i16_arena;    // f32 -> i16 converted samples (for the host MD5)
};
cudaError_t launch_flac_measure(const FlacLaunch &p, uint8_t *rice_k, uint32_t max_ch, uint32_t max_bs,
                                cudaStream_t s);
// words of global scratch the emit pass needs (0 when the bit buffer fits in shared memory)
size_t flac_emit_scratch_words(uint64_t n_blocks, uint32_t max_ch, uint32_t max_bs, uint32_t max_frame_bytes,
                               int sm_count);
cudaError_t launch_flac_emit(const FlacLaunch &p, const uint8_t *rice_k, uint32_t max_ch, uint32_t max_bs,
                             uint32_t max_frame_bytes, const uint64_t *frame_off, uint8_t *out_arena,
                             uint32_t *scratch, int sm_count, cudaStream_t s);
uint32_t flac_slot_bytes(uint32_t block_size, uint32_t channels);
// flac::encode_flac_with_level over host PCM (pcm != nullptr) or device-resident PCM (file i at d_base + d_off[i])
glc_status flac_encode_impl(glc_ctx *ctx, uint32_t n_files, const float *const *pcm, const float *d_base,
                            const uint64_t *d_off, const uint64_t *n_samples, const uint32_t *sample_rate,
                            const uint16_t *channels, uint8_t level, uint8_t **bytes, uint64_t *len);

// ---- host-side plumbing shared by the API translation units ----
glc_status set_error(glc_status st, const char *fmt, ...);
void *pinned_alloc(glc_ctx *ctx, size_t bytes);
void pinned_release(glc_ctx *ctx, void *p);
cudaError_t dev_alloc(glc_ctx *ctx, void **out, size_t bytes, cudaStream_t s); // context-owned device pool
void dev_free(glc_ctx *ctx, void *p, cudaStream_t s);
int ctx_device(glc_ctx *ctx);
glc_ctx *decoder_ctx(glc_decoder *dec);
cudaStream_t ctx_compute_stream(glc_ctx *ctx);
cudaStream_t ctx_d2h_stream(glc_ctx *ctx);
void ctx_count_launch(glc_ctx *ctx, int kernel_id, uint64_t n);
void ctx_count_bytes(glc_ctx *ctx, uint64_t h2d, uint64_t d2h);
// host -> device; pageable sources are staged through a ring of pinned chunks filled by a few host threads
cudaError_t ctx_h2d(glc_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t s);
void ctx_time_begin(glc_ctx *ctx, int kernel_id, void **token);
void ctx_time_end(glc_ctx *ctx, void *token);

// Host image of one or more encoded streams sharing allocations (ref-counted by their boxes).
struct EncodedBlock
{
    glc_ctx *ctx;
    std::atomic<int> refs; // the outputs of one batch call share a block and may be freed from different threads
    std::vector<void *> pinned;
    std::vector<void *> heap;
};
struct EncodedBox
{
    glc_encoded pub; // must stay first: the ABI hands out &pub
    EncodedBlock *blk;
};

} // namespace glc
