// glc_fast_kernels.cu -- FAST transform mode (GLC_MODE_FAST), decode side: FFT-based IMDCT fused with the
// dequantiser + synthesis window, sm_100a.  (The encode side is glc_fast_encode.cu; the transform notes
// below apply to both.)
//
// This is the "fused MDCT+quantize kernel" of BASELINE.json's north_star: window -> fold -> DCT-IV
// by a 512-point complex FFT (warp-level, register radix butterflies, one shared-memory exchange)
// -> scale / masking thresholds / quantise / ordered compaction, one pass over the PCM.  It computes
// the TRUE MDCT, whereas the reference multiplies by an f32 table that is up to 6.85e-4 away from
// the true basis and accumulates sequentially, so its parity class is
// TOLERANCE, not bit-exact: EXACT mode (glc_exact_gemm.cu) stays the parity-gated default.  The
// stream layout, frame counts, gapless metadata and sample counts are identical in both modes.
//
// Transform (N = 1024 coefficients from 2N windowed samples b[i] = x[i]*w[i]):
//   fold      u[m]      = -b[3N/2-1-m] - b[3N/2+m]          m <  N/2
//             u[N/2+m]  =  b[m]        - b[N-1-m]
//   DCT-IV    z[n] = (u[2n] + i u[N-1-2n]) * exp(-i pi (4n+1)/(4N)),  Z = FFT_512(z),
//             y[k] = Z[k] * exp(-i pi k/N),  X[2k] = Re y[k],  X[N-1-2k] = -Im y[k]
//   (the IMDCT is the transpose: v = DCT-IV(c), then the fold is undone with the same signs)
// FFT_512 = 16 x 32: n = n1 + 32 n2, k = k2 + 16 k1.  A warp transforms TWO frame-channels at once:
//   pass 1  lane = n1: 16-point FFT over n2 in registers (one per frame-channel), times
//           T[k2] = exp(-i pi (4 n1+1)/(4N)) * w512^(n1 k2) (per-lane constants), to shared memory;
//   pass 2  lane = (frame-channel, k2): 32-point FFT over n1 in registers, post-twiddle, results
//           to shared memory in natural coefficient order.
// Compiled with FMA contraction ON (this file only).
#include <math.h>

#include <algorithm>

#include "glc_fft_gen.cuh"
#include "glc_internal.cuh"

namespace glc
{

namespace
{

constexpr int kFastThreads = 128;            // 4 warps
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastFcs = 8;                  // frame-channels per CTA (4 pairs)
constexpr int kCoefStride = kHop;            // floats per frame-channel in the coefficient buffer

struct FastSmem
{
    // folded + windowed input as (u[2n], u[N-1-2n]) pairs, 4 KiB per frame-channel.  The two rows of
    // a pair are reused in place, first as the FFT exchange buffer ([fc][k2][n1 ^ k2], XOR-swizzled
    // instead of padded so that it fits exactly), then as the coefficient buffer ([fc][1024] floats).
    float2 u[kFastFcs][kHop / 2];
    float inv_w[kHop];
    float band_base[kFastWarps][kMaxBands];
    float band_base2[kFastWarps][kMaxBands]; // second frame-channel of the warp's pair
    float band_fac[kMaxBands];  // 0.01 * compression_factor * perceptual_factor        src/codec.rs:221-223
    float band_rcnt[kMaxBands]; // 1 / bins in the band
    int16_t band_lo[kMaxBands], band_hi[kMaxBands];
    uint8_t band_of[kHop];
    uint32_t frame_nnz[kFastFcs];
    int n_bands;
    // staging descriptors of the group's frame-channels
    const float *st_ptr[kFastFcs];
    long long st_base[kFastFcs];
    unsigned long long st_row[kFastFcs]; // output row (frame-channel index in the batch)
    int st_interior[kFastFcs];
    int st_lf[kFastFcs];                 // frame of the group the frame-channel belongs to
    int st_off[kFastFcs];                // element offset of the frame-channel's window from st_ptr[0]
    // the group being processed (filled by thread 0)
    const float *g_src;
    long long g_len;
    unsigned long long g_first_row, g_first_frame, g_frame0;
    uint32_t g_ch, g_n_frames;
};

// per-lane twiddle constants, computed once per thread
struct LaneTw
{
    float2 t[16]; // pass-1 output twiddle for lane n1, k2 = 0..15
    float2 qb;    // post-twiddle base for k2 = lane & 15, times `norm`
};

// The table is built on the host in double precision (fast_twiddle_table) and uploaded once per
// context: [32 lanes][16] pass-1 twiddles, then [16] post-twiddle bases (times norm).
__device__ __forceinline__ void load_lane_tw(LaneTw &tw, const float2 *table, int lane)
{
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2)
        tw.t[k2] = __ldg(table + lane * 16 + k2);
    tw.qb = __ldg(table + 32 * 16 + (lane & 15));
}

// DCT-IV of two frame-channels whose (u[2n], u[N-1-2n]) pairs are the two consecutive rows at `pair`
// (shared memory, 2 x 512 float2).  Results (times `norm`) replace them in place: coefficients of the
// first at ((float*)pair)[0..1023], of the second at [1024..2047].  All 32 lanes must call.
__device__ __forceinline__ void dct4_pair(float2 *pair, const LaneTw &tw, int lane)
{
    using namespace fastfft;
    // ---- pass 1: lane = n1, 16-point FFT over n2 for each of the two inputs (rolled: the kernel is
    //      instruction-cache bound, one copy of the butterfly code is enough) ----
#pragma unroll 1
    for (int f = 0; f < 2; ++f)
    {
        float2 *u = pair + f * (kHop / 2);
        float re[16], im[16];
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2)
        {
            const float2 v = u[lane + 32 * n2];
            re[n2] = v.x * kPreStepRe[n2] - v.y * kPreStepIm[n2];
            im[n2] = v.x * kPreStepIm[n2] + v.y * kPreStepRe[n2];
        }
        fft16(re, im);
        __syncwarp(); // every lane holds its column: the row can be overwritten
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2)
        {
            const int s = kBitrev16[k2];
            u[k2 * 32 + (lane ^ k2)] =
                make_float2(re[s] * tw.t[k2].x - im[s] * tw.t[k2].y, re[s] * tw.t[k2].y + im[s] * tw.t[k2].x);
        }
    }
    __syncwarp();
    // ---- pass 2: lane = (f, k2), 32-point FFT over n1 ----
    const int f = lane >> 4, k2 = lane & 15;
    float re[32], im[32];
    {
        const float2 *x = pair + f * (kHop / 2) + k2 * 32;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1)
        {
            const float2 v = x[n1 ^ k2];
            re[n1] = v.x;
            im[n1] = v.y;
        }
    }
    fft32(re, im);
    __syncwarp(); // every lane has its column: the buffer can be overwritten with coefficients
    float *coef = reinterpret_cast<float *>(pair) + f * kCoefStride;
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1)
    {
        const int s = kBitrev32[k1];
        // post-twiddle exp(-i pi k/N) * norm, k = k2 + 16 k1
        const float qr = kPostStepRe[k1] * tw.qb.x - kPostStepIm[k1] * tw.qb.y;
        const float qi = kPostStepRe[k1] * tw.qb.y + kPostStepIm[k1] * tw.qb.x;
        const float yr = re[s] * qr - im[s] * qi;
        const float yi = re[s] * qi + im[s] * qr;
        const int k = k2 + 16 * k1;
        coef[2 * k] = yr;
        coef[kHop - 1 - 2 * k] = -yi;
    }
    __syncwarp();
}

// ------------------------------------------------------------------ decode

// One CTA per 8 rows: dequantise into the (c[2n], c[N-1-2n]) layout, DCT-IV, unfold, synthesis window.
__global__ void __launch_bounds__(kFastThreads, 5) fast_decode_kernel(const FastDecodeLaunch p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem &sm = *reinterpret_cast<FastSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t row0 = p.row_begin + (uint64_t)blockIdx.x * kFastFcs;
    if (row0 >= p.row_end)
        return;
    const uint32_t rows_here = (uint32_t)min((uint64_t)kFastFcs, p.row_end - row0);
    __shared__ int s_live[kFastFcs];

    // which rows are transformed: sparse frames with at least one pair
    if (tid < kFastFcs)
    {
        int live = 0;
        if ((uint32_t)tid < rows_here)
        {
            const uint64_t row = row0 + tid;
            uint32_t lo = 0, hi = p.n_files - 1;
            while (lo < hi)
            {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (p.files[mid].first_row <= row)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            const DecFileDesc &fd = p.files[lo];
            const uint64_t frame = fd.first_frame + (row - fd.first_row) / fd.channels;
            live = (!p.is_raw[frame] && p.pair_off[row + 1] > p.pair_off[row]) ? 1 : 0;
            p.row_slot[row] = live ? (int32_t)(p.slot_base + (row - p.row_begin)) : -1;
        }
        s_live[tid] = live;
    }
    for (uint32_t e = tid; e < ((rows_here + 1u) & ~1u) * (kHop / 2); e += kFastThreads)
        sm.u[e >> 9][e & 511] = make_float2(0.f, 0.f);
    __syncthreads();
    // dequantise (src/codec.rs:651-665).  "Later duplicates overwrite": lane-ordered replay when the
    // indices are not strictly ascending.
    for (uint32_t r = warp; r < rows_here; r += kFastWarps)
    {
        if (!s_live[r])
            continue;
        const uint64_t row = row0 + r;
        const uint64_t b = p.pair_off[row];
        const uint32_t n = (uint32_t)(p.pair_off[row + 1] - b);
        const glc_pair *pr = p.pairs + b;
        const float scale = fmaxf(p.scales[row], 1e-12f) * (1.0f / 32768.0f);
        bool ascending = true;
        for (uint32_t j = lane; j + 1 < n; j += 32)
            ascending = ascending && (pr[j].idx < pr[j + 1].idx);
        ascending = __all_sync(0xffffffffu, ascending);
        float *uf = reinterpret_cast<float *>(sm.u[r]);
        auto put = [&](uint32_t j) {
            const glc_pair q = pr[j];
            if (q.idx < kHop)
            {
                const uint32_t k = q.idx;
                // c[2n] -> u[n].x ; c[N-1-2n] -> u[n].y
                const uint32_t slot = (k & 1u) ? (((kHop - 1 - k) >> 1) * 2 + 1) : ((k >> 1) * 2);
                uf[slot] = (float)q.q * scale;
            }
        };
        if (ascending)
            for (uint32_t j = lane; j < n; j += 32)
                put(j);
        else if (lane == 0)
            for (uint32_t j = 0; j < n; ++j)
                put(j);
    }
    __syncthreads();
    LaneTw tw;
    load_lane_tw(tw, p.twiddles, lane);
    for (uint32_t prn = warp; prn * 2 < rows_here; prn += kFastWarps)
    {
        const uint32_t fa = prn * 2, fb = min(prn * 2 + 1, rows_here - 1);
        if (!s_live[fa] && !s_live[fb])
            continue;
        dct4_pair(sm.u[fa], tw, lane);
        const uint32_t n_here = (prn * 2 + 1 < rows_here) ? 2u : 1u;
        for (uint32_t h = 0; h < n_here; ++h)
        {
            if (!s_live[fa + h])
                continue;
            const float *v = reinterpret_cast<const float *>(sm.u[fa]) + h * kCoefStride;
            float *out = p.blocks + (p.slot_base + (row0 - p.row_begin) + fa + h) * kFrame;
            // unfold (transpose of the fold) + synthesis window           src/codec.rs:672-675
            // four consecutive outputs per lane: [0,512) = v[512+i], [512,1536) = -v[1535-i] (read reversed),
            // [1536,2048) = -v[i-1536]; the segment is uniform over the warp in every iteration
#pragma unroll 4
            for (int it = 0; it < kFrame / 128; ++it)
            {
                const int i0 = it * 128 + lane * 4;
                const float4 w = __ldg(reinterpret_cast<const float4 *>(p.window + i0));
                float4 r;
                if (it < 4)
                {
                    const float4 q = *reinterpret_cast<const float4 *>(v + 512 + i0);
                    r = make_float4(q.x * w.x, q.y * w.y, q.z * w.z, q.w * w.w);
                }
                else if (it < 12)
                {
                    const float4 q = *reinterpret_cast<const float4 *>(v + 1532 - i0);
                    r = make_float4(-q.w * w.x, -q.z * w.y, -q.y * w.z, -q.x * w.w);
                }
                else
                {
                    const float4 q = *reinterpret_cast<const float4 *>(v + i0 - 1536);
                    r = make_float4(-q.x * w.x, -q.y * w.y, -q.z * w.z, -q.w * w.w);
                }
                *reinterpret_cast<float4 *>(out + i0) = r;
            }
        }
        __syncwarp();
    }
}

} // namespace

void fast_twiddle_table(float norm, float *out /* kFastTwiddleFloats */)
{
    const double pi = 3.14159265358979323846;
    for (int n1 = 0; n1 < 32; ++n1)
        for (int k2 = 0; k2 < 16; ++k2)
        {
            // exp(-i pi (4 n1 + 1)/4096) * exp(-2 pi i n1 k2 / 512)
            const double a = -pi * ((4.0 * n1 + 1.0) / 4096.0 + (2.0 * n1 * k2) / 512.0);
            out[2 * (n1 * 16 + k2)] = (float)cos(a);
            out[2 * (n1 * 16 + k2) + 1] = (float)sin(a);
        }
    for (int k2 = 0; k2 < 16; ++k2)
    {
        const double a = -pi * k2 / 1024.0; // exp(-i pi k2/N), times norm
        out[2 * (512 + k2)] = (float)(cos(a) * (double)norm);
        out[2 * (512 + k2) + 1] = (float)(sin(a) * (double)norm);
    }
}

cudaError_t launch_fast_decode(const FastDecodeLaunch &p, cudaStream_t s)
{
    if (p.row_end <= p.row_begin)
        return cudaSuccess;
    GLC_SET_MAX_DYN_SMEM_ONCE(fast_decode_kernel, sizeof(FastSmem));
    const uint64_t n = p.row_end - p.row_begin;
    fast_decode_kernel<<<(unsigned)((n + kFastFcs - 1) / kFastFcs), kFastThreads, sizeof(FastSmem), s>>>(p);
    return cudaGetLastError();
}

} // namespace glc
