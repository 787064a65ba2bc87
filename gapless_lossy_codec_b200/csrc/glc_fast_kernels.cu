// glc_fast_kernels.cu -- FAST transform mode (GLC_MODE_FAST): FFT-based MDCT / IMDCT fused with the
// quantiser (encode) and the dequantiser + synthesis window (decode), sm_100a.
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

struct GroupGeom
{
    uint32_t file;
    uint32_t n_frames; // frames of this group (<= frames_per_group)
    uint64_t frame0;   // first frame of the group, local to the file
};

__device__ __forceinline__ GroupGeom locate_group(const uint64_t *first_group, const FileDesc *files, uint32_t n_files,
                                                  uint64_t g)
{
    uint32_t lo = 0, hi = n_files - 1;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (first_group[mid] <= g)
            lo = mid;
        else
            hi = mid - 1;
    }
    GroupGeom gg;
    gg.file = lo;
    const uint32_t ch = files[lo].channels;
    const uint32_t fpg = kFastFcs / ch ? kFastFcs / ch : 1u;
    gg.frame0 = (g - first_group[lo]) * fpg;
    const uint64_t left = files[lo].n_frames - gg.frame0;
    gg.n_frames = (uint32_t)(left < fpg ? left : fpg);
    return gg;
}

// ------------------------------------------------------------------ encode

// Persistent CTAs: the perceptual tables and the lane twiddles are loaded once, then the CTA walks
// frame groups with a grid stride.
__global__ void __launch_bounds__(kFastThreads, 4) fast_encode_kernel(const FastEncodeLaunch p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastSmem &sm = *reinterpret_cast<FastSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const DevPerceptual &pm = *p.perc;
        for (int k = tid; k < kHop; k += kFastThreads)
        {
            sm.inv_w[k] = pm.inv_w[k];
            sm.band_of[k] = pm.band_of[k];
        }
        const int nb = pm.n_edges - 1;
        if (tid < nb)
        {
            sm.band_lo[tid] = (int16_t)pm.band_edges[tid];
            sm.band_hi[tid] = (int16_t)pm.band_edges[tid + 1];
            sm.band_fac[tid] = 0.01f * pm.cf * pm.band_pf[tid];
            sm.band_rcnt[tid] = 1.0f / pm.band_cnt[tid];
        }
        if (tid == 0)
            sm.n_bands = nb;
    }
    const float noise_floor_factor = p.perc->noise_floor_factor;
    LaneTw tw;
    load_lane_tw(tw, p.twiddles, lane);

    for (uint64_t g = p.group_begin + blockIdx.x; g < p.group_end; g += gridDim.x)
    {
        __syncthreads(); // tables ready / previous group done with shared memory
        if (tid == 0)
        {
            const GroupGeom gg0 = locate_group(p.first_group, p.files, p.n_files, g);
            const FileDesc &fd0 = p.files[gg0.file];
            sm.g_src = p.pcm_arena + fd0.pcm_off;
            sm.g_len = (long long)fd0.len;
            sm.g_first_row = fd0.first_row;
            sm.g_first_frame = fd0.first_frame;
            sm.g_frame0 = gg0.frame0;
            sm.g_ch = fd0.channels;
            sm.g_n_frames = gg0.n_frames;
        }
        if (tid < kFastFcs)
            sm.frame_nnz[tid] = 0;
        __syncthreads();
        struct
        {
            uint64_t frame0;
            uint32_t n_frames;
        } gg{sm.g_frame0, sm.g_n_frames};
        struct
        {
            uint64_t first_row, first_frame;
        } fd{sm.g_first_row, sm.g_first_frame};
        const uint32_t ch = sm.g_ch;
        const uint32_t n_fc = gg.n_frames * ch; // may exceed kFastFcs only when ch > 8 (then processed in rounds)
        const float *src = sm.g_src;
        const long long len = sm.g_len;
        const int n_bands = sm.n_bands;

        for (uint32_t fc0 = 0; fc0 < n_fc; fc0 += kFastFcs)
        {
            const uint32_t fcs_here = min((uint32_t)kFastFcs, n_fc - fc0);
            __syncthreads();
            // ---- stage: fold + window, (u[2n], u[N-1-2n]) per n, over the reference's padded signal
            //      (512 zeros + data + zero tail, src/codec.rs:433-447).  A thread takes the same n of
            //      every frame-channel of the group: the two window values are loaded once and up to
            //      32 PCM loads are in flight before the first use. ----
            if (tid < (int)fcs_here)
            {
                const uint32_t lf = (fc0 + tid) / ch, c = (fc0 + tid) - lf * ch;
                const long long base = (long long)((gg.frame0 + lf) * kHop) - kHop / 2; // sample index of i = 0
                sm.st_base[tid] = base;
                sm.st_ptr[tid] = src + base * (long long)ch + c; // only dereferenced in range
                sm.st_interior[tid] = base >= 0 && base + kFrame <= len; // no padding inside this frame
                sm.st_row[tid] = fd.first_row + (gg.frame0 + lf) * ch + c;
                sm.st_lf[tid] = (int)lf;
                const uint32_t lf0 = fc0 / ch, c0 = fc0 - lf0 * ch; // frame-channel 0 of this round
                sm.st_off[tid] = (int)(lf - lf0) * (int)(kHop * ch) + ((int)c - (int)c0);
            }
            __syncthreads();
            const int ich = (int)ch;
            // per-thread copies of the round's descriptors: one base pointer, 32-bit offsets, a validity mask
            const float *gb = sm.st_ptr[0];
            int foff[kFastFcs];
            unsigned interior_mask = 0;
#pragma unroll
            for (int fc = 0; fc < kFastFcs; ++fc)
            {
                foff[fc] = sm.st_off[fc];
                if ((uint32_t)fc < fcs_here && sm.st_interior[fc])
                    interior_mask |= 1u << fc;
            }
            // Common case (every frame-channel of a full group lies inside its file): no validity tests,
            // no per-frame-channel descriptors, 32 independent loads in flight per thread.
            const bool all_interior = fcs_here == (uint32_t)kFastFcs && interior_mask == (1u << kFastFcs) - 1u;
            if (all_interior)
            {
#pragma unroll 1
                for (int it = 0; it < (kHop / 2) / kFastThreads; ++it)
                {
                    const int n = tid + it * kFastThreads;
                    const float wa = __ldg(p.window + 512 + 2 * n);
                    const float wo = __ldg(p.window + (n < 256 ? 511 - 2 * n : 2 * n - 512));
                    const bool lo_half = n < 256;
                    // lo_half: u0 = -b[i0] wa - b[i1] wo, u1 =  b[i2] wo - b[i3] wa
                    // else   : u0 =  b[i0] wo - b[i1] wa, u1 = -b[i2] wa - b[i3] wo
                    const int i0 = lo_half ? 1535 - 2 * n : 2 * n - 512;
                    const int i1 = lo_half ? 1536 + 2 * n : 1535 - 2 * n;
                    const int i2 = lo_half ? 511 - 2 * n : 512 + 2 * n;
                    const int i3 = lo_half ? 512 + 2 * n : 2559 - 2 * n;
                    const float *q0 = gb + i0 * ich, *q1 = gb + i1 * ich, *q2 = gb + i2 * ich, *q3 = gb + i3 * ich;
                    const float c00 = lo_half ? -wa : wo, c01 = lo_half ? -wo : -wa;
                    const float c10 = lo_half ? wo : -wa, c11 = lo_half ? -wa : -wo;
                    float x[kFastFcs][4];
                    if (ich == 2)
                    {
                        // stereo: the two channels of a frame sit side by side, one 8-byte load serves both
#pragma unroll
                        for (int fc = 0; fc < kFastFcs; fc += 2)
                        {
                            const float2 a0 = __ldg(reinterpret_cast<const float2 *>(q0 + foff[fc]));
                            const float2 a1 = __ldg(reinterpret_cast<const float2 *>(q1 + foff[fc]));
                            const float2 a2 = __ldg(reinterpret_cast<const float2 *>(q2 + foff[fc]));
                            const float2 a3 = __ldg(reinterpret_cast<const float2 *>(q3 + foff[fc]));
                            x[fc][0] = a0.x, x[fc + 1][0] = a0.y;
                            x[fc][1] = a1.x, x[fc + 1][1] = a1.y;
                            x[fc][2] = a2.x, x[fc + 1][2] = a2.y;
                            x[fc][3] = a3.x, x[fc + 1][3] = a3.y;
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int fc = 0; fc < kFastFcs; ++fc)
                        {
                            x[fc][0] = __ldg(q0 + foff[fc]);
                            x[fc][1] = __ldg(q1 + foff[fc]);
                            x[fc][2] = __ldg(q2 + foff[fc]);
                            x[fc][3] = __ldg(q3 + foff[fc]);
                        }
                    }
#pragma unroll
                    for (int fc = 0; fc < kFastFcs; ++fc)
                        sm.u[fc][n] = make_float2(x[fc][0] * c00 + x[fc][1] * c01, x[fc][2] * c10 + x[fc][3] * c11);
                }
            }
            else
#pragma unroll 1
            for (int it = 0; it < (kHop / 2) / kFastThreads; ++it)
            {
                const int n = tid + it * kFastThreads;
                // window symmetry w[2047-i] = w[i] leaves two distinct values per n
                const float wa = __ldg(p.window + 512 + 2 * n);
                const float wo = __ldg(p.window + (n < 256 ? 511 - 2 * n : 2 * n - 512));
                int i0, i1, i2, i3;
                if (n < 256)
                {
                    i0 = 1535 - 2 * n; // u0 = -b[i0] - b[i1], u1 = b[i2] - b[i3]
                    i1 = 1536 + 2 * n;
                    i2 = 511 - 2 * n;
                    i3 = 512 + 2 * n;
                }
                else
                {
                    i0 = 2 * n - 512;  // u0 = b[i0] - b[i1], u1 = -b[i2] - b[i3]
                    i1 = 1535 - 2 * n;
                    i2 = 512 + 2 * n;
                    i3 = 2559 - 2 * n;
                }
                const int o0 = i0 * ich, o1 = i1 * ich, o2 = i2 * ich, o3 = i3 * ich;
                float x[kFastFcs][4];
#pragma unroll
                for (int fc = 0; fc < kFastFcs; ++fc)
                {
                    x[fc][0] = x[fc][1] = x[fc][2] = x[fc][3] = 0.0f;
                    if ((interior_mask >> fc) & 1u)
                    {
                        x[fc][0] = __ldg(gb + (foff[fc] + o0));
                        x[fc][1] = __ldg(gb + (foff[fc] + o1));
                        x[fc][2] = __ldg(gb + (foff[fc] + o2));
                        x[fc][3] = __ldg(gb + (foff[fc] + o3));
                    }
                }
#pragma unroll
                for (int fc = 0; fc < kFastFcs; ++fc)
                    if ((uint32_t)fc < fcs_here)
                    {
                        float u0, u1;
                        if (n < 256)
                        {
                            u0 = -x[fc][0] * wa - x[fc][1] * wo; // windows: w[i0] = wa, w[i1] = wo, w[i2] = wo, w[i3] = wa
                            u1 = x[fc][2] * wo - x[fc][3] * wa;
                        }
                        else
                        {
                            u0 = x[fc][0] * wo - x[fc][1] * wa;  // windows: w[i0] = wo, w[i1] = wa, w[i2] = wa, w[i3] = wo
                            u1 = -x[fc][2] * wa - x[fc][3] * wo;
                        }
                        sm.u[fc][n] = make_float2(u0, u1);
                    }
                // frames that touch the 512-zero lead-in or the zero tail (first / last frames of a file)
#pragma unroll 1
                for (uint32_t fc = 0; fc < fcs_here; ++fc)
                {
                    if (sm.st_interior[fc])
                        continue;
                    const float *fb = sm.st_ptr[fc];
                    const long long base = sm.st_base[fc];
                    auto smp = [&](int i) -> float {
                        const long long pos = base + i;
                        return (pos >= 0 && pos < len) ? __ldg(fb + (long long)i * ich) : 0.0f;
                    };
                    const float y0 = smp(i0), y1 = smp(i1), y2 = smp(i2), y3 = smp(i3);
                    float u0, u1;
                    if (n < 256)
                    {
                        u0 = -y0 * wa - y1 * wo;
                        u1 = y2 * wo - y3 * wa;
                    }
                    else
                    {
                        u0 = y0 * wo - y1 * wa;
                        u1 = -y2 * wa - y3 * wo;
                    }
                    sm.u[fc][n] = make_float2(u0, u1);
                }
            }
            if (fcs_here & 1u) // the transform works on pairs: an odd group gets a silent partner
                for (uint32_t n = tid; n < kHop / 2; n += kFastThreads)
                    sm.u[fcs_here][n] = make_float2(0.f, 0.f);
            __syncthreads();

            for (uint32_t pr = warp; pr * 2 < fcs_here; pr += kFastWarps)
            {
                const uint32_t fa = pr * 2;
                dct4_pair(sm.u[fa], tw, lane);
                // Both frame-channels of the pair are quantised in lock step: the two instruction streams
                // are independent, which doubles the work in flight of this latency-bound phase.  (For an
                // odd group the partner is silence; its results are not stored.)
                const bool two = pr * 2 + 1 < fcs_here;
                const float *coef0 = reinterpret_cast<const float *>(sm.u[fa]);
                const float *coef1 = coef0 + kCoefStride;
                float m0 = 0.0f, m1 = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                {
                    const float4 v0 = *reinterpret_cast<const float4 *>(coef0 + j * 128 + lane * 4);
                    const float4 v1 = *reinterpret_cast<const float4 *>(coef1 + j * 128 + lane * 4);
                    m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))));
                    m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))));
                }
                // scale = max|c| .max(1e-10)                                  src/codec.rs:488-489
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
                    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
                }
                const float gmax0 = fmaxf(m0, 1e-10f), gmax1 = fmaxf(m1, 1e-10f);
                const float scale0 = gmax0, scale1 = gmax1;
                // band energies -> per-band base threshold (already times scale, :288)   :205-224
                float *base0 = sm.band_base[warp], *base1 = sm.band_base2[warp];
                for (int b0 = 0; b0 < n_bands; b0 += 32)
                {
                    const int b = b0 + lane;
                    int lo = 0, hi = 0;
                    if (b < n_bands)
                    {
                        lo = sm.band_lo[b];
                        hi = sm.band_hi[b];
                    }
                    const bool wide = (hi - lo) > 32;
                    float acc0 = 0.0f, acc1 = 0.0f;
                    if (!wide)
                        for (int k = lo; k < hi; ++k)
                        {
                            acc0 = fmaf(coef0[k], coef0[k], acc0);
                            acc1 = fmaf(coef1[k], coef1[k], acc1);
                        }
                    unsigned wide_mask = __ballot_sync(0xffffffffu, wide);
                    while (wide_mask)
                    {
                        const int src_lane = __ffs(wide_mask) - 1;
                        wide_mask &= wide_mask - 1;
                        const int wlo = __shfl_sync(0xffffffffu, lo, src_lane), whi = __shfl_sync(0xffffffffu, hi, src_lane);
                        float p0 = 0.0f, p1 = 0.0f;
                        for (int k = wlo + lane; k < whi; k += 32)
                        {
                            p0 = fmaf(coef0[k], coef0[k], p0);
                            p1 = fmaf(coef1[k], coef1[k], p1);
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
                        {
                            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
                            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                        }
                        if (lane == src_lane)
                        {
                            acc0 = p0;
                            acc1 = p1;
                        }
                    }
                    if (b < n_bands)
                    {
                        const float f = sm.band_fac[b], rc = sm.band_rcnt[b];
                        base0[b] = sqrtf(acc0 * rc) * f * scale0;
                        base1[b] = sqrtf(acc1 * rc) * f * scale1;
                    }
                }
                __syncwarp();
                // thresholds + quantiser + ordered compaction               src/codec.rs:226-235, 277-307
                const float nf0 = noise_floor_factor * scale0, nf1 = noise_floor_factor * scale1;
                const float gate0 = 0.3f * gmax0, gate1 = 0.3f * gmax1;
                const float cap0 = 0.05f * gmax0 * scale0, cap1 = 0.05f * gmax1 * scale1;
                const float qmul0 = 32768.0f / scale0, qmul1 = 32768.0f / scale1;
                const uint64_t row0 = sm.st_row[fa], row1 = sm.st_row[two ? fa + 1 : fa];
                glc_pair *dst0 = p.slots + row0 * kHop, *dst1 = p.slots + row1 * kHop;
                uint32_t total0 = 0, total1 = 0;
                auto quant1 = [&](float v, float bs, float iw, float nf, float gate, float cap, float qmul) -> int {
                    const float a = fabsf(v);
                    float th = bs * iw;
                    th = fminf(th, a > gate ? cap : th);
                    const float x = v * qmul;
                    // round half away from zero (f32::round) as trunc(x + copysign(0.5, x))
                    const int q = __float2int_rz(fminf(fmaxf(x + copysignf(0.5f, x), -32768.0f), 32767.0f));
                    return a > fmaxf(nf, th) ? q : 0;
                };
#pragma unroll 1
                for (int j = 0; j < 8; ++j)
                {
                    const float4 c40 = *reinterpret_cast<const float4 *>(coef0 + j * 128 + lane * 4);
                    const float4 c41 = *reinterpret_cast<const float4 *>(coef1 + j * 128 + lane * 4);
                    const float4 iw4 = *reinterpret_cast<const float4 *>(sm.inv_w + j * 128 + lane * 4);
                    const uchar4 bo4 = *reinterpret_cast<const uchar4 *>(sm.band_of + j * 128 + lane * 4);
                    int q0[4], q1[4];
                    q0[0] = quant1(c40.x, base0[bo4.x], iw4.x, nf0, gate0, cap0, qmul0);
                    q1[0] = quant1(c41.x, base1[bo4.x], iw4.x, nf1, gate1, cap1, qmul1);
                    q0[1] = quant1(c40.y, base0[bo4.y], iw4.y, nf0, gate0, cap0, qmul0);
                    q1[1] = quant1(c41.y, base1[bo4.y], iw4.y, nf1, gate1, cap1, qmul1);
                    q0[2] = quant1(c40.z, base0[bo4.z], iw4.z, nf0, gate0, cap0, qmul0);
                    q1[2] = quant1(c41.z, base1[bo4.z], iw4.z, nf1, gate1, cap1, qmul1);
                    q0[3] = quant1(c40.w, base0[bo4.w], iw4.w, nf0, gate0, cap0, qmul0);
                    q1[3] = quant1(c41.w, base1[bo4.w], iw4.w, nf1, gate1, cap1, qmul1);
                    const uint32_t cnt0 = (q0[0] != 0) + (q0[1] != 0) + (q0[2] != 0) + (q0[3] != 0);
                    const uint32_t cnt1 = (q1[0] != 0) + (q1[1] != 0) + (q1[2] != 0) + (q1[3] != 0);
                    // one scan for both: the two counts (<= 4 per lane, <= 128 per warp) share a register
                    uint32_t incl = cnt0 | (cnt1 << 16);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        const uint32_t nn = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o)
                            incl += nn;
                    }
                    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                    uint32_t pos0 = total0 + (incl & 0xffffu) - cnt0, pos1 = total1 + (incl >> 16) - cnt1;
                    const uint32_t kbase = (uint32_t)(j * 128 + lane * 4);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                    {
                        if (q0[e] != 0)
                        {
                            glc_pair pr2;
                            pr2.idx = (uint16_t)(kbase + e);
                            pr2.q = (int16_t)q0[e];
                            dst0[pos0++] = pr2;
                        }
                        if (two && q1[e] != 0)
                        {
                            glc_pair pr2;
                            pr2.idx = (uint16_t)(kbase + e);
                            pr2.q = (int16_t)q1[e];
                            dst1[pos1++] = pr2;
                        }
                    }
                    total0 += tot & 0xffffu;
                    total1 += tot >> 16;
                }
                if (lane == 0)
                {
                    p.nnz[row0] = total0;
                    p.scales[row0] = scale0;
                    atomicAdd(&sm.frame_nnz[sm.st_lf[fa]], total0); // frame index < frames per group <= kFastFcs
                    if (two)
                    {
                        p.nnz[row1] = total1;
                        p.scales[row1] = scale1;
                        atomicAdd(&sm.frame_nnz[sm.st_lf[fa + 1]], total1);
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // raw-PCM / sparse decision per frame                                 src/codec.rs:505-540
        if (tid < gg.n_frames)
        {
            const uint64_t frame = fd.first_frame + gg.frame0 + tid;
            const uint64_t row_f = fd.first_row + (gg.frame0 + tid) * ch;
            const uint64_t compressed = (uint64_t)ch * 8 + (uint64_t)sm.frame_nnz[tid] * 4 + 8 + (uint64_t)ch * 4 + 64;
            const float rhs = (float)((uint64_t)kFrame * ch * 2) * 0.85f;
            const bool raw = (float)compressed >= rhs;
            p.is_raw[frame] = raw ? 1 : 0;
            p.raw_len[frame] = raw ? (uint32_t)(kFrame * ch) : 0u;
            if (raw)
                for (uint32_t c = 0; c < ch; ++c)
                {
                    p.nnz[row_f + c] = 0;
                    p.scales[row_f + c] = 0.0f;
                }
        }
    }
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
            p.row_slot[row] = live ? (int32_t)row : -1;
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
            float *out = p.blocks + (row0 + fa + h) * kFrame;
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

uint64_t fast_groups_for(uint32_t n_frames, uint32_t channels)
{
    const uint32_t fpg = kFastFcs / channels ? kFastFcs / channels : 1u;
    return ((uint64_t)n_frames + fpg - 1) / fpg;
}

cudaError_t launch_fast_encode(const FastEncodeLaunch &p, cudaStream_t s)
{
    if (p.group_end <= p.group_begin)
        return cudaSuccess;
    GLC_SET_MAX_DYN_SMEM_ONCE(fast_encode_kernel, sizeof(FastSmem));
    // persistent CTAs: 4 resident per SM, grid-stride over the frame groups
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t grid = std::min<uint64_t>(p.group_end - p.group_begin, (uint64_t)sms * 4);
    fast_encode_kernel<<<(unsigned)grid, kFastThreads, sizeof(FastSmem), s>>>(p);
    return cudaGetLastError();
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
