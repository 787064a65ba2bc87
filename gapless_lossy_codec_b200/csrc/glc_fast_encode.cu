// glc_fast_encode.cu -- FAST transform mode, encode side (GLC_MODE_FAST), sm_100a.
//
// The "fused MDCT+quantize kernel" of BASELINE.json's north_star, second version.  fast_encode_kernel makes ONE
// pass over the PCM: padding + window + fold + pre-twiddle, DCT-IV by a 512-point complex FFT, scale / masking
// thresholds, keep mask + counts, raw/sparse decision per frame (src/codec.rs:505-540), then quantise + ordered
// compaction (sparse frames) or the raw-PCM body (raw frames).  A frame group writes its result COMPACT into its
// own scratch slot; per-group totals are scanned (one small launch) and fast_place_kernel moves every group's
// block to its final position in the stream and fills the offset tables.  (A decoupled look-back inside the
// kernel was measured first: with persistent CTAs every group waits for the slowest of the ~600 groups in
// flight before it, 28 % of all stall samples; the two-step placement has no inter-CTA dependency at all.)
//
// It computes the TRUE MDCT, whereas the reference multiplies by an f32 table that is up to 6.85e-4 away
// from the true basis: parity class TOLERANCE (tests/test_gpu_fast.py,
// profiles/*fast_flip_report*).  Everything structural (frame counts, gapless metadata, stream layout,
// raw-frame bodies, the raw/sparse rule) is identical to EXACT mode.
//
// Mapping (persistent CTAs drawing groups from a ticket counter, 256 threads = 8 warps, 64 registers,
// 4 CTAs = 32 warps per SM):
//   group  = max(1, 8 / channels) consecutive frames of one file = up to 8 frame-channels (fcs)
//   stage  all threads: thread t makes z'[n] = (u[2n] + i u[N-1-2n]) exp(-i pi (4n+1)/(4N)) for n = t, t + 256
//          straight from the PCM arena (8-byte loads for stereo).  Every PCM sample is read by the two frames
//          that overlap it; the second read is an L1/L2 hit.
//   fft    one warp per fc: 512 = 16 (n2) x 32 (n1).  pass 1: lane = n1, 16-point FFT over n2 in
//          registers; pass 2: lane = (h, k2), 16-point FFT over the even (h = 0) or odd (h = 1) n1, the
//          last radix-2 step between the two half-warps by shuffle.  32 live values per lane instead of
//          the 64 of a 32-point register FFT: this is what lets the kernel run at 64 registers.
//   quant  one warp per fc: max + last-band energy in one pass, band energies one lane per band, keep mask,
//          counts (all 8 x 128-bin segments scanned at once: eight 8-bit counters packed in two registers), then
//          -- only for frames that stay sparse -- quantise straight into the group's slot in global memory.
//
// Compiled with FMA contraction ON (tolerance class).
#include <math.h>

#include <algorithm>

#include "glc_fft_gen.cuh"
#include "glc_internal.cuh"

namespace glc
{

namespace
{

constexpr int kThreads = 256;
constexpr int kFcs = 8; // frame-channels per round (one per warp)

// exp(-2 pi i kappa / 32), kappa = 0..15: the radix-2 step that joins the two half-warps of pass 2
__device__ constexpr float kW32Re[16] = {1.0f, 0.980785251f, 0.923879504f, 0.831469595f, 0.707106769f, 0.555570245f,
                                         0.382683426f, 0.195090324f, 0.0f, -0.195090324f, -0.382683426f, -0.555570245f,
                                         -0.707106769f, -0.831469595f, -0.923879504f, -0.980785251f};
__device__ constexpr float kW32Im[16] = {0.0f, -0.195090324f, -0.382683426f, -0.555570245f, -0.707106769f, -0.831469595f,
                                         -0.923879504f, -0.980785251f, -1.0f, -0.980785251f, -0.923879504f, -0.831469595f,
                                         -0.707106769f, -0.555570245f, -0.382683426f, -0.195090324f};

typedef unsigned long long u64;
constexpr int kNarrowMax = 32; // bands wider than this are summed by the whole warp

struct Smem
{
    float2 u[kFcs][kHop / 2]; // per fc: pre-twiddled folded input -> FFT exchange -> coefficients
    float inv_w[kHop];
    float2 pre[kHop / 2]; // exp(-i pi (4n+1)/(4N)): DCT-IV pre-twiddle, applied while staging
    float2 tw1[16][32];   // pass-1 output twiddles w512^(n1 k2), [k2][n1]
    float2 qb2[32];       // post-twiddle base of pass-2 lane (h, k2), times norm
    float band_base[kFcs][kMaxBands];
    float band_fac[kMaxBands];  // 0.01 * compression_factor * perceptual_factor        src/codec.rs:221-223
    float band_rcnt[kMaxBands]; // 1 / bins in the band
    int16_t band_lo[kMaxBands], band_hi[kMaxBands];
    uint8_t band_of[kHop];
    int n_bands;
    int top_lo;         // first bin of the last (widest) band: every bin from here on shares one threshold base
    int narrow_trip[2]; // widest narrow band among bands 0..31 / 32..63
    // the group / step being processed
    const float *g_src;
    long long g_len;
    u64 g_first_row, g_first_frame, g_frame0;
    uint32_t g_ch, g_n_frames;
    long long fc_base[kFcs]; // sample index (per channel) of i = 0 of the fc's frame
    uint32_t fc_chan[kFcs];
    uint32_t fc_lf[kFcs];
    uint32_t fc_nnz[kFcs];
    float fc_scale[kFcs];
    uint32_t fc_seg[kFcs][2]; // kept bins per 128-bin segment (eight 8-bit fields)
    uint32_t frame_nnz[kFcs];
    uint32_t frame_raw[kFcs];
    u64 g_next;    // the next group this CTA holds (ticket)
};

// DCT-IV of ONE frame-channel.  `u` (shared memory, 512 float2) holds z'[n] = (u[2n] + i u[N-1-2n]) *
// exp(-i pi (4n+1)/(4N)); the 1024 coefficients (times norm) replace it in place as floats.  All 32 lanes call.
__device__ __forceinline__ void dct4_single(float2 *u, const Smem &sm, int lane)
{
    using namespace fastfft;
    float re[16], im[16];
    // ---- pass 1: lane = n1, 16-point FFT over n2 ----
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2)
    {
        const float2 v = u[lane + 32 * n2];
        re[n2] = v.x;
        im[n2] = v.y;
    }
    fft16(re, im);
    __syncwarp(); // every lane holds its column: the row can be overwritten
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2)
    {
        const int s = kBitrev16[k2];
        const float2 t = sm.tw1[k2][lane];
        u[k2 * 32 + (lane ^ k2)] = make_float2(re[s] * t.x - im[s] * t.y, re[s] * t.y + im[s] * t.x);
    }
    __syncwarp();
    // ---- pass 2: lane = (h, k2): 16-point FFT over m of the samples n1 = 2m + h, then
    //      Z[k2 + 16 (kappa + 16 s)] = F0[kappa] + (-1)^s W32^kappa F1[kappa] across the half-warps ----
    const int h = lane >> 4, k2 = lane & 15;
    {
        const float2 *x = u + k2 * 32;
#pragma unroll
        for (int m = 0; m < 16; ++m)
        {
            const float2 v = x[(2 * m + h) ^ k2];
            re[m] = v.x;
            im[m] = v.y;
        }
    }
    fft16(re, im);
    __syncwarp(); // every lane has its inputs: the buffer can be overwritten with coefficients
    const float sg = h ? -1.0f : 1.0f;
    const float2 qb = sm.qb2[lane];
    float *coef = reinterpret_cast<float *>(u);
    float *c_even = coef + 2 * (k2 + 256 * h);             // coef[2k],     k = k2 + 16 kappa + 256 h
    float *c_odd = coef + (kHop - 1) - 2 * (k2 + 256 * h); // coef[N-1-2k]
#pragma unroll
    for (int kp = 0; kp < 16; ++kp)
    {
        const int s = kBitrev16[kp];
        // own contribution: F0 (h = 0) or W32^kappa F1 (h = 1)
        const float wr = h ? kW32Re[kp] : 1.0f, wi = h ? kW32Im[kp] : 0.0f;
        const float gr = re[s] * wr - im[s] * wi;
        const float gi = re[s] * wi + im[s] * wr;
        const float orr = __shfl_xor_sync(0xffffffffu, gr, 16);
        const float oi = __shfl_xor_sync(0xffffffffu, gi, 16);
        const float zr = fmaf(sg, gr, orr); // h = 0: F0 + G1;  h = 1: F0 - G1
        const float zi = fmaf(sg, gi, oi);
        // post-twiddle exp(-i pi k/N) * norm, k = k2 + 256 h + 16 kappa
        const float qr = kPostStepRe[kp] * qb.x - kPostStepIm[kp] * qb.y;
        const float qi = kPostStepRe[kp] * qb.y + kPostStepIm[kp] * qb.x;
        c_even[32 * kp] = zr * qr - zi * qi;
        c_odd[-32 * kp] = -(zr * qi + zi * qr);
    }
    __syncwarp();
}

// ((x*w)*32767).clamp(-32768,32767) as i16 (src/codec.rs:498-502): the float -> s16 conversion truncates,
// saturates and maps NaN to 0, exactly like Rust's `as i16` after the clamp
__device__ __forceinline__ short raw_i16(float x, float w)
{
    const float sc = __fmul_rn(__fmul_rn(x, w), 32767.0f);
    short q;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(q) : "f"(sc));
    return q;
}

// number of files whose first group is <= g, minus one: warp-cooperative 33-ary search (all lanes call)
__device__ __forceinline__ uint32_t locate_file(const uint64_t *first_group, uint32_t n_files, uint64_t g, int lane)
{
    uint32_t lo = 0, hi = n_files - 1; // answer in [lo, hi]
    while (lo < hi)
    {
        const uint32_t span = hi - lo; // probe lo + 1 + floor(lane * span / 32) in (lo, hi]
        const uint32_t probe = lo + 1 + (uint32_t)(((uint64_t)lane * span) >> 5);
        const bool le = probe <= hi && __ldg(first_group + probe) <= g;
        const unsigned bal = __ballot_sync(0xffffffffu, le);
        // probes are non-decreasing in lane, so `le` is a prefix of ones
        const int n_le = __popc(bal);
        const uint32_t new_lo = n_le ? __shfl_sync(0xffffffffu, probe, n_le - 1) : lo;
        const uint32_t new_hi = n_le < 32 ? __shfl_sync(0xffffffffu, probe, n_le) - 1 : hi;
        lo = new_lo;
        hi = new_hi < hi ? new_hi : hi;
    }
    return lo;
}

__global__ void __launch_bounds__(kThreads, 4) fast_encode_kernel(const FastEncodeLaunch p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- tables, once per CTA ----
    {
        const DevPerceptual &pm = *p.perc;
        for (int k = tid; k < kHop; k += kThreads)
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
        {
            sm.n_bands = nb;
            sm.top_lo = pm.band_edges[nb - 1];
            int t0 = 0, t1 = 0;
            for (int b = 0; b < nb; ++b)
            {
                const int wdt = pm.band_edges[b + 1] - pm.band_edges[b];
                if (wdt <= kNarrowMax)
                    (b < 32 ? t0 : t1) = max(b < 32 ? t0 : t1, wdt);
            }
            sm.narrow_trip[0] = t0;
            sm.narrow_trip[1] = t1;
        }
        // host table (fast_twiddle_table): [n1][k2] = exp(-i pi (4 n1+1)/4096) w512^(n1 k2), then [16] post-twiddle
        // bases (times norm).  The per-lane factor (entry k2 = 0) moves into the staging pre-twiddle.
        for (int e = tid; e < 512; e += kThreads)
        {
            const float2 t = __ldg(p.twiddles + e), l = __ldg(p.twiddles + (e & ~15));
            sm.tw1[e & 15][e >> 4] = make_float2(t.x * l.x + t.y * l.y, t.y * l.x - t.x * l.y); // t * conj(l)
        }
        for (int n = tid; n < kHop / 2; n += kThreads)
        {
            float sn, cs;
            sincospif(-(float)(4 * n + 1) * (1.0f / 4096.0f), &sn, &cs);
            sm.pre[n] = make_float2(cs, sn);
        }
        if (tid < 32)
        {
            const float2 q = __ldg(p.twiddles + 512 + (tid & 15));
            // lanes of the upper half-warp handle k + 256: one more factor exp(-i pi/4)
            const float c = 0.70710678118654752f;
            sm.qb2[tid] = (tid >> 4) ? make_float2(c * (q.x + q.y), c * (q.y - q.x)) : q;
        }
        if (tid == 0)
            sm.g_next = p.group_begin + atomicAdd(p.ticket, 1u);
    }
    const float noise_floor_factor = p.perc->noise_floor_factor;

// group geometry stays in shared memory and is re-read where needed: holding it in registers across the
// transform (32 live values per lane) is what makes a 64-register build spill
#define G_FRAME0 (sm.g_frame0)
#define G_FIRST_ROW (sm.g_first_row)
#define G_FIRST_FRAME (sm.g_first_frame)
#define G_SRC (sm.g_src)
#define G_LEN (sm.g_len)

    // Groups are handed out by a ticket counter, i.e. in the order CTAs actually run: every group before the
    // one a CTA holds is then owned by a CTA that is executing, whatever the residency of the grid -- the
    // condition under which the look-back cannot deadlock.
    for (;;)
    {
        __syncthreads(); // [A] tables / ticket ready, previous group done with shared memory
        const uint64_t g = sm.g_next;
        if (g >= p.group_end)
            break;
        if (warp == 0)
        {
            const uint32_t file = locate_file(p.first_group, p.n_files, g, lane);
            const FileDesc &fd0 = p.files[file];
            const uint32_t ch0 = fd0.channels;
            const uint32_t fpg = kFcs / ch0 ? kFcs / ch0 : 1u;
            const uint64_t fr0 = (g - p.first_group[file]) * fpg;
            const uint64_t left = fd0.n_frames - fr0;
            const uint32_t nfr = (uint32_t)(left < fpg ? left : fpg);
            if (lane == 0)
            {
                sm.g_src = p.pcm_arena + fd0.pcm_off;
                sm.g_len = (long long)fd0.len;
                sm.g_first_row = fd0.first_row;
                sm.g_first_frame = fd0.first_frame;
                sm.g_frame0 = fr0;
                sm.g_ch = ch0;
                sm.g_n_frames = nfr;
            }
            if (lane < kFcs)
            {
                sm.frame_nnz[lane] = 0;
                sm.frame_raw[lane] = 0;
                if ((uint32_t)lane < nfr * ch0) // descriptors of step 0
                {
                    const uint32_t lf = lane / ch0, c = lane - lf * ch0;
                    sm.fc_base[lane] = (long long)((fr0 + lf) * kHop) - kHop / 2;
                    sm.fc_chan[lane] = c;
                    sm.fc_lf[lane] = lf;
                }
            }
        }
        __syncthreads(); // [B]
        const uint32_t n_frames = sm.g_n_frames, ch = sm.g_ch;
        const uint32_t n_fc = n_frames * ch; // exceeds kFcs only when ch > 8 (then one frame, several rounds)
        const bool multi_round = n_fc > (uint32_t)kFcs;
        const uint32_t n_rounds = (n_fc + kFcs - 1) / kFcs;

        // Steps of up to 8 frame-channels (one per warp).  ch <= 8: ONE step that counts and emits.  ch > 8 (one
        // frame per group, several rounds of 8 channels): the rounds are first run to COUNT (per-row nnz / scale
        // parked in the output tables), the last of them decides raw / sparse and fetches the output offset; if
        // the frame stays sparse the rounds run again to EMIT (the keep masks live in registers, one round at a
        // time).
        const uint32_t n_steps = multi_round ? 2 * n_rounds : 1;
        uint32_t running_pairs = 0; // EMIT rounds of a wide frame: pairs of the rounds before this one
        for (uint32_t step = 0; step < n_steps; ++step)
        {
            const bool count_step = !multi_round || step < n_rounds;
            const bool emit_step = !multi_round || step >= n_rounds;
            const uint32_t round = multi_round ? step % n_rounds : 0;
            const bool last_count = count_step && round + 1 == n_rounds;
            const uint32_t fc0 = round * kFcs;
            const uint32_t fcs_here = min((uint32_t)kFcs, n_fc - fc0);
            if (step > 0)
            {
                __syncthreads(); // the previous step is done with shared memory
                if (emit_step && sm.frame_raw[0])
                    break; // a raw frame was written out completely by its last COUNT step
                if (tid < (int)fcs_here)
                {
                    sm.fc_base[tid] = (long long)(G_FRAME0 * kHop) - kHop / 2;
                    sm.fc_chan[tid] = fc0 + tid;
                    sm.fc_lf[tid] = 0;
                }
                __syncthreads();
            }

            // ---------------------------------------------------------------- stage: window + fold + pre-twiddle
            // Thread t makes z'[n] for n = t and n = t + 256 of every fc of the step (b = padded PCM of the frame,
            // w[2047-i] = w[i]):
            //   n < 256 : u0 = -b[1535-2n] wa - b[1536+2n] wo,  u1 =  b[511-2n] wo - b[512+2n] wa     wa = w[512+2n]
            //   n >= 256: u0 =  b[2n-512] wo - b[1535-2n] wa,   u1 = -b[512+2n] wa - b[2559-2n] wo    wo = w[511-2n] / w[2n-512]
            //   z'[n] = (u0 + i u1) * exp(-i pi (4n+1)/(4N))
            {
                const int na = tid, nb = tid + 256;
                const float wa_a = __ldg(p.window + 512 + 2 * na), wo_a = __ldg(p.window + 511 - 2 * na);
                const float wa_b = __ldg(p.window + 512 + 2 * nb), wo_b = __ldg(p.window + 2 * nb - 512);
                const float2 pa = sm.pre[na], pb = sm.pre[nb];
                // sample positions inside the frame
                const int a0 = 1535 - 2 * na, a1 = 1536 + 2 * na, a2 = 511 - 2 * na, a3 = 512 + 2 * na;
                const int b0 = 2 * nb - 512, b1 = 1535 - 2 * nb, b2 = 512 + 2 * nb, b3 = 2559 - 2 * nb;
                auto emit = [&](float2 *z, float x0, float x1, float x2, float x3, float y0, float y1, float y2, float y3) {
                    const float ua0 = -x0 * wa_a - x1 * wo_a, ua1 = x2 * wo_a - x3 * wa_a;
                    const float ub0 = y0 * wo_b - y1 * wa_b, ub1 = -y2 * wa_b - y3 * wo_b;
                    z[na] = make_float2(ua0 * pa.x - ua1 * pa.y, ua0 * pa.y + ua1 * pa.x);
                    z[nb] = make_float2(ub0 * pb.x - ub1 * pb.y, ub0 * pb.y + ub1 * pb.x);
                };
                if (ch <= 2)
                {
                    const uint32_t frames_here = fcs_here / ch;
                    for (uint32_t lf = 0; lf < frames_here; ++lf)
                    {
                        const long long base = sm.fc_base[lf * ch];
                        float2 *z0 = sm.u[lf * ch];
                        if (base >= 0 && base + kFrame <= G_LEN)
                        {
                            if (ch == 2)
                            {
                                // (L, R) per sample frame; one 64-bit pointer per frame and thread, the eight
                                // positions are small constant offsets from it
                                const float2 *q = reinterpret_cast<const float2 *>(G_SRC) + base + 2 * tid;
                                const float2 x0 = __ldg(q + (1535 - 4 * tid)), x1 = __ldg(q + 1536), x2 = __ldg(q + (511 - 4 * tid)),
                                             x3 = __ldg(q + 512);
                                const float2 y0 = __ldg(q), y1 = __ldg(q + (1023 - 4 * tid)), y2 = __ldg(q + 1024),
                                             y3 = __ldg(q + (2047 - 4 * tid));
                                emit(z0, x0.x, x1.x, x2.x, x3.x, y0.x, y1.x, y2.x, y3.x);
                                emit(z0 + kHop / 2, x0.y, x1.y, x2.y, x3.y, y0.y, y1.y, y2.y, y3.y);
                            }
                            else
                            {
                                const float *q = G_SRC + base;
                                emit(z0, __ldg(q + a0), __ldg(q + a1), __ldg(q + a2), __ldg(q + a3), __ldg(q + b0), __ldg(q + b1),
                                     __ldg(q + b2), __ldg(q + b3));
                            }
                        }
                        else
                        {
                            // first / last frames of a file: the 512-zero lead-in and the zero tail (src/codec.rs:433-447)
                            for (uint32_t c = 0; c < ch; ++c)
                            {
                                const float *fb = G_SRC + c;
                                const long long len = G_LEN;
                                auto smp = [&](int i) -> float {
                                    const long long pos = base + i;
                                    return (pos >= 0 && pos < len) ? __ldg(fb + pos * (long long)ch) : 0.0f;
                                };
                                emit(z0 + c * (kHop / 2), smp(a0), smp(a1), smp(a2), smp(a3), smp(b0), smp(b1), smp(b2), smp(b3));
                            }
                        }
                    }
                }
                else
                {
                    // 3 and more channels: strided scalar loads, the same fold
                    for (uint32_t fc = 0; fc < fcs_here; ++fc)
                    {
                        const long long base = sm.fc_base[fc];
                        const float *fb = G_SRC + sm.fc_chan[fc];
                        const long long len = G_LEN;
                        const bool interior = base >= 0 && base + kFrame <= len;
                        auto smp = [&](int i) -> float {
                            const long long pos = base + i;
                            return (interior || (pos >= 0 && pos < len)) ? __ldg(fb + pos * (long long)ch) : 0.0f;
                        };
                        emit(sm.u[fc], smp(a0), smp(a1), smp(a2), smp(a3), smp(b0), smp(b1), smp(b2), smp(b3));
                    }
                }
            }
            __syncthreads(); // [C]

            // ---------------------------------------------------------------- transform + keep mask, one warp per fc
            // what survives the barriers in registers: the keep mask and this lane's offsets; the coefficients
            // are re-read from shared memory and per-fc scalars (scale, nnz, segment totals) wait there
            unsigned keep = 0;               // bit 4j+e: bin 128 j + 4 lane + e is kept
            unsigned pos_lo = 0, pos_hi = 0; // this lane's exclusive offset inside segment j (eight 8-bit fields)
            const bool active = (uint32_t)warp < fcs_here;
            if (active)
            {
                dct4_single(sm.u[warp], sm, lane);
                const float *coef = reinterpret_cast<const float *>(sm.u[warp]);
                const int top_lo = sm.top_lo, n_bands = sm.n_bands;
                // max |c| and, in the same pass, the energy of the last band (bins top_lo..1023: more than half of
                // the row at every common sample rate)
                float m = 0.0f, e_top = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                {
                    const float4 v = *reinterpret_cast<const float4 *>(coef + j * 128 + lane * 4);
                    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                    const int k0 = j * 128 + lane * 4;
                    if (j * 128 >= top_lo) // uniform: the whole segment lies in the last band
                        e_top = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, e_top))));
                    else if ((j + 1) * 128 > top_lo) // uniform: the segment straddles the band edge
                    {
                        e_top = fmaf(k0 >= top_lo ? v.x : 0.0f, v.x, e_top);
                        e_top = fmaf(k0 + 1 >= top_lo ? v.y : 0.0f, v.y, e_top);
                        e_top = fmaf(k0 + 2 >= top_lo ? v.z : 0.0f, v.z, e_top);
                        e_top = fmaf(k0 + 3 >= top_lo ? v.w : 0.0f, v.w, e_top);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                    e_top += __shfl_xor_sync(0xffffffffu, e_top, o);
                }
                const float gmax = fmaxf(m, 1e-10f); // scale = max|c| .max(1e-10)              src/codec.rs:488-489
                const float scale = gmax;
                // band energies -> per-band base threshold (already times scale, :288)           :205-224
                // one lane per band, a uniform trip count (the widest narrow band of the batch)
                float *base = sm.band_base[warp];
                for (int b0 = 0; b0 < n_bands; b0 += 32)
                {
                    const int b = b0 + lane;
                    int lo = 0, hi = 0;
                    if (b < n_bands - 1) // the last band was summed above
                    {
                        lo = sm.band_lo[b];
                        hi = sm.band_hi[b];
                    }
                    const bool wide = (hi - lo) > kNarrowMax;
                    if (wide)
                        hi = lo;
                    float acc = 0.0f;
                    const int trip = sm.narrow_trip[b0 >> 5];
                    for (int i = 0; i < trip; i += 4)
                    {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                        {
                            const int k = lo + i + e;
                            const float c = k < hi ? coef[k] : 0.0f;
                            acc = fmaf(c, c, acc);
                        }
                    }
                    // bands other than the last that are wider than kNarrowMax (odd sample rates): the whole warp
                    unsigned wide_mask = __ballot_sync(0xffffffffu, wide);
                    while (wide_mask)
                    {
                        const int src_lane = __ffs(wide_mask) - 1;
                        wide_mask &= wide_mask - 1;
                        const int wlo = __shfl_sync(0xffffffffu, lo, src_lane);
                        const int whi = __shfl_sync(0xffffffffu, (int)sm.band_hi[b0 + src_lane], src_lane);
                        float part = 0.0f;
                        for (int k = wlo + lane; k < whi; k += 32)
                            part = fmaf(coef[k], coef[k], part);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
                            part += __shfl_xor_sync(0xffffffffu, part, o);
                        if (lane == src_lane)
                            acc = part;
                    }
                    if (b == n_bands - 1)
                        acc = e_top;
                    if (b < n_bands)
                        base[b] = sqrtf(acc * sm.band_rcnt[b]) * sm.band_fac[b] * scale;
                }
                __syncwarp();
                // thresholds -> keep mask                                                  src/codec.rs:226-235, 277-296
                //   th = base * inv_w;  if |c| > 0.3 gmax: th = min(th, 0.05 gmax * scale);  keep = |c| > nf && |c| > th
                //   <=>  keep = |c| > max(nf, th)  ||  |c| > max(nf, cap, gate)          (one row constant instead of a select)
                const float nf = noise_floor_factor * scale;
                const float t2 = fmaxf(fmaxf(nf, 0.05f * gmax * scale), 0.3f * gmax);
                const float base_top = base[n_bands - 1];
                unsigned cnt_lo = 0, cnt_hi = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                {
                    const float4 c4 = *reinterpret_cast<const float4 *>(coef + j * 128 + lane * 4);
                    const float4 iw4 = *reinterpret_cast<const float4 *>(sm.inv_w + j * 128 + lane * 4);
                    const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
                    const float iw[4] = {iw4.x, iw4.y, iw4.z, iw4.w};
                    float bs[4] = {base_top, base_top, base_top, base_top};
                    if (j * 128 < top_lo) // uniform: below the last band every bin looks its band up
                    {
                        const uchar4 bo4 = *reinterpret_cast<const uchar4 *>(sm.band_of + j * 128 + lane * 4);
                        bs[0] = base[bo4.x];
                        bs[1] = base[bo4.y];
                        bs[2] = base[bo4.z];
                        bs[3] = base[bo4.w];
                    }
                    unsigned nib = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                    {
                        const float a = fabsf(cv[e]);
                        // a kept bin never quantises to 0: a > nf = 10^-2.4 scale means |q| >= 130
                        nib |= (a > fmaxf(nf, bs[e] * iw[e]) || a > t2) ? (1u << e) : 0u;
                    }
                    keep |= nib << (4 * j);
                    const unsigned c = __popc(nib);
                    if (j < 4)
                        cnt_lo |= c << (8 * j);
                    else
                        cnt_hi |= c << (8 * (j - 4));
                }
                // one warp scan serves four segments: a segment holds at most 4 x 32 = 128 kept bins, which fits
                // its 8-bit field, so no carry ever crosses into the neighbouring field
                unsigned inc_lo = cnt_lo, inc_hi = cnt_hi;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const unsigned a = __shfl_up_sync(0xffffffffu, inc_lo, o);
                    const unsigned b = __shfl_up_sync(0xffffffffu, inc_hi, o);
                    if (lane >= o)
                    {
                        inc_lo += a;
                        inc_hi += b;
                    }
                }
                pos_lo = inc_lo - cnt_lo;
                pos_hi = inc_hi - cnt_hi;
                if (lane == 31)
                {
                    // kept bins of every segment over the whole warp (at most 128 each) and their sum
                    const unsigned s4 = (inc_lo & 0x00ff00ffu) + ((inc_lo >> 8) & 0x00ff00ffu) + (inc_hi & 0x00ff00ffu) +
                                        ((inc_hi >> 8) & 0x00ff00ffu);
                    const uint32_t nnz = (s4 & 0xffffu) + (s4 >> 16);
                    sm.fc_seg[warp][0] = inc_lo;
                    sm.fc_seg[warp][1] = inc_hi;
                    sm.fc_nnz[warp] = nnz;
                    sm.fc_scale[warp] = scale;
                    if (count_step)
                        atomicAdd(&sm.frame_nnz[sm.fc_lf[warp]], nnz);
                }
            }
            __syncthreads(); // [D]

            // ---------------------------------------------------------------- raw / sparse decision
            // src/codec.rs:505-540, a pure function of the frame's kept-bin count: every thread evaluates it itself,
            // so nothing below needs another barrier
            unsigned raw_mask = 0; // bit lf: frame lf of the group is stored as raw PCM
            {
                const float rhs = (float)((u64)kFrame * ch * 2) * 0.85f;
                for (uint32_t lf = 0; lf < n_frames; ++lf)
                {
                    const u64 compressed = (u64)ch * 8 + (u64)sm.frame_nnz[lf] * 4 + 8 + (u64)ch * 4 + 64;
                    raw_mask |= ((float)compressed >= rhs) ? (1u << lf) : 0u;
                }
            }
            auto is_raw_frame = [&](uint32_t lf) -> bool { return (raw_mask >> lf) & 1u; };
            // The group's slot: its rows' 4 KiB slots, contiguous.  Pairs go compact to the front (row order),
            // raw frame bodies compact to the back (frame order): sum(nnz) * 4 + raw fcs * 4096 <= fcs * 4096.
            const u64 row0 = G_FIRST_ROW + G_FRAME0 * ch;
            glc_pair *slot = p.slots + row0 * kHop;
            uint32_t raw_units_total = 0; // raw frame-channels of the group (valid once every channel is counted)
            if (last_count)
            {
                uint32_t pairs_total = 0;
                for (uint32_t lf = 0; lf < n_frames; ++lf)
                {
                    if (is_raw_frame(lf))
                        raw_units_total += ch;
                    else
                        pairs_total += sm.frame_nnz[lf];
                }
                if (tid == 0)
                {
                    p.grp_pairs[g] = pairs_total;
                    p.grp_raw[g] = raw_units_total;
                }
                if (tid < (int)n_frames)
                    p.is_raw[G_FIRST_FRAME + G_FRAME0 + tid] = is_raw_frame(tid) ? 1 : 0;
                // raw-PCM frame bodies: ((x*w)*32767) as i16, planar [ch][2048]             src/codec.rs:498-502
                uint32_t units = n_fc - raw_units_total;
                for (uint32_t lf = 0; lf < n_frames; ++lf)
                {
                    if (!is_raw_frame(lf))
                        continue;
                    int16_t *dst = reinterpret_cast<int16_t *>(slot) + (size_t)units * kFrame;
                    units += ch;
                    const long long base = (long long)((G_FRAME0 + lf) * kHop) - kHop / 2;
                    const float *src = G_SRC;
                    if (ch <= 2 && base >= 0 && base + kFrame <= G_LEN)
                    {
                        // four sample frames per thread: 16-byte loads, 8-byte stores per plane
                        for (uint32_t i = tid * 4; i < (uint32_t)kFrame; i += kThreads * 4)
                        {
                            const float4 w = __ldg(reinterpret_cast<const float4 *>(p.window + i));
                            if (ch == 1)
                            {
                                const float4 x = __ldg(reinterpret_cast<const float4 *>(src + base + i));
                                *reinterpret_cast<short4 *>(dst + i) =
                                    make_short4(raw_i16(x.x, w.x), raw_i16(x.y, w.y), raw_i16(x.z, w.z), raw_i16(x.w, w.w));
                            }
                            else
                            {
                                const float4 a = __ldg(reinterpret_cast<const float4 *>(src + (base + i) * 2));
                                const float4 b = __ldg(reinterpret_cast<const float4 *>(src + (base + i) * 2 + 4));
                                *reinterpret_cast<short4 *>(dst + i) =
                                    make_short4(raw_i16(a.x, w.x), raw_i16(a.z, w.y), raw_i16(b.x, w.z), raw_i16(b.z, w.w));
                                *reinterpret_cast<short4 *>(dst + kFrame + i) =
                                    make_short4(raw_i16(a.y, w.x), raw_i16(a.w, w.y), raw_i16(b.y, w.z), raw_i16(b.w, w.w));
                            }
                        }
                    }
                    else
                    {
                        const long long len = G_LEN;
                        for (uint32_t e = tid; e < (uint32_t)kFrame * ch; e += kThreads)
                        {
                            // reads are [pos][c]-ordered for coalescing, the store goes to the planar slot
                            const uint32_t i = e / ch, c = e - i * ch;
                            const long long pos = base + i;
                            const float x = (pos >= 0 && pos < len) ? __ldg(src + pos * (long long)ch + c) : 0.0f;
                            dst[(size_t)c * kFrame + i] = raw_i16(x, __ldg(p.window + i));
                        }
                    }
                }
                if (multi_round && raw_units_total)
                    for (uint32_t c = tid; c < ch; c += kThreads)
                    {
                        // a raw frame wider than one round: its parked rows carry no pairs
                        p.nnz[row0 + c] = 0;
                        p.scales[row0 + c] = 0.0f;
                    }
            }
            const bool group_raw0 = multi_round && last_count && raw_units_total != 0; // wide frame turned out raw
            if (tid == 0 && (!multi_round || group_raw0 || (emit_step && step + 1 == n_steps)))
                sm.g_next = p.group_begin + atomicAdd(p.ticket, 1u); // this is the group's last step
            if (multi_round && last_count && tid == 0)
                sm.frame_raw[0] = raw_units_total ? 1u : 0u; // read by the EMIT steps after their first barrier
            if (multi_round && count_step && active && lane == 0 && !group_raw0)
            {
                // parked: final unless the frame turns out raw (then the last COUNT step has just zeroed every row
                // of the frame, this round's included: it must not park on top of that)
                p.nnz[row0 + fc0 + warp] = sm.fc_nnz[warp];
                p.scales[row0 + fc0 + warp] = sm.fc_scale[warp];
            }
            if (emit_step && active)
            {
                // per-row table entries, then quantise the kept bins straight into the slot:
                // q = round(v / scale * 32768) (nearest; ties are measure-zero in this tolerance class), saturated
                // to i16                                                                   src/codec.rs:298-303
                const bool fc_raw = is_raw_frame(sm.fc_lf[warp]);
                uint32_t off = running_pairs;
                for (int f = 0; f < warp; ++f)
                    off += is_raw_frame(sm.fc_lf[f]) ? 0u : sm.fc_nnz[f];
                if (lane == 0 && !multi_round)
                {
                    p.nnz[row0 + warp] = fc_raw ? 0u : sm.fc_nnz[warp];
                    p.scales[row0 + warp] = fc_raw ? 0.0f : sm.fc_scale[warp];
                }
                if (!fc_raw)
                {
                    const float qmul = 32768.0f / sm.fc_scale[warp];
                    const unsigned seg_lo = sm.fc_seg[warp][0], seg_hi = sm.fc_seg[warp][1];
                    const float *coef = reinterpret_cast<const float *>(sm.u[warp]);
                    uint32_t *dst = reinterpret_cast<uint32_t *>(slot + off);
                    unsigned segbase = 0; // exclusive start of segment j (uniform over the warp)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                    {
                        const unsigned nib = (keep >> (4 * j)) & 15u;
                        const unsigned segtot = ((j < 4 ? seg_lo : seg_hi) >> (8 * (j & 3))) & 0xffu;
                        if (segtot) // uniform: nothing kept in this segment by any lane
                        {
                            const float4 c4 = *reinterpret_cast<const float4 *>(coef + j * 128 + lane * 4);
                            const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
                            unsigned pos = segbase + (((j < 4 ? pos_lo : pos_hi) >> (8 * (j & 3))) & 0xffu);
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                            {
                                short q;
                                asm("cvt.rni.s16.f32 %0, %1;" : "=h"(q) : "f"(cv[e] * qmul));
                                if ((nib >> e) & 1u)
                                    dst[pos++] = (uint32_t)(j * 128 + lane * 4 + e) | ((uint32_t)(unsigned short)q << 16);
                            }
                        }
                        segbase += segtot;
                    }
                }
            }
            if (multi_round && emit_step)
                for (uint32_t f = 0; f < fcs_here; ++f)
                    running_pairs += sm.fc_nnz[f]; // uniform: published before barrier [D]
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// fast_place_kernel: one WARP per frame group (8 groups per CTA, no block-wide barrier).  Moves the group's
// compact block of pairs and its raw bodies from the slot to their final positions (exclusive scans of the
// per-group totals) and writes the offset tables.
__global__ void __launch_bounds__(kThreads) fast_place_kernel(const FastEncodeLaunch p)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 g = p.group_begin + (u64)blockIdx.x * (kThreads / 32) + warp;
    if (g >= p.group_end)
        return;
    const uint32_t file = locate_file(p.first_group, p.n_files, g, lane);
    const FileDesc &fd = p.files[file];
    const uint32_t ch = fd.channels;
    const uint32_t fpg = kFcs / ch ? kFcs / ch : 1u;
    const u64 fr0 = (g - p.first_group[file]) * fpg;
    const u64 left = fd.n_frames - fr0;
    const uint32_t n_frames = (uint32_t)(left < fpg ? left : fpg), n_fc = n_frames * ch;
    const u64 row0 = fd.first_row + fr0 * ch, frame0 = fd.first_frame + fr0;
    const u64 pair_base = p.grp_pair_off[g];
    const u64 raw_base = p.grp_raw_off[g]; // units of 2048 i16
    const uint32_t n_pairs = p.grp_pairs[g], n_raw_units = p.grp_raw[g];
    // tables: exclusive scans inside the group (a warp scan per 32 rows / frames, carried)
    u64 carry = pair_base;
    for (uint32_t r0 = 0; r0 < n_fc; r0 += 32)
    {
        const uint32_t r = r0 + lane;
        const uint32_t v = r < n_fc ? p.nnz[row0 + r] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (r < n_fc)
            p.pair_off[row0 + r] = carry + (incl - v);
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    {
        const uint32_t v = ((uint32_t)lane < n_frames && p.is_raw[frame0 + lane]) ? ch : 0u; // at most 8 frames per group
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if ((uint32_t)lane < n_frames)
            p.raw_off[frame0 + lane] = (raw_base + (incl - v)) * (u64)kFrame;
    }
    if (lane == 0 && g + 1 == p.group_end)
    {
        // end markers of this launch (the next wave writes the same values as its first entries)
        p.pair_off[row0 + n_fc] = pair_base + n_pairs;
        p.raw_off[frame0 + n_frames] = (raw_base + n_raw_units) * (u64)kFrame;
    }
    // pairs: one contiguous block (the destination is only 4-byte aligned)
    const uint32_t *src = reinterpret_cast<const uint32_t *>(p.slots + row0 * kHop);
    uint32_t *dst = reinterpret_cast<uint32_t *>(p.pairs + (pair_base - (p.ring ? p.grp_pair_off[p.group_begin] : 0ull)));
    {
        // eight independent loads in flight per lane: the copy is latency-bound otherwise
        uint32_t k = lane;
        for (; k + 7 * 32 < n_pairs; k += 8 * 32)
        {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = __ldcs(src + k + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                dst[k + 32 * u] = v[u];
        }
        for (; k < n_pairs; k += 32)
            dst[k] = __ldcs(src + k);
    }
    // raw bodies: the back of the slot, 4 KiB units on both sides
    if (n_raw_units)
    {
        const int4 *rs = reinterpret_cast<const int4 *>(reinterpret_cast<const int16_t *>(p.slots + row0 * kHop) +
                                                        (size_t)(n_fc - n_raw_units) * kFrame);
        int4 *rd = reinterpret_cast<int4 *>(p.raw + (raw_base - (p.ring ? p.grp_raw_off[p.group_begin] : 0ull)) * (u64)kFrame);
        const uint32_t n16 = n_raw_units * (kFrame * 2 / 16); // a multiple of 256
        for (uint32_t k = lane; k < n16; k += 8 * 32)
        {
            int4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = __ldcs(rs + k + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                rd[k + 32 * u] = v[u];
        }
    }
}

#undef G_FRAME0
#undef G_FIRST_ROW
#undef G_FIRST_FRAME
#undef G_SRC
#undef G_LEN

} // namespace

uint64_t fast_groups_for(uint32_t n_frames, uint32_t channels)
{
    const uint32_t fpg = kFcs / channels ? kFcs / channels : 1u;
    return ((uint64_t)n_frames + fpg - 1) / fpg;
}

cudaError_t launch_fast_encode(const FastEncodeLaunch &p, cudaStream_t s)
{
    if (p.group_end <= p.group_begin)
        return cudaSuccess;
    GLC_SET_MAX_DYN_SMEM_ONCE(fast_encode_kernel, sizeof(Smem));
    // persistent CTAs drawing groups from the ticket counter: grid = SMs x occupancy (queried once per device:
    // the occupancy call costs tens of microseconds of host time per launch otherwise)
    static std::atomic<int> resident_ctas[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int resident = resident_ctas[dev & 63].load(std::memory_order_acquire);
    if (resident == 0)
    {
        int sms = 148, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fast_encode_kernel, kThreads, sizeof(Smem));
        if (e != cudaSuccess)
            return e;
        if (per_sm < 1)
            return cudaErrorLaunchOutOfResources;
        resident = sms * per_sm;
        resident_ctas[dev & 63].store(resident, std::memory_order_release);
    }
    const uint64_t grid = std::min<uint64_t>(p.group_end - p.group_begin, (uint64_t)resident);
    fast_encode_kernel<<<(unsigned)grid, kThreads, sizeof(Smem), s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fast_place(const FastEncodeLaunch &p, cudaStream_t s)
{
    if (p.group_end <= p.group_begin)
        return cudaSuccess;
    if (p.group_end - p.group_begin > 0x7fffffffull)
        return cudaErrorInvalidValue;
    fast_place_kernel<<<(unsigned)((p.group_end - p.group_begin + kThreads / 32 - 1) / (kThreads / 32)), kThreads, 0, s>>>(p);
    return cudaGetLastError();
}

} // namespace glc
