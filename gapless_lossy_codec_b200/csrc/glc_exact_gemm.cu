// glc_exact_gemm.cu -- EXACT-mode transform kernels for sm_100a.
//
// What the reference computes (src/codec.rs:359-390):
//     MDCT   out[k] = (sum_{i=0..2047} block[i] * tab[k][i]) * norm      block[i] = x[i]*window[i]
//     IMDCT  out[i] = (sum_{k=0..1023} coef[k]  * tab[k][i]) * norm      then out[i] *= window[i]
// with `s` starting at +0.0, strictly ascending reduction index, the product rounded to f32 and
// then the sum rounded to f32 (no FMA).  Bit-exact parity needs exactly that chain per output, so
// this is a dense contraction  C[row][n] = sum_j A[j][row] * T[k(j)][n]  whose reduction order is
// fixed and whose inner operation is a rounded multiply followed by a rounded add.  Tensor cores, FMA
// contraction and split-K are all ruled out; the binding roof is the FP32 pipe
// at two instructions per multiply-add, not HBM.  Since round 2 both kernels issue those two instructions
// as PACKED pairs (f32x2: two outputs per issue slot, the FMA pipe busy for two cycles), which takes the
// operand loads and bookkeeping off the critical issue port: see mul_then_add_x2 below for the exact forms
// and for why their constants are kernel parameters.  GLC_MDCT_X2=0 / GLC_IMDCT_X2=0 build the scalar
// FMUL + FADD loops (bit-identical, 39.7 instead of 37.9 ms and 13.8 instead of 12.6 ms per hour of stereo).
//
// Mapping:
//   * CTA tile = 128 rows (frame-channels) x 128 outputs, 256 threads, 8x8 outputs per thread,
//     every accumulator an independent sequential chain; 2 CTAs per SM (128 registers).
//   * BOTH operands arrive by the TMA engine (cp.async.bulk -> UBLKCP) into a 3-slot shared-memory
//     ring guarded by full/empty mbarriers; there is no __syncthreads and no operand-producing
//     code in the main loop.  A stage is 32 reduction steps: A = [32][128 rows] (16 KiB) and
//     T = [32][128 outputs] (16 KiB).
//       - MDCT: A tiles are written by window_tile_kernel (padding rules of src/codec.rs:433-447 +
//         window multiply, i.e. block[i]), one contiguous 16 KiB block per (row tile, stage); T comes
//         from a host re-tiled copy of the table, also one contiguous block per (output block, stage).
//   * IMDCT (imdct_sparse_kernel below) keeps the ring and the TMA feed but is SPARSE at warp
//     granularity: a warp owns 2 rows and executes only the reduction steps at which one of ITS rows
//     has a coefficient (a step mask per warp and stage, delivered with the A stage).  Skipping a k
//     whose coefficient is zero in a row is exact: the product is +-0 and the running sum (which starts
//     at +0.0 and can never become -0) is unchanged.
//   * Epilogue: * norm (and * window for IMDCT), two float4 stores per row.
#include "glc_internal.cuh"

namespace glc
{

namespace
{

constexpr int kStageFloats = kKC * kBN;        // 4096 floats = 16 KiB per operand per stage
constexpr int kStageBytes = kStageFloats * 4;
constexpr int kRing = 3;
// Outputs per thread of the MDCT contraction.  Measured on the hour-long bench signal: 8 (256 threads,
// 126 registers, 16 warps/SM) 39.6 ms (scalar build); 4 (512 threads, 64 registers, 32 warps/SM) 41.7 ms -- twice the
// warps do not make up for 3 instead of 4 operand loads per 64 instead of 128 arithmetic instructions.
#ifndef GLC_MDCT_NC
#define GLC_MDCT_NC 8
#endif
constexpr int kMdctNC = GLC_MDCT_NC;

__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_addr(bar);
    uint32_t done;
    do
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

#pragma nv_diag_suppress 177 // which of the helpers below are used depends on the compile-time variant
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Packed pairs of f32 (sm_100: mul/add.rn.f32x2 work on two lanes of a 64-bit register, each lane rounded to nearest
// exactly like the scalar instruction, and an explicit .rn is never contracted into an fma): half the issue slots
// for the same arithmetic, which matters where the kernel is bound by issue slots and not by the FMA pipe.
__device__ __forceinline__ void lds128_x2(uint32_t addr, unsigned long long &lo, unsigned long long &hi)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr));
}
__device__ __forceinline__ unsigned long long pack_x2(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack_x2(unsigned long long v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long mul_then_add_x2(unsigned long long acc, unsigned long long a, unsigned long long t,
                                                           unsigned long long one2, unsigned long long neg_zero2)
{
    // ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad false (seen in the SASS), which
    // rounds once instead of twice, and it does the same to fma(a, t, -0.0) followed by fma(p, 1.0, acc) when it can
    // see the constants.  With the constants arriving as kernel parameters the two FMAs stay, and they are exact
    // restatements:  a * t + (-0.0) == RN(a * t) with the sign of a zero product kept;  p * 1.0 + acc == RN(acc + p)
    unsigned long long p, r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(a), "l"(t), "l"(neg_zero2));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p), "l"(one2), "l"(acc));
    return r;
}

#pragma nv_diag_default 177

// One TMA-engine bulk copy global -> shared, completing `bytes` on the mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

struct GemmParams
{
    const float *a_tiles;    // [m_tile][stage][kKC][kBM]
    const float *tab;        // MDCT: re-tiled [n_block][stage][kKC][kBN]; IMDCT: natural [1024][2048]
    const float *window;     // IMDCT epilogue
    const uint8_t *stage_list; // IMDCT: [m_tile][kImdctStages] ascending stages that hold a coefficient
    const uint32_t *n_k;     // IMDCT: [m_tile] number of listed stages (0 = nothing to do)
    const uint32_t *n_tiles; // IMDCT: device-side count of live row tiles (grid is sized for the worst case)
    uint64_t tile_begin;     // first row tile of this launch
    uint64_t n_rows;         // rows (MDCT: frame-channels; IMDCT: compacted slots) that exist; stores are clipped
    float norm;
    float *out;              // MDCT: coefs[row][1024]; IMDCT: blocks[slot][2048]
    // packed (f32x2) constants handed over as PARAMETERS so that ptxas cannot see their values: {1.0f, 1.0f} and
    // {-0.0f, -0.0f} (see mul_then_add_x2)
    unsigned long long x2_one, x2_neg_zero;
};

struct Smem
{
    float a[kRing][kStageFloats];
    float t[kRing][kStageFloats];
    uint64_t full[kRing];
    uint64_t empty[kRing];
    uint32_t released[kRing]; // warps that have left the slot (running count)
};

// MDCT: reduce over i (64 stages of 32 steps).  NC = outputs per thread (8 rows x NC outputs):
// NC = 8 -> 256 threads, 16 warps/SM; NC = 4 -> 512 threads, 64 registers, 32 warps/SM.
template <int NC>
__global__ void __launch_bounds__(kGemmThreads * 8 / NC, 2) exact_gemm_kernel(const __grid_constant__ GemmParams p)
{
    constexpr int kThreads = kGemmThreads * 8 / NC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int kNBlocks = kHop / kBN;
    constexpr int n_stages = kFrame / kKC;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tx = NC == 8 ? (tid & 15) : (tid & 31);
    const int ty = NC == 8 ? (tid >> 4) : (tid >> 5);
    // linear grid, output block fastest: the CTAs that run together share A tiles in L2
    const int n_block = (int)(blockIdx.x % kNBlocks);
    const uint64_t m_tile = p.tile_begin + blockIdx.x / kNBlocks;

    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < kRing; ++s)
        {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kThreads / 32);
            sm.released[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const float *a_src = p.a_tiles + (size_t)m_tile * (kFrame * kBM);
    const float *t_src = p.tab + (size_t)n_block * (kFrame / kKC) * kStageFloats;

    // Fill ring slot `s % kRing` with stage s (two bulk copies).  The first kRing stages are issued by
    // warp 0; stage s + kRing is issued by the last warp that leaves stage s, at the moment the slot
    // becomes free, so no warp ever blocks on the others in order to feed the ring.
    auto issue = [&](int s) {
        const int slot = s % kRing;
        if (lane == 0)
        {
            mbar_expect_tx(&sm.full[slot], 2 * kStageBytes);
            bulk_g2s(sm.a[slot], a_src + (size_t)s * kStageFloats, kStageBytes, &sm.full[slot]);
            bulk_g2s(sm.t[slot], t_src + (size_t)s * kStageFloats, kStageBytes, &sm.full[slot]);
        }
        __syncwarp();
    };

    if (warp == 0)
        for (int s = 0; s < kRing; ++s)
            issue(s);

#if GLC_MDCT_X2
    static_assert(!GLC_MDCT_X2 || NC == 8, "packed variant: 8 x 8 outputs per thread");
    unsigned long long acc2[8][NC / 2]; // pairs of adjacent outputs, +0.0 in both lanes
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < NC / 2; ++c)
            acc2[r][c] = 0ull;
#else
    float acc[8][NC];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c)
            acc[r][c] = 0.0f;
#endif

    for (int s = 0; s < n_stages; ++s)
    {
        const int slot = s % kRing;
        mbar_wait(&sm.full[slot], (uint32_t)(s / kRing) & 1u);

        const float *As = sm.a[slot] + ty * 8;
        const float *Ts = sm.t[slot] + tx * 4;
#if GLC_MDCT_X2
#pragma unroll 4
        for (int ii = 0; ii < kKC; ++ii)
        {
            const float4 a_lo = *reinterpret_cast<const float4 *>(As + ii * kBM);
            const float4 a_hi = *reinterpret_cast<const float4 *>(As + ii * kBM + 4);
            const ulonglong2 t_lo = *reinterpret_cast<const ulonglong2 *>(Ts + ii * kBN);
            const ulonglong2 t_hi = *reinterpret_cast<const ulonglong2 *>(Ts + ii * kBN + 64);
            const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const unsigned long long t2[4] = {t_lo.x, t_lo.y, t_hi.x, t_hi.y};
#pragma unroll
            for (int r = 0; r < 8; ++r)
            {
                const unsigned long long ar = pack_x2(a[r], a[r]);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    acc2[r][c] = mul_then_add_x2(acc2[r][c], ar, t2[c], p.x2_one, p.x2_neg_zero);
            }
        }
#else
#pragma unroll 4
        for (int ii = 0; ii < kKC; ++ii)
        {
            const float4 a_lo = *reinterpret_cast<const float4 *>(As + ii * kBM);
            const float4 a_hi = *reinterpret_cast<const float4 *>(As + ii * kBM + 4);
            const float4 t_lo = *reinterpret_cast<const float4 *>(Ts + ii * kBN);
            float4 t_hi = t_lo;
            if (NC == 8)
                t_hi = *reinterpret_cast<const float4 *>(Ts + ii * kBN + 64);
            const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const float t[8] = {t_lo.x, t_lo.y, t_lo.z, t_lo.w, t_hi.x, t_hi.y, t_hi.z, t_hi.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    acc[r][c] = __fadd_rn(acc[r][c], __fmul_rn(a[r], t[c]));
        }
#endif
        __syncwarp();
        uint32_t nth = 0;
        if (lane == 0)
        {
            mbar_arrive(&sm.empty[slot]);
            nth = atomicAdd(&sm.released[slot], 1u);
        }
        nth = __shfl_sync(0xffffffffu, nth, 0);
        if (nth % (kThreads / 32) == kThreads / 32 - 1 && s + kRing < n_stages)
        {
            // last warp out: all arrivals precede their counts, so this wait returns at once; it orders
            // the other warps' reads of the slot before the refill
            mbar_wait(&sm.empty[slot], (uint32_t)(s / kRing) & 1u);
            issue(s + kRing);
        }
    }

    // ---- epilogue: * norm, two float4 stores per row ----
#if GLC_MDCT_X2
    float acc[8][NC];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < NC / 2; ++c)
            unpack_x2(acc2[r][c], acc[r][2 * c], acc[r][2 * c + 1]);
#endif
    const int n_lo = n_block * kBN + tx * 4;
    const int n_hi = n_lo + 64;
    const uint64_t row0 = m_tile * kBM + ty * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        const uint64_t row = row0 + r;
        if (row >= p.n_rows)
            continue;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < NC; ++c)
            v[c] = __fmul_rn(acc[r][c], p.norm);
        float *orow = p.out + row * kHop;
        *reinterpret_cast<float4 *>(orow + n_lo) = make_float4(v[0], v[1], v[2], v[3]);
        if (NC == 8)
            *reinterpret_cast<float4 *>(orow + n_hi) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

cudaError_t launch_gemm(const GemmParams &p, uint64_t m_tiles, cudaStream_t s)
{
    const size_t smem = sizeof(Smem);
    GLC_SET_MAX_DYN_SMEM_ONCE(exact_gemm_kernel<kMdctNC>, smem);
    if (m_tiles == 0)
        return cudaSuccess;
    const uint64_t n_ctas = m_tiles * (kHop / kBN);
    if (n_ctas > 0x7fffffffull)
        return cudaErrorInvalidValue;
    exact_gemm_kernel<kMdctNC><<<(unsigned)n_ctas, kGemmThreads * 8 / kMdctNC, smem, s>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// imdct_sparse_kernel: direct IMDCT (src/codec.rs:377-390) + synthesis window (:672-675) over the
// compacted rows of a decode wave.
//   CTA   = 32 rows x 256 outputs, 16 warps, 40 registers: 3 CTAs = 48 warps per SM.
//   warp  = 2 rows x 256 outputs (thread: 2 rows x 2 float4 of outputs), so a reduction step is needed
//           by the whole warp or by none of it: the warp walks the set bits of its step mask and the
//           branch is uniform.  On the bench workload a row holds ~258 of the 1024 indices, the union of
//           2 rows ~344, of 4 rows ~446, of 8 rows ~560, of 128 rows ~945 (the dense contraction of the
//           first version).
//   stage = 32 consecutive coefficient indices: A = [32][32 rows] values + 16 masks (4 160 B) written by
//           dequant_tile_kernel, T = 32 rows x 256 outputs of the table, contiguous in the re-tiled copy
//           (tile_table_for_imdct): two bulk copies per stage.  Stages in which the tile has no
//           coefficient at all are not listed and never loaded.
//   ring  = 2 slots (3 CTAs x 74 KB fill the SM; the same memory as 4 stages of 16 indices or 8 of 8
//           measured slower: 14.8 and 17.3 ms, the per-stage bookkeeping counts), full/empty mbarriers; a warp with nothing to do in a stage releases it at once;
//           there is no producer warp: a slot is refilled by the last warp that leaves it.
// Measured on the hour-long bench signal: 26.5 ms (dense over the 128-row union) -> 13.8 ms (8-row warps
// 20.1, 4-row warps 15.7, 2-row warps with 16-index stages 14.8).  What is left is the per-step bookkeeping (12 of 44 instructions) and warps of
// one CTA waiting for each other at the ring; with every mask forced to all-ones the same kernel runs
// at the full issue rate, i.e. the pipeline itself is not the limit.
#ifndef GLC_IMDCT_RING
#define GLC_IMDCT_RING 2
#endif
constexpr int kImdctRing = GLC_IMDCT_RING;
constexpr int kImdctNQ = kImdctBN / 128; // float4 chunks of outputs per thread (lane*4 + 128 q)
struct ImdctSmem
{
    float a[kImdctRing][kImdctAStageFloats];
    float t[kImdctRing][kImdctKC * kImdctBN];
    uint64_t full[kImdctRing];
    uint64_t empty[kImdctRing];
    uint32_t released[kImdctRing]; // consumer warps that have left the slot (running count)
    uint8_t stage_list[kImdctStages]; // the tile's stages
};
static_assert((kImdctAStageFloats * 4) % 16 == 0, "bulk copies need 16-byte granularity");
static_assert(kImdctRowsPerWarp == 1 || kImdctRowsPerWarp == 2, "a warp owns one row or a pair of rows");

#ifndef GLC_IMDCT_MINB
#define GLC_IMDCT_MINB (GLC_IMDCT_RW == 2 && GLC_IMDCT_BN == 256 ? 3 : 1)
#endif
__global__ void __launch_bounds__(kImdctThreads, GLC_IMDCT_MINB) imdct_sparse_kernel(const __grid_constant__ GemmParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ImdctSmem &sm = *reinterpret_cast<ImdctSmem *>(smem_raw);
    constexpr int kNBlocks = kFrame / kImdctBN;
    constexpr uint32_t kTxBytes = kImdctAStageFloats * 4 + kImdctKC * kImdctBN * 4;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // linear grid, output block fastest: the CTAs that run together share A tiles in L2
    const int n_block = (int)(blockIdx.x % kNBlocks);
    const uint64_t m_tile = p.tile_begin + blockIdx.x / kNBlocks;
    if (m_tile >= (uint64_t)__ldg(p.n_tiles))
        return;
    const int n_stages = (int)__ldg(p.n_k + m_tile);

    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < kImdctRing; ++s)
        {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kImdctWarps);
            sm.released[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < n_stages)
        sm.stage_list[tid] = __ldg(p.stage_list + (size_t)m_tile * kImdctStages + tid);
    __syncthreads();

    const float *a_src = p.a_tiles + (size_t)m_tile * kImdctATileFloats;
    const float *t_src = p.tab + (size_t)n_block * kHop * kImdctBN;

    // Fill ring slot `j % ring` with the j-th listed stage: two bulk copies (the table is re-tiled so that
    // the kImdctKC rows of a stage are contiguous for every output block).  There is no producer warp:
    // the first `ring` stages are issued by warp 0 and stage j + ring by whichever warp is the last to
    // leave stage j, i.e. at the moment the slot becomes free.
    auto issue = [&](int j) {
        if (lane == 0)
        {
            const int slot = j % kImdctRing;
            const uint32_t st = sm.stage_list[j];
            mbar_expect_tx(&sm.full[slot], kTxBytes);
            bulk_g2s(sm.a[slot], a_src + (size_t)st * kImdctAStageFloats, kImdctAStageFloats * 4, &sm.full[slot]);
            bulk_g2s(sm.t[slot], t_src + (size_t)st * kImdctKC * kImdctBN, kImdctKC * kImdctBN * 4, &sm.full[slot]);
        }
        __syncwarp();
    };

    if (warp == 0)
        for (int s = 0; s < kImdctRing && s < n_stages; ++s)
            issue(s);

    constexpr int RW = kImdctRowsPerWarp;
    constexpr int NO = 4 * kImdctNQ; // outputs per thread and row
#if GLC_IMDCT_X2
    static_assert(!GLC_IMDCT_X2 || (kImdctRowsPerWarp == 2 && !GLC_IMDCT_ROWMASK), "packed variant: two rows per warp, union masks");
    unsigned long long acc2[RW][NO / 2]; // pairs of adjacent outputs; +0.0 in both lanes
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int c = 0; c < NO / 2; ++c)
            acc2[r][c] = 0ull;
#else
    float acc[RW][NO];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int c = 0; c < NO; ++c)
            acc[r][c] = 0.0f;
#endif

    for (int s = 0; s < n_stages; ++s)
    {
        const int slot = s % kImdctRing;
        mbar_wait(&sm.full[slot], (uint32_t)(s / kImdctRing) & 1u);

        // bit-reversed step mask: the next step in ascending order is clz(rm).  Operand addresses are
        // formed as 32-bit shared-memory addresses (one shift-add each per step).
#if GLC_IMDCT_ROWMASK
        static_assert(!GLC_IMDCT_ROWMASK || kImdctRowsPerWarp == 2, "row masks are written for two rows per warp");
        const uint2 rmw = reinterpret_cast<const uint2 *>(sm.a[slot] + kImdctKC * kImdctBM)[warp];
        const uint32_t rm0 = __brev(rmw.x), rm1 = __brev(rmw.y);
        uint32_t rm = rm0 | rm1;
#else
        uint32_t rm = __brev(reinterpret_cast<const uint32_t *>(sm.a[slot] + kImdctKC * kImdctBM)[warp]);
#endif
        const uint32_t a_base = smem_addr(sm.a[slot] + warp * RW);
        const uint32_t t_base = smem_addr(sm.t[slot] + lane * 4);
        while (rm)
        {
            const uint32_t ii = (uint32_t)__clz((int)rm);
            rm ^= 0x80000000u >> ii; // the bit is known to be set
            float a[RW];
            if constexpr (RW == 2)
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a[0]), "=f"(a[1]) : "r"(a_base + ii * (kImdctBM * 4)));
            else
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a[0]) : "r"(a_base + ii * (kImdctBM * 4)));
#if GLC_IMDCT_X2
            {
                const unsigned long long a0 = pack_x2(a[0], a[0]), a1 = pack_x2(a[1], a[1]);
#pragma unroll
                for (int q = 0; q < kImdctNQ; ++q)
                {
                    unsigned long long t0, t1;
                    lds128_x2(t_base + ii * (kImdctBN * 4) + q * 512, t0, t1);
                    acc2[0][2 * q + 0] = mul_then_add_x2(acc2[0][2 * q + 0], a0, t0, p.x2_one, p.x2_neg_zero);
                    acc2[0][2 * q + 1] = mul_then_add_x2(acc2[0][2 * q + 1], a0, t1, p.x2_one, p.x2_neg_zero);
                    acc2[1][2 * q + 0] = mul_then_add_x2(acc2[1][2 * q + 0], a1, t0, p.x2_one, p.x2_neg_zero);
                    acc2[1][2 * q + 1] = mul_then_add_x2(acc2[1][2 * q + 1], a1, t1, p.x2_one, p.x2_neg_zero);
                }
            }
#else
            float t[NO];
#pragma unroll
            for (int q = 0; q < kImdctNQ; ++q)
            {
                const float4 v = lds128(t_base + ii * (kImdctBN * 4) + q * 512);
                t[4 * q + 0] = v.x;
                t[4 * q + 1] = v.y;
                t[4 * q + 2] = v.z;
                t[4 * q + 3] = v.w;
            }
#if GLC_IMDCT_ROWMASK
            // skipping a row at a step where it has no pair is exact: the product would be +-0 and the running
            // sum, which starts at +0.0, is never -0
            const uint32_t bit = 0x80000000u >> ii;
            if (rm0 & bit)
            {
#pragma unroll
                for (int c = 0; c < NO; ++c)
                    acc[0][c] = __fadd_rn(acc[0][c], __fmul_rn(a[0], t[c]));
            }
            if (rm1 & bit)
            {
#pragma unroll
                for (int c = 0; c < NO; ++c)
                    acc[1][c] = __fadd_rn(acc[1][c], __fmul_rn(a[1], t[c]));
            }
#else
#pragma unroll
            for (int r = 0; r < RW; ++r)
#pragma unroll
                for (int c = 0; c < NO; ++c)
                    acc[r][c] = __fadd_rn(acc[r][c], __fmul_rn(a[r], t[c]));
#endif
#endif // GLC_IMDCT_X2
        }
        __syncwarp();
        uint32_t nth = 0;
        if (lane == 0)
        {
            mbar_arrive(&sm.empty[slot]);
            nth = atomicAdd(&sm.released[slot], 1u);
        }
        nth = __shfl_sync(0xffffffffu, nth, 0);
        if (nth % kImdctWarps == kImdctWarps - 1 && s + kImdctRing < n_stages)
        {
            // last warp out: every arrival on `empty` precedes its own count, so this wait returns at once
            // and orders the other warps' reads of the slot before the refill
            mbar_wait(&sm.empty[slot], (uint32_t)(s / kImdctRing) & 1u);
            issue(s + kImdctRing);
        }
    }

    // ---- epilogue: * norm, * window, float4 per 128 outputs (a warp writes contiguous 512-byte runs) ----
#if GLC_IMDCT_X2
    float acc[RW][NO];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int c = 0; c < NO / 2; ++c)
            unpack_x2(acc2[r][c], acc[r][2 * c], acc[r][2 * c + 1]);
#endif
    const uint64_t row0 = m_tile * kImdctBM + warp * RW;
#pragma unroll
    for (int q = 0; q < kImdctNQ; ++q)
    {
        const int n0 = n_block * kImdctBN + lane * 4 + 128 * q;
        const float4 w = __ldg(reinterpret_cast<const float4 *>(p.window + n0));
#pragma unroll
        for (int r = 0; r < RW; ++r)
        {
            const uint64_t row = row0 + r;
            if (row >= p.n_rows)
                continue;
            float4 v;
            v.x = __fmul_rn(__fmul_rn(acc[r][4 * q + 0], p.norm), w.x);
            v.y = __fmul_rn(__fmul_rn(acc[r][4 * q + 1], p.norm), w.y);
            v.z = __fmul_rn(__fmul_rn(acc[r][4 * q + 2], p.norm), w.z);
            v.w = __fmul_rn(__fmul_rn(acc[r][4 * q + 3], p.norm), w.w);
            *reinterpret_cast<float4 *>(p.out + row * kFrame + n0) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// window_tile_kernel: the A operand of the MDCT.  block[i] = x[i] * window[i] over the reference's
// virtual padded signal (512 zeros + data + zeros to a multiple of 1024 + 512 zeros,
// src/codec.rs:433-447, 476-481), written as [row tile][stage][32 reduction steps][128 rows].
// One CTA per (row tile, group of 8 stages): coalesced reads along the sample axis, transpose in
// shared memory, 16 KiB contiguous writes.
struct RowSrc
{
    const float *base; // first sample of this row's channel
    long long start;   // sample index (per channel) of reduction step 0; negative inside the 512-zero lead-in
    long long len;     // samples per channel
    int stride;        // channels
    int valid;
};

constexpr int kWtStagesPerCta = 8;

__global__ void __launch_bounds__(256) window_tile_kernel(const MdctLaunch p, float *a_tiles)
{
    __shared__ RowSrc rows[kBM];
    __shared__ float tile[kKC][kBM + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t m_tile_rel = blockIdx.x / (kFrame / kKC / kWtStagesPerCta);
    const int sgroup = (int)(blockIdx.x % (kFrame / kKC / kWtStagesPerCta));
    const uint64_t row0 = p.row_begin + m_tile_rel * kBM;

    if (tid < kBM)
    {
        RowSrc rs;
        const uint64_t row = row0 + tid;
        rs.valid = row < p.row_end;
        rs.base = nullptr;
        rs.start = 0;
        rs.len = 0;
        rs.stride = 1;
        if (rs.valid)
        {
            uint32_t lo = 0, hi = p.n_files - 1; // the file whose row range holds `row`
            while (lo < hi)
            {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (p.files[mid].first_row <= row)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            const FileDesc fd = p.files[lo];
            const uint64_t local = row - fd.first_row;
            const uint64_t f = local / fd.channels;
            const uint32_t c = (uint32_t)(local - f * fd.channels);
            rs.base = p.pcm_arena + fd.pcm_off + c;
            rs.start = (long long)(f * kHop) - (kHop / 2); // 512 leading zeros, src/codec.rs:438
            rs.len = (long long)fd.len;
            rs.stride = (int)fd.channels;
        }
        rows[tid] = rs;
    }
    __syncthreads();

    float *dst_tile = a_tiles + (size_t)m_tile_rel * (kFrame * kBM);
    for (int s = sgroup * kWtStagesPerCta; s < (sgroup + 1) * kWtStagesPerCta; ++s)
    {
        const int r = s * kKC + lane;
        const float w = __ldg(p.window + r);
#pragma unroll 4
        for (int j = 0; j < 16; ++j)
        {
            const RowSrc &rs = rows[warp * 16 + j];
            const long long pos = rs.start + r;
            float v = 0.0f;
            if (rs.valid && pos >= 0 && pos < rs.len)
                v = __ldg(rs.base + pos * rs.stride);
            tile[lane][warp * 16 + j] = __fmul_rn(v, w); // block[i] = x[i] * window[i]
        }
        __syncthreads();
        float *dst = dst_tile + (size_t)s * kStageFloats;
#pragma unroll
        for (int e = tid; e < kStageFloats; e += 256)
            dst[e] = tile[e >> 7][e & 127];
        __syncthreads();
    }
}

} // namespace

size_t mdct_a_tile_floats(uint64_t n_rows)
{
    return (size_t)((n_rows + kBM - 1) / kBM) * kFrame * kBM;
}

cudaError_t launch_window_tiles(const MdctLaunch &l, cudaStream_t s)
{
    if (l.row_end <= l.row_begin)
        return cudaSuccess;
    const uint64_t m_tiles = (l.row_end - l.row_begin + kBM - 1) / kBM;
    window_tile_kernel<<<(unsigned)(m_tiles * (kFrame / kKC / kWtStagesPerCta)), 256, 0, s>>>(l, l.a_tiles);
    return cudaGetLastError();
}

cudaError_t launch_mdct_exact(const MdctLaunch &l, cudaStream_t s)
{
    if (l.row_end <= l.row_begin)
        return cudaSuccess;
    const uint64_t m_tiles = (l.row_end - l.row_begin + kBM - 1) / kBM;
    GemmParams p{};
    p.a_tiles = l.a_tiles;
    p.tab = l.tab_tiled;
    p.tile_begin = 0;
    p.n_rows = l.row_end - l.row_begin;
    p.norm = l.norm;
    p.out = l.coefs + l.row_begin * kHop; // rows of this launch are numbered from 0 inside the kernel
    p.x2_one = 0x3f8000003f800000ull;
    p.x2_neg_zero = 0x8000000080000000ull;
    return launch_gemm(p, m_tiles, s);
}

cudaError_t launch_imdct_exact(const ImdctLaunch &l, cudaStream_t s)
{
    GemmParams p{};
    p.a_tiles = l.a_tiles;
    p.tab = l.tab;
    p.window = l.window;
    p.stage_list = l.stage_list;
    p.n_k = l.n_stages;
    p.n_tiles = l.n_tiles;
    p.tile_begin = 0;
    p.n_rows = l.max_slots;
    p.norm = l.norm;
    p.out = l.blocks;
    p.x2_one = 0x3f8000003f800000ull;
    p.x2_neg_zero = 0x8000000080000000ull;
    GLC_SET_MAX_DYN_SMEM_ONCE(imdct_sparse_kernel, sizeof(ImdctSmem));
    const uint64_t m_tiles = (l.max_slots + kImdctBM - 1) / kImdctBM;
    if (m_tiles == 0)
        return cudaSuccess;
    const uint64_t n_ctas = m_tiles * (kFrame / kImdctBN);
    if (n_ctas > 0x7fffffffull)
        return cudaErrorInvalidValue;
    imdct_sparse_kernel<<<(unsigned)n_ctas, kImdctThreads, sizeof(ImdctSmem), s>>>(p);
    return cudaGetLastError();
}

} // namespace glc
