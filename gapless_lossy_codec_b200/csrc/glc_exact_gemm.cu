// glc_exact_gemm.cu -- EXACT-mode transform kernels for sm_100a.
//
// What the reference computes (src/codec.rs:359-390):
//     MDCT   out[k] = (sum_{i=0..2047} block[i] * tab[k][i]) * norm      block[i] = x[i]*window[i]
//     IMDCT  out[i] = (sum_{k=0..1023} coef[k]  * tab[k][i]) * norm      then out[i] *= window[i]
// with `s` starting at +0.0, strictly ascending reduction index, the product rounded to f32 and
// then the sum rounded to f32 (no FMA).  Bit-exact parity needs exactly that chain per output, so
// this is a dense contraction  C[row][n] = sum_r A[row][r] * T[r][n]  whose reduction order is
// fixed and whose inner operation is FMUL followed by FADD.  Tensor cores, FMA contraction and
// split-K are all ruled out; the binding roof is FP32 issue, not HBM.
//
// Mapping:
//   * CTA tile = 128 rows (frame-channels) x 128 outputs, 256 threads, 8x8 outputs per thread,
//     every accumulator an independent sequential chain; 2 CTAs per SM (<=128 registers).
//   * The cosine table is re-tiled on the host so that a stage (32 reduction steps x 128 outputs,
//     16 KiB) is one contiguous block, moved by ONE bulk-copy (TMA engine, cp.async.bulk) that
//     completes on an mbarrier; double buffered.  The table (8 MiB) stays resident in the 126 MB L2.
//   * The A operand is produced in the kernel: MDCT reads interleaved PCM straight from HBM, applies
//     padding rules (512 leading zeros, zero tail), multiplies by the window and stores [r][row] in
//     shared memory (the "fused window" of the north star); IMDCT reads dense dequantised rows.
//   * VARIANT 0: scalar FMUL + FADD.  VARIANT 1/2: packed f32x2 (FMUL2/FADD2/FFMA2, sm_100+): two
//     outputs per instruction, halving issue slots.  ptxas 12.9 contracts mul.rn.f32x2 +
//     add.rn.f32x2 into FFMA2 even with -fmad=false, so the packed variants route one of the two
//     steps through an FFMA2 against a RUN-TIME constant (x*1+acc or a*b+(-0)), which is exact and
//     cannot be folded.
//   * IMDCT skips a whole stage when no row of the tile has a non-zero coefficient in that k-chunk:
//     adding +-0 to a running sum that can never be -0 is the identity, so skipping is bit-exact.
#include "glc_internal.cuh"

namespace glc
{

namespace
{

constexpr int kAStride = kBM + 4;          // floats per reduction step in the A stage (16B aligned, bank-skewed)
constexpr int kAStrideDup = 2 * kBM + 8;   // duplicated (a,a) layout for the packed variants
constexpr int kStageBytesT = kKC * kBN * 4;

typedef unsigned long long u64;

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

// One TMA-engine bulk copy global -> shared, completing `bytes` on the mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

struct RowSrc
{
    const float *base; // first sample of this row's channel
    long long start;   // sample index (per channel) of reduction step 0; negative inside the 512-zero lead-in
    long long len;     // samples per channel
    int stride;        // channels
    int valid;
};

struct GemmParams
{
    // MDCT
    const float *pcm_arena;
    const FileDesc *files;
    uint32_t n_files;
    // IMDCT
    const float *coefs_in;
    const uint32_t *stage_mask;
    // common
    uint64_t row_begin; // multiple of kBM
    uint64_t n_rows;    // = row_end
    const float *tab_tiled;
    const float *window;
    float norm;
    float *out;
    u64 rt_const; // (1.0f,1.0f) for VARIANT 1, (-0.0f,-0.0f) for VARIANT 2
};

template <int MODE, int VARIANT>
struct Smem
{
    static constexpr int kAS = (VARIANT == 0) ? kAStride : kAStrideDup;
    float a[2][kKC * kAS];
    float t[2][kKC * kBN];
    uint64_t bar[2];
    RowSrc rows[kBM];
};

// MODE 0 = MDCT (reduce over i, 64 stages), MODE 1 = IMDCT (reduce over k, 32 stages)
template <int MODE, int VARIANT>
__global__ void __launch_bounds__(kGemmThreads, 2) exact_gemm_kernel(const __grid_constant__ GemmParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef Smem<MODE, VARIANT> S;
    S &sm = *reinterpret_cast<S *>(smem_raw);
    constexpr int kAS = S::kAS;
    constexpr int kStages = (MODE == 0 ? kFrame : kHop) / kKC;
    constexpr int kNOut = (MODE == 0 ? kHop : kFrame);
    constexpr int kRed = (MODE == 0 ? kFrame : kHop);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tx = tid & 15;
    const int ty = tid >> 4;
    // linear grid, output block fastest: the CTAs of one wave share few A rows (L2-friendly)
    constexpr int kNBlocks = (MODE == 0 ? kHop : kFrame) / kBN;
    const int n_block = (int)(blockIdx.x % kNBlocks);
    const uint64_t m_tile = blockIdx.x / kNBlocks;
    const uint64_t row0 = p.row_begin + m_tile * kBM;

    if (tid == 0)
    {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- per-row source descriptors ----
    if (tid < kBM)
    {
        RowSrc rs;
        const uint64_t row = row0 + tid;
        rs.valid = row < p.n_rows;
        rs.base = nullptr;
        rs.start = 0;
        rs.len = 0;
        rs.stride = 1;
        if (rs.valid)
        {
            if (MODE == 0)
            {
                // binary search the file whose row range holds `row`
                uint32_t lo = 0, hi = p.n_files - 1;
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
            else
            {
                rs.base = p.coefs_in + row * kHop;
                rs.len = kHop;
            }
        }
        sm.rows[tid] = rs;
    }

    uint32_t mask = 0xffffffffu;
    if (MODE == 1)
        mask = p.stage_mask[row0 / kBM];
    __syncthreads();

    const float *tab_nb = p.tab_tiled + (size_t)n_block * kStages * (kKC * kBN);

    // ---- accumulators ----
    float acc[8][8];
    u64 acc2[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            acc[r][c] = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            acc2[r][c] = 0ull;
    }

    // stage iteration (all stages for MDCT; the set bits of `mask` for IMDCT)
    auto next_stage = [&](int after) -> int {
        if (MODE == 0)
            return after + 1 < kStages ? after + 1 : -1;
        const uint32_t rest = (after >= 31) ? 0u : (mask & (0xffffffffu << (after + 1)));
        return rest ? (__ffs(rest) - 1) : -1;
    };
    int cur = (MODE == 0) ? 0 : (mask ? (__ffs(mask) - 1) : -1);

    // A-operand producer.  Thread (warp, lane) covers reduction step `lane` of a stage for rows
    // warp*16 .. warp*16+15, two rows at a time so that only two registers stay live across the
    // math of a 4-step chunk (the loads are issued before the chunk, the stores after it).
    float areg[2];
    auto fetch_a2 = [&](int s, int j) {
        const int r = s * kKC + lane;
#pragma unroll
        for (int u = 0; u < 2; ++u)
        {
            const RowSrc &rs = sm.rows[warp * 16 + j + u];
            float v = 0.0f;
            if (MODE == 0)
            {
                const long long pos = rs.start + r;
                if (rs.valid && pos >= 0 && pos < rs.len)
                    v = __ldg(rs.base + pos * rs.stride);
            }
            else
            {
                if (rs.valid)
                    v = __ldg(rs.base + r);
            }
            areg[u] = v;
        }
    };
    auto store_a2 = [&](int buf, int j, float w) {
        float *dst = &sm.a[buf][lane * kAS];
        const float v0 = (MODE == 0) ? __fmul_rn(areg[0], w) : areg[0]; // block[i] = x[i]*window[i]
        const float v1 = (MODE == 0) ? __fmul_rn(areg[1], w) : areg[1];
        if (VARIANT == 0)
            *reinterpret_cast<float2 *>(dst + warp * 16 + j) = make_float2(v0, v1);
        else
            *reinterpret_cast<float4 *>(dst + 2 * (warp * 16 + j)) = make_float4(v0, v0, v1, v1);
    };
    auto window_at = [&](int s) -> float { return (MODE == 0) ? __ldg(p.window + s * kKC + lane) : 1.0f; };

    uint32_t phase = 0u; // bit b = parity to wait for on bar[b]
    int buf = 0;
    if (cur >= 0)
    {
        if (tid == 0)
        {
            mbar_expect_tx(&sm.bar[0], kStageBytesT);
            bulk_g2s(sm.t[0], tab_nb + (size_t)cur * (kKC * kBN), kStageBytesT, &sm.bar[0]);
        }
        const float w = window_at(cur);
#pragma unroll
        for (int j = 0; j < 16; j += 2)
        {
            fetch_a2(cur, j);
            store_a2(0, j, w);
        }
        __syncthreads();
    }

    while (cur >= 0)
    {
        const int nxt = next_stage(cur);
        float wn = 1.0f;
        if (nxt >= 0)
        {
            if (tid == 0)
            {
                mbar_expect_tx(&sm.bar[buf ^ 1], kStageBytesT);
                bulk_g2s(sm.t[buf ^ 1], tab_nb + (size_t)nxt * (kKC * kBN), kStageBytesT, &sm.bar[buf ^ 1]);
            }
            wn = window_at(nxt);
        }
        mbar_wait(&sm.bar[buf], (phase >> buf) & 1u);
        phase ^= 1u << buf;

        const float *As = sm.a[buf];
        const float *Ts = sm.t[buf];
#pragma unroll 1
        for (int q = 0; q < kKC / 4; ++q)
        {
            if (nxt >= 0)
                fetch_a2(nxt, 2 * q);
            if (VARIANT == 0)
            {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                {
                    const int ii = q * 4 + u;
                    const float4 a_lo = *reinterpret_cast<const float4 *>(As + ii * kAS + ty * 8);
                    const float4 a_hi = *reinterpret_cast<const float4 *>(As + ii * kAS + ty * 8 + 4);
                    const float4 t_lo = *reinterpret_cast<const float4 *>(Ts + ii * kBN + tx * 4);
                    const float4 t_hi = *reinterpret_cast<const float4 *>(Ts + ii * kBN + 64 + tx * 4);
                    const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
                    const float t[8] = {t_lo.x, t_lo.y, t_lo.z, t_lo.w, t_hi.x, t_hi.y, t_hi.z, t_hi.w};
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            acc[r][c] = __fadd_rn(acc[r][c], __fmul_rn(a[r], t[c]));
                }
            }
            else
            {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                {
                    const int ii = q * 4 + u;
                    const ulonglong2 a01 = *reinterpret_cast<const ulonglong2 *>(As + ii * kAS + ty * 16);
                    const ulonglong2 a23 = *reinterpret_cast<const ulonglong2 *>(As + ii * kAS + ty * 16 + 4);
                    const ulonglong2 a45 = *reinterpret_cast<const ulonglong2 *>(As + ii * kAS + ty * 16 + 8);
                    const ulonglong2 a67 = *reinterpret_cast<const ulonglong2 *>(As + ii * kAS + ty * 16 + 12);
                    const ulonglong2 t_lo = *reinterpret_cast<const ulonglong2 *>(Ts + ii * kBN + tx * 4);
                    const ulonglong2 t_hi = *reinterpret_cast<const ulonglong2 *>(Ts + ii * kBN + 64 + tx * 4);
                    const u64 a[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
                    const u64 t[4] = {t_lo.x, t_lo.y, t_hi.x, t_hi.y};
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                        {
                            if (VARIANT == 1) // FMUL2 then FFMA2(prod, 1.0, acc): prod*1 is exact
                                acc2[r][c] = fma2(mul2(a[r], t[c]), p.rt_const, acc2[r][c]);
                            else // FFMA2(a, t, -0.0) == RN(a*t) exactly, then FADD2
                                acc2[r][c] = add2(acc2[r][c], fma2(a[r], t[c], p.rt_const));
                        }
                }
            }
            if (nxt >= 0)
                store_a2(buf ^ 1, 2 * q, wn);
        }
        __syncthreads();
        buf ^= 1;
        cur = nxt;
    }

    // ---- epilogue: * norm (and * window for IMDCT), 2 x float4 per row ----
    const int n_lo = n_block * kBN + tx * 4;
    const int n_hi = n_lo + 64;
    float4 w_lo = make_float4(1.f, 1.f, 1.f, 1.f), w_hi = w_lo;
    if (MODE == 1)
    {
        w_lo = __ldg(reinterpret_cast<const float4 *>(p.window + n_lo));
        w_hi = __ldg(reinterpret_cast<const float4 *>(p.window + n_hi));
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        const uint64_t row = row0 + ty * 8 + r;
        if (row >= p.n_rows)
            continue;
        float v[8];
        if (VARIANT == 0)
        {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                v[c] = acc[r][c];
        }
        else
        {
#pragma unroll
            for (int c = 0; c < 4; ++c)
            {
                v[2 * c] = __uint_as_float((uint32_t)(acc2[r][c] & 0xffffffffull));
                v[2 * c + 1] = __uint_as_float((uint32_t)(acc2[r][c] >> 32));
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
            v[c] = __fmul_rn(v[c], p.norm);
        if (MODE == 1)
        {
            v[0] = __fmul_rn(v[0], w_lo.x);
            v[1] = __fmul_rn(v[1], w_lo.y);
            v[2] = __fmul_rn(v[2], w_lo.z);
            v[3] = __fmul_rn(v[3], w_lo.w);
            v[4] = __fmul_rn(v[4], w_hi.x);
            v[5] = __fmul_rn(v[5], w_hi.y);
            v[6] = __fmul_rn(v[6], w_hi.z);
            v[7] = __fmul_rn(v[7], w_hi.w);
        }
        float *orow = p.out + row * kNOut;
        *reinterpret_cast<float4 *>(orow + n_lo) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(orow + n_hi) = make_float4(v[4], v[5], v[6], v[7]);
    }
    (void)kRed;
}

template <int MODE, int VARIANT>
cudaError_t launch_gemm(const GemmParams &p, cudaStream_t s)
{
    typedef Smem<MODE, VARIANT> S;
    static bool configured = false;
    const size_t smem = sizeof(S);
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(exact_gemm_kernel<MODE, VARIANT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
        configured = true;
    }
    if (p.n_rows <= p.row_begin)
        return cudaSuccess;
    const uint64_t m_tiles = (p.n_rows - p.row_begin + kBM - 1) / kBM;
    if (m_tiles == 0)
        return cudaSuccess;
    constexpr int kNBlocks = (MODE == 0 ? kHop : kFrame) / kBN;
    const uint64_t n_ctas = m_tiles * kNBlocks;
    if (n_ctas > 0x7fffffffull)
        return cudaErrorInvalidValue;
    exact_gemm_kernel<MODE, VARIANT><<<(unsigned)n_ctas, kGemmThreads, smem, s>>>(p);
    return cudaGetLastError();
}

} // namespace

cudaError_t launch_mdct_exact(const MdctLaunch &l, cudaStream_t s)
{
    GemmParams p{};
    p.pcm_arena = l.pcm_arena;
    p.files = l.files;
    p.n_files = l.n_files;
    p.row_begin = l.row_begin;
    p.n_rows = l.row_end;
    p.tab_tiled = l.tab_tiled;
    p.window = l.window;
    p.norm = l.norm;
    p.out = l.coefs;
    switch (l.variant)
    {
    case 1:
        p.rt_const = 0x3f8000003f800000ull;
        return launch_gemm<0, 1>(p, s);
    case 2:
        p.rt_const = 0x8000000080000000ull;
        return launch_gemm<0, 2>(p, s);
    default:
        return launch_gemm<0, 0>(p, s);
    }
}

cudaError_t launch_imdct_exact(const ImdctLaunch &l, cudaStream_t s)
{
    GemmParams p{};
    p.coefs_in = l.coefs;
    p.stage_mask = l.stage_mask;
    p.row_begin = l.row_begin;
    p.n_rows = l.row_end;
    p.tab_tiled = l.tab_tiled;
    p.window = l.window;
    p.norm = l.norm;
    p.out = l.blocks;
    switch (l.variant)
    {
    case 1:
        p.rt_const = 0x3f8000003f800000ull;
        return launch_gemm<1, 1>(p, s);
    case 2:
        p.rt_const = 0x8000000080000000ull;
        return launch_gemm<1, 2>(p, s);
    default:
        return launch_gemm<1, 0>(p, s);
    }
}

} // namespace glc
