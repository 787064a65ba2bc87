// glc_flac_kernels.cu -- FLAC encoder of the reference (src/flac.rs), one CTA per FLAC block (sm_100a).
//
// The reference encoder is a fixed function of its input: fixed predictor
// whose order depends only on the level (src/flac.rs:692-700), partition order from level and block
// size (:590-608), Rice parameter floor(log2(mean|r|)) capped at 14 (:515-552), independent
// channels, 16-bit samples.  Blocks are independent, so the path is two streaming passes:
//
//   pass A  "measure": f32 -> i16 (:955-958), residuals, per-partition Rice parameters, exact frame
//           size in bytes.  Output: i16 arena, parameters, frame_bytes[].
//   (scan of frame_bytes on the device gives every frame its final byte offset)
//   pass B  "emit":    every thread owns a RUN of 16 consecutive samples of a channel: code lengths, run
//           totals, block-wide exclusive scan of the 256 run totals, then the thread assembles its run's
//           bits word by word in a register and stores whole words into a zeroed shared-memory bit buffer
//           (plain stores; atomicOr only for the first and last word of the run, which neighbours share),
//           CRC-8 of the header, parallel CRC-16 of the frame (per-thread partial CRCs combined with
//           x^(8n) mod P multiplications), word-wise copy to the frame's final position.
//
// Integer work only: bit-exact against the oracle is the bar.  HBM-bound: 4 B/sample read in pass
// A, 2 B/sample written, 2 B/sample + bitstream in pass B.
#include <algorithm>
#include <type_traits>

#include "glc_internal.cuh"

namespace glc
{

namespace
{

constexpr int kFlacThreads = 256;
constexpr int kMaxParts = 64; // partition order <= 6

struct BlockGeom
{
    uint32_t file;       // index into the file table
    uint32_t frame_no;   // frame number inside the file
    uint32_t bs;         // samples per channel in this block
    uint32_t bs_pad;     // channel stride of the planar shared-memory copy: bs rounded up to 8 samples (16 bytes)
    uint32_t ch;
    uint32_t rate;
    uint64_t smp_off;    // interleaved sample offset of the block inside the file
};

__device__ __forceinline__ BlockGeom locate_block(const FlacFileDesc *files, uint32_t n_files, uint64_t b)
{
    uint32_t lo = 0, hi = n_files - 1;
    while (lo < hi)
    {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (files[mid].first_block <= b)
            lo = mid;
        else
            hi = mid - 1;
    }
    const FlacFileDesc &fd = files[lo];
    BlockGeom g;
    g.file = lo;
    g.frame_no = (uint32_t)(b - fd.first_block);
    g.ch = fd.channels;
    g.rate = fd.sample_rate;
    g.smp_off = (uint64_t)g.frame_no * fd.block_size * fd.channels;
    const uint64_t remaining = fd.n_samples - g.smp_off;
    const uint64_t per_ch = remaining / fd.channels; // src/flac.rs:1026-1027
    g.bs = (uint32_t)(per_ch < fd.block_size ? per_ch : fd.block_size);
    g.bs_pad = (g.bs + 7u) & ~7u;
    return g;
}

// src/flac.rs:692-700
__device__ __forceinline__ int predictor_order(int level, uint32_t bs)
{
    if (level == 0)
        return 0;
    if (level == 1)
        return bs >= 1 ? 1 : 0;
    if (level == 2)
        return bs >= 2 ? 2 : 0;
    if (level <= 4)
        return bs >= 3 ? 3 : 0;
    return bs >= 4 ? 4 : 0;
}

// src/flac.rs:590-608
__device__ __forceinline__ int partition_order(int level, uint32_t bs, int order)
{
    int tz = bs ? (__ffs(bs) - 1) : 32;
    if (tz > 8)
        tz = 8;
    const int cap = level == 0 ? 0 : (level <= 2 ? 2 : (level <= 5 ? 4 : 6));
    int po = cap < tz ? cap : tz;
    while (po > 0)
    {
        const uint32_t ps = bs >> po;
        if (ps > (uint32_t)order && ps >= 4)
            break;
        --po;
    }
    return po;
}

// (s * 32767.0).clamp(-32768.0, 32767.0) as i16          src/flac.rs:955-958
__device__ __forceinline__ int f32_to_i16(float s)
{
    // one instruction: truncate toward zero, saturate to i16, NaN -> 0 (= Rust's clamp then `as i16`)
    short q;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(q) : "f"(__fmul_rn(s, 32767.0f)));
    return (int)q;
}

// residual of the fixed predictors, src/flac.rs:498-507 (no overflow for 16-bit input)

// residual of the fixed predictors, src/flac.rs:498-507 (no overflow for 16-bit input); used by the compact
// (not unrolled) paths of ragged tail blocks, which read the planar copy directly
__device__ __forceinline__ int residual_at(const int16_t *s, uint32_t i, int order)
{
    switch (order)
    {
    case 1:
        return (int)s[i] - (int)s[i - 1];
    case 2:
        return (int)s[i] - (2 * (int)s[i - 1] - (int)s[i - 2]);
    case 3:
        return (int)s[i] - (3 * (int)s[i - 1] - 3 * (int)s[i - 2] + (int)s[i - 3]);
    case 4:
        return (int)s[i] - (4 * (int)s[i - 1] - 6 * (int)s[i - 2] + 4 * (int)s[i - 3] - (int)s[i - 4]);
    default:
        return (int)s[i];
    }
}

__device__ __forceinline__ uint32_t zigzag(int r) // src/flac.rs:560-567
{
    return r >= 0 ? ((uint32_t)r << 1) : ((((uint32_t)(-(r + 1))) << 1) | 1u);
}

__device__ __forceinline__ uint32_t header_bytes(uint32_t bs, uint32_t frame_no)
{
    uint32_t n = 4;
    n += frame_no < 0x80 ? 1 : (frame_no < 0x800 ? 2 : (frame_no < 0x10000 ? 3 : (frame_no < 0x200000 ? 4 : (frame_no < 0x4000000 ? 5 : (frame_no < 0x80000000u ? 6 : 7)))));
    uint32_t code;
    switch (bs)
    {
    case 192: case 576: case 1152: case 2304: case 4608:
    case 256: case 512: case 1024: case 2048: case 4096: case 8192: case 16384: case 32768:
        code = 1;
        break;
    default:
        code = bs < 256 ? 6 : 7;
        break;
    }
    if (code == 6)
        n += 1;
    else if (code == 7)
        n += 2;
    return n + 1; // + CRC-8
}

// loads the block's samples as planar i16 into shared memory: s[c*bs_pad + i].  Four samples per load
// (float4 from the PCM arena / 8-byte loads from the i16 arena), all of a thread's loads issued before
// the first use so that several are in flight.  File starts are 4-sample aligned in both arenas and
// every block before a file's last one holds a multiple of 4 samples, so the vector path applies
// whenever (block offset) % 4 == 0.
__device__ __forceinline__ void scatter_sample(int16_t *s_smp, const BlockGeom &g, uint32_t e, int q)
{
    uint32_t i, c;
    if (g.ch == 1)
    {
        i = e;
        c = 0;
    }
    else if (g.ch == 2)
    {
        i = e >> 1;
        c = e & 1u;
    }
    else
    {
        i = e / g.ch;
        c = e - i * g.ch;
    }
    s_smp[c * g.bs_pad + i] = (int16_t)q;
}

template <bool FROM_F32>
__device__ __forceinline__ void load_block(const FlacLaunch &p, const FlacFileDesc &fd, const BlockGeom &g,
                                           int16_t *s_smp)
{
    const uint32_t n = g.bs * g.ch;
    const float *src = p.pcm_arena + fd.pcm_off + g.smp_off;
    int16_t *arena = p.i16_arena + fd.i16_off + g.smp_off;
    // vectorisable quads: the block must start on a 4-sample boundary of the arena it is read from
    const uint64_t start = (FROM_F32 ? fd.pcm_off : fd.i16_off) + g.smp_off;
    const bool arena_ok = FROM_F32 ? ((fd.i16_off + g.smp_off) & 3ull) == 0 : true; // short4 stores into the i16 arena
    const uint32_t n4 = ((start & 3ull) == 0 && arena_ok) ? n / 4 : 0;
    constexpr int kBatch = 4;
    for (uint32_t q0 = threadIdx.x; q0 < n4; q0 += blockDim.x * kBatch)
    {
        if (FROM_F32)
        {
            float4 v[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
            {
                const uint32_t q = q0 + u * blockDim.x;
                if (q < n4)
                    v[u] = __ldg(reinterpret_cast<const float4 *>(src) + q);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
            {
                const uint32_t q = q0 + u * blockDim.x;
                if (q < n4)
                {
                    const int a = f32_to_i16(v[u].x), b = f32_to_i16(v[u].y), c = f32_to_i16(v[u].z), d = f32_to_i16(v[u].w);
                    short4 o;
                    o.x = (short)a;
                    o.y = (short)b;
                    o.z = (short)c;
                    o.w = (short)d;
                    reinterpret_cast<short4 *>(arena)[q] = o;
                    scatter_sample(s_smp, g, 4 * q, a);
                    scatter_sample(s_smp, g, 4 * q + 1, b);
                    scatter_sample(s_smp, g, 4 * q + 2, c);
                    scatter_sample(s_smp, g, 4 * q + 3, d);
                }
            }
        }
        else
        {
            short4 v[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
            {
                const uint32_t q = q0 + u * blockDim.x;
                if (q < n4)
                    v[u] = reinterpret_cast<const short4 *>(arena)[q];
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
            {
                const uint32_t q = q0 + u * blockDim.x;
                if (q < n4)
                {
                    scatter_sample(s_smp, g, 4 * q, v[u].x);
                    scatter_sample(s_smp, g, 4 * q + 1, v[u].y);
                    scatter_sample(s_smp, g, 4 * q + 2, v[u].z);
                    scatter_sample(s_smp, g, 4 * q + 3, v[u].w);
                }
            }
        }
    }
    for (uint32_t e = 4 * n4 + threadIdx.x; e < n; e += blockDim.x)
    {
        int q;
        if (FROM_F32)
        {
            q = f32_to_i16(__ldg(src + e));
            arena[e] = (int16_t)q;
        }
        else
            q = arena[e];
        scatter_sample(s_smp, g, e, q);
    }
    if (FROM_F32 && g.frame_no + 1 == fd.n_blocks)
    {
        // The MD5 covers every converted sample (src/flac.rs:1004), also a trailing partial sample
        // frame that no block encodes (input length not a multiple of the channel count).
        const uint64_t done = g.smp_off + n;
        for (uint64_t e = done + threadIdx.x; e < fd.n_samples; e += blockDim.x)
            p.i16_arena[fd.i16_off + e] = (int16_t)f32_to_i16(__ldg(p.pcm_arena + fd.pcm_off + e));
    }
}

constexpr int kRunLen = 16; // consecutive samples per thread and channel: 256 x 16 = the largest block (src/flac.rs:983-995)
static_assert(kFlacThreads * kRunLen == 4096, "a thread run times the CTA size must cover the largest block");

__device__ __forceinline__ uint32_t rice_param32(uint32_t sum_abs, uint32_t n) // src/flac.rs:515-552
{
    if (n == 0)
        return 0;
    const uint32_t mean = sum_abs / n;
    if (mean == 0)
        return 0;
    const int lg = 31 - __clz((int)mean);
    return lg < 14 ? (uint32_t)lg : 14u;
}

struct ChannelPlan
{
    int order, po;
    uint32_t dps, nparts;
    bool run_uniform; // every thread's run of kRunLen samples lies inside one partition
    int dps_shift;    // log2(dps) when dps is a power of two, else -1
    __device__ __forceinline__ uint32_t part_of(uint32_t i) const { return dps_shift >= 0 ? i >> dps_shift : i / dps; }
};

__device__ __forceinline__ ChannelPlan plan_channel(int level, uint32_t bs)
{
    ChannelPlan cp;
    cp.order = predictor_order(level, bs);
    cp.po = cp.order ? partition_order(level, bs, cp.order) : 0;
    cp.dps = bs >> cp.po;
    cp.nparts = 1u << cp.po;
    cp.run_uniform = (cp.dps % kRunLen) == 0;
    cp.dps_shift = (cp.dps & (cp.dps - 1u)) == 0 ? (__ffs(cp.dps) - 1) : -1;
    return cp;
}

// The thread's run of kRunLen samples plus the four before it (the predictor's history), as ints.
// s = channel base in the planar copy (16-byte aligned, stride bs_pad), i0 = first sample of the run (a
// multiple of 16).  Samples outside the block read as 0 and are never used.
__device__ __forceinline__ void load_run(const int16_t *s, uint32_t i0, uint32_t bs_pad, int (&v)[kRunLen + 4])
{
    // history: samples i0-4 .. i0-1 (8 bytes, aligned)
    if (i0 >= 4)
    {
        const uint2 h = *reinterpret_cast<const uint2 *>(s + i0 - 4);
        v[0] = (int)(short)(h.x & 0xffffu);
        v[1] = (int)(short)(h.x >> 16);
        v[2] = (int)(short)(h.y & 0xffffu);
        v[3] = (int)(short)(h.y >> 16);
    }
    else
        v[0] = v[1] = v[2] = v[3] = 0;
#pragma unroll
    for (int q = 0; q < kRunLen / 8; ++q)
    {
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (i0 + 8 * q < bs_pad)
            w = *reinterpret_cast<const uint4 *>(s + i0 + 8 * q);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
        {
            v[4 + 8 * q + 2 * e] = (int)(short)(ww[e] & 0xffffu);
            v[4 + 8 * q + 2 * e + 1] = (int)(short)(ww[e] >> 16);
        }
    }
}

// residual of the fixed predictors at run position j (sample i0 + j), src/flac.rs:498-507; ORDER known at compile
// time (the hot paths are instantiated per predictor order: no per-sample switch, and one launch runs ONE
// instantiation, which is what keeps its loop inside the instruction cache)
template <int ORDER>
__device__ __forceinline__ int run_residual_t(const int (&v)[kRunLen + 4], int j)
{
    const int x0 = v[4 + j], x1 = v[3 + j], x2 = v[2 + j], x3 = v[1 + j], x4 = v[j];
    if (ORDER == 1)
        return x0 - x1;
    if (ORDER == 2)
        return x0 - (2 * x1 - x2);
    if (ORDER == 3)
        return x0 - (3 * x1 - 3 * x2 + x3);
    if (ORDER == 4)
        return x0 - (4 * x1 - 6 * x2 + 4 * x3 - x4);
    return x0;
}

// ------------------------------------------------------------------ pass A

__global__ void __launch_bounds__(kFlacThreads, 4) flac_measure_kernel(const FlacLaunch p, uint8_t *rice_k /* [blocks][ch][64] */,
                                                                       uint32_t max_ch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int16_t *s_smp = reinterpret_cast<int16_t *>(smem_raw);
    __shared__ uint32_t s_psum[kMaxParts];
    __shared__ uint32_t s_k[kMaxParts];
    __shared__ uint32_t s_bits;
    __shared__ unsigned long long s_total_bits;

    const uint64_t b = blockIdx.x;
    if (b >= p.n_blocks_total)
        return;
    const BlockGeom g = locate_block(p.files, p.n_files, b);
    const FlacFileDesc &fd = p.files[g.file];
    load_block<true>(p, fd, g, s_smp);
    const int tid = threadIdx.x, lane = tid & 31;
    const ChannelPlan cp = plan_channel(p.level, g.bs);
    const uint32_t i0 = (uint32_t)tid * kRunLen;
    if (tid == 0)
        s_total_bits = 0;

    for (uint32_t c = 0; c < g.ch; ++c)
    {
        __syncthreads();
        if (cp.order == 0)
        {
            if (tid == 0)
                s_total_bits += 8ull + 16ull * g.bs; // verbatim subframe, src/flac.rs:722-729
            continue;
        }
        if (tid < kMaxParts)
            s_psum[tid] = 0;
        if (tid == 0)
            s_bits = 0;
        __syncthreads();
        // ---- residuals of the thread's run and per-partition sums of |r| ----
        // fast path (every block but a ragged tail): the whole run lies inside the block and inside one
        // partition; unrolled, everything in registers.  Otherwise a compact loop over the planar copy.
        const int16_t *s = s_smp + c * g.bs_pad;
        const bool fastp = cp.run_uniform && i0 + kRunLen <= g.bs;
        uint32_t zz[kRunLen];
        if (fastp)
        {
            int v[kRunLen + 4];
            load_run(s, i0, g.bs_pad, v);
            uint32_t run_abs = 0;
            auto body = [&](auto order_tag) {
                constexpr int ORDER = decltype(order_tag)::value;
#pragma unroll
                for (int j = 0; j < kRunLen; ++j)
                {
                    const int r = run_residual_t<ORDER>(v, j);
                    const bool valid = i0 != 0 || j >= ORDER; // only thread 0 holds warm-up samples
                    zz[j] = zigzag(r);
                    run_abs += valid ? (uint32_t)(r < 0 ? -r : r) : 0u;
                }
            };
            switch (cp.order)
            {
            case 1: body(std::integral_constant<int, 1>()); break;
            case 2: body(std::integral_constant<int, 2>()); break;
            case 3: body(std::integral_constant<int, 3>()); break;
            default: body(std::integral_constant<int, 4>()); break;
            }
            atomicAdd(&s_psum[cp.part_of(i0)], run_abs);
        }
        else
            for (uint32_t i = max(i0, (uint32_t)cp.order); i < min(i0 + kRunLen, g.bs); ++i)
            {
                const int r = residual_at(s, i, cp.order);
                atomicAdd(&s_psum[cp.part_of(i)], (uint32_t)(r < 0 ? -r : r));
            }
        __syncthreads();
        uint8_t *kout = rice_k + (b * max_ch + c) * kMaxParts;
        if ((uint32_t)tid < cp.nparts)
        {
            const uint32_t cnt = tid == 0 ? cp.dps - (uint32_t)cp.order : cp.dps;
            const uint32_t k = rice_param32(s_psum[tid], cnt);
            s_k[tid] = k;
            kout[tid] = (uint8_t)k;
            if (cnt)
                atomicAdd(&s_bits, 4u); // an empty first partition writes nothing (:632-635)
        }
        __syncthreads();
        // ---- code lengths ----
        uint32_t mine = 0;
        if (fastp)
        {
            const uint32_t k = s_k[cp.part_of(i0)];
            const int first_valid = i0 == 0 ? cp.order : 0;
#pragma unroll
            for (int j = 0; j < kRunLen; ++j)
                mine += j >= first_valid ? (zz[j] >> k) + 1u + k : 0u;
        }
        else
            for (uint32_t i = max(i0, (uint32_t)cp.order); i < min(i0 + kRunLen, g.bs); ++i)
            {
                const uint32_t k = s_k[cp.part_of(i)];
                mine += (zigzag(residual_at(s, i, cp.order)) >> k) + 1u + k;
            }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0 && mine)
            atomicAdd(&s_bits, mine);
        __syncthreads();
        if (tid == 0)
            s_total_bits += 8ull + 16ull * cp.order + 6ull + s_bits;
    }
    __syncthreads();
    if (tid == 0)
    {
        const unsigned long long body = (s_total_bits + 7ull) >> 3;
        p.frame_bytes[b] = (uint32_t)(header_bytes(g.bs, g.frame_no) + body + 2ull);
    }
}

// ------------------------------------------------------------------ pass B

// OR the low `nbits` (<= 32) bits of v into the big-endian bit buffer at bit position `pos`
__device__ __forceinline__ void put_bits(uint32_t *buf, unsigned long long pos, uint32_t v, int nbits)
{
    const uint32_t word = (uint32_t)(pos >> 5);
    const int off = (int)(pos & 31);
    const unsigned long long w = (unsigned long long)v << (64 - nbits - off);
    const uint32_t hi = (uint32_t)(w >> 32), lo = (uint32_t)w;
    if (hi)
        atomicOr(buf + word, hi);
    if (lo)
        atomicOr(buf + word + 1, lo);
}

// A thread's contiguous bit range of the frame, written front to back through a 64-bit accumulator: codes
// are appended below the bits already collected and every time 32 of them are complete the word goes to the
// (zeroed) buffer.  Only the first word of the range (its leading bits belong to the previous range) and the
// final partial word (its trailing bits belong to the next range) are merged with atomicOr; every word in
// between belongs to this thread alone and is a plain store.
struct RunWriter
{
    uint32_t *buf;
    unsigned long long acc;
    uint32_t wi, nb;
    bool first;
    __device__ __forceinline__ void begin(uint32_t *b, uint32_t first_bit)
    {
        buf = b;
        wi = first_bit >> 5;
        nb = first_bit & 31u; // the leading bits of the first word stay zero in this thread's copy
        acc = 0;
        first = true;
    }
    __device__ __forceinline__ void emit(uint32_t word)
    {
        if (first)
        {
            atomicOr(buf + wi, word);
            first = false;
        }
        else
            buf[wi] = word;
        ++wi;
    }
    // n (1..32) bits; `code` has no bits set at or above bit n
    __device__ __forceinline__ void append(uint32_t code, uint32_t n)
    {
        acc = (acc << n) | code;
        nb += n;
        if (nb >= 32u)
        {
            emit((uint32_t)(acc >> (nb - 32u)));
            nb -= 32u;
        }
    }
    // a Rice code: q zeros, the stop bit and the k low bits (code = stop | low, k + 1 bits)
    __device__ __forceinline__ void append_rice(uint32_t q, uint32_t code, uint32_t k1)
    {
        if (q + k1 <= 32u)
        {
            append(code, q + k1); // the common case: the zeros ride in front of the code
            return;
        }
        while (q >= 32u) // long unary run: whole words of zeros
        {
            append(0u, 32u);
            q -= 32u;
        }
        if (q)
            append(0u, q);
        append(code, k1);
    }
    __device__ __forceinline__ void finish()
    {
        if (nb)
            atomicOr(buf + wi, (uint32_t)(acc << (32u - nb)));
    }
};

__device__ __forceinline__ uint32_t gf16_mul(uint32_t a, uint32_t b) // mod x^16+x^15+x^2+1
{
    uint32_t r = 0;
#pragma unroll
    for (int i = 15; i >= 0; --i)
    {
        r <<= 1;
        if (r & 0x10000u)
            r ^= 0x18005u;
        if ((b >> i) & 1u)
            r ^= a;
    }
    return r & 0xffffu;
}

__device__ __forceinline__ uint32_t buf_byte(const uint32_t *buf, uint32_t j)
{
    return (buf[j >> 2] >> (24 - 8 * (j & 3))) & 0xffu;
}

__device__ void write_frame_header(uint32_t *bitbuf, const BlockGeom &g)
{
    // frame header, src/flac.rs:759-871
    uint8_t h[16];
    uint32_t n = 0;
    h[n++] = 0xFF;
    h[n++] = 0xF8;
    uint32_t bsb;
    switch (g.bs)
    {
    case 192: bsb = 1; break;
    case 576: bsb = 2; break;
    case 1152: bsb = 3; break;
    case 2304: bsb = 4; break;
    case 4608: bsb = 5; break;
    case 256: bsb = 8; break;
    case 512: bsb = 9; break;
    case 1024: bsb = 10; break;
    case 2048: bsb = 11; break;
    case 4096: bsb = 12; break;
    case 8192: bsb = 13; break;
    case 16384: bsb = 14; break;
    case 32768: bsb = 15; break;
    default: bsb = g.bs < 256 ? 6 : 7; break;
    }
    uint32_t srb;
    switch (g.rate)
    {
    case 88200: srb = 1; break;
    case 176400: srb = 2; break;
    case 192000: srb = 3; break;
    case 8000: srb = 4; break;
    case 16000: srb = 5; break;
    case 22050: srb = 6; break;
    case 24000: srb = 7; break;
    case 32000: srb = 8; break;
    case 44100: srb = 9; break;
    case 48000: srb = 10; break;
    case 96000: srb = 11; break;
    default: srb = 0; break;
    }
    h[n++] = (uint8_t)((bsb << 4) | srb);
    const uint32_t chb = g.ch == 1 ? 0 : (g.ch == 2 ? 1 : ((g.ch - 1) & 0xF));
    h[n++] = (uint8_t)((chb << 4) | (4u << 1)); // 16 bits per sample -> 0b100, reserved 0
    const uint32_t v = g.frame_no; // UTF-8 style number, src/flac.rs:427-478
    if (v < 0x80)
        h[n++] = (uint8_t)v;
    else if (v < 0x800)
    {
        h[n++] = (uint8_t)(0xC0 | ((v >> 6) & 0x1F));
        h[n++] = (uint8_t)(0x80 | (v & 0x3F));
    }
    else if (v < 0x10000)
    {
        h[n++] = (uint8_t)(0xE0 | ((v >> 12) & 0x0F));
        h[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F));
        h[n++] = (uint8_t)(0x80 | (v & 0x3F));
    }
    else if (v < 0x200000)
    {
        h[n++] = (uint8_t)(0xF0 | ((v >> 18) & 0x07));
        h[n++] = (uint8_t)(0x80 | ((v >> 12) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F));
        h[n++] = (uint8_t)(0x80 | (v & 0x3F));
    }
    else if (v < 0x4000000)
    {
        h[n++] = (uint8_t)(0xF8 | ((v >> 24) & 0x03));
        h[n++] = (uint8_t)(0x80 | ((v >> 18) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 12) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F));
        h[n++] = (uint8_t)(0x80 | (v & 0x3F));
    }
    else if (v < 0x80000000u)
    {
        h[n++] = (uint8_t)(0xFC | ((v >> 30) & 0x01));
        h[n++] = (uint8_t)(0x80 | ((v >> 24) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 18) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 12) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F));
        h[n++] = (uint8_t)(0x80 | (v & 0x3F));
    }
    else
    {
        h[n++] = 0xFE;
        h[n++] = (uint8_t)(0x80 | ((v >> 30) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 24) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 18) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 12) & 0x3F));
        h[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F));
        h[n++] = (uint8_t)(0x80 | (v & 0x3F));
    }
    if (bsb == 6)
        h[n++] = (uint8_t)((g.bs - 1) & 0xFF);
    else if (bsb == 7)
    {
        h[n++] = (uint8_t)(((g.bs - 1) >> 8) & 0xFF);
        h[n++] = (uint8_t)((g.bs - 1) & 0xFF);
    }
    uint32_t c8 = 0; // CRC-8 poly 0x07, src/flac.rs:19-51
    for (uint32_t j = 0; j < n; ++j)
    {
        c8 ^= h[j];
        for (int i = 0; i < 8; ++i)
            c8 = (c8 & 0x80u) ? (((c8 << 1) ^ 0x07u) & 0xFFu) : ((c8 << 1) & 0xFFu);
    }
    h[n++] = (uint8_t)c8;
    for (uint32_t j = 0; j < n; ++j)
        put_bits(bitbuf, 8ull * j, h[j], 8);
}

// GLOBAL_SCRATCH: the bit buffer lives in global memory (frames too large for shared memory); the common
// instantiation addresses it as shared memory outright instead of through generic pointers
template <bool GLOBAL_SCRATCH>
__global__ void __launch_bounds__(kFlacThreads, 4) flac_emit_kernel(const FlacLaunch p, const uint8_t *rice_k, uint32_t max_ch,
                                                                 const uint64_t *frame_off, uint8_t *out_arena,
                                                                 uint32_t smp_bytes, uint32_t buf_words,
                                                                 uint32_t *g_scratch /* null = bit buffer in smem */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int16_t *s_smp = reinterpret_cast<int16_t *>(smem_raw);
    uint32_t *bitbuf = GLOBAL_SCRATCH ? g_scratch + (size_t)blockIdx.x * buf_words
                                      : reinterpret_cast<uint32_t *>(smem_raw + smp_bytes);
    constexpr int kWarps = kFlacThreads / 32;
    __shared__ uint32_t s_wsum[2][kWarps]; // code bits per warp (runs in sample order), double-buffered by channel parity
    __shared__ uint16_t s_crc_tab[4][256]; // [k][x] = CRC-16 of byte x followed by k zero bytes (slicing by 4)
    __shared__ uint16_t s_xpow[32];
    __shared__ uint32_t s_crc_part[kWarps];
    // multiplication by x^(8 * chunk * 2^L) mod P as two byte-indexed tables per merge level L (the CRC is linear:
    // a * b = low byte of a times b XOR high byte of a times b); rebuilt only when the chunk size changes
    __shared__ uint16_t s_mul[6][2][256];
    __shared__ uint16_t s_bitmul[6][16];
    int cur_m = -1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        // CRC-16 table (poly 0x8005, MSB first, init 0; src/flac.rs:54-80) and x^(8*2^j) mod P
        uint32_t crc = (uint32_t)tid << 8;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x8005u) : (crc << 1);
        s_crc_tab[0][tid] = (uint16_t)crc;
        // one more zero byte per table: crc <- (crc << 8) ^ tab0[crc >> 8]; tab0[y] is recomputed bitwise here
        for (int t = 1; t < 4; ++t)
        {
            uint32_t hi = crc & 0xff00u;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                hi = (hi & 0x8000u) ? ((hi << 1) ^ 0x8005u) : (hi << 1);
            crc = ((crc << 8) & 0xffffu) ^ (hi & 0xffffu);
            s_crc_tab[t][tid] = (uint16_t)crc;
        }
        if (tid == 0)
        {
            uint32_t x = 0x0100u; // x^8
            for (int j = 0; j < 32; ++j)
            {
                s_xpow[j] = (uint16_t)x;
                x = gf16_mul(x, x);
            }
        }
    }
    __syncthreads();
    const uint32_t i0 = (uint32_t)tid * kRunLen;

    for (uint64_t b = blockIdx.x; b < p.n_blocks_total; b += gridDim.x)
    {
        const BlockGeom g = locate_block(p.files, p.n_files, b);
        const FlacFileDesc &fd = p.files[g.file];
        const uint32_t fbytes = p.frame_bytes[b];
        const uint32_t fwords = (fbytes + 3) >> 2;
        for (uint32_t w = tid; w < fwords + 1 && w < buf_words; w += kFlacThreads)
            bitbuf[w] = 0;
        load_block<false>(p, fd, g, s_smp);
        __syncthreads();

        const uint32_t hbytes = header_bytes(g.bs, g.frame_no);
        if (tid == 0)
            write_frame_header(bitbuf, g);

        const ChannelPlan cp = plan_channel(p.level, g.bs);
        const int order = cp.order;
        uint32_t sub0 = hbytes * 8u; // first bit of the current subframe (the same value in every thread)

        for (uint32_t c = 0; c < g.ch; ++c)
        {
            const int16_t *s = s_smp + c * g.bs_pad;
            const uint8_t *kin = rice_k + (b * max_ch + c) * kMaxParts;
            const bool have = i0 < g.bs;
            // ---- bits of the thread's run ----
            // fast path: the whole run lies inside the block and inside one partition (every block but a ragged
            // tail): unrolled, samples and codes in registers, one Rice parameter per run, and a partition can
            // only start at the run's first sample (or at sample `order` of the block, in thread 0's run).
            // Everything else (tail blocks, verbatim subframes) goes through compact loops over the planar copy.
            const bool fastp = order != 0 && cp.run_uniform && i0 + kRunLen <= g.bs;
            const uint32_t fixed = order == 0 ? 8u : 8u + 16u * (uint32_t)order + 6u; // bits before the first run
            uint32_t zz[kRunLen];
            uint32_t run_bits = 0;
            uint32_t k_run = 0;
            int pj = -1; // run position at which a partition starts (its 4-bit parameter goes in front), or -1
            if (fastp)
            {
                int v[kRunLen + 4];
                load_run(s, i0, g.bs_pad, v);
                k_run = (uint32_t)kin[cp.part_of(i0)];
                const bool starts = cp.dps_shift >= 0 ? (i0 & (cp.dps - 1u)) == 0 : (i0 % cp.dps) == 0;
                pj = i0 == 0 ? order : (starts ? 0 : -1);
                auto body = [&](auto order_tag) {
                    constexpr int ORDER = decltype(order_tag)::value;
#pragma unroll
                    for (int j = 0; j < kRunLen; ++j)
                    {
                        const bool valid = i0 != 0 || j >= ORDER;
                        zz[j] = zigzag(run_residual_t<ORDER>(v, j));
                        run_bits += valid ? (zz[j] >> k_run) + 1u + k_run : 0u;
                    }
                };
                switch (order)
                {
                case 1: body(std::integral_constant<int, 1>()); break;
                case 2: body(std::integral_constant<int, 2>()); break;
                case 3: body(std::integral_constant<int, 3>()); break;
                default: body(std::integral_constant<int, 4>()); break;
                }
                run_bits += pj >= 0 ? 4u : 0u;
            }
            else if (order == 0)
                run_bits = have ? 16u * min((uint32_t)kRunLen, g.bs - i0) : 0u; // verbatim samples (src/flac.rs:722-729)
            else
                for (uint32_t i = max(i0, (uint32_t)order); i < min(i0 + kRunLen, g.bs); ++i)
                {
                    const uint32_t part = cp.part_of(i);
                    const uint32_t k = (uint32_t)kin[part];
                    const uint32_t pstart = part == 0 ? (uint32_t)order : part * cp.dps;
                    run_bits += (zigzag(residual_at(s, i, order)) >> k) + 1u + k + (i == pstart ? 4u : 0u);
                }
            // ---- exclusive scan of the 256 run totals ----
            uint32_t incl = run_bits;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += t;
            }
            if (lane == 31)
                s_wsum[c & 1][warp] = incl;
            __syncthreads(); // the only barrier per channel (s_wsum alternates between two copies)
            uint32_t before = 0, chan_bits = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w)
            {
                const uint32_t t = s_wsum[c & 1][w];
                before += w < warp ? t : 0u;
                chan_bits += t;
            }
            const uint32_t first_bit = sub0 + fixed + before + (incl - run_bits);
            if (tid == 0)
            {
                if (order == 0)
                    put_bits(bitbuf, sub0, 0x02u, 8); // verbatim subframe: 0 | 000001 | 0 (src/flac.rs:704-720)
                else
                {
                    // subframe header + warm-up + residual header, src/flac.rs:704-720, 733-736, 611-614
                    put_bits(bitbuf, sub0, (0x08u | (uint32_t)order) << 1, 8); // 0 | 001ooo | 0
                    for (int i = 0; i < order; ++i)
                        put_bits(bitbuf, sub0 + 8ull + 16ull * i, (uint32_t)(uint16_t)s[i], 16);
                    put_bits(bitbuf, sub0 + 8ull + 16ull * order, (uint32_t)cp.po, 6); // method 00 + partition order
                }
            }
            if (run_bits)
            {
                RunWriter rw;
                rw.begin(bitbuf, first_bit);
                if (fastp)
                {
                    const uint32_t kmask = (1u << k_run) - 1u, stop = 1u << k_run, nb_code = k_run + 1u;
                    const int first_valid = i0 == 0 ? order : 0;
#pragma unroll
                    for (int j = 0; j < kRunLen; ++j)
                    {
                        if (j == pj)
                            rw.append(k_run, 4u); // Rice parameter of the partition (src/flac.rs:676)
                        if (j >= first_valid)
                            rw.append_rice(zz[j] >> k_run, stop | (zz[j] & kmask), nb_code); // unary zeros, stop bit, k low bits
                    }
                }
                else if (order == 0)
                    for (uint32_t i = i0; i < min(i0 + kRunLen, g.bs); ++i)
                        rw.append((uint32_t)(uint16_t)s[i], 16u);
                else
                    for (uint32_t i = max(i0, (uint32_t)order); i < min(i0 + kRunLen, g.bs); ++i)
                    {
                        const uint32_t part = cp.part_of(i);
                        const uint32_t k = (uint32_t)kin[part];
                        const uint32_t pstart = part == 0 ? (uint32_t)order : part * cp.dps;
                        if (i == pstart)
                            rw.append(k, 4u);
                        const uint32_t z = zigzag(residual_at(s, i, order));
                        rw.append_rice(z >> k, (1u << k) | (z & ((1u << k) - 1u)), k + 1u);
                    }
                rw.finish();
            }
            sub0 += fixed + chan_bits;
        }
        __syncthreads();

        // ---- CRC-16 over bytes [0, nb): the frame is right-aligned in a virtual buffer of 256 chunks of
        //      2^m bytes (leading zero bytes do not change a CRC with init 0), every thread hashes its
        //      chunk with the table, then chunks are merged pairwise: crc(A|B) = crc(A) x^(8|B|) + crc(B),
        //      with x^(8 * 2^j) mod P tabulated ----
        const uint32_t nb = fbytes - 2;
        {
            // one chunk size for the whole launch (from the largest frame): smaller frames simply have more
            // leading zero padding, and the merge tables are built once per CTA
            uint32_t m = 0;
            while (((uint32_t)kFlacThreads << m) < (buf_words - 4u) * 4u)
                ++m;
            if ((int)m != cur_m) // first block of this CTA
            {
                cur_m = (int)m;
                if (tid < 6)
                {
                    uint32_t bm = s_xpow[m + tid]; // x^(8 * 2^(m + L)); times x^i for i = 0..15
                    for (int i = 0; i < 16; ++i)
                    {
                        s_bitmul[tid][i] = (uint16_t)bm;
                        bm = ((bm << 1) ^ ((bm & 0x8000u) ? 0x18005u : 0u)) & 0xffffu;
                    }
                }
                __syncthreads();
                for (uint32_t e = tid; e < 6u * 512u; e += kFlacThreads)
                {
                    const uint32_t L = e >> 9, hi = (e >> 8) & 1u, x = e & 255u;
                    uint32_t acc = 0;
#pragma unroll
                    for (int jb = 0; jb < 8; ++jb)
                        acc ^= ((x >> jb) & 1u) ? (uint32_t)s_bitmul[L][jb + 8 * hi] : 0u;
                    s_mul[L][hi][x] = (uint16_t)acc;
                }
                __syncthreads();
            }
            const uint32_t chunk = 1u << m;
            const uint32_t pad = kFlacThreads * chunk - nb;
            const uint32_t v0 = (uint32_t)tid * chunk, v1 = v0 + chunk;
            const uint32_t b0 = v0 > pad ? v0 - pad : 0u, b1 = v1 > pad ? v1 - pad : 0u;
            uint32_t crc = 0;
            uint32_t j = b0;
            auto step = [&](uint32_t byte) { crc = ((crc << 8) & 0xffffu) ^ s_crc_tab[0][((crc >> 8) ^ byte) & 0xffu]; };
            for (; j < b1 && (j & 3u); ++j)
                step(buf_byte(bitbuf, j));
            for (; j + 4 <= b1; j += 4)
            {
                // four bytes at once: the running CRC is folded into the first two, every byte looks up the table
                // for its distance from the end of the word
                const uint32_t w = bitbuf[j >> 2] ^ (crc << 16);
                crc = (uint32_t)s_crc_tab[3][w >> 24] ^ s_crc_tab[2][(w >> 16) & 0xffu] ^ s_crc_tab[1][(w >> 8) & 0xffu] ^
                      s_crc_tab[0][w & 0xffu];
            }
            for (; j < b1; ++j)
                step(buf_byte(bitbuf, j));
#pragma unroll
            for (int l = 0; l < 5; ++l)
            {
                const uint32_t other = __shfl_down_sync(0xffffffffu, crc, 1 << l);
                crc = (uint32_t)s_mul[l][0][crc & 0xffu] ^ s_mul[l][1][(crc >> 8) & 0xffu] ^ other;
            }
            if (lane == 0)
                s_crc_part[warp] = crc;
            __syncthreads();
            if (tid == 0)
            {
                uint32_t acc = 0;
                for (int w = 0; w < kWarps; ++w)
                    acc = ((uint32_t)s_mul[5][0][acc & 0xffu] ^ s_mul[5][1][(acc >> 8) & 0xffu]) ^ s_crc_part[w];
                put_bits(bitbuf, (unsigned long long)nb * 8ull, acc & 0xffffu, 16);
            }
        }
        __syncthreads();
        // ---- copy out: bytes up to the first 4-byte boundary of the destination, then whole words (two
        //      big-endian buffer words funnel-shifted and byte-swapped per store), then the tail ----
        uint8_t *dst = out_arena + frame_off[b];
        const uint32_t head = min(fbytes, (uint32_t)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u));
        const uint32_t n_words = (fbytes - head) >> 2;
        if ((uint32_t)tid < head)
            dst[tid] = (uint8_t)buf_byte(bitbuf, tid);
        uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
        for (uint32_t k = tid; k < n_words; k += kFlacThreads)
        {
            const uint32_t j = head + 4u * k; // first source byte of this word
            const uint32_t w0 = bitbuf[j >> 2], w1 = bitbuf[(j >> 2) + 1];
            const uint32_t be = __funnelshift_l(w1, w0, 8u * (j & 3u)); // bytes j..j+3, most significant first
            dw[k] = __byte_perm(be, 0u, 0x0123);
        }
        for (uint32_t j = head + 4u * n_words + tid; j < fbytes; j += kFlacThreads)
            dst[j] = (uint8_t)buf_byte(bitbuf, j);
        __syncthreads();
    }
}

} // namespace

uint32_t flac_slot_bytes(uint32_t block_size, uint32_t channels)
{
    // worst case per sample with 16-bit input: k capped at 14, |r| < 2^19 -> about 79 bits
    return 32u + channels * (block_size * 10u + 64u);
}

cudaError_t launch_flac_measure(const FlacLaunch &p, uint8_t *rice_k, uint32_t max_ch, uint32_t max_bs, cudaStream_t s)
{
    if (p.n_blocks_total == 0)
        return cudaSuccess;
    const size_t smem = (size_t)((max_bs + 7u) & ~7u) * max_ch * sizeof(int16_t) + 16;
    cudaError_t e = cudaFuncSetAttribute(flac_measure_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return e;
    flac_measure_kernel<<<(unsigned)p.n_blocks_total, kFlacThreads, smem, s>>>(p, rice_k, max_ch);
    return cudaGetLastError();
}

namespace
{
constexpr size_t kEmitSmemLimit = 200 * 1024;
struct EmitPlan
{
    uint32_t smp_bytes, buf_words;
    size_t smem;
    unsigned grid;
    bool global_scratch;
};
EmitPlan plan_emit(uint64_t n_blocks, uint32_t max_ch, uint32_t max_bs, uint32_t max_frame_bytes, int sm_count)
{
    EmitPlan pl;
    pl.smp_bytes = (uint32_t)(((size_t)((max_bs + 7u) & ~7u) * max_ch * sizeof(int16_t) + 31) & ~(size_t)15);
    pl.buf_words = ((max_frame_bytes + 3) >> 2) + 4;
    pl.smem = (size_t)pl.smp_bytes + (size_t)pl.buf_words * 4;
    pl.global_scratch = pl.smem > kEmitSmemLimit;
    if (pl.global_scratch)
    {
        // frames too large for shared memory (pathological residuals / many channels): the bit
        // buffer moves to a per-CTA global scratch area, same code path through generic pointers
        pl.smem = pl.smp_bytes;
        pl.grid = (unsigned)std::min<uint64_t>(n_blocks, (uint64_t)sm_count * 2);
    }
    else
    {
        // exactly one resident set of CTAs (each loops over blocks with a grid stride): more would queue behind
        // the first set and run the tail of the launch at partial occupancy
        int per_sm = 0;
        (void)cudaFuncSetAttribute(flac_emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flac_emit_kernel<false>, kFlacThreads, pl.smem) != cudaSuccess ||
            per_sm < 1)
        {
            (void)cudaGetLastError();
            per_sm = 1;
        }
        pl.grid = (unsigned)std::min<uint64_t>(n_blocks, (uint64_t)sm_count * (unsigned)per_sm);
    }
    return pl;
}
} // namespace

size_t flac_emit_scratch_words(uint64_t n_blocks, uint32_t max_ch, uint32_t max_bs, uint32_t max_frame_bytes,
                               int sm_count)
{
    if (n_blocks == 0)
        return 0;
    const EmitPlan pl = plan_emit(n_blocks, max_ch, max_bs, max_frame_bytes, sm_count);
    return pl.global_scratch ? (size_t)pl.grid * pl.buf_words : 0;
}

cudaError_t launch_flac_emit(const FlacLaunch &p, const uint8_t *rice_k, uint32_t max_ch, uint32_t max_bs,
                             uint32_t max_frame_bytes, const uint64_t *frame_off, uint8_t *out_arena,
                             uint32_t *scratch, int sm_count, cudaStream_t s)
{
    if (p.n_blocks_total == 0)
        return cudaSuccess;
    const EmitPlan pl = plan_emit(p.n_blocks_total, max_ch, max_bs, max_frame_bytes, sm_count);
    if (pl.global_scratch && !scratch)
        return cudaErrorInvalidValue;
    if (pl.global_scratch)
    {
        cudaError_t e = cudaFuncSetAttribute(flac_emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess)
            return e;
        flac_emit_kernel<true><<<pl.grid, kFlacThreads, pl.smem, s>>>(p, rice_k, max_ch, frame_off, out_arena, pl.smp_bytes,
                                                                      pl.buf_words, scratch);
    }
    else
    {
        cudaError_t e = cudaFuncSetAttribute(flac_emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
        if (e != cudaSuccess)
            return e;
        flac_emit_kernel<false><<<pl.grid, kFlacThreads, pl.smem, s>>>(p, rice_k, max_ch, frame_off, out_arena, pl.smp_bytes,
                                                                       pl.buf_words, nullptr);
    }
    return cudaGetLastError();
}

} // namespace glc
