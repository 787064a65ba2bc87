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
    // RFC 1321 compression function, fully unrolled (the chain a->d->c->b is the serial critical path
    // of the whole FLAC encode of one file: about 5 cycles per step)
    void block(const uint8_t *p)
    {
        uint32_t x[16];
        memcpy(x, p, 64); // little-endian host
        uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#define MD5_F(b, c, d) ((d) ^ ((b) & ((c) ^ (d))))
#define MD5_G(b, c, d) ((c) ^ ((d) & ((b) ^ (c))))
#define MD5_H(b, c, d) ((b) ^ (c) ^ (d))
#define MD5_I(b, c, d) ((c) ^ ((b) | ~(d)))
#define MD5_STEP(f, a, b, c, d, xk, s, k) \
    a += f(b, c, d) + (xk) + (k);          \
    a = rol(a, s) + b;
        MD5_STEP(MD5_F, a, b, c, d, x[0], 7, 0xd76aa478u)
        MD5_STEP(MD5_F, d, a, b, c, x[1], 12, 0xe8c7b756u)
        MD5_STEP(MD5_F, c, d, a, b, x[2], 17, 0x242070dbu)
        MD5_STEP(MD5_F, b, c, d, a, x[3], 22, 0xc1bdceeeu)
        MD5_STEP(MD5_F, a, b, c, d, x[4], 7, 0xf57c0fafu)
        MD5_STEP(MD5_F, d, a, b, c, x[5], 12, 0x4787c62au)
        MD5_STEP(MD5_F, c, d, a, b, x[6], 17, 0xa8304613u)
        MD5_STEP(MD5_F, b, c, d, a, x[7], 22, 0xfd469501u)
        MD5_STEP(MD5_F, a, b, c, d, x[8], 7, 0x698098d8u)
        MD5_STEP(MD5_F, d, a, b, c, x[9], 12, 0x8b44f7afu)
        MD5_STEP(MD5_F, c, d, a, b, x[10], 17, 0xffff5bb1u)
        MD5_STEP(MD5_F, b, c, d, a, x[11], 22, 0x895cd7beu)
        MD5_STEP(MD5_F, a, b, c, d, x[12], 7, 0x6b901122u)
        MD5_STEP(MD5_F, d, a, b, c, x[13], 12, 0xfd987193u)
        MD5_STEP(MD5_F, c, d, a, b, x[14], 17, 0xa679438eu)
        MD5_STEP(MD5_F, b, c, d, a, x[15], 22, 0x49b40821u)
        MD5_STEP(MD5_G, a, b, c, d, x[1], 5, 0xf61e2562u)
        MD5_STEP(MD5_G, d, a, b, c, x[6], 9, 0xc040b340u)
        MD5_STEP(MD5_G, c, d, a, b, x[11], 14, 0x265e5a51u)
        MD5_STEP(MD5_G, b, c, d, a, x[0], 20, 0xe9b6c7aau)
        MD5_STEP(MD5_G, a, b, c, d, x[5], 5, 0xd62f105du)
        MD5_STEP(MD5_G, d, a, b, c, x[10], 9, 0x02441453u)
        MD5_STEP(MD5_G, c, d, a, b, x[15], 14, 0xd8a1e681u)
        MD5_STEP(MD5_G, b, c, d, a, x[4], 20, 0xe7d3fbc8u)
        MD5_STEP(MD5_G, a, b, c, d, x[9], 5, 0x21e1cde6u)
        MD5_STEP(MD5_G, d, a, b, c, x[14], 9, 0xc33707d6u)
        MD5_STEP(MD5_G, c, d, a, b, x[3], 14, 0xf4d50d87u)
        MD5_STEP(MD5_G, b, c, d, a, x[8], 20, 0x455a14edu)
        MD5_STEP(MD5_G, a, b, c, d, x[13], 5, 0xa9e3e905u)
        MD5_STEP(MD5_G, d, a, b, c, x[2], 9, 0xfcefa3f8u)
        MD5_STEP(MD5_G, c, d, a, b, x[7], 14, 0x676f02d9u)
        MD5_STEP(MD5_G, b, c, d, a, x[12], 20, 0x8d2a4c8au)
        MD5_STEP(MD5_H, a, b, c, d, x[5], 4, 0xfffa3942u)
        MD5_STEP(MD5_H, d, a, b, c, x[8], 11, 0x8771f681u)
        MD5_STEP(MD5_H, c, d, a, b, x[11], 16, 0x6d9d6122u)
        MD5_STEP(MD5_H, b, c, d, a, x[14], 23, 0xfde5380cu)
        MD5_STEP(MD5_H, a, b, c, d, x[1], 4, 0xa4beea44u)
        MD5_STEP(MD5_H, d, a, b, c, x[4], 11, 0x4bdecfa9u)
        MD5_STEP(MD5_H, c, d, a, b, x[7], 16, 0xf6bb4b60u)
        MD5_STEP(MD5_H, b, c, d, a, x[10], 23, 0xbebfbc70u)
        MD5_STEP(MD5_H, a, b, c, d, x[13], 4, 0x289b7ec6u)
        MD5_STEP(MD5_H, d, a, b, c, x[0], 11, 0xeaa127fau)
        MD5_STEP(MD5_H, c, d, a, b, x[3], 16, 0xd4ef3085u)
        MD5_STEP(MD5_H, b, c, d, a, x[6], 23, 0x04881d05u)
        MD5_STEP(MD5_H, a, b, c, d, x[9], 4, 0xd9d4d039u)
        MD5_STEP(MD5_H, d, a, b, c, x[12], 11, 0xe6db99e5u)
        MD5_STEP(MD5_H, c, d, a, b, x[15], 16, 0x1fa27cf8u)
        MD5_STEP(MD5_H, b, c, d, a, x[2], 23, 0xc4ac5665u)
        MD5_STEP(MD5_I, a, b, c, d, x[0], 6, 0xf4292244u)
        MD5_STEP(MD5_I, d, a, b, c, x[7], 10, 0x432aff97u)
        MD5_STEP(MD5_I, c, d, a, b, x[14], 15, 0xab9423a7u)
        MD5_STEP(MD5_I, b, c, d, a, x[5], 21, 0xfc93a039u)
        MD5_STEP(MD5_I, a, b, c, d, x[12], 6, 0x655b59c3u)
        MD5_STEP(MD5_I, d, a, b, c, x[3], 10, 0x8f0ccc92u)
        MD5_STEP(MD5_I, c, d, a, b, x[10], 15, 0xffeff47du)
        MD5_STEP(MD5_I, b, c, d, a, x[1], 21, 0x85845dd1u)
        MD5_STEP(MD5_I, a, b, c, d, x[8], 6, 0x6fa87e4fu)
        MD5_STEP(MD5_I, d, a, b, c, x[15], 10, 0xfe2ce6e0u)
        MD5_STEP(MD5_I, c, d, a, b, x[6], 15, 0xa3014314u)
        MD5_STEP(MD5_I, b, c, d, a, x[13], 21, 0x4e0811a1u)
        MD5_STEP(MD5_I, a, b, c, d, x[4], 6, 0xf7537e82u)
        MD5_STEP(MD5_I, d, a, b, c, x[11], 10, 0xbd3af235u)
        MD5_STEP(MD5_I, c, d, a, b, x[2], 15, 0x2ad7d2bbu)
        MD5_STEP(MD5_I, b, c, d, a, x[9], 21, 0xeb86d391u)
#undef MD5_STEP
#undef MD5_F
#undef MD5_G
#undef MD5_H
#undef MD5_I
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

// pcm == nullptr: the samples are already on the device, file i at d_base + d_off[i] (16-byte aligned).
glc_status glc::flac_encode_impl(glc_ctx *ctx, uint32_t n_files, const float *const *pcm, const float *d_base,
                                 const uint64_t *d_off, const uint64_t *n_samples, const uint32_t *sample_rate,
                                 const uint16_t *channels, uint8_t level, uint8_t **bytes, uint64_t *len)
{
    if (!ctx || (!pcm && !(d_base && d_off)) || !n_samples || !sample_rate || !channels || !bytes || !len || n_files == 0)
        return set_error(GLC_ERR_INVALID_ARG, "null/empty argument");
    // ---- checks, in the reference's order: length first (:963), then level (:972) ----
    for (uint32_t i = 0; i < n_files; ++i)
    {
        if (pcm && !pcm[i] && n_samples[i])
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
    // (cudaGetDeviceProperties costs milliseconds per call; one attribute is all that is needed)
    int sm_count = 148;
    FL_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx_device(ctx)));

    std::vector<FlacFileDesc> files(n_files);
    uint64_t tot_pcm = 0, tot_blocks = 0;
    uint32_t max_ch = 1, max_bs = 16;
    for (uint32_t i = 0; i < n_files; ++i)
    {
        FlacFileDesc &f = files[i];
        const uint64_t total = n_samples[i] / channels[i];
        uint64_t bs = level <= 2 ? 1152 : 4096; // src/flac.rs:983-995
        bs = std::max<uint64_t>(std::min<uint64_t>(bs, total), 16);
        f.pcm_off = pcm ? tot_pcm : d_off[i];
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

    // ---- MD5 (src/flac.rs:1004 -> :305-318) is a serial chain per file.  The f32 -> i16 conversion is
    //      already done by pass A on the device, so the converted samples come back in 32 MB chunks on
    //      the d2h stream while pass B runs, and host threads (one per file) hash each chunk as soon
    //      as its event fires: the hash is the only thing left on the host's critical path. ----
    std::vector<std::vector<uint8_t>> md5(n_files, std::vector<uint8_t>(16));
    std::vector<std::thread> workers;
    struct CopyOp
    {
        uint64_t off, cnt; // samples within the file
        uint32_t event;    // index of the event that covers this chunk
    };
    std::vector<std::vector<CopyOp>> ops(n_files);
    std::vector<cudaEvent_t> chunk_events;
    const uint64_t kInlineMd5Samples = 4ull << 20; // 8 MB of 16-bit samples: about 12 ms of hashing
    bool inline_md5 = false;
    int16_t *h_i16 = nullptr;
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
        if (pcm)
            FL_STEP(dev_alloc(ctx, (void **)&d_pcm, std::max<uint64_t>(tot_pcm, 1) * 4, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_i16, std::max<uint64_t>(tot_pcm, 1) * 2, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_files, sizeof(FlacFileDesc) * n_files, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_k, std::max<uint64_t>(tot_blocks, 1) * max_ch * 64, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_fbytes, std::max<uint64_t>(tot_blocks, 1) * 4, cs));
        FL_STEP(dev_alloc(ctx, (void **)&d_foff, (tot_blocks + 1) * 8, cs));
        FL_STEP(cudaMemcpyAsync(d_files, files.data(), sizeof(FlacFileDesc) * n_files, cudaMemcpyHostToDevice, cs));
        bool bad = false;
        for (uint32_t i = 0; pcm && i < n_files && !bad; ++i)
            if (ctx_h2d(ctx, d_pcm + files[i].pcm_off, pcm[i], n_samples[i] * 4, cs) !=
                cudaSuccess)
                bad = true;
        if (bad)
        {
            st = set_error(GLC_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (pcm)
            ctx_count_bytes(ctx, tot_pcm * 4, 0);

        FlacLaunch fl{};
        fl.pcm_arena = pcm ? d_pcm : d_base;
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

        {
            // converted samples -> pinned host memory, chunk by chunk, behind pass A
            const uint64_t kChunk = 16ull << 20; // samples (32 MB)
            h_i16 = (int16_t *)pinned_alloc(ctx, std::max<uint64_t>(tot_pcm, 1) * 2);
            if (!h_i16)
            {
                st = set_error(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
                break;
            }
            cudaStream_t ds = ctx_d2h_stream(ctx);
            cudaEvent_t ev_a;
            FL_STEP(cudaEventCreateWithFlags(&ev_a, cudaEventDisableTiming));
            chunk_events.push_back(ev_a);
            FL_STEP(cudaEventRecord(ev_a, cs));
            FL_STEP(cudaStreamWaitEvent(ds, ev_a, 0));
            uint64_t pending = 0;
            bool bad2 = false;
            for (uint32_t i = 0; i < n_files && !bad2; ++i)
                for (uint64_t off = 0; off < n_samples[i] && !bad2; off += kChunk)
                {
                    const uint64_t cnt = std::min<uint64_t>(kChunk, n_samples[i] - off);
                    if (cudaMemcpyAsync(h_i16 + files[i].i16_off + off, d_i16 + files[i].i16_off + off, cnt * 2,
                                        cudaMemcpyDeviceToHost, ds) != cudaSuccess)
                        bad2 = true;
                    pending += cnt;
                    ops[i].push_back({off, cnt, (uint32_t)chunk_events.size()});
                    const bool last = (i + 1 == n_files) && (off + cnt >= n_samples[i]);
                    if (pending >= kChunk || last)
                    {
                        cudaEvent_t ev;
                        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess ||
                            cudaEventRecord(ev, ds) != cudaSuccess)
                            bad2 = true;
                        chunk_events.push_back(ev);
                        pending = 0;
                    }
                }
            if (bad2)
            {
                st = set_error(GLC_ERR_CUDA, "D2H of converted samples failed: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
            ctx_count_bytes(ctx, 0, tot_pcm * 2);
            // Small calls hash on the calling thread once everything else is queued (a new thread costs
            // milliseconds: creation plus binding the CUDA context, against microseconds of kernels).
            inline_md5 = tot_pcm <= kInlineMd5Samples;
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            const unsigned nthreads = inline_md5 ? 0u : std::min<unsigned>(hw, n_files);
            const int device = ctx_device(ctx);
            for (unsigned t = 0; t < nthreads; ++t)
                workers.emplace_back([&, t, nthreads, device]() {
                    cudaSetDevice(device);
                    for (uint32_t i = t; i < n_files; i += nthreads)
                    {
                        Md5 m;
                        for (const CopyOp &op : ops[i])
                        {
                            cudaEventSynchronize(chunk_events[op.event]);
                            m.update(reinterpret_cast<const uint8_t *>(h_i16 + files[i].i16_off + op.off), op.cnt * 2);
                        }
                        m.finish(md5[i].data());
                    }
                });
        }

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
        if (const size_t sw = flac_emit_scratch_words(tot_blocks, max_ch, max_bs, max_frame, sm_count))
            FL_STEP(dev_alloc(ctx, (void **)&d_scratch, sw * 4, cs));
        ctx_count_launch(ctx, GLC_K_FLAC_GATHER, 1);
        ctx_time_begin(ctx, GLC_K_FLAC_GATHER, &tok);
        FL_STEP(launch_flac_emit(fl, d_k, max_ch, max_bs, max_frame, d_foff, d_out, d_scratch,
                                 sm_count, cs));
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
        if (inline_md5)
            for (uint32_t i = 0; i < n_files; ++i)
            {
                Md5 m;
                for (const CopyOp &op : ops[i])
                {
                    cudaEventSynchronize(chunk_events[op.event]);
                    m.update(reinterpret_cast<const uint8_t *>(h_i16 + files[i].i16_off + op.off), op.cnt * 2);
                }
                m.finish(md5[i].data());
            }
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
    if (h_i16)
        pinned_release(ctx, h_i16);
    for (cudaEvent_t e : chunk_events)
        cudaEventDestroy(e);
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

extern "C" glc_status glc_flac_encode_batch(glc_ctx *ctx, uint32_t n_files, const float *const *pcm,
                                            const uint64_t *n_samples, const uint32_t *sample_rate,
                                            const uint16_t *channels, uint8_t level, uint8_t **bytes, uint64_t *len)
{
    if (!pcm)
        return set_error(GLC_ERR_INVALID_ARG, "null/empty argument");
    return flac_encode_impl(ctx, n_files, pcm, nullptr, nullptr, n_samples, sample_rate, channels, level, bytes, len);
}

extern "C" glc_status glc_flac_encode(glc_ctx *ctx, const float *pcm, uint64_t n_samples, uint32_t sample_rate,
                                      uint16_t channels, uint8_t level, uint8_t **bytes, uint64_t *len)
{
    if (!bytes || !len)
        return set_error(GLC_ERR_INVALID_ARG, "null argument");
    const float *files[1] = {pcm};
    return glc_flac_encode_batch(ctx, 1, files, &n_samples, &sample_rate, &channels, level, bytes, len);
}
