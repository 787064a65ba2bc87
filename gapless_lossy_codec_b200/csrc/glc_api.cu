// glc_api.cu -- host side of libglc_b200.so: contexts, memory pools, batching, the C ABI of
// include/glc.h.  Raw CUDA runtime calls only (no PyTorch, no CPU compute fallback: when there is
// no device every entry point fails with GLC_ERR_NO_DEVICE).
//
// Host-side structure of one encode (reference Encoder::encode, src/codec.rs:421-565):
//   files -> FileDesc table (row/frame numbering across the whole batch)
//   waves of frames:  H2D of the wave's PCM on the copy stream  ||  mdct_exact + quant_pack of the
//                     previous wave on the compute stream (events order them)
//   scan (nnz, raw lengths) -> gather into the compact stream -> D2H into pooled pinned memory.
// One decode (Decoder::decode, src/codec.rs:595-768):
//   H2D stream -> dequant (+ per-tile k-chunk masks) -> imdct_exact -> overlap-add/interleave -> D2H,
//   the gapless trim being an offset/length applied to the D2H copy.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "glc_internal.cuh"

using namespace glc;

// ------------------------------------------------------------------ errors

static thread_local std::string g_last_error;

static glc_status fail(glc_status st, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return st;
}

#define CUDA_TRY(expr)                                                                                   \
    do                                                                                                   \
    {                                                                                                    \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(GLC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

#define GLC_TRY(expr)              \
    do                             \
    {                              \
        glc_status _s = (expr);    \
        if (_s != GLC_OK)          \
            return _s;             \
    } while (0)

extern "C" const char *glc_last_error(void) { return g_last_error.c_str(); }
extern "C" uint32_t glc_abi_version(void) { return GLC_ABI_VERSION; }

extern "C" glc_status glc_device_count(int *count)
{
    if (!count)
        return fail(GLC_ERR_INVALID_ARG, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
    {
        *count = 0;
        (void)cudaGetLastError();
        return fail(GLC_ERR_NO_DEVICE, "no CUDA device available (%s); libglc_b200 has no CPU fallback",
                    e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    }
    *count = n;
    return GLC_OK;
}

// Size classes of the two pools: large requests are rounded up to 1/8 of their power of two (at most
// 12.5 % slack), so that batches of similar but not identical size reuse the same blocks instead of
// going back to the driver (a pinned allocation costs about 0.2 s per GB).
static size_t pool_size_class(size_t bytes)
{
    if (bytes <= ((size_t)1 << 20))
        return (bytes + 4095) & ~(size_t)4095;
    int lg = 63 - __builtin_clzll((unsigned long long)bytes);
    const size_t step = (size_t)1 << (lg - 3);
    return (bytes + step - 1) / step * step;
}

// ----------------------------------------------------------- pinned pool

struct PinnedBlock
{
    void *p;
    size_t cap;
    bool used;
};

struct PinnedPool
{
    std::vector<PinnedBlock> blocks;
    std::mutex mu;
    uint64_t n_grow = 0, grow_bytes = 0; // cudaHostAlloc calls (statistics)

    void *alloc(size_t bytes)
    {
        if (bytes == 0)
            bytes = 16;
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        for (size_t i = 0; i < blocks.size(); ++i)
            if (!blocks[i].used && blocks[i].cap >= bytes && blocks[i].cap <= bytes * 2 + (1u << 20))
                if (best < 0 || blocks[i].cap < blocks[best].cap)
                    best = (int)i;
        if (best >= 0)
        {
            blocks[best].used = true;
            return blocks[best].p;
        }
        // Growth is expensive (cudaHostAlloc: about 0.2 s per GB, and serialised between the processes of a box),
        // and the batches of a workload differ by a few per cent: large blocks get 25 % headroom so that the
        // next, slightly larger request still fits (a block serves requests down to half its size)
        size_t cap = pool_size_class(bytes >= ((size_t)16 << 20) ? bytes + bytes / 4 : bytes);
        void *p = nullptr;
        cudaError_t e = cudaHostAlloc(&p, cap, cudaHostAllocDefault);
        if (e != cudaSuccess && cap > pool_size_class(bytes))
        {
            (void)cudaGetLastError();
            cap = pool_size_class(bytes); // no room for the headroom: the exact class
            e = cudaHostAlloc(&p, cap, cudaHostAllocDefault);
        }
        if (e != cudaSuccess)
        {
            (void)cudaGetLastError();
            return nullptr;
        }
        blocks.push_back({p, cap, true});
        ++n_grow;
        grow_bytes += cap;
        return p;
    }
    bool release(void *p)
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &b : blocks)
            if (b.p == p)
            {
                b.used = false;
                return true;
            }
        return false;
    }
    bool owns(const void *p) // p points into one of the pool's blocks
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto &b : blocks)
            if ((const char *)p >= (const char *)b.p && (const char *)p < (const char *)b.p + b.cap)
                return true;
        return false;
    }
    void destroy()
    {
        for (auto &b : blocks)
            cudaFreeHost(b.p);
        blocks.clear();
    }
};

// ----------------------------------------------------------- device pool
//
// Grow-only cache of cudaMalloc blocks owned by the context.  The hot path allocates and frees
// multi-GB scratch and output buffers on every call; the driver's stream-ordered allocator
// (cudaMallocAsync / cudaFreeAsync) put tens of milliseconds of idle time on the GPU timeline of an
// hour-long batch (measured), so steady-state calls never enter the driver:
// a freed block goes back to this list and the next request of a similar size takes it.
// Reuse is safe without events because every consumer is stream-ordered: a block last used on
// stream A is handed to stream A again without a wait, and to another stream only after A drained.
struct DevBlock
{
    void *p;
    size_t cap;
    bool used;
    cudaStream_t last;
};

struct DevicePool
{
    std::vector<DevBlock> blocks;
    std::mutex mu;
    size_t total = 0;
    uint64_t n_grow = 0, grow_bytes = 0; // cudaMalloc calls (statistics)

    cudaError_t alloc(void **out, size_t bytes, cudaStream_t s)
    {
        *out = nullptr;
        if (bytes == 0)
            bytes = 256;
        std::lock_guard<std::mutex> lk(mu);
        int best = -1;
        for (int pass = 0; pass < 2 && best < 0; ++pass)
            for (size_t i = 0; i < blocks.size(); ++i)
            {
                const DevBlock &b = blocks[i];
                if (b.used || b.cap < bytes || b.cap > bytes + bytes / 2 + ((size_t)4 << 20))
                    continue;
                if (pass == 0 && b.last != s)
                    continue; // prefer a block whose previous user ran on the same stream
                if (best < 0 || b.cap < blocks[best].cap)
                    best = (int)i;
            }
        if (best >= 0)
        {
            DevBlock &b = blocks[best];
            if (b.last != s && b.last != nullptr)
                cudaStreamSynchronize(b.last);
            b.used = true;
            b.last = s;
            *out = b.p;
            return cudaSuccess;
        }
        // large blocks get 12.5 % headroom: batches of a workload differ by a few per cent (a block serves
        // requests down to two thirds of its size, see above)
        size_t cap = bytes >= ((size_t)16 << 20)  ? pool_size_class(bytes + bytes / 8)
                     : bytes >= ((size_t)1 << 20) ? pool_size_class(bytes)
                                                  : ((bytes + 511) & ~(size_t)511);
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, cap);
        if (e != cudaSuccess && cap > pool_size_class(bytes))
        {
            (void)cudaGetLastError();
            cap = pool_size_class(bytes); // no room for the headroom: the exact class
            e = cudaMalloc(&p, cap);
        }
        if (e != cudaSuccess)
        {
            // out of memory: give every idle block back to the driver and retry once
            (void)cudaGetLastError();
            cudaDeviceSynchronize();
            for (size_t i = 0; i < blocks.size();)
                if (!blocks[i].used)
                {
                    cudaFree(blocks[i].p);
                    total -= blocks[i].cap;
                    blocks.erase(blocks.begin() + (long)i);
                }
                else
                    ++i;
            e = cudaMalloc(&p, cap);
            if (e != cudaSuccess)
                return e;
        }
        blocks.push_back({p, cap, true, s});
        total += cap;
        ++n_grow;
        grow_bytes += cap;
        *out = p;
        return cudaSuccess;
    }
    void release(void *p, cudaStream_t s)
    {
        if (!p)
            return;
        std::lock_guard<std::mutex> lk(mu);
        for (auto &b : blocks)
            if (b.p == p)
            {
                b.used = false;
                b.last = s;
                return;
            }
    }
    void destroy()
    {
        for (auto &b : blocks)
            cudaFree(b.p);
        blocks.clear();
        total = 0;
    }
};

// ------------------------------------------------ pageable host memory -> device
//
// The reference's entry points take borrowed slices (`Encoder::encode(&mut self, samples: &[f32], ..)`,
// src/codec.rs:421; `Decoder::decode(&EncodedAudio)`, :744): ordinary pageable memory.  cudaMemcpyAsync from
// pageable memory is staged by the driver through one internal buffer, synchronously, at a fraction of the
// PCIe rate.  Here a few host threads copy the caller's bytes into a ring of pinned chunks while the DMA
// engine drains the chunks filled before: the transfer then runs at min(host memcpy rate, PCIe rate) and
// overlaps the kernels of the previous wave like a pinned transfer does.
class ParallelMemcpy
{
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    const char *src = nullptr;
    char *dst = nullptr;
    size_t bytes = 0;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;

    static void part(const char *s, char *d, size_t n, int i, int k)
    {
        const size_t a = (n * (size_t)i / (size_t)k) & ~(size_t)63, b = i + 1 == k ? n : ((n * (size_t)(i + 1) / (size_t)k) & ~(size_t)63);
        if (b > a)
            memcpy(d + a, s + a, b - a);
    }
    void run(int idx)
    {
        uint64_t seen = 0;
        for (;;)
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stop || generation != seen; });
            if (stop)
                return;
            seen = generation;
            const char *s = src;
            char *d = dst;
            const size_t n = bytes;
            const int k = (int)workers.size() + 1;
            lk.unlock();
            part(s, d, n, idx + 1, k);
            lk.lock();
            if (--pending == 0)
                cv_done.notify_one();
        }
    }

  public:
    void start(int n_workers)
    {
        for (int i = 0; i < n_workers; ++i)
            workers.emplace_back([this, i] { run(i); });
    }
    // copies on the calling thread plus the workers; returns when every byte is in place
    void copy(void *d, const void *s, size_t n)
    {
        if (workers.empty() || n < ((size_t)1 << 20))
        {
            memcpy(d, s, n);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            src = (const char *)s;
            dst = (char *)d;
            bytes = n;
            pending = (int)workers.size();
            ++generation;
        }
        cv.notify_all();
        part((const char *)s, (char *)d, n, 0, (int)workers.size() + 1);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~ParallelMemcpy()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : workers)
            t.join();
    }
};

struct HostStager
{
    static constexpr size_t kChunk = (size_t)8 << 20; // bytes per pinned chunk
    static constexpr int kSlots = 4;
    void *slot[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done[kSlots];
    bool busy[kSlots] = {false, false, false, false};
    int next = 0;
    bool ready = false;
    ParallelMemcpy pm;
};

// ------------------------------------------------------------------ context

struct TimedLaunch
{
    int id;
    cudaEvent_t a, b;
};

struct glc_ctx
{
    int device;
    glc_mode mode;
    cudaStream_t compute, copy, d2h;
    cudaEvent_t t0, t1;
    HostTables host;
    float *d_tab_mdct, *d_tab_imdct, *d_window, *d_fast_tw;
    PinnedPool pool;
    DevicePool dpool;
    // Batched calls return many buffers carved out of one pinned slab: pointer -> slab base, and the
    // number of live pointers per slab.  glc_free consults these before the pool.
    double hwm_pairs_per_row = 0.0, hwm_raw_per_row = 0.0; // densest encodes seen: sizes the host arenas
    std::unordered_map<void *, void *> slab_of;
    std::unordered_map<void *, size_t> slab_refs;
    std::mutex slab_mu;
    glc_stats stats;
    bool timing;
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> ev_free;
    int gemm_variant;
    uint64_t wave_frames;
    float *d_flush;
    size_t flush_floats;
    HostStager *stager = nullptr; // created on the first transfer from pageable memory
    uint64_t staged_bytes = 0;    // bytes that went through the stager (statistics)
    uint64_t staged_base = 0, pool_base[4] = {0, 0, 0, 0}; // counters at the last glc_stats_reset
};

struct glc_encoder
{
    glc_ctx *ctx;
    uint32_t sample_rate;
    DevPerceptual *d_perc;
};

struct glc_decoder
{
    glc_ctx *ctx;
    uint32_t channels, sample_rate;
};

static cudaEvent_t get_event(glc_ctx *c)
{
    if (!c->ev_free.empty())
    {
        cudaEvent_t e = c->ev_free.back();
        c->ev_free.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

// Events of one call: they go back to the context's free list on every exit path (an error return used to
// drop them).
struct EventScope
{
    glc_ctx *c;
    std::vector<cudaEvent_t> held;
    explicit EventScope(glc_ctx *ctx) : c(ctx) {}
    EventScope(const EventScope &) = delete;
    EventScope &operator=(const EventScope &) = delete;
    cudaEvent_t get()
    {
        cudaEvent_t e = get_event(c);
        held.push_back(e);
        return e;
    }
    ~EventScope()
    {
        for (cudaEvent_t e : held)
            c->ev_free.push_back(e);
    }
};

// Brackets one kernel launch for the per-kernel statistics.
struct LaunchScope
{
    glc_ctx *c;
    int id;
    cudaStream_t s;
    cudaEvent_t a;
    LaunchScope(glc_ctx *ctx, int kid, cudaStream_t st, uint64_t n_launches = 1) : c(ctx), id(kid), s(st), a(nullptr)
    {
        c->stats.launches[id] += n_launches;
        if (c->timing)
        {
            a = get_event(c);
            cudaEventRecord(a, s);
        }
    }
    ~LaunchScope()
    {
        if (c->timing)
        {
            cudaEvent_t b = get_event(c);
            cudaEventRecord(b, s);
            c->timed.push_back({id, a, b});
        }
    }
};

static void drain_timed(glc_ctx *c)
{
    for (auto &t : c->timed)
    {
        cudaEventSynchronize(t.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess)
            c->stats.kernel_ms[t.id] += ms;
        c->ev_free.push_back(t.a);
        c->ev_free.push_back(t.b);
    }
    c->timed.clear();
}

extern "C" glc_status glc_ctx_create(int device, glc_mode mode, glc_ctx **out)
{
    if (!out)
        return fail(GLC_ERR_INVALID_ARG, "out is null");
    *out = nullptr;
    int n = 0;
    GLC_TRY(glc_device_count(&n));
    if (device < 0 || device >= n)
        return fail(GLC_ERR_INVALID_ARG, "device %d out of range (have %d)", device, n);
    if (mode != GLC_MODE_EXACT && mode != GLC_MODE_FAST)
        return fail(GLC_ERR_INVALID_ARG, "unknown transform mode %d", (int)mode);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(GLC_ERR_NO_DEVICE, "device %d is sm_%d%d; libglc_b200 is built for sm_100a only", device,
                    prop.major, prop.minor);
    glc_ctx *c = new (std::nothrow) glc_ctx();
    if (!c)
        return fail(GLC_ERR_NO_MEMORY, "out of host memory");
    c->device = device;
    c->mode = mode;
    c->timing = false;
    memset(&c->stats, 0, sizeof c->stats);
    c->gemm_variant = 0;
    if (const char *v = getenv("GLC_GEMM_VARIANT"))
        c->gemm_variant = atoi(v);
    c->wave_frames = 0; // 0 = automatic wave sizing (encode_core); otherwise rows per wave
    if (const char *v = getenv("GLC_WAVE_ROWS"))
        c->wave_frames = (uint64_t)atoll(v) > 0 ? (uint64_t)atoll(v) : 0;
    c->d_flush = nullptr;
    c->flush_floats = 0;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->t0));
    CUDA_TRY(cudaEventCreate(&c->t1));
    // tables: host libm -> tiled copies -> device
    build_host_tables(&c->host);
    const size_t tab_bytes = sizeof(float) * (size_t)kHop * kFrame;
    float *tiled = (float *)malloc(tab_bytes);
    if (!tiled)
        return fail(GLC_ERR_NO_MEMORY, "out of host memory");
    CUDA_TRY(cudaMalloc(&c->d_tab_mdct, tab_bytes));
    CUDA_TRY(cudaMalloc(&c->d_tab_imdct, tab_bytes));
    CUDA_TRY(cudaMalloc(&c->d_window, sizeof(float) * kFrame));
    tile_table_for_mdct(c->host.cos_tab, tiled);
    CUDA_TRY(cudaMemcpy(c->d_tab_mdct, tiled, tab_bytes, cudaMemcpyHostToDevice));
    tile_table_for_imdct(c->host.cos_tab, tiled);
    CUDA_TRY(cudaMemcpy(c->d_tab_imdct, tiled, tab_bytes, cudaMemcpyHostToDevice));
    free(tiled);
    CUDA_TRY(cudaMemcpy(c->d_window, c->host.window, sizeof(float) * kFrame, cudaMemcpyHostToDevice));
    {
        float tw[kFastTwiddleFloats];
        fast_twiddle_table(c->host.norm, tw);
        CUDA_TRY(cudaMalloc(&c->d_fast_tw, sizeof tw));
        CUDA_TRY(cudaMemcpy(c->d_fast_tw, tw, sizeof tw, cudaMemcpyHostToDevice));
    }
    *out = c;
    return GLC_OK;
}

extern "C" void glc_ctx_destroy(glc_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    drain_timed(c);
    for (auto e : c->ev_free)
        cudaEventDestroy(e);
    cudaFree(c->d_tab_mdct);
    cudaFree(c->d_tab_imdct);
    cudaFree(c->d_window);
    cudaFree(c->d_fast_tw);
    if (c->d_flush)
        cudaFree(c->d_flush);
    if (c->stager)
    {
        for (int i = 0; i < HostStager::kSlots; ++i)
            if (c->stager->slot[i])
            {
                cudaEventDestroy(c->stager->done[i]);
                cudaFreeHost(c->stager->slot[i]);
            }
        delete c->stager;
    }
    c->pool.destroy();
    c->dpool.destroy();
    free_host_tables(&c->host);
    cudaEventDestroy(c->t0);
    cudaEventDestroy(c->t1);
    cudaStreamDestroy(c->compute);
    cudaStreamDestroy(c->copy);
    cudaStreamDestroy(c->d2h);
    delete c;
}

extern "C" glc_status glc_ctx_set_tuning(glc_ctx *c, int gemm_variant, uint64_t wave_frames)
{
    if (!c)
        return fail(GLC_ERR_INVALID_ARG, "ctx is null");
    (void)gemm_variant; // reserved: the packed f32x2 variants measured no faster and were removed
    c->wave_frames = wave_frames; // rows per encode wave, 0 = automatic
    return GLC_OK;
}

extern "C" glc_status glc_host_alloc(glc_ctx *c, size_t bytes, void **out)
{
    if (!c || !out)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    cudaSetDevice(c->device);
    *out = c->pool.alloc(bytes);
    if (!*out)
        return fail(GLC_ERR_NO_MEMORY, "pinned allocation of %zu bytes failed", bytes);
    return GLC_OK;
}

extern "C" void glc_host_free(glc_ctx *c, void *p)
{
    if (c && p)
        c->pool.release(p);
}

extern "C" void glc_free(glc_ctx *c, void *p)
{
    if (!c || !p)
        return;
    {
        std::lock_guard<std::mutex> lk(c->slab_mu);
        auto it = c->slab_of.find(p);
        if (it != c->slab_of.end())
        {
            void *base = it->second;
            c->slab_of.erase(it);
            if (--c->slab_refs[base] == 0)
            {
                c->slab_refs.erase(base);
                c->pool.release(base);
            }
            return;
        }
    }
    if (!c->pool.release(p))
        free(p);
}

// One pinned slab for `n` buffers of the given byte sizes (each 256-byte aligned); every returned
// pointer is released individually with glc_free.
static bool slab_alloc(glc_ctx *c, uint32_t n, const uint64_t *bytes, void **out)
{
    uint64_t total = 0;
    std::vector<uint64_t> off(n);
    for (uint32_t i = 0; i < n; ++i)
    {
        off[i] = total;
        total += (std::max<uint64_t>(bytes[i], 1) + 255) & ~(uint64_t)255;
    }
    char *base = (char *)c->pool.alloc(total);
    if (!base)
        return false;
    std::lock_guard<std::mutex> lk(c->slab_mu);
    c->slab_refs[base] = n;
    for (uint32_t i = 0; i < n; ++i)
    {
        out[i] = base + off[i];
        c->slab_of[out[i]] = base;
    }
    return true;
}

extern "C" void glc_stats_reset(glc_ctx *c)
{
    if (!c)
        return;
    drain_timed(c);
    memset(&c->stats, 0, sizeof c->stats);
    c->staged_base = c->staged_bytes;
    c->pool_base[0] = c->pool.n_grow;
    c->pool_base[1] = c->pool.grow_bytes;
    c->pool_base[2] = c->dpool.n_grow;
    c->pool_base[3] = c->dpool.grow_bytes;
}

extern "C" void glc_stats_get(glc_ctx *c, glc_stats *out)
{
    if (!c || !out)
        return;
    cudaSetDevice(c->device);
    drain_timed(c);
    *out = c->stats;
    out->staged_bytes = c->staged_bytes - c->staged_base;
    out->pinned_allocs = c->pool.n_grow - c->pool_base[0];
    out->pinned_alloc_bytes = c->pool.grow_bytes - c->pool_base[1];
    out->dev_allocs = c->dpool.n_grow - c->pool_base[2];
    out->dev_alloc_bytes = c->dpool.grow_bytes - c->pool_base[3];
}

extern "C" void glc_stats_enable_kernel_timing(glc_ctx *c, int on)
{
    if (c)
    {
        drain_timed(c);
        c->timing = on != 0;
    }
}

extern "C" glc_status glc_timer_begin(glc_ctx *c)
{
    if (!c)
        return fail(GLC_ERR_INVALID_ARG, "ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->t0, c->compute));
    return GLC_OK;
}

extern "C" glc_status glc_timer_end(glc_ctx *c, float *ms)
{
    if (!c || !ms)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->t1, c->compute));
    CUDA_TRY(cudaEventSynchronize(c->t1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->t0, c->t1));
    return GLC_OK;
}

extern "C" glc_status glc_ctx_sync(glc_ctx *c)
{
    if (!c)
        return fail(GLC_ERR_INVALID_ARG, "ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->copy));
    CUDA_TRY(cudaStreamSynchronize(c->compute));
    CUDA_TRY(cudaStreamSynchronize(c->d2h));
    return GLC_OK;
}

extern "C" glc_status glc_flush_l2(glc_ctx *c)
{
    if (!c)
        return fail(GLC_ERR_INVALID_ARG, "ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->d_flush)
    {
        c->flush_floats = (size_t)64 << 20; // 256 MiB > 126 MB L2
        CUDA_TRY(cudaMalloc(&c->d_flush, c->flush_floats * sizeof(float)));
    }
    LaunchScope ls(c, GLC_K_MISC, c->compute);
    CUDA_TRY(launch_fill(c->d_flush, c->flush_floats, 0.0f, c->compute));
    return GLC_OK;
}

extern "C" glc_status glc_measure_fp32_issue(glc_ctx *c, int packed, double *tera)
{
    if (!c || !tera)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
    float *sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 64));
    const int blocks = prop.multiProcessorCount * 8;
    const int iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep)
    {
        CUDA_TRY(cudaEventRecord(c->t0, c->compute));
        CUDA_TRY(launch_fp32_issue_bench(packed, iters, sink, blocks, c->compute));
        CUDA_TRY(cudaEventRecord(c->t1, c->compute));
        CUDA_TRY(cudaEventSynchronize(c->t1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, c->t0, c->t1));
        // per thread per iteration: 16 chains x (mul + add) [x2 lanes when packed]
        const double ops = (double)blocks * 256.0 * iters * 16.0 * 2.0 * (packed ? 2.0 : 1.0);
        const double t = ops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best)
            best = t;
    }
    cudaFree(sink);
    *tera = best;
    return GLC_OK;
}

// ---- plumbing exported to the other translation units ----
namespace glc
{
glc_status set_error(glc_status st, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return st;
}
void *pinned_alloc(glc_ctx *ctx, size_t bytes) { return ctx->pool.alloc(bytes); }
void pinned_release(glc_ctx *ctx, void *p) { ctx->pool.release(p); }
cudaError_t dev_alloc(glc_ctx *ctx, void **out, size_t bytes, cudaStream_t s) { return ctx->dpool.alloc(out, bytes, s); }
void dev_free(glc_ctx *ctx, void *p, cudaStream_t s) { ctx->dpool.release(p, s); }
int ctx_device(glc_ctx *ctx) { return ctx->device; }
glc_ctx *decoder_ctx(glc_decoder *dec) { return dec->ctx; }
cudaStream_t ctx_compute_stream(glc_ctx *ctx) { return ctx->compute; }
cudaStream_t ctx_d2h_stream(glc_ctx *ctx) { return ctx->d2h; }
void ctx_count_launch(glc_ctx *ctx, int kernel_id, uint64_t n) { ctx->stats.launches[kernel_id] += n; }
cudaError_t ctx_h2d(glc_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t s);
void ctx_count_bytes(glc_ctx *ctx, uint64_t h2d, uint64_t d2h)
{
    ctx->stats.h2d_bytes += h2d;
    ctx->stats.d2h_bytes += d2h;
}
void ctx_time_begin(glc_ctx *ctx, int kernel_id, void **token)
{
    *token = nullptr;
    if (!ctx->timing)
        return;
    TimedLaunch *t = new TimedLaunch();
    t->id = kernel_id;
    t->a = get_event(ctx);
    t->b = nullptr;
    cudaEventRecord(t->a, ctx->compute);
    *token = t;
}
void ctx_time_end(glc_ctx *ctx, void *token)
{
    if (!token)
        return;
    TimedLaunch *t = (TimedLaunch *)token;
    t->b = get_event(ctx);
    cudaEventRecord(t->b, ctx->compute);
    ctx->timed.push_back(*t);
    delete t;
}
} // namespace glc


// true when the DMA engine can read `p` directly: the library's own pinned pool, or memory the caller pinned
static bool host_ptr_is_pinned(glc_ctx *c, const void *p)
{
    if (c->pool.owns(p))
        return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Host -> device on `stream`.  Pinned sources go to the DMA engine as they are; pageable ones above 256 KiB
// pass through the stager's ring (see HostStager).  Returns after the LAST byte has been handed to a pinned
// chunk, i.e. the caller's buffer is no longer needed once the function returns... for pageable sources; for
// pinned ones it must stay valid until the stream reaches the copy, as with cudaMemcpyAsync.
static cudaError_t h2d_async(glc_ctx *c, void *dst, const void *src, size_t bytes, bool pinned, cudaStream_t stream)
{
    if (bytes == 0)
        return cudaSuccess;
    if (pinned || bytes < ((size_t)256 << 10))
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
    if (!c->stager)
        c->stager = new HostStager();
    HostStager &st = *c->stager;
    if (!st.ready)
    {
        for (int i = 0; i < HostStager::kSlots; ++i)
        {
            if (st.slot[i]) // a previous attempt got this far before it ran out of memory
                continue;
            void *p = nullptr;
            cudaError_t e = cudaHostAlloc(&p, HostStager::kChunk, cudaHostAllocDefault);
            if (e != cudaSuccess)
                return e;
            e = cudaEventCreateWithFlags(&st.done[i], cudaEventDisableTiming);
            if (e != cudaSuccess)
            {
                cudaFreeHost(p);
                return e;
            }
            st.slot[i] = p; // slot and event exist together (the context's teardown relies on it)
        }
        int n = 3; // + the calling thread
        if (const char *v = getenv("GLC_COPY_THREADS"))
            n = std::max(0, atoi(v) - 1);
        st.pm.start(n);
        st.ready = true;
    }
    const char *s = (const char *)src;
    char *d = (char *)dst;
    while (bytes)
    {
        const size_t n = std::min(bytes, HostStager::kChunk);
        const int i = st.next;
        st.next = (st.next + 1) % HostStager::kSlots;
        if (st.busy[i])
        {
            cudaError_t e = cudaEventSynchronize(st.done[i]); // the DMA engine is done with this chunk
            if (e != cudaSuccess)
                return e;
        }
        st.pm.copy(st.slot[i], s, n);
        cudaError_t e = cudaMemcpyAsync(d, st.slot[i], n, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess)
            e = cudaEventRecord(st.done[i], stream);
        if (e != cudaSuccess)
            return e;
        st.busy[i] = true;
        c->staged_bytes += n;
        s += n;
        d += n;
        bytes -= n;
    }
    return cudaSuccess;
}

namespace glc
{
cudaError_t ctx_h2d(glc_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t s)
{
    return h2d_async(ctx, dst, src, bytes, host_ptr_is_pinned(ctx, src), s);
}
} // namespace glc

// ------------------------------------------------------------ small helpers

// GLC_TRACE=1: print host-side phase timings (each phase boundary synchronises the stream, so the
// numbers are only for finding overheads, never for benchmarks).
struct PhaseTrace
{
    // GLC_TRACE=1: sync at every phase boundary and print host/GPU-drain times
    // GLC_TRACE=2: host timestamps only
    // GLC_TRACE=3: CUDA events at phase boundaries (no added sync), printed when the scope ends
    int level;
    cudaStream_t s;
    const char *what;
    double t_last;
    std::vector<std::pair<const char *, cudaEvent_t>> evs;
    static double now()
    {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    void ev(const char *name)
    {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        evs.push_back({name, e});
    }
    PhaseTrace(const char *w, cudaStream_t st) : s(st), what(w)
    {
        static const int lv = getenv("GLC_TRACE") ? atoi(getenv("GLC_TRACE")) : 0;
        level = lv;
        if (level == 1)
            cudaStreamSynchronize(s);
        if (level == 3)
            ev("begin");
        if (level)
            t_last = now();
    }
    void mark(const char *phase)
    {
        if (!level)
            return;
        if (level == 3)
        {
            ev(phase);
            return;
        }
        const double t_host = now();
        if (level == 1)
            cudaStreamSynchronize(s);
        const double t = now();
        fprintf(stderr, "[glc trace] %s: %-14s host %.3f ms, +gpu drain %.3f ms\n", what, phase, t_host - t_last,
                t - t_host);
        t_last = t;
    }
    ~PhaseTrace()
    {
        if (level != 3 || evs.empty())
            return;
        cudaEventSynchronize(evs.back().second);
        for (size_t i = 1; i < evs.size(); ++i)
        {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, evs[i - 1].second, evs[i].second);
            fprintf(stderr, "[glc trace] %s: %-14s gpu %.3f ms\n", what, evs[i].first, ms);
        }
        for (auto &e : evs)
            cudaEventDestroy(e.second);
    }
};

#define dmalloc(pp, count, s) dmalloc_impl(c, pp, count, s)
#define dfree(p, s) c->dpool.release(p, s)

template <typename T>
static cudaError_t dmalloc_impl(glc_ctx *c, T **p, size_t count, cudaStream_t s)
{
    *p = nullptr;
    return c->dpool.alloc((void **)p, std::max<size_t>(count, 1) * sizeof(T), s);
}

// Device-pool blocks owned by one call: everything still held is returned to the pool when the scope
// ends, on every exit path (a failing CUDA_TRY included); keep() hands a block to the caller instead.
struct DevScope
{
    glc_ctx *c;
    cudaStream_t s;
    std::vector<void *> held;
    DevScope(glc_ctx *ctx, cudaStream_t st) : c(ctx), s(st) {}
    DevScope(const DevScope &) = delete;
    DevScope &operator=(const DevScope &) = delete;
    template <typename T>
    cudaError_t alloc(T **p, size_t count)
    {
        cudaError_t e = dmalloc_impl(c, p, count, s);
        if (e == cudaSuccess)
            held.push_back((void *)*p);
        return e;
    }
    void keep(void *p)
    {
        for (auto &h : held)
            if (h == p)
                h = nullptr;
    }
    ~DevScope()
    {
        for (void *h : held)
            if (h)
                c->dpool.release(h, s);
    }
};

static uint64_t padded_len(uint64_t L)
{
    uint64_t v = 512 + L; // HOP/2 zeros + data           src/codec.rs:438-439
    const uint64_t rem = v % kHop;
    if (rem)
        v += kHop - rem;  // to a multiple of HOP           :440-444
    return v + 512;       // + HOP/2 zeros                  :445
}

static uint64_t frames_for(uint64_t L)
{
    return (padded_len(L) - kFrame) / kHop + 1; // src/codec.rs:449-455 (L > 512 guaranteed by caller)
}

// Contiguous frame ranges of a batch ("waves"): f0..f1 frames, r0..r1 rows.  Waves are sized in
// multiples of 37 row tiles so that the MDCT grids fill every resident CTA slot (see encode_core).
struct Wave
{
    uint64_t f0, f1, r0, r1;
};

// With group_align, wave boundaries inside a file fall on multiples of that file's frame-group size
// (quant_frames_per_group), so that the per-group kernels never see half a group.
template <typename Desc>
// first_rows (0 = like the others): size of the first wave.  With host buffers the pipeline can only start
// once the first wave has crossed PCIe, so that one is kept short.
static std::vector<Wave> plan_waves(const std::vector<Desc> &files, uint64_t tot_frames, uint64_t tot_rows,
                                    uint64_t target_rows, uint64_t *max_wave_rows, bool group_align = false,
                                    uint64_t first_rows = 0)
{
    const uint32_t n_files = (uint32_t)files.size();
    auto file_of_frame = [&](uint64_t fr) -> uint32_t {
        uint32_t lo = 0, hi = n_files - 1;
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (files[mid].first_frame <= fr)
                lo = mid;
            else
                hi = mid - 1;
        }
        return lo;
    };
    auto row_of_frame = [&](uint64_t fr) -> uint64_t {
        if (fr >= tot_frames)
            return tot_rows;
        uint32_t lo = 0, hi = n_files - 1;
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (files[mid].first_frame <= fr)
                lo = mid;
            else
                hi = mid - 1;
        }
        return files[lo].first_row + (fr - files[lo].first_frame) * files[lo].channels;
    };
    std::vector<Wave> waves;
    uint64_t mx = 0, f = 0;
    while (f < tot_frames)
    {
        const uint64_t r0 = row_of_frame(f);
        uint64_t lo = f + 1, hi = tot_frames; // largest f1 whose rows fit the target (at least one frame)
        while (lo < hi)
        {
            const uint64_t mid = (lo + hi + 1) >> 1;
            if (row_of_frame(mid) - r0 <= ((waves.empty() && first_rows) ? first_rows : target_rows))
                lo = mid;
            else
                hi = mid - 1;
        }
        if (group_align && lo < tot_frames)
        {
            const uint32_t fi = file_of_frame(lo);
            const uint64_t fpg = quant_frames_per_group(files[fi].channels);
            const uint64_t local = lo - files[fi].first_frame;
            uint64_t aligned = files[fi].first_frame + local / fpg * fpg;
            if (aligned <= f) // would make an empty wave: round up instead (never past the file's end)
                aligned = std::min<uint64_t>(files[fi].first_frame + (local / fpg + 1) * fpg,
                                             files[fi].first_frame + files[fi].n_frames);
            lo = aligned;
        }
        Wave w{f, lo, r0, row_of_frame(lo)};
        mx = std::max(mx, w.r1 - w.r0);
        waves.push_back(w);
        f = lo;
    }
    *max_wave_rows = mx;
    return waves;
}

static const uint64_t kRowQuantum = 37ull * kBM;

extern "C" glc_status glc_dma_probe(glc_ctx *c, uint64_t h2d_bytes, uint64_t d2h_bytes, int concurrent, float *elapsed_ms)
{
    if (!c || !elapsed_ms)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    DevScope ds(c, c->compute);
    char *d_up = nullptr, *d_down = nullptr;
    CUDA_TRY(ds.alloc(&d_up, (size_t)std::max<uint64_t>(h2d_bytes, 16)));
    CUDA_TRY(ds.alloc(&d_down, (size_t)std::max<uint64_t>(d2h_bytes, 16)));
    void *h_up = c->pool.alloc(std::max<uint64_t>(h2d_bytes, 16)), *h_down = c->pool.alloc(std::max<uint64_t>(d2h_bytes, 16));
    if (!h_up || !h_down)
    {
        if (h_up)
            c->pool.release(h_up);
        if (h_down)
            c->pool.release(h_down);
        return fail(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
    }
    memset(h_up, 1, (size_t)std::max<uint64_t>(h2d_bytes, 16)); // touch: the pages exist before the clock starts
    memset(h_down, 1, (size_t)std::max<uint64_t>(d2h_bytes, 16));
    // timed on the device: one start event, one end event per direction; the call's figure is the later end
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    cudaError_t e = cudaStreamSynchronize(c->compute);
    for (int i = 0; i < 3 && e == cudaSuccess; ++i)
        e = cudaEventCreate(&ev[i]);
    if (e == cudaSuccess)
        e = cudaEventRecord(ev[0], c->copy);
    if (e == cudaSuccess && h2d_bytes)
        e = cudaMemcpyAsync(d_up, h_up, h2d_bytes, cudaMemcpyHostToDevice, c->copy);
    if (e == cudaSuccess)
        e = cudaEventRecord(ev[1], c->copy);
    // concurrent: the other direction starts together with the first; otherwise after it
    if (e == cudaSuccess)
        e = cudaStreamWaitEvent(c->d2h, concurrent ? ev[0] : ev[1], 0);
    if (e == cudaSuccess && d2h_bytes)
        e = cudaMemcpyAsync(h_down, d_down, d2h_bytes, cudaMemcpyDeviceToHost, c->d2h);
    if (e == cudaSuccess)
        e = cudaEventRecord(ev[2], c->d2h);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->copy);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->d2h);
    float up_ms = 0.0f, down_ms = 0.0f;
    if (e == cudaSuccess)
        e = cudaEventElapsedTime(&up_ms, ev[0], ev[1]);
    if (e == cudaSuccess)
        e = cudaEventElapsedTime(&down_ms, ev[0], ev[2]);
    for (cudaEvent_t x : ev)
        if (x)
            cudaEventDestroy(x);
    c->pool.release(h_up);
    c->pool.release(h_down);
    if (e != cudaSuccess)
        return fail(GLC_ERR_CUDA, "DMA probe failed: %s", cudaGetErrorString(e));
    *elapsed_ms = std::max(up_ms, down_ms);
    return GLC_OK;
}

// ------------------------------------------------- device-resident objects

struct glc_dev_pcm
{
    glc_ctx *ctx;
    float *d;          // interleaved samples (for decode output: untrimmed stream)
    uint64_t n;        // valid interleaved samples after trim
    uint64_t trim_off; // first valid value (decode output)
    uint64_t n_alloc;
    uint16_t channels;
};

struct glc_dev_encoded
{
    glc_ctx *ctx;
    uint32_t sample_rate;
    std::vector<FileDesc> files; // one entry per stream in the batch
    uint64_t n_rows, n_frames;
    uint8_t *d_is_raw;
    uint32_t *d_nnz;
    uint64_t *d_pair_off;
    glc_pair *d_pairs;
    float *d_scales;
    uint64_t *d_raw_off;
    int16_t *d_raw;
    uint64_t n_pairs, n_raw; // valid once totals_known
    bool totals_known;
};

// Pinned host copies of the two variable-length arrays, filled wave by wave while later waves are
// still being computed (host-input encodes only).  The arenas grow like a vector: the first wave's
// density sizes them, a later overflow reallocates and copies.
struct EncodeHostOut
{
    glc_pair *h_pairs = nullptr;
    uint64_t pairs_cap = 0;
    int16_t *h_raw = nullptr;
    uint64_t raw_cap = 0;
    uint64_t n_pairs = 0, n_raw = 0;
};

extern "C" void glc_dev_pcm_free(glc_dev_pcm *p)
{
    if (!p)
        return;
    glc_ctx *c = p->ctx;
    dfree(p->d, c->compute);
    delete p;
}

extern "C" void glc_dev_encoded_free(glc_dev_encoded *e)
{
    if (!e)
        return;
    glc_ctx *c = e->ctx;
    cudaStream_t s = c->compute;
    dfree(e->d_is_raw, s);
    dfree(e->d_nnz, s);
    dfree(e->d_pair_off, s);
    dfree(e->d_pairs, s);
    dfree(e->d_scales, s);
    dfree(e->d_raw_off, s);
    dfree(e->d_raw, s);
    delete e;
}

// ------------------------------------------------------------------ encoder

extern "C" glc_status glc_encoder_new(glc_ctx *c, uint32_t sample_rate, glc_encoder **out)
{
    if (!c || !out)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    if (sample_rate == 0)
        return fail(GLC_ERR_INVALID_ARG, "sample_rate is 0");
    CUDA_TRY(cudaSetDevice(c->device));
    HostPerceptual hp;
    build_host_perceptual(sample_rate, &hp);
    DevPerceptual *dp = new DevPerceptual();
    memset(dp, 0, sizeof *dp);
    memcpy(dp->inv_w, hp.inv_w, sizeof hp.inv_w);
    memcpy(dp->band_edges, hp.band_edges, sizeof hp.band_edges);
    memcpy(dp->band_pf, hp.band_pf, sizeof hp.band_pf);
    memcpy(dp->band_cnt, hp.band_cnt, sizeof hp.band_cnt);
    for (int b = 0; b + 1 < hp.n_edges; ++b)
        for (int k = hp.band_edges[b]; k < hp.band_edges[b + 1]; ++k)
            dp->band_of[k] = (uint8_t)b;
    dp->n_edges = hp.n_edges;
    dp->cf = hp.cf;
    dp->noise_floor_factor = c->host.noise_floor_factor;
    glc_encoder *e = new glc_encoder();
    e->ctx = c;
    e->sample_rate = sample_rate;
    cudaError_t ce = cudaMalloc(&e->d_perc, sizeof(DevPerceptual));
    if (ce == cudaSuccess)
        ce = cudaMemcpy(e->d_perc, dp, sizeof(DevPerceptual), cudaMemcpyHostToDevice);
    delete dp;
    if (ce != cudaSuccess)
    {
        delete e;
        return fail(GLC_ERR_CUDA, "encoder table upload failed: %s", cudaGetErrorString(ce));
    }
    *out = e;
    return GLC_OK;
}

extern "C" void glc_encoder_free(glc_encoder *e)
{
    if (!e)
        return;
    cudaSetDevice(e->ctx->device);
    cudaFree(e->d_perc);
    delete e;
}

// Builds the batch numbering.  Returns GLC_ERR_TOO_SHORT for inputs the reference panics on.
static glc_status build_file_table(uint32_t n_files, const uint64_t *n_samples, const uint16_t *channels,
                                   std::vector<FileDesc> &files, uint64_t *tot_rows, uint64_t *tot_frames,
                                   uint64_t *tot_pcm)
{
    files.resize(n_files);
    uint64_t rows = 0, frames = 0, pcm = 0;
    for (uint32_t i = 0; i < n_files; ++i)
    {
        const uint32_t ch = channels[i];
        if (ch == 0)
            return fail(GLC_ERR_INVALID_ARG, "file %u: channels is 0", i);
        if (n_samples[i] % ch)
            return fail(GLC_ERR_INVALID_ARG, "file %u: %llu samples is not a multiple of %u channels", i,
                        (unsigned long long)n_samples[i], ch);
        const uint64_t L = n_samples[i] / ch;
        if (L <= 512)
            return fail(GLC_ERR_TOO_SHORT,
                        "file %u: %llu samples per channel; the reference panics for <= 512 "
                        "(src/codec.rs:449-452,474)",
                        i, (unsigned long long)L);
        FileDesc &f = files[i];
        f.pcm_off = pcm;
        f.len = L;
        f.first_row = rows;
        f.first_frame = frames;
        f.channels = ch;
        f.n_frames = (uint32_t)frames_for(L);
        rows += (uint64_t)f.n_frames * ch;
        frames += f.n_frames;
        pcm += n_samples[i];
        // keep every file's PCM 16-byte aligned in the arena
        pcm = (pcm + 3) & ~(uint64_t)3;
    }
    *tot_rows = rows;
    *tot_frames = frames;
    *tot_pcm = pcm;
    return GLC_OK;
}

// Host PCM of a batched encode: f32 as the reference's Encoder::encode takes it, or the integer
// samples a WAV/FLAC loader would have divided by 2^(bits-1) (src/audio.rs:39-83) -- then the
// integers cross PCIe (half the bytes for 16-bit sources) and the division runs on the device.
struct HostPcm
{
    const void *const *ptr; // [n_files]
    int elem_bytes;         // 4 = f32 or i32 container, 2 = i16 container
    bool is_int;
    float inv_max;          // 1 / 2^(bits-1)
    void *d_stage;          // device staging arena for integer input (same offsets as the f32 arena)
};

// Core: everything after "PCM is (being put) in the arena".  host_pcm == nullptr means the arena
// is already fully resident; otherwise the H2D copies are issued here, wave by wave.
static glc_status encode_core(glc_encoder *enc, const std::vector<FileDesc> &files, uint64_t tot_rows,
                              uint64_t tot_frames, float *d_arena, const HostPcm *host_pcm,
                              const uint64_t *n_samples, EncodeHostOut *ho, glc_dev_encoded **out)
{
    glc_ctx *c = enc->ctx;
    cudaStream_t cs = c->compute;
    const uint32_t n_files = (uint32_t)files.size();
    PhaseTrace tr("encode", cs);

    DevScope ds(c, cs); // scratch of this call: back to the pool on every exit path
    FileDesc *d_files = nullptr;
    CUDA_TRY(ds.alloc(&d_files, n_files));
    CUDA_TRY(cudaMemcpyAsync(d_files, files.data(), sizeof(FileDesc) * n_files, cudaMemcpyHostToDevice, cs));
    c->stats.h2d_bytes += sizeof(FileDesc) * n_files;

    glc_dev_encoded *de = new glc_dev_encoded();
    *out = de; // owned by the caller from here on, also when a later step fails (it frees what is set)
    de->ctx = c;
    de->sample_rate = enc->sample_rate;
    de->files = files;
    de->n_rows = tot_rows;
    de->n_frames = tot_frames;
    de->d_is_raw = nullptr;
    de->d_nnz = nullptr;
    de->d_pair_off = nullptr;
    de->d_scales = nullptr;
    de->d_raw_off = nullptr;
    de->d_pairs = nullptr;
    de->d_raw = nullptr;
    de->n_pairs = de->n_raw = 0;
    de->totals_known = false;
    glc_pair *d_slots = nullptr;
    uint32_t *d_raw_len = nullptr;
    float *d_coefs = nullptr;
    CUDA_TRY(dmalloc(&de->d_is_raw, tot_frames, cs));
    CUDA_TRY(dmalloc(&de->d_nnz, tot_rows, cs));
    CUDA_TRY(dmalloc(&de->d_pair_off, tot_rows + 1, cs));
    CUDA_TRY(dmalloc(&de->d_scales, tot_rows, cs));
    CUDA_TRY(dmalloc(&de->d_raw_off, tot_frames + 1, cs));
    const bool fast = c->mode == GLC_MODE_FAST;
    if (!fast)
        CUDA_TRY(ds.alloc(&d_raw_len, tot_frames));

    // Wave plan: contiguous frame ranges.  A wave's MDCT grid is (rows/128) x 8 CTAs and 2 x 148 CTAs
    // are resident at a time, so waves are sized in multiples of 37 row tiles (4 736 rows): every
    // resident slot then runs the same number of CTAs and no partial "CTA wave" idles the GPU.
    // Host input: small waves so that the H2D of wave w+1 hides behind the MDCT of wave w.
    // Device-resident input: large waves (fewer launches, scratch still bounded).
    uint64_t target_rows = kRowQuantum * (host_pcm ? 4 : 32);
    if (fast && !host_pcm)
        target_rows = kRowQuantum * 110; // device-resident FAST: one launch per ~521 000 rows (2 GiB of slots at most)
    if (c->wave_frames)
        target_rows = std::max<uint64_t>(c->wave_frames, 1); // explicit tuning: rows per wave
    uint64_t max_wave_rows = 0;
    const std::vector<Wave> waves = plan_waves(files, tot_frames, tot_rows, target_rows, &max_wave_rows, true,
                                               (host_pcm && !c->wave_frames) ? kRowQuantum : 0);
    // The compact outputs are produced wave by wave and their totals are only known on the device, so a wave's
    // share is sized for the worst case (every coefficient kept / every frame raw).  Device-resident results keep
    // all of it; a host-bound encode drains every wave over PCIe one wave behind, so it holds a ring of TWO waves
    // (the gathers subtract the wave's first offsets, GatherLaunch::pair_bias) instead of 8 KiB per row of the batch.
    glc_pair *d_ring_pairs = nullptr;
    int16_t *d_ring_raw = nullptr;
    const uint64_t ring_pairs = max_wave_rows * kHop, ring_raw = max_wave_rows * kFrame; // elements per half
    cudaEvent_t ev_ring[2] = {nullptr, nullptr}; // the D2H copies out of each half have been queued up to here
    if (ho)
    {
        CUDA_TRY(ds.alloc(&d_ring_pairs, 2 * std::max<uint64_t>(ring_pairs, 1)));
        CUDA_TRY(ds.alloc(&d_ring_raw, 2 * std::max<uint64_t>(ring_raw, 1)));
    }
    else
    {
        CUDA_TRY(dmalloc(&de->d_pairs, tot_rows * kHop, cs));
        CUDA_TRY(dmalloc(&de->d_raw, tot_rows * kFrame, cs));
    }
    float *d_atiles = nullptr;
    uint64_t *d_first_group = nullptr;
    // frame groups (max(1, 8/ch) frames of one file): the unit of work of quant_pack and of the FAST kernel
    std::vector<uint64_t> first_group(n_files + 1);
    uint64_t n_groups = 0;
    for (uint32_t i = 0; i < n_files; ++i)
    {
        first_group[i] = n_groups;
        n_groups += quant_groups_for(files[i].n_frames, files[i].channels);
    }
    first_group[n_files] = n_groups;
    auto group_of_frame = [&](uint64_t fr) -> uint64_t { // fr is group-aligned or a file boundary
        if (fr >= tot_frames)
            return n_groups;
        uint32_t lo = 0, hi = n_files - 1;
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi + 1) >> 1;
            if (files[mid].first_frame <= fr)
                lo = mid;
            else
                hi = mid - 1;
        }
        return first_group[lo] + (fr - files[lo].first_frame) / quant_frames_per_group(files[lo].channels);
    };
    CUDA_TRY(ds.alloc(&d_first_group, n_files + 1));
    CUDA_TRY(cudaMemcpyAsync(d_first_group, first_group.data(), 8 * (n_files + 1), cudaMemcpyHostToDevice, cs));
    uint32_t *d_grp_pairs = nullptr, *d_grp_raw = nullptr;
    uint64_t *d_grp_pair_off = nullptr, *d_grp_raw_off = nullptr;
    unsigned int *d_tickets = nullptr;
    // slots live for one wave only (EXACT: per-row slots read by the gather of the same wave; FAST: per-group
    // compact blocks read by the placement of the same wave): sized by the largest wave, indexed by absolute row
    CUDA_TRY(ds.alloc(&d_slots, max_wave_rows * kHop));
    if (!fast)
    {
        CUDA_TRY(ds.alloc(&d_coefs, max_wave_rows * kHop));
        CUDA_TRY(ds.alloc(&d_atiles, mdct_a_tile_floats(max_wave_rows)));
    }
    else
    {
        // per-group totals and their exclusive scans (placement of the groups' compact blocks), one ticket
        // counter per wave
        CUDA_TRY(ds.alloc(&d_grp_pairs, n_groups));
        CUDA_TRY(ds.alloc(&d_grp_raw, n_groups));
        CUDA_TRY(ds.alloc(&d_grp_pair_off, n_groups + 1));
        CUDA_TRY(ds.alloc(&d_grp_raw_off, n_groups + 1));
        CUDA_TRY(ds.alloc(&d_tickets, waves.size()));
        CUDA_TRY(cudaMemsetAsync(d_tickets, 0, sizeof(unsigned int) * waves.size(), cs));
    }
    tr.mark("alloc");

    // H2D plan: file i is needed by the first wave that touches it; copy whole files in order,
    // splitting big files at wave boundaries so that copy and compute overlap.
    EventScope events(c);
    cudaEvent_t ev_copy = nullptr;
    if (host_pcm)
        ev_copy = events.get();
    // host output: per-wave D2H of the compacted pairs / raw bodies on the d2h stream
    struct PinnedGuard // releases the block on every exit path
    {
        PinnedPool &pool;
        void *p;
        ~PinnedGuard()
        {
            if (p)
                pool.release(p);
        }
    } tot_guard{c->pool, nullptr};
    uint64_t *h_tot = nullptr; // [2 * waves] running totals after each wave (pinned)
    std::vector<cudaEvent_t> wave_done;
    if (ho)
    {
        h_tot = (uint64_t *)c->pool.alloc(16 * waves.size());
        tot_guard.p = h_tot;
        if (!h_tot)
            return fail(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
    }
    auto grow = [&](void **buf, uint64_t *cap, uint64_t need, uint64_t keep, size_t elem, uint64_t rows_done,
                    double hwm) -> bool {
        if (need <= *cap)
            return true;
        // density so far (or the densest encode this context has seen) extrapolated to the whole
        // batch, +25 % and 1 Mi elements of slack
        const double dens = std::max(rows_done ? (double)need / (double)rows_done : 0.0, hwm);
        uint64_t ncap = std::max<uint64_t>(need, (uint64_t)(dens * (double)tot_rows * 1.25) + (1u << 20));
        void *nb = c->pool.alloc(ncap * elem);
        if (!nb)
            return false;
        if (keep)
        {
            cudaStreamSynchronize(c->d2h); // earlier waves must have landed before they are moved
            memcpy(nb, *buf, keep * elem);
        }
        if (*buf)
            c->pool.release(*buf);
        *buf = nb;
        *cap = ncap;
        return true;
    };
    auto drain_wave = [&](size_t w) -> glc_status {
        CUDA_TRY(cudaEventSynchronize(wave_done[w]));
        const uint64_t p1 = h_tot[2 * w], q1 = h_tot[2 * w + 1];
        const uint64_t p0 = w ? h_tot[2 * (w - 1)] : 0, q0 = w ? h_tot[2 * (w - 1) + 1] : 0;
        if (!grow((void **)&ho->h_pairs, &ho->pairs_cap, p1, p0, sizeof(glc_pair), waves[w].r1, c->hwm_pairs_per_row) ||
            !grow((void **)&ho->h_raw, &ho->raw_cap, q1, q0, sizeof(int16_t), waves[w].r1, c->hwm_raw_per_row))
            return fail(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
        CUDA_TRY(cudaStreamWaitEvent(c->d2h, wave_done[w], 0));
        if (p1 > p0)
            CUDA_TRY(cudaMemcpyAsync(ho->h_pairs + p0, d_ring_pairs + (w & 1) * ring_pairs, (p1 - p0) * sizeof(glc_pair),
                                     cudaMemcpyDeviceToHost, c->d2h));
        if (q1 > q0)
            CUDA_TRY(cudaMemcpyAsync(ho->h_raw + q0, d_ring_raw + (w & 1) * ring_raw, (q1 - q0) * sizeof(int16_t),
                                     cudaMemcpyDeviceToHost, c->d2h));
        c->stats.d2h_bytes += (p1 - p0) * sizeof(glc_pair) + (q1 - q0) * sizeof(int16_t);
        // wave w + 2 writes the same half of the ring: it waits for these copies
        if (!ev_ring[w & 1])
            ev_ring[w & 1] = events.get();
        CUDA_TRY(cudaEventRecord(ev_ring[w & 1], c->d2h));
        return GLC_OK;
    };
    std::vector<char> file_pinned; // per file: can the DMA engine read the caller's buffer directly?
    if (host_pcm)
    {
        file_pinned.resize(n_files);
        for (uint32_t i = 0; i < n_files; ++i)
            file_pinned[i] = host_ptr_is_pinned(c, host_pcm->ptr[i]) ? 1 : 0;
    }
    std::vector<std::pair<uint64_t, uint64_t>> pending_convert; // (arena offset, count) of integer samples to convert
    uint32_t copy_file = 0;      // next file with bytes left to copy
    uint64_t copy_done = 0;      // interleaved samples of copy_file already enqueued
    for (size_t wi = 0; wi < waves.size(); ++wi)
    {
        const Wave &w = waves[wi];
        if (ho && wi >= 2) // this wave's half of the output ring is free once wave wi - 2 has left for the host
            CUDA_TRY(cudaStreamWaitEvent(cs, ev_ring[wi & 1], 0));
        if (host_pcm)
        {
            // everything up to the last sample any frame < w.f1 can touch
            bool issued = false;
            while (copy_file < n_files)
            {
                const FileDesc &fd = files[copy_file];
                uint64_t need; // interleaved samples of this file needed by frames < w.f1
                if (fd.first_frame >= w.f1)
                    break;
                const uint64_t last_local = std::min<uint64_t>(w.f1 - fd.first_frame, fd.n_frames);
                const uint64_t last_pos = last_local * kHop + kHop; // exclusive, padded coordinates
                const uint64_t smp = last_pos > 512 ? std::min<uint64_t>(last_pos - 512, fd.len) : 0;
                need = smp * fd.channels;
                if (need > copy_done)
                {
                    const size_t eb = (size_t)host_pcm->elem_bytes;
                    char *dst_h2d = host_pcm->is_int ? (char *)host_pcm->d_stage : (char *)d_arena;
                    CUDA_TRY(h2d_async(c, dst_h2d + (fd.pcm_off + copy_done) * eb,
                                       (const char *)host_pcm->ptr[copy_file] + copy_done * eb, (need - copy_done) * eb,
                                       file_pinned[copy_file] != 0, c->copy));
                    c->stats.h2d_bytes += (need - copy_done) * eb;
                    if (host_pcm->is_int)
                    {
                        // one conversion launch per run of files: a file's PCM starts at most 3 elements (the
                        // 16-byte alignment) after the previous one ends, and nobody reads the gap
                        const uint64_t a = fd.pcm_off + copy_done, n = need - copy_done;
                        if (!pending_convert.empty() &&
                            a - (pending_convert.back().first + pending_convert.back().second) <= 4)
                            pending_convert.back().second = a + n - pending_convert.back().first;
                        else
                            pending_convert.push_back({a, n});
                    }
                    copy_done = need;
                    issued = true;
                }
                if (copy_done >= n_samples[copy_file])
                {
                    ++copy_file;
                    copy_done = 0;
                }
                else
                    break;
            }
            if (issued || wi == 0)
            {
                CUDA_TRY(cudaEventRecord(ev_copy, c->copy));
                CUDA_TRY(cudaStreamWaitEvent(cs, ev_copy, 0));
            }
            // integer input: sample as f32 / 2^(bits-1) (src/audio.rs:51-59) for what just arrived
            for (const auto &seg : pending_convert)
            {
                LaunchScope ls(c, GLC_K_MISC, cs);
                CUDA_TRY(launch_pcm_convert(host_pcm->d_stage, host_pcm->elem_bytes, seg.first, seg.second,
                                            host_pcm->inv_max, d_arena, cs));
            }
            pending_convert.clear();
        }
        if (fast)
        {
            FastEncodeLaunch fe{};
            fe.pcm_arena = d_arena;
            fe.files = d_files;
            fe.n_files = n_files;
            fe.first_group = d_first_group;
            fe.group_begin = group_of_frame(w.f0);
            fe.group_end = group_of_frame(w.f1);
            fe.window = c->d_window;
            fe.twiddles = reinterpret_cast<const float2 *>(c->d_fast_tw);
            fe.norm = c->host.norm;
            fe.perc = enc->d_perc;
            fe.nnz = de->d_nnz;
            fe.scales = de->d_scales;
            fe.is_raw = de->d_is_raw;
            fe.slots = d_slots - w.r0 * kHop;
            fe.grp_pairs = d_grp_pairs;
            fe.grp_raw = d_grp_raw;
            fe.ticket = d_tickets + wi;
            fe.grp_pair_off = d_grp_pair_off;
            fe.grp_raw_off = d_grp_raw_off;
            fe.pair_off = de->d_pair_off;
            fe.raw_off = de->d_raw_off;
            fe.pairs = ho ? d_ring_pairs + (wi & 1) * ring_pairs : de->d_pairs;
            fe.raw = ho ? d_ring_raw + (wi & 1) * ring_raw : de->d_raw;
            fe.ring = ho != nullptr;
            {
                LaunchScope ls(c, GLC_K_FAST_ENCODE, cs);
                CUDA_TRY(launch_fast_encode(fe, cs));
            }
            // placement: scans of the per-group totals (continuing from the previous wave), then the move
            {
                LaunchScope ls2(c, GLC_K_SCAN, cs, 2);
                const uint64_t g0 = fe.group_begin, gn = fe.group_end - fe.group_begin;
                CUDA_TRY(launch_scan_u32_u64(d_grp_pairs + g0, d_grp_pair_off + g0, gn, cs, wi ? d_grp_pair_off + g0 : nullptr));
                CUDA_TRY(launch_scan_u32_u64(d_grp_raw + g0, d_grp_raw_off + g0, gn, cs, wi ? d_grp_raw_off + g0 : nullptr));
            }
            {
                LaunchScope ls3(c, GLC_K_GATHER, cs, 1);
                CUDA_TRY(launch_fast_place(fe, cs));
            }
        }
        MdctLaunch m{};
        if (!fast)
        {
            LaunchScope ls(c, GLC_K_WINDOW_TILE, cs);
            m.pcm_arena = d_arena;
            m.files = d_files;
            m.n_files = n_files;
            m.row_begin = w.r0;
            m.row_end = w.r1;
            m.tab_tiled = c->d_tab_mdct;
            m.window = c->d_window;
            m.norm = c->host.norm;
            m.coefs = d_coefs - w.r0 * kHop; // kernels index coefficients by absolute row
            m.a_tiles = d_atiles;
            CUDA_TRY(launch_window_tiles(m, cs));
        }
        if (!fast)
        {
            LaunchScope ls(c, GLC_K_MDCT_EXACT, cs);
            CUDA_TRY(launch_mdct_exact(m, cs));
        }
        if (!fast)
        {
            LaunchScope ls(c, GLC_K_QUANT_PACK, cs);
            QuantPackLaunch q{};
            q.coefs = d_coefs - w.r0 * kHop;
            q.files = d_files;
            q.n_files = n_files;
            q.first_group = d_first_group;
            q.group_begin = group_of_frame(w.f0);
            q.group_end = group_of_frame(w.f1);
            q.perc = enc->d_perc;
            q.slots = d_slots - w.r0 * kHop;
            q.nnz = de->d_nnz;
            q.scales = de->d_scales;
            q.is_raw = de->d_is_raw;
            q.raw_len = d_raw_len;
            CUDA_TRY(launch_quant_pack(q, cs));
        }
        // ---- variable-length layout of this wave: the scans continue from the previous wave's totals ----
        if (!fast)
        {
            LaunchScope ls(c, GLC_K_SCAN, cs, 2);
            CUDA_TRY(launch_scan_u32_u64(de->d_nnz + w.r0, de->d_pair_off + w.r0, w.r1 - w.r0, cs,
                                         wi ? de->d_pair_off + w.r0 : nullptr));
            CUDA_TRY(launch_scan_u32_u64(d_raw_len + w.f0, de->d_raw_off + w.f0, w.f1 - w.f0, cs,
                                         wi ? de->d_raw_off + w.f0 : nullptr));
        }
        if (!fast)
        {
            LaunchScope ls(c, GLC_K_GATHER, cs, 2);
            GatherLaunch g{};
            g.slots = d_slots - w.r0 * kHop;
            g.nnz = de->d_nnz;
            g.pair_off = de->d_pair_off;
            g.pairs = ho ? d_ring_pairs + (wi & 1) * ring_pairs : de->d_pairs;
            g.is_raw = de->d_is_raw;
            g.raw_off = de->d_raw_off;
            g.raw = ho ? d_ring_raw + (wi & 1) * ring_raw : de->d_raw;
            g.pair_bias = ho ? de->d_pair_off + w.r0 : nullptr;
            g.raw_bias = ho ? de->d_raw_off + w.f0 : nullptr;
            g.pcm_arena = d_arena;
            g.files = d_files;
            g.n_files = n_files;
            g.window = c->d_window;
            g.row_begin = w.r0;
            g.row_end = w.r1;
            g.frame_begin = w.f0;
            g.frame_end = w.f1;
            CUDA_TRY(launch_gather(g, cs));
        }
        if (ho)
        {
            CUDA_TRY(cudaMemcpyAsync(h_tot + 2 * wi, de->d_pair_off + w.r1, 8, cudaMemcpyDeviceToHost, cs));
            CUDA_TRY(cudaMemcpyAsync(h_tot + 2 * wi + 1, de->d_raw_off + w.f1, 8, cudaMemcpyDeviceToHost, cs));
            cudaEvent_t ev = events.get();
            wave_done.push_back(ev);
            CUDA_TRY(cudaEventRecord(ev, cs));
            if (wi >= 1)
                GLC_TRY(drain_wave(wi - 1)); // one wave behind: the GPU already has wave wi queued
        }
    }
    if (ho)
    {
        GLC_TRY(drain_wave(waves.size() - 1));
        ho->n_pairs = h_tot[2 * (waves.size() - 1)];
        ho->n_raw = h_tot[2 * (waves.size() - 1) + 1];
        de->n_pairs = ho->n_pairs;
        de->n_raw = ho->n_raw;
        de->totals_known = true;
        c->hwm_pairs_per_row = std::max(c->hwm_pairs_per_row, (double)ho->n_pairs / (double)tot_rows);
        c->hwm_raw_per_row = std::max(c->hwm_raw_per_row, (double)ho->n_raw / (double)tot_rows);
        c->stats.d2h_bytes += 16 * waves.size();
        // the ring goes back to the pool in compute-stream order: the last copies out of it come first
        for (cudaEvent_t e : ev_ring)
            if (e)
                CUDA_TRY(cudaStreamWaitEvent(cs, e, 0));
    }
    tr.mark("waves");
    return GLC_OK;
}

static void block_unref(EncodedBlock *b)
{
    if (b->refs.fetch_sub(1, std::memory_order_acq_rel) > 1)
        return;
    for (void *p : b->pinned)
        b->ctx->pool.release(p);
    for (void *p : b->heap)
        free(p);
    delete b;
}

extern "C" void glc_encoded_free(glc_ctx *c, glc_encoded *e)
{
    (void)c;
    if (!e)
        return;
    EncodedBox *box = reinterpret_cast<EncodedBox *>(e);
    block_unref(box->blk);
    delete box;
}

static glc_status download_encoded(glc_dev_encoded *de, EncodeHostOut *ho, glc_encoded **out /* [n_files] */)
{
    glc_ctx *c = de->ctx;
    cudaStream_t cs = c->compute;
    const uint64_t R = de->n_rows, F = de->n_frames;
    if (!de->totals_known)
    {
        uint64_t totals[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(&totals[0], de->d_pair_off + R, 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaMemcpyAsync(&totals[1], de->d_raw_off + F, 8, cudaMemcpyDeviceToHost, cs));
        CUDA_TRY(cudaStreamSynchronize(cs));
        de->n_pairs = totals[0];
        de->n_raw = totals[1];
        de->totals_known = true;
    }
    EncodedBlock *blk = new EncodedBlock();
    blk->ctx = c;
    blk->refs = 0;
    size_t n_boxes = 0;
    // a failure below releases the block (pinned buffers included) and the boxes already handed out
    auto fail_cleanup = [&]() {
        for (size_t i = 0; i < n_boxes; ++i)
        {
            delete reinterpret_cast<EncodedBox *>(out[i]);
            out[i] = nullptr;
        }
        blk->refs = 1;
        block_unref(blk);
    };
#define DL_TRY(expr)                                                                                       \
    do                                                                                                     \
    {                                                                                                      \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
        {                                                                                                  \
            fail_cleanup();                                                                                \
            return fail(GLC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                        __LINE__);                                                                         \
        }                                                                                                  \
    } while (0)
    auto pin = [&](size_t bytes) -> void * {
        void *p = c->pool.alloc(bytes);
        if (p)
            blk->pinned.push_back(p);
        return p;
    };
    uint8_t *h_is_raw = (uint8_t *)pin(F);
    uint32_t *h_nnz = (uint32_t *)pin(R * 4);
    uint64_t *h_pair_off = (uint64_t *)pin((R + 1) * 8);
    // the two big arrays may already have come down wave by wave (host-input encode)
    glc_pair *h_pairs = nullptr;
    int16_t *h_raw = nullptr;
    if (ho)
    {
        // an array that never grew is empty (no pairs / no raw frame in the whole batch): the device side of a
        // host-bound encode is a ring of two waves, there is nothing left to fetch from it
        if (ho->h_pairs)
            blk->pinned.push_back(h_pairs = ho->h_pairs);
        else
            h_pairs = (glc_pair *)pin(0);
        if (ho->h_raw)
            blk->pinned.push_back(h_raw = ho->h_raw);
        else
            h_raw = (int16_t *)pin(0);
        ho->h_pairs = nullptr;
        ho->h_raw = nullptr;
    }
    else
    {
        h_pairs = (glc_pair *)pin(de->n_pairs * 4);
        h_raw = (int16_t *)pin(de->n_raw * 2);
    }
    float *h_scales = (float *)pin(R * 4);
    uint64_t *h_raw_off = (uint64_t *)pin((F + 1) * 8);
    if (!h_is_raw || !h_nnz || !h_pair_off || !h_pairs || !h_scales || !h_raw_off || !h_raw)
    {
        blk->refs = 1;
        block_unref(blk);
        return fail(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
    }
    DL_TRY(cudaMemcpyAsync(h_is_raw, de->d_is_raw, F, cudaMemcpyDeviceToHost, cs));
    DL_TRY(cudaMemcpyAsync(h_nnz, de->d_nnz, R * 4, cudaMemcpyDeviceToHost, cs));
    DL_TRY(cudaMemcpyAsync(h_pair_off, de->d_pair_off, (R + 1) * 8, cudaMemcpyDeviceToHost, cs));
    DL_TRY(cudaMemcpyAsync(h_scales, de->d_scales, R * 4, cudaMemcpyDeviceToHost, cs));
    DL_TRY(cudaMemcpyAsync(h_raw_off, de->d_raw_off, (F + 1) * 8, cudaMemcpyDeviceToHost, cs));
    if (!ho)
    {
        if (de->n_pairs)
            DL_TRY(cudaMemcpyAsync(h_pairs, de->d_pairs, de->n_pairs * 4, cudaMemcpyDeviceToHost, cs));
        if (de->n_raw)
            DL_TRY(cudaMemcpyAsync(h_raw, de->d_raw, de->n_raw * 2, cudaMemcpyDeviceToHost, cs));
        c->stats.d2h_bytes += de->n_pairs * 4 + de->n_raw * 2;
    }
    DL_TRY(cudaStreamSynchronize(cs));
    DL_TRY(cudaStreamSynchronize(c->d2h));
    c->stats.d2h_bytes += F + R * 4 + (R + 1) * 8 + R * 4 + (F + 1) * 8;

    const size_t nf = de->files.size();
    for (size_t i = 0; i < nf; ++i)
    {
        const FileDesc &fd = de->files[i];
        EncodedBox *box = new EncodedBox();
        box->blk = blk;
        blk->refs++;
        glc_encoded &e = box->pub;
        memset(&e, 0, sizeof e);
        const uint64_t rows = (uint64_t)fd.n_frames * fd.channels;
        e.sample_rate = de->sample_rate;
        e.channels = (uint16_t)fd.channels;
        e.total_samples = fd.len * fd.channels;
        e.encoder_delay = kHop / 2;                                   // src/codec.rs:547
        e.padding = (uint32_t)(padded_len(fd.len) - fd.len - kHop / 2); // :546
        e.original_length = e.total_samples;                          // :562
        e.n_frames = fd.n_frames;
        e.frame_is_raw = h_is_raw + fd.first_frame;
        e.nnz = h_nnz + fd.first_row;
        e.scales = h_scales + fd.first_row;
        e.pairs = h_pairs + h_pair_off[fd.first_row];
        e.raw = h_raw + h_raw_off[fd.first_frame];
        if (nf == 1)
        {
            e.pair_offset = h_pair_off;
            e.raw_offset = h_raw_off;
        }
        else
        {
            // rebase the exclusive scans so that each file's offsets start at 0
            uint64_t *po = (uint64_t *)malloc((rows + 1) * 8);
            uint64_t *ro = (uint64_t *)malloc(((uint64_t)fd.n_frames + 1) * 8);
            if (!po || !ro)
            {
                free(po);
                free(ro);
                delete box;
                blk->refs--;
                fail_cleanup();
                return fail(GLC_ERR_NO_MEMORY, "out of host memory");
            }
            blk->heap.push_back(po);
            blk->heap.push_back(ro);
            const uint64_t pb = h_pair_off[fd.first_row], rb = h_raw_off[fd.first_frame];
            for (uint64_t r = 0; r <= rows; ++r)
                po[r] = h_pair_off[fd.first_row + r] - pb;
            for (uint64_t f = 0; f <= fd.n_frames; ++f)
                ro[f] = h_raw_off[fd.first_frame + f] - rb;
            e.pair_offset = po;
            e.raw_offset = ro;
        }
        out[i] = &box->pub;
        n_boxes = i + 1;
    }
    return GLC_OK;
#undef DL_TRY
}

static glc_status encode_batch_impl(glc_encoder *enc, uint32_t n_files, const void *const *pcm, int elem_bytes,
                                    bool is_int, uint32_t bits, const uint64_t *n_samples, const uint16_t *channels,
                                    glc_encoded **out)
{
    if (!enc || !pcm || !n_samples || !channels || !out || n_files == 0)
        return fail(GLC_ERR_INVALID_ARG, "null/empty argument");
    if (is_int && (bits < 1 || bits > (uint32_t)elem_bytes * 8))
        return fail(GLC_ERR_INVALID_ARG, "bits_per_sample %u does not fit a %d-byte container", bits, elem_bytes);
    glc_ctx *c = enc->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    std::vector<FileDesc> files;
    uint64_t rows, frames, tot_pcm;
    GLC_TRY(build_file_table(n_files, n_samples, channels, files, &rows, &frames, &tot_pcm));
    for (uint32_t i = 0; i < n_files; ++i)
        if (!pcm[i])
            return fail(GLC_ERR_INVALID_ARG, "file %u: pcm is null", i);
    DevScope arena_scope(c, c->copy);
    float *d_arena = nullptr;
    CUDA_TRY(arena_scope.alloc(&d_arena, tot_pcm));
    HostPcm hp{};
    hp.ptr = pcm;
    hp.elem_bytes = elem_bytes;
    hp.is_int = is_int;
    hp.inv_max = is_int ? 1.0f / (float)(1ull << (bits - 1)) : 1.0f;
    char *d_stage = nullptr;
    if (is_int)
        CUDA_TRY(arena_scope.alloc(&d_stage, tot_pcm * (size_t)elem_bytes));
    hp.d_stage = d_stage;
    glc_dev_encoded *de = nullptr;
    EncodeHostOut ho;
    glc_status st = encode_core(enc, files, rows, frames, d_arena, &hp, n_samples, &ho, &de);
    if (st == GLC_OK)
        st = download_encoded(de, &ho, out);
    cudaStreamSynchronize(c->d2h);
    if (ho.h_pairs)
        c->pool.release(ho.h_pairs);
    if (ho.h_raw)
        c->pool.release(ho.h_raw);
    if (de)
        glc_dev_encoded_free(de);
    cudaStreamSynchronize(c->compute); // the arenas (arena_scope) are idle when they return to the pool
    return st;
}

extern "C" glc_status glc_encode_batch(glc_encoder *enc, uint32_t n_files, const float *const *pcm,
                                       const uint64_t *n_samples, const uint16_t *channels, glc_encoded **out)
{
    return encode_batch_impl(enc, n_files, (const void *const *)pcm, 4, false, 0, n_samples, channels, out);
}

extern "C" glc_status glc_encode_batch_i16(glc_encoder *enc, uint32_t n_files, const int16_t *const *pcm,
                                           const uint64_t *n_samples, const uint16_t *channels, glc_encoded **out)
{
    return encode_batch_impl(enc, n_files, (const void *const *)pcm, 2, true, 16, n_samples, channels, out);
}

extern "C" glc_status glc_encode_i16(glc_encoder *enc, const int16_t *pcm, uint64_t n_samples, uint16_t channels,
                                     glc_encoded **out)
{
    if (!out)
        return fail(GLC_ERR_INVALID_ARG, "out is null");
    const void *files[1] = {pcm};
    return encode_batch_impl(enc, 1, files, 2, true, 16, &n_samples, &channels, out);
}

extern "C" glc_status glc_encode_i32(glc_encoder *enc, const int32_t *pcm, uint64_t n_samples, uint16_t channels,
                                     uint32_t bits_per_sample, glc_encoded **out)
{
    if (!out)
        return fail(GLC_ERR_INVALID_ARG, "out is null");
    const void *files[1] = {pcm};
    return encode_batch_impl(enc, 1, files, 4, true, bits_per_sample, &n_samples, &channels, out);
}

extern "C" glc_status glc_encode(glc_encoder *enc, const float *pcm, uint64_t n_samples, uint16_t channels,
                                 glc_encoded **out)
{
    if (!out)
        return fail(GLC_ERR_INVALID_ARG, "out is null");
    const float *files[1] = {pcm};
    return glc_encode_batch(enc, 1, files, &n_samples, &channels, out);
}

extern "C" glc_status glc_dev_upload(glc_ctx *c, const float *pcm, uint64_t n_samples, uint16_t channels,
                                     glc_dev_pcm **out)
{
    if (!c || !pcm || !out || channels == 0)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    glc_dev_pcm *p = new glc_dev_pcm();
    p->ctx = c;
    p->n = n_samples;
    p->trim_off = 0;
    p->n_alloc = n_samples;
    p->channels = channels;
    CUDA_TRY(dmalloc(&p->d, n_samples, c->compute));
    CUDA_TRY(cudaMemcpyAsync(p->d, pcm, n_samples * sizeof(float), cudaMemcpyHostToDevice, c->compute));
    CUDA_TRY(cudaStreamSynchronize(c->compute));
    c->stats.h2d_bytes += n_samples * sizeof(float);
    *out = p;
    return GLC_OK;
}

extern "C" glc_status glc_dev_encode(glc_encoder *enc, const glc_dev_pcm *pcm, glc_dev_encoded **out)
{
    if (!enc || !pcm || !out)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    glc_ctx *c = enc->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    std::vector<FileDesc> files;
    uint64_t rows, frames, tot_pcm;
    const uint64_t n = pcm->n;
    const uint16_t ch = pcm->channels;
    GLC_TRY(build_file_table(1, &n, &ch, files, &rows, &frames, &tot_pcm));
    *out = nullptr;
    const glc_status st = encode_core(enc, files, rows, frames, pcm->d + pcm->trim_off, nullptr, nullptr, nullptr, out);
    if (st != GLC_OK && *out)
    {
        glc_dev_encoded_free(*out);
        *out = nullptr;
    }
    return st;
}

extern "C" glc_status glc_dev_encoded_download(const glc_dev_encoded *de, glc_encoded **out)
{
    if (!de || !out)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(de->ctx->device));
    return download_encoded(const_cast<glc_dev_encoded *>(de), nullptr, out);
}

// ------------------------------------------------------------------ decoder

extern "C" glc_status glc_decoder_new(glc_ctx *c, uint32_t channels, uint32_t sample_rate, glc_decoder **out)
{
    if (!c || !out)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    glc_decoder *d = new glc_decoder();
    d->ctx = c;
    d->channels = channels;
    d->sample_rate = sample_rate;
    *out = d;
    return GLC_OK;
}

extern "C" void glc_decoder_free(glc_decoder *d) { delete d; }

static glc_status validate_encoded(const glc_encoded *e, uint32_t idx)
{
    if (!e)
        return fail(GLC_ERR_INVALID_ARG, "stream %u is null", idx);
    if (e->channels == 0)
        return fail(GLC_ERR_CORRUPT, "stream %u: 0 channels", idx);
    if (e->n_frames && (!e->frame_is_raw || !e->nnz || !e->pair_offset || !e->scales || !e->raw_offset))
        return fail(GLC_ERR_CORRUPT, "stream %u: missing arrays", idx);
    return GLC_OK;
}

// Host side of a decode whose streams and output live in host memory: per wave, the pairs / raw
// frames go up on the copy stream and the finished PCM comes down on the d2h stream while the
// compute stream works on the waves in between.
struct DecodeHostIO
{
    const glc_encoded *const *enc;  // [n_files]
    const uint64_t *h_pair_off;     // [rows+1] batch-wide exclusive scan (host copy)
    const uint64_t *h_raw_off;      // [frames+1]
    float *const *h_out;            // [n_files] pinned outputs
    bool out16;                     // outputs are 16-bit PCM (h_out entries are int16_t*): converted on the device
    const uint64_t *win_off;        // [n_files] first untrimmed value that is kept (gapless trim)
    const uint64_t *win_len;        // [n_files] number of values kept
    glc_pair *d_pairs;              // device arrays being filled wave by wave
    int16_t *d_raw;
    const char *pinned;             // [n_files] the stream's pairs AND raw arrays can be read by the DMA engine
};

// Device part of a batched decode.  With io == nullptr the stream arrays are already on the device
// and the output stays there.
static glc_status decode_core(glc_ctx *c, const std::vector<DecFileDesc> &files, uint64_t tot_rows,
                              uint64_t tot_frames, uint64_t total_out, const uint8_t *d_is_raw,
                              const uint64_t *d_pair_off, const glc_pair *d_pairs, const float *d_scales,
                              const uint64_t *d_raw_off, const int16_t *d_raw, const DecodeHostIO *io,
                              float **d_out_ret)
{
    cudaStream_t cs = c->compute;
    const uint32_t n_files = (uint32_t)files.size();
    DecFileDesc *d_files = nullptr;
    float *d_atiles = nullptr, *d_blocks = nullptr, *d_out = nullptr;
    uint32_t *d_flags = nullptr, *d_active = nullptr, *d_ntiles = nullptr, *d_nk = nullptr;
    uint64_t *d_slot_off = nullptr;
    int32_t *d_row_slot = nullptr;
    uint8_t *d_stage_list = nullptr;
    PhaseTrace tr("decode", cs);

    uint64_t target_rows = kRowQuantum * (io ? 4 : 32);
    if (c->wave_frames)
        target_rows = std::max<uint64_t>(c->wave_frames, 1);
    uint64_t max_wave_rows = 0;
    std::vector<Wave> waves = plan_waves(files, tot_frames, tot_rows, target_rows, &max_wave_rows, false,
                                         (io && !c->wave_frames) ? kRowQuantum : 0);
    if (waves.empty())
        waves.push_back(Wave{0, 0, 0, 0}); // only streams without frames: their hops (zeros) are still produced
    const uint64_t wave_tiles = (max_wave_rows + kImdctBM - 1) / kImdctBM;

    DevScope ds(c, cs); // every scratch block below goes back to the pool on any exit path
    CUDA_TRY(ds.alloc(&d_files, n_files));
    CUDA_TRY(cudaMemcpyAsync(d_files, files.data(), sizeof(DecFileDesc) * n_files, cudaMemcpyHostToDevice, cs));
    c->stats.h2d_bytes += sizeof(DecFileDesc) * n_files;
    // Worst-case sizes per WAVE (every row transformed, every k present); the live counts stay on the device,
    // so the decode needs no host round trip.  `blocks` is a ring of two waves: hop h needs frames h-1 and h,
    // i.e. the overlap-add of a wave reads its own blocks and the last frame of the wave before it, never
    // anything older (8 KiB per row of two waves instead of 8 KiB per row of the whole batch).
    const uint64_t ring_cap = (max_wave_rows + kImdctBM - 1) / kImdctBM * kImdctBM + kBM; // slots per ring half
    CUDA_TRY(ds.alloc(&d_atiles, wave_tiles * kImdctATileFloats));
    CUDA_TRY(ds.alloc(&d_blocks, 2 * ring_cap * kFrame));
    CUDA_TRY(ds.alloc(&d_flags, max_wave_rows));
    CUDA_TRY(ds.alloc(&d_slot_off, max_wave_rows + 1));
    CUDA_TRY(ds.alloc(&d_row_slot, tot_rows));
    CUDA_TRY(ds.alloc(&d_active, max_wave_rows));
    CUDA_TRY(ds.alloc(&d_ntiles, 1));
    CUDA_TRY(ds.alloc(&d_nk, wave_tiles));
    CUDA_TRY(ds.alloc(&d_stage_list, wave_tiles * kImdctStages));
    CUDA_TRY(ds.alloc(&d_out, total_out));
    int16_t *d_out16 = nullptr;
    if (io && io->h_out && io->out16)
        CUDA_TRY(ds.alloc(&d_out16, total_out));
    tr.mark("alloc");

    // A frame boundary `fr` belongs to the LOWEST-numbered file that starts there (streams without frames
    // share their first_frame with the next file and still own one hop, the all-zero final overlap of
    // src/codec.rs:723-729), else to the file that contains it.
    auto file_at_boundary = [&](uint64_t fr) -> uint32_t {
        uint32_t lo = 0, hi = n_files; // first file with first_frame >= fr
        while (lo < hi)
        {
            const uint32_t mid = (lo + hi) >> 1;
            if (files[mid].first_frame < fr)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo < n_files && files[lo].first_frame == fr)
            return lo;
        return lo - 1; // fr > 0 here, so lo >= 1
    };
    auto out_index = [&](uint64_t fr) -> uint64_t { // first output value that needs frame `fr`
        if (fr >= tot_frames)
            return total_out;
        const uint32_t i = file_at_boundary(fr);
        return files[i].out_off + (fr - files[i].first_frame) * kHop * files[i].channels;
    };
    auto hop_id = [&](uint64_t fr) -> uint64_t { // batch-wide hop number of the hop that starts at frame `fr`
        if (fr >= tot_frames)
            return tot_frames + n_files;
        return fr + file_at_boundary(fr);
    };
    uint32_t ola_tile_channels = 0;
    for (const DecFileDesc &f : files)
        if (f.channels >= 3 && f.channels <= 8)
            ola_tile_channels = std::max(ola_tile_channels, f.channels);
    EventScope events(c);
    if (io)
    {
        // the copy stream writes buffers that were handed out in compute-stream order
        cudaEvent_t ev = events.get();
        CUDA_TRY(cudaEventRecord(ev, cs));
        CUDA_TRY(cudaStreamWaitEvent(c->copy, ev, 0));
    }
    for (size_t wi = 0; wi < waves.size(); ++wi)
    {
        const Wave &w = waves[wi];
        const uint64_t slot_base = (wi & 1) * ring_cap; // this wave's half of the block ring
        if (io)
        {
            // H2D of this wave's pairs and raw frames, file segment by file segment
            bool any = false;
            for (uint32_t i = 0; i < n_files; ++i)
            {
                const DecFileDesc &f = files[i];
                const uint64_t fr_rows = f.n_frames * f.channels;
                const uint64_t a = std::max(w.r0, f.first_row), b = std::min(w.r1, f.first_row + fr_rows);
                if (a < b)
                {
                    const uint64_t p0 = io->h_pair_off[a], p1 = io->h_pair_off[b], pf = io->h_pair_off[f.first_row];
                    if (p1 > p0)
                    {
                        CUDA_TRY(h2d_async(c, io->d_pairs + p0, io->enc[i]->pairs + (p0 - pf), (p1 - p0) * 4,
                                           io->pinned[i] != 0, c->copy));
                        c->stats.h2d_bytes += (p1 - p0) * 4;
                        any = true;
                    }
                }
                const uint64_t fa = std::max(w.f0, f.first_frame), fb = std::min(w.f1, f.first_frame + f.n_frames);
                if (fa < fb)
                {
                    const uint64_t q0 = io->h_raw_off[fa], q1 = io->h_raw_off[fb], qf = io->h_raw_off[f.first_frame];
                    if (q1 > q0)
                    {
                        CUDA_TRY(h2d_async(c, io->d_raw + q0, io->enc[i]->raw + (q0 - qf), (q1 - q0) * 2,
                                           io->pinned[i] != 0, c->copy));
                        c->stats.h2d_bytes += (q1 - q0) * 2;
                        any = true;
                    }
                }
            }
            if (any || wi == 0)
            {
                cudaEvent_t ev = events.get();
                CUDA_TRY(cudaEventRecord(ev, c->copy));
                CUDA_TRY(cudaStreamWaitEvent(cs, ev, 0));
            }
        }
        if (c->mode == GLC_MODE_FAST)
        {
            LaunchScope ls(c, GLC_K_FAST_DECODE, cs);
            FastDecodeLaunch fdl{};
            fdl.pairs = d_pairs;
            fdl.pair_off = d_pair_off;
            fdl.scales = d_scales;
            fdl.is_raw = d_is_raw;
            fdl.files = d_files;
            fdl.n_files = n_files;
            fdl.row_begin = w.r0;
            fdl.row_end = w.r1;
            fdl.window = c->d_window;
            fdl.twiddles = reinterpret_cast<const float2 *>(c->d_fast_tw);
            fdl.norm = c->host.norm;
            fdl.row_slot = d_row_slot;
            fdl.slot_base = slot_base;
            fdl.blocks = d_blocks;
            CUDA_TRY(launch_fast_decode(fdl, cs));
        }
        else
        {
            LaunchScope ls(c, GLC_K_DEQUANT, cs, 4);
            DequantLaunch q{};
            q.pairs = d_pairs;
            q.pair_off = d_pair_off;
            q.scales = d_scales;
            q.is_raw = d_is_raw;
            q.files = d_files;
            q.n_files = n_files;
            q.row_begin = w.r0;
            q.row_end = w.r1;
            q.slot_base = slot_base; // slots <= rows of the wave: they fit its half of the ring
            q.row_slot = d_row_slot;
            q.flags = d_flags;
            q.slot_off = d_slot_off;
            q.active_rows = d_active;
            q.n_tiles = d_ntiles;
            q.stage_list = d_stage_list;
            q.n_stages = d_nk;
            q.a_tiles = d_atiles;
            CUDA_TRY(launch_dequant(q, cs));
        }
        if (c->mode != GLC_MODE_FAST)
        {
            LaunchScope ls(c, GLC_K_IMDCT_EXACT, cs);
            ImdctLaunch m{};
            m.a_tiles = d_atiles;
            m.stage_list = d_stage_list;
            m.n_stages = d_nk;
            m.n_tiles = d_ntiles;
            m.max_slots = (w.r1 - w.r0 + kImdctBM - 1) / kImdctBM * kImdctBM;
            m.tab = c->d_tab_imdct;
            m.window = c->d_window;
            m.norm = c->host.norm;
            m.blocks = d_blocks + slot_base * kFrame;
            CUDA_TRY(launch_imdct_exact(m, cs));
        }
        const uint64_t o0 = wi ? out_index(w.f0) : 0, o1 = out_index(w.f1);
        {
            LaunchScope ls(c, GLC_K_OLA, cs);
            OlaLaunch o{};
            o.blocks = d_blocks;
            o.row_slot = d_row_slot;
            o.is_raw = d_is_raw;
            o.raw_off = d_raw_off;
            o.raw = d_raw;
            o.files = d_files;
            o.n_files = n_files;
            o.hop_begin = wi ? hop_id(w.f0) : 0;
            o.hop_end = hop_id(w.f1);
            o.out = d_out;
            o.tile_channels = ola_tile_channels;
            CUDA_TRY(launch_ola(o, cs));
        }
        if (d_out16 && o1 > o0)
        {
            LaunchScope ls(c, GLC_K_MISC, cs);
            CUDA_TRY(launch_pcm_to_i16(d_out + o0, d_out16 + o0, o1 - o0, cs));
        }
        if (io && io->h_out)
        {
            // D2H of the finished range, clipped to every file's gapless window
            cudaEvent_t ev = events.get();
            CUDA_TRY(cudaEventRecord(ev, cs));
            CUDA_TRY(cudaStreamWaitEvent(c->d2h, ev, 0));
            for (uint32_t i = 0; i < n_files; ++i)
            {
                const uint64_t w0 = files[i].out_off + io->win_off[i], w1 = w0 + io->win_len[i];
                const uint64_t a = std::max(o0, w0), b = std::min(o1, w1);
                if (a < b && d_out16)
                {
                    CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<int16_t *>(io->h_out[i]) + (a - w0), d_out16 + a,
                                             (b - a) * 2, cudaMemcpyDeviceToHost, c->d2h));
                    c->stats.d2h_bytes += (b - a) * 2;
                }
                else if (a < b)
                {
                    CUDA_TRY(cudaMemcpyAsync(io->h_out[i] + (a - w0), d_out + a, (b - a) * 4, cudaMemcpyDeviceToHost,
                                             c->d2h));
                    c->stats.d2h_bytes += (b - a) * 4;
                }
            }
        }
    }
    tr.mark("waves");
    if (io)
    {
        CUDA_TRY(cudaStreamSynchronize(c->d2h));
        CUDA_TRY(cudaStreamSynchronize(cs));
    }
    ds.keep(d_out); // the caller owns the output; the scratch blocks are released by `ds`
    tr.mark("free");
    *d_out_ret = d_out;
    return GLC_OK;
}

// Trim rule of Decoder::decode, src/codec.rs:755-765
static void trim_window(uint64_t untrimmed, uint32_t delay, uint64_t original_length, uint64_t *off, uint64_t *len)
{
    uint64_t o = 0, n = untrimmed;
    if (n > delay)
    {
        o = delay;
        n -= delay;
    }
    if (n > original_length)
        n = original_length;
    *off = o;
    *len = n;
}

// Where a decode leaves its PCM when the caller wants it to stay in HBM (fused decode -> FLAC).
struct DeviceSink
{
    float *d_out = nullptr;          // untrimmed streams of all files, caller frees with dfree(.., compute)
    std::vector<uint64_t> off, len;  // per file: first kept value (absolute in d_out) and kept count
};

static glc_status decode_batch_impl(glc_decoder *dec, uint32_t n_files, const glc_encoded *const *enc, bool trim,
                                    float **pcm, uint64_t *n_out, DeviceSink *sink = nullptr, bool out16 = false)
{
    if (!dec || !enc || (!sink && (!pcm || !n_out)) || n_files == 0)
        return fail(GLC_ERR_INVALID_ARG, "null/empty argument");
    glc_ctx *c = dec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t cs = c->compute;
    std::vector<DecFileDesc> files(n_files);
    uint64_t rows = 0, frames = 0, outv = 0, npairs = 0, nraw = 0;
    for (uint32_t i = 0; i < n_files; ++i)
    {
        GLC_TRY(validate_encoded(enc[i], i));
        const glc_encoded *e = enc[i];
        DecFileDesc &f = files[i];
        f.first_row = rows;
        f.first_frame = frames;
        f.out_off = outv;
        f.n_frames = e->n_frames;
        f.channels = e->channels;
        f.pad = 0;
        const uint64_t r = e->n_frames * e->channels;
        rows += r;
        frames += e->n_frames;
        outv += (e->n_frames + 1) * kHop * e->channels;
        npairs += e->n_frames ? e->pair_offset[r] : 0;
        nraw += e->n_frames ? e->raw_offset[e->n_frames] : 0;
    }
    // ---- everything this call owns, released on every exit path ----
    std::vector<void *> pinned_tmp, dev_tmp;
    std::vector<float *> h_out(n_files, nullptr);
    auto cleanup = [&](bool keep_outputs) {
        for (void *p : pinned_tmp)
            c->pool.release(p);
        for (void *p : dev_tmp)
            dfree(p, cs);
        if (!keep_outputs)
            for (float *p : h_out)
                if (p)
                    glc_free(c, p);
    };
    auto pin = [&](size_t bytes) -> void * {
        void *p = c->pool.alloc(bytes);
        if (p)
            pinned_tmp.push_back(p);
        return p;
    };
#define DEC_TRY(expr)                                                                                      \
    do                                                                                                     \
    {                                                                                                      \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
        {                                                                                                  \
            cleanup(false);                                                                                \
            return fail(GLC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                        __LINE__);                                                                         \
        }                                                                                                  \
    } while (0)

    // ---- small per-row / per-frame metadata: concatenate into pinned buffers, one H2D each ----
    uint8_t *h_is_raw = (uint8_t *)pin(frames);
    uint64_t *h_pair_off = (uint64_t *)pin((rows + 1) * 8);
    float *h_scales = (float *)pin(rows * 4);
    uint64_t *h_raw_off = (uint64_t *)pin((frames + 1) * 8);
    if (!h_is_raw || !h_pair_off || !h_scales || !h_raw_off)
    {
        cleanup(false);
        return fail(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
    }
    {
        uint64_t pb = 0, rb = 0;
        for (uint32_t i = 0; i < n_files; ++i)
        {
            const glc_encoded *e = enc[i];
            const DecFileDesc &f = files[i];
            const uint64_t r = e->n_frames * e->channels;
            if (e->n_frames == 0)
                continue;
            memcpy(h_is_raw + f.first_frame, e->frame_is_raw, e->n_frames);
            memcpy(h_scales + f.first_row, e->scales, r * 4);
            for (uint64_t k = 0; k < r; ++k)
            {
                // nnz is authoritative for the per-row count (pair_offset may come from a foreign producer)
                h_pair_off[f.first_row + k] = pb + e->pair_offset[k];
                if (e->pair_offset[k + 1] < e->pair_offset[k] ||
                    e->pair_offset[k + 1] - e->pair_offset[k] != e->nnz[k])
                {
                    cleanup(false);
                    return fail(GLC_ERR_CORRUPT, "stream %u: pair_offset/nnz mismatch at row %llu", i,
                                (unsigned long long)k);
                }
            }
            for (uint64_t k = 0; k < e->n_frames; ++k)
            {
                h_raw_off[f.first_frame + k] = rb + e->raw_offset[k];
                const uint64_t rl = e->raw_offset[k + 1] - e->raw_offset[k];
                // raw_pcm = Some(v) may hold any number of values, also none (the reference reads what is
                // there and leaves the rest of the block at zero, src/codec.rs:633-640); None holds nothing
                if (e->raw_offset[k + 1] < e->raw_offset[k] || (e->frame_is_raw[k] == 0 && rl != 0))
                {
                    cleanup(false);
                    return fail(GLC_ERR_CORRUPT, "stream %u: frame %llu raw flag/length mismatch", i,
                                (unsigned long long)k);
                }
            }
            pb += e->pair_offset[r];
            rb += e->raw_offset[e->n_frames];
        }
        h_pair_off[rows] = pb;
        h_raw_off[frames] = rb;
    }
    uint8_t *d_is_raw = nullptr;
    uint64_t *d_pair_off = nullptr, *d_raw_off = nullptr;
    glc_pair *d_pairs = nullptr;
    float *d_scales = nullptr;
    int16_t *d_raw = nullptr;
    auto dev = [&](auto **pp, size_t count) -> cudaError_t {
        cudaError_t e = dmalloc(pp, count, cs);
        if (e == cudaSuccess)
            dev_tmp.push_back((void *)*pp);
        return e;
    };
    DEC_TRY(dev(&d_is_raw, frames));
    DEC_TRY(dev(&d_pair_off, rows + 1));
    DEC_TRY(dev(&d_scales, rows));
    DEC_TRY(dev(&d_raw_off, frames + 1));
    DEC_TRY(dev(&d_pairs, npairs));
    DEC_TRY(dev(&d_raw, nraw));
    DEC_TRY(cudaMemcpyAsync(d_is_raw, h_is_raw, frames, cudaMemcpyHostToDevice, cs));
    DEC_TRY(cudaMemcpyAsync(d_pair_off, h_pair_off, (rows + 1) * 8, cudaMemcpyHostToDevice, cs));
    DEC_TRY(cudaMemcpyAsync(d_scales, h_scales, rows * 4, cudaMemcpyHostToDevice, cs));
    DEC_TRY(cudaMemcpyAsync(d_raw_off, h_raw_off, (frames + 1) * 8, cudaMemcpyHostToDevice, cs));
    c->stats.h2d_bytes += frames + (rows + 1) * 8 + rows * 4 + (frames + 1) * 8;

    // ---- outputs: the gapless trim is a window on each file's untrimmed stream ----
    std::vector<uint64_t> win_off(n_files), win_len(n_files);
    for (uint32_t i = 0; i < n_files; ++i)
    {
        const glc_encoded *e = enc[i];
        const uint64_t untrimmed = (e->n_frames + 1) * kHop * e->channels;
        win_off[i] = 0;
        win_len[i] = untrimmed;
        if (trim)
            trim_window(untrimmed, e->encoder_delay, e->original_length, &win_off[i], &win_len[i]);
    }
    if (!sink)
    {
        std::vector<uint64_t> bytes(n_files);
        for (uint32_t i = 0; i < n_files; ++i)
            bytes[i] = win_len[i] * (out16 ? 2 : 4);
        if (!slab_alloc(c, n_files, bytes.data(), (void **)h_out.data()))
        {
            cleanup(false);
            return fail(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
        }
    }
    DecodeHostIO io{};
    io.enc = enc;
    io.h_pair_off = h_pair_off;
    io.h_raw_off = h_raw_off;
    io.h_out = sink ? nullptr : h_out.data();
    io.out16 = out16;
    io.win_off = win_off.data();
    io.win_len = win_len.data();
    io.d_pairs = d_pairs;
    io.d_raw = d_raw;
    std::vector<char> enc_pinned(n_files);
    for (uint32_t i = 0; i < n_files; ++i)
        enc_pinned[i] = (enc[i]->n_frames == 0 || ((!enc[i]->pairs || host_ptr_is_pinned(c, enc[i]->pairs)) &&
                                                    (!enc[i]->raw || host_ptr_is_pinned(c, enc[i]->raw))))
                            ? 1
                            : 0;
    io.pinned = enc_pinned.data();
    float *d_out = nullptr;
    glc_status st = decode_core(c, files, rows, frames, outv, d_is_raw, d_pair_off, d_pairs, d_scales, d_raw_off,
                                d_raw, &io, &d_out);
    cudaError_t se = cudaStreamSynchronize(c->d2h);
    if (se == cudaSuccess)
        se = cudaStreamSynchronize(cs);
    if (se == cudaSuccess)
        se = cudaStreamSynchronize(c->copy);
    if (st == GLC_OK && se != cudaSuccess)
        st = fail(GLC_ERR_CUDA, "decode failed: %s", cudaGetErrorString(se));
    if (sink && st == GLC_OK)
    {
        sink->d_out = d_out;
        sink->off.resize(n_files);
        sink->len.resize(n_files);
        for (uint32_t i = 0; i < n_files; ++i)
        {
            sink->off[i] = files[i].out_off + win_off[i];
            sink->len[i] = win_len[i];
        }
        d_out = nullptr;
    }
    if (d_out)
        dfree(d_out, cs);
    cleanup(st == GLC_OK);
    if (st == GLC_OK && !sink)
        for (uint32_t i = 0; i < n_files; ++i)
        {
            pcm[i] = h_out[i];
            n_out[i] = win_len[i];
        }
    return st;
#undef DEC_TRY
}

// The CLI's `glc -d file.glc --flac-level N` path (src/main.rs:55-113): Decoder::decode, then
// flac::export_to_flac_with_level on the decoded samples -- here the PCM never leaves HBM.
extern "C" glc_status glc_decode_to_flac_batch(glc_decoder *dec, uint32_t n_files, const glc_encoded *const *enc,
                                               uint8_t level, uint8_t **bytes, uint64_t *len)
{
    if (!dec || !enc || !bytes || !len || n_files == 0)
        return fail(GLC_ERR_INVALID_ARG, "null/empty argument");
    glc_ctx *c = dec->ctx;
    DeviceSink sink;
    GLC_TRY(decode_batch_impl(dec, n_files, enc, true, nullptr, nullptr, &sink));
    std::vector<uint32_t> rates(n_files);
    std::vector<uint16_t> chs(n_files);
    for (uint32_t i = 0; i < n_files; ++i)
    {
        rates[i] = enc[i]->sample_rate; // the stream header wins (src/codec.rs:598; main.rs:62-66)
        chs[i] = enc[i]->channels;
    }
    glc_status st = flac_encode_impl(c, n_files, nullptr, sink.d_out, sink.off.data(), sink.len.data(), rates.data(),
                                     chs.data(), level, bytes, len);
    cudaStreamSynchronize(c->compute);
    dfree(sink.d_out, c->compute);
    return st;
}

extern "C" glc_status glc_decode_to_flac(glc_decoder *dec, const glc_encoded *enc, uint8_t level, uint8_t **bytes,
                                         uint64_t *len)
{
    return glc_decode_to_flac_batch(dec, 1, &enc, level, bytes, len);
}

// The CLI's default `glc -d file.glc` output (src/main.rs:95-105): Decoder::decode, then audio::export_to_wav,
// whose first step is convert_f32_to_i16 (src/audio.rs:11-16).  The conversion runs on the device and
// 16-bit samples cross PCIe: half the D2H bytes of glc_decode.
extern "C" glc_status glc_decode_batch_i16(glc_decoder *dec, uint32_t n_files, const glc_encoded *const *enc,
                                           int16_t **pcm, uint64_t *n_samples)
{
    return decode_batch_impl(dec, n_files, enc, true, reinterpret_cast<float **>(pcm), n_samples, nullptr, true);
}

extern "C" glc_status glc_decode_i16(glc_decoder *dec, const glc_encoded *enc, int16_t **pcm, uint64_t *n_samples)
{
    return decode_batch_impl(dec, 1, &enc, true, reinterpret_cast<float **>(pcm), n_samples, nullptr, true);
}

extern "C" glc_status glc_decode_batch(glc_decoder *dec, uint32_t n_files, const glc_encoded *const *enc,
                                       float **pcm, uint64_t *n_samples)
{
    return decode_batch_impl(dec, n_files, enc, true, pcm, n_samples);
}

extern "C" glc_status glc_decode(glc_decoder *dec, const glc_encoded *enc, float **pcm, uint64_t *n_samples)
{
    return decode_batch_impl(dec, 1, &enc, true, pcm, n_samples);
}

extern "C" glc_status glc_decode_untrimmed(glc_decoder *dec, const glc_encoded *enc, float **pcm,
                                           uint64_t *n_samples)
{
    return decode_batch_impl(dec, 1, &enc, false, pcm, n_samples);
}

extern "C" glc_status glc_dev_decode(glc_decoder *dec, const glc_dev_encoded *de, glc_dev_pcm **out)
{
    if (!dec || !de || !out)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    if (de->files.size() != 1)
        return fail(GLC_ERR_UNSUPPORTED, "glc_dev_decode handles single-stream objects");
    glc_ctx *c = dec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    const FileDesc &fd = de->files[0];
    std::vector<DecFileDesc> files(1);
    files[0].first_row = 0;
    files[0].first_frame = 0;
    files[0].out_off = 0;
    files[0].n_frames = fd.n_frames;
    files[0].channels = fd.channels;
    files[0].pad = 0;
    const uint64_t untrimmed = ((uint64_t)fd.n_frames + 1) * kHop * fd.channels;
    float *d_out = nullptr;
    GLC_TRY(decode_core(c, files, de->n_rows, de->n_frames, untrimmed, de->d_is_raw, de->d_pair_off, de->d_pairs,
                        de->d_scales, de->d_raw_off, de->d_raw, nullptr, &d_out));
    glc_dev_pcm *p = new glc_dev_pcm();
    p->ctx = c;
    p->d = d_out;
    p->n_alloc = untrimmed;
    p->channels = (uint16_t)fd.channels;
    trim_window(untrimmed, kHop / 2, fd.len * fd.channels, &p->trim_off, &p->n);
    *out = p;
    return GLC_OK;
}

extern "C" glc_status glc_dev_pcm_download(const glc_dev_pcm *p, float **pcm, uint64_t *n_samples)
{
    if (!p || !pcm || !n_samples)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    glc_ctx *c = p->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    float *h = (float *)c->pool.alloc(p->n * 4);
    if (!h)
        return fail(GLC_ERR_NO_MEMORY, "pinned host allocation failed");
    CUDA_TRY(cudaMemcpyAsync(h, p->d + p->trim_off, p->n * 4, cudaMemcpyDeviceToHost, c->compute));
    CUDA_TRY(cudaStreamSynchronize(c->compute));
    c->stats.d2h_bytes += p->n * 4;
    *pcm = h;
    *n_samples = p->n;
    return GLC_OK;
}

// ------------------------------------------------------- streaming decode
//
// Decoder::decode_streaming (src/codec.rs:595-741) hands out chunks of exactly 500 frames while a
// background thread keeps decoding; here every _next decodes just the frames of its chunk on the
// device (plus the one frame before it, whose second half is the chunk's first overlap), so the
// first audio is available after one chunk's worth of work and the host holds one chunk at a time.

struct glc_stream
{
    glc_decoder *dec;
    const glc_encoded *enc; // borrowed: must stay alive until _close (the reference shares an Arc)
    float *chunk;           // pinned memory of the current chunk (from glc_decode_untrimmed)
    uint64_t next_frame;    // frames already handed out
    bool finished;
    std::vector<uint64_t> pair_off, raw_off; // rebased offsets of the sub-range being decoded
};

extern "C" glc_status glc_decode_stream_open(glc_decoder *dec, const glc_encoded *enc, glc_stream **out)
{
    if (!dec || !enc || !out)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    GLC_TRY(validate_encoded(enc, 0));
    glc_stream *s = new glc_stream();
    s->dec = dec;
    s->enc = enc;
    s->chunk = nullptr;
    s->next_frame = 0;
    s->finished = false;
    *out = s;
    return GLC_OK;
}

extern "C" glc_status glc_decode_stream_next(glc_stream *s, const float **samples, uint64_t *n_samples,
                                             int *is_last, float *progress_percent)
{
    if (!s || !samples || !n_samples || !is_last)
        return fail(GLC_ERR_INVALID_ARG, "null argument");
    if (s->finished)
        return fail(GLC_ERR_INVALID_ARG, "stream already delivered its last chunk");
    const glc_encoded *e = s->enc;
    const uint64_t ch = e->channels, per_frame = (uint64_t)kHop * ch;
    const uint64_t remaining = e->n_frames - s->next_frame;
    const bool last = remaining < GLC_FRAMES_PER_CHUNK;
    const uint64_t a = s->next_frame, b = last ? e->n_frames : a + GLC_FRAMES_PER_CHUNK;
    const uint64_t a0 = a ? a - 1 : 0; // one frame of history for the first overlap
    if (s->chunk)
    {
        glc_free(s->dec->ctx, s->chunk);
        s->chunk = nullptr;
    }
    // view of frames [a0, b) as a stream of its own
    glc_encoded sub = *e;
    sub.n_frames = b - a0;
    sub.frame_is_raw = e->frame_is_raw + a0;
    sub.nnz = e->nnz + a0 * ch;
    sub.scales = e->scales + a0 * ch;
    const uint64_t rows = (b - a0) * ch;
    uint64_t pbase = 0, rbase = 0;
    if (e->n_frames)
    {
        pbase = e->pair_offset[a0 * ch];
        rbase = e->raw_offset[a0];
    }
    s->pair_off.resize(rows + 1);
    s->raw_off.resize(b - a0 + 1);
    for (uint64_t r = 0; r <= rows; ++r)
        s->pair_off[r] = e->n_frames ? e->pair_offset[a0 * ch + r] - pbase : 0;
    for (uint64_t f = 0; f <= b - a0; ++f)
        s->raw_off[f] = e->n_frames ? e->raw_offset[a0 + f] - rbase : 0;
    sub.pair_offset = s->pair_off.data();
    sub.raw_offset = s->raw_off.data();
    sub.pairs = e->pairs + pbase;
    sub.raw = e->raw + rbase;
    float *pcm = nullptr;
    uint64_t n = 0;
    GLC_TRY(glc_decode_untrimmed(s->dec, &sub, &pcm, &n)); // hops of frames a0..b-1, then the final overlap
    s->chunk = pcm;
    const uint64_t skip = (a - a0) * per_frame; // the history frame's own hop was delivered with the previous chunk
    *samples = pcm + skip;
    if (!last)
    {
        // a full chunk is flushed as soon as 500 frames are buffered (src/codec.rs:708-717);
        // `idx` there is the index of the frame that completed the chunk
        *n_samples = GLC_FRAMES_PER_CHUNK * per_frame;
        *is_last = 0;
        const uint64_t idx = a + GLC_FRAMES_PER_CHUNK - 1;
        if (progress_percent)
            *progress_percent = (float)idx / (float)e->n_frames * 100.0f;
        s->next_frame = b;
    }
    else
    {
        *n_samples = (remaining + 1) * per_frame; // leftover frames + final overlap (:723-732)
        *is_last = 1;
        if (progress_percent)
            *progress_percent = -1.0f;
        s->next_frame = e->n_frames;
        s->finished = true;
    }
    return GLC_OK;
}

extern "C" void glc_decode_stream_close(glc_stream *s)
{
    if (!s)
        return;
    if (s->chunk)
        glc_free(s->dec->ctx, s->chunk);
    delete s;
}
