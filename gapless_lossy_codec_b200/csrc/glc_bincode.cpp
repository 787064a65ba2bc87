// glc_bincode.cpp -- the `.glc` container image: bincode 1.3 serialisation of EncodedAudio, as
// written by save_encoded / read by load_encoded (reference src/codec.rs:774-786).
//
// bincode is an external crate (Cargo.toml:16, `bincode = "1.3"`, un-vendored).  Its default 1.x
// wire format is little-endian with fixed-width integers: a u64 element count in front of every
// Vec, a one-byte 0/1 tag in front of every Option, and struct/tuple fields back to back in
// declaration order.  For the derives at src/codec.rs:31-69 that gives
//   header{u32 rate, u16 channels, u64 total} , u64 n_frames , frames... , gapless{u32,u32,u64}
//   frame = Vec<Vec<(u16,i16)>> , Vec<f32> , Option<Vec<i16>>
// A raw frame carries two empty Vecs and Some(raw); a sparse frame carries None.
#include <stdlib.h>
#include <algorithm>
#include <string.h>

#include <new>

#include "glc_internal.cuh"

using namespace glc;

namespace
{
struct Sink
{
    uint8_t *p = nullptr;
    uint64_t n = 0, cap = 0;
    bool ok = true;
    void put(const void *src, uint64_t len)
    {
        if (!ok)
            return;
        if (n + len > cap)
        {
            uint64_t nc = cap ? cap : 1 << 16;
            while (nc < n + len)
                nc *= 2;
            uint8_t *q = (uint8_t *)realloc(p, nc);
            if (!q)
            {
                ok = false;
                return;
            }
            p = q;
            cap = nc;
        }
        memcpy(p + n, src, len);
        n += len;
    }
    template <typename T> void le(T v) { put(&v, sizeof v); } // x86-64 / little-endian hosts only
};

struct Source
{
    const uint8_t *p;
    uint64_t n, pos = 0;
    bool ok = true;
    bool get(void *dst, uint64_t len)
    {
        if (!ok || len > n - pos)
        {
            ok = false;
            return false;
        }
        memcpy(dst, p + pos, len);
        pos += len;
        return true;
    }
    template <typename T> T le()
    {
        T v{};
        get(&v, sizeof v);
        return v;
    }
};
} // namespace

extern "C" glc_status glc_encoded_to_bincode(glc_ctx *ctx, const glc_encoded *e, uint8_t **bytes, uint64_t *len)
{
    (void)ctx;
    if (!e || !bytes || !len)
        return set_error(GLC_ERR_INVALID_ARG, "null argument");
    Sink w;
    const uint64_t ch = e->channels;
    w.le<uint32_t>(e->sample_rate);
    w.le<uint16_t>(e->channels);
    w.le<uint64_t>(e->total_samples);
    w.le<uint64_t>(e->n_frames);
    for (uint64_t f = 0; f < e->n_frames; ++f)
    {
        if (e->frame_is_raw[f])
        {
            w.le<uint64_t>(0); // sparse_coeffs_per_channel: Vec::new()
            w.le<uint64_t>(0); // scale_factors: Vec::new()
            w.le<uint8_t>(1);  // Some(raw_pcm)
            const uint64_t cnt = e->raw_offset[f + 1] - e->raw_offset[f];
            w.le<uint64_t>(cnt);
            w.put(e->raw + e->raw_offset[f], cnt * sizeof(int16_t));
        }
        else
        {
            w.le<uint64_t>(ch);
            for (uint64_t c = 0; c < ch; ++c)
            {
                const uint64_t row = f * ch + c;
                w.le<uint64_t>(e->nnz[row]);
                w.put(e->pairs + e->pair_offset[row], (uint64_t)e->nnz[row] * sizeof(glc_pair));
            }
            w.le<uint64_t>(ch);
            w.put(e->scales + f * ch, ch * sizeof(float));
            w.le<uint8_t>(0); // None
        }
    }
    w.le<uint32_t>(e->encoder_delay);
    w.le<uint32_t>(e->padding);
    w.le<uint64_t>(e->original_length);
    if (!w.ok)
    {
        free(w.p);
        return set_error(GLC_ERR_NO_MEMORY, "out of host memory");
    }
    *bytes = w.p; // released with glc_free (falls through to free())
    *len = w.n;
    return GLC_OK;
}

extern "C" glc_status glc_encoded_from_bincode(glc_ctx *ctx, const uint8_t *bytes, uint64_t len, glc_encoded **out)
{
    if (!bytes || !out)
        return set_error(GLC_ERR_INVALID_ARG, "null argument");
    Source r{bytes, len};
    EncodedBox *box = new (std::nothrow) EncodedBox();
    EncodedBlock *blk = new (std::nothrow) EncodedBlock();
    if (!box || !blk)
    {
        delete box; // whichever of the two was allocated
        delete blk;
        return set_error(GLC_ERR_NO_MEMORY, "out of host memory");
    }
    blk->ctx = ctx;
    blk->refs = 1;
    box->blk = blk;
    glc_encoded &e = box->pub;
    memset(&e, 0, sizeof e);
    auto bail = [&](glc_status st, const char *msg) {
        for (void *p : blk->heap)
            free(p);
        delete blk;
        delete box;
        return set_error(st, "%s", msg);
    };
    e.sample_rate = r.le<uint32_t>();
    e.channels = r.le<uint16_t>();
    e.total_samples = r.le<uint64_t>();
    e.n_frames = r.le<uint64_t>();
    // Sizes come from an untrusted header: a frame serialises to at least 17 bytes (two u64 lengths and the
    // Option tag) after the 22-byte header and before the 16-byte gapless block, and a sparse frame to at
    // least 12 bytes per channel, so both counts are bounded by the image size before anything is allocated.
    // (Raw frames of a foreign producer may carry fewer values than channels; the per-row tables are
    // capped at 64 bytes per image byte, 64 MiB for small images.)
    if (!r.ok || e.channels == 0 || len < 38 || e.n_frames > (len - 38) / 17)
        return bail(GLC_ERR_CORRUPT, "truncated or invalid .glc header");
    const uint64_t ch = e.channels, rows = e.n_frames * ch;
    if (rows * 16 > std::max<uint64_t>((uint64_t)64 << 20, 64 * len))
        return bail(GLC_ERR_CORRUPT, ".glc header claims more frame-channels than the image can hold");
    auto heap = [&](uint64_t bytes_) -> void * {
        void *p = calloc(bytes_ ? bytes_ : 1, 1);
        if (p)
            blk->heap.push_back(p);
        return p;
    };
    e.frame_is_raw = (uint8_t *)heap(e.n_frames);
    e.nnz = (uint32_t *)heap(rows * 4);
    e.pair_offset = (uint64_t *)heap((rows + 1) * 8);
    e.scales = (float *)heap(rows * 4);
    e.raw_offset = (uint64_t *)heap((e.n_frames + 1) * 8);
    e.pairs = (glc_pair *)heap(len); // payloads cannot exceed the file size
    e.raw = (int16_t *)heap(len);
    if (!e.frame_is_raw || !e.nnz || !e.pair_offset || !e.scales || !e.raw_offset || !e.pairs || !e.raw)
        return bail(GLC_ERR_NO_MEMORY, "out of host memory");
    uint64_t np = 0, nr = 0;
    for (uint64_t f = 0; f < e.n_frames; ++f)
    {
        const uint64_t n_vecs = r.le<uint64_t>();
        if (!r.ok || (n_vecs != 0 && n_vecs != ch))
            return bail(GLC_ERR_CORRUPT, "frame with an unexpected number of channel vectors");
        for (uint64_t c = 0; c < ch; ++c)
        {
            const uint64_t row = f * ch + c;
            e.pair_offset[row] = np;
            if (c < n_vecs)
            {
                const uint64_t cnt = r.le<uint64_t>();
                if (!r.ok || cnt > (len - r.pos) / sizeof(glc_pair))
                    return bail(GLC_ERR_CORRUPT, "truncated coefficient vector");
                e.nnz[row] = (uint32_t)cnt;
                r.get(e.pairs + np, cnt * sizeof(glc_pair));
                np += cnt;
            }
        }
        const uint64_t n_scales = r.le<uint64_t>();
        if (!r.ok || n_scales != n_vecs)
            return bail(GLC_ERR_CORRUPT, "scale vector length differs from channel vectors");
        r.get(e.scales + f * ch, n_scales * sizeof(float));
        const uint8_t tag = r.le<uint8_t>();
        e.raw_offset[f] = nr;
        if (tag == 1)
        {
            const uint64_t cnt = r.le<uint64_t>();
            if (!r.ok || cnt > (len - r.pos) / sizeof(int16_t) || n_vecs != 0)
                return bail(GLC_ERR_CORRUPT, "bad raw frame");
            r.get(e.raw + nr, cnt * sizeof(int16_t));
            nr += cnt;
            e.frame_is_raw[f] = 1;
        }
        else if (tag != 0 || n_vecs == 0 || !r.ok)
            return bail(GLC_ERR_CORRUPT, "bad Option tag / empty sparse frame");
    }
    e.pair_offset[rows] = np;
    e.raw_offset[e.n_frames] = nr;
    e.encoder_delay = r.le<uint32_t>();
    e.padding = r.le<uint32_t>();
    e.original_length = r.le<uint64_t>();
    if (!r.ok || r.pos != len)
        return bail(GLC_ERR_CORRUPT, "truncated or oversized .glc image");
    *out = &box->pub;
    return GLC_OK;
}
