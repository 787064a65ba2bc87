/*
 * oracle/bincode_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * Byte image of `bincode::serialize(&EncodedAudio)` as used by save_encoded /
 * load_encoded (src/codec.rs:774-786).  bincode is an un-vendored dependency
 * (Cargo.toml:16 `bincode = "1.3"`, no Cargo.lock in the reference); its 1.x
 * default configuration is: little-endian, fixed-width integers, u64 length prefix
 * for every Vec, one u8 tag (0/1) for Option, struct and tuple fields concatenated
 * in declaration order.  Applied to the derives at src/codec.rs:31-69:
 *
 *   u32 sample_rate | u16 channels | u64 total_samples
 *   u64 n_frames | per frame {
 *       u64 n_channel_vecs | per channel { u64 n | n x (u16 idx, i16 q) }
 *       u64 n_scales | n_scales x f32
 *       u8 tag | if tag==1 { u64 n | n x i16 }
 *   }
 *   u32 encoder_delay | u32 padding | u64 original_length
 *
 * Anchors: the encoder's own size estimate (src/codec.rs:505-515: 8 bytes per Vec
 * length, 4 per pair) and README.md:56-60 (7014 bytes for a 1 s stereo example).
 * "parity unpinned": no .glc file from the reference exists to compare with.
 */
#include "oracle_common.h"

#include <stdlib.h>
#include <string.h>

typedef struct
{
    uint8_t *b;
    uint64_t n, cap;
} wb_t;

static void wb_put(wb_t *w, const void *src, uint64_t len)
{
    if (w->n + len > w->cap)
    {
        uint64_t nc = w->cap ? w->cap * 2 : 4096;
        while (nc < w->n + len)
            nc *= 2;
        w->b = (uint8_t *)realloc(w->b, nc);
        w->cap = nc;
    }
    memcpy(w->b + w->n, src, len);
    w->n += len;
}
static void wb_u64(wb_t *w, uint64_t v) { wb_put(w, &v, 8); }
static void wb_u32(wb_t *w, uint32_t v) { wb_put(w, &v, 4); }
static void wb_u16(wb_t *w, uint16_t v) { wb_put(w, &v, 2); }
static void wb_u8(wb_t *w, uint8_t v) { wb_put(w, &v, 1); }

int orc_bincode_serialize(const orc_encoded *e, uint8_t **bytes, uint64_t *len)
{
    if (!e || !bytes || !len)
        return 1;
    wb_t w = {0, 0, 0};
    const uint32_t ch = e->channels;
    wb_u32(&w, e->sample_rate);
    wb_u16(&w, e->channels);
    wb_u64(&w, e->total_samples);
    wb_u64(&w, e->n_frames);
    for (uint64_t f = 0; f < e->n_frames; ++f)
    {
        if (e->frame_is_raw[f])
        {
            wb_u64(&w, 0);
            wb_u64(&w, 0);
            wb_u8(&w, 1);
            uint64_t n = e->raw_offset[f + 1] - e->raw_offset[f];
            wb_u64(&w, n);
            wb_put(&w, e->raw + e->raw_offset[f], n * 2);
        }
        else
        {
            wb_u64(&w, ch);
            for (uint32_t c = 0; c < ch; ++c)
            {
                uint64_t fc = f * ch + c;
                wb_u64(&w, e->nnz[fc]);
                wb_put(&w, e->pairs + e->pair_offset[fc], (uint64_t)e->nnz[fc] * 4);
            }
            wb_u64(&w, ch);
            wb_put(&w, e->scales + f * ch, (uint64_t)ch * 4);
            wb_u8(&w, 0);
        }
    }
    wb_u32(&w, e->encoder_delay);
    wb_u32(&w, e->padding);
    wb_u64(&w, e->original_length);
    *bytes = w.b;
    *len = w.n;
    return 0;
}

typedef struct
{
    const uint8_t *p;
    uint64_t n, pos;
    int err;
} rb_t;
static void rb_get(rb_t *r, void *dst, uint64_t len)
{
    if (r->pos + len > r->n)
    {
        r->err = 1;
        memset(dst, 0, len);
        return;
    }
    memcpy(dst, r->p + r->pos, len);
    r->pos += len;
}
static uint64_t rb_u64(rb_t *r) { uint64_t v; rb_get(r, &v, 8); return v; }
static uint32_t rb_u32(rb_t *r) { uint32_t v; rb_get(r, &v, 4); return v; }
static uint16_t rb_u16(rb_t *r) { uint16_t v; rb_get(r, &v, 2); return v; }
static uint8_t rb_u8(rb_t *r) { uint8_t v; rb_get(r, &v, 1); return v; }

/* Accepts the shapes the reference encoder produces: a frame is either sparse
 * (channels coefficient vecs + channels scales, tag 0) or raw (tag 1). */
int orc_bincode_deserialize(const uint8_t *bytes, uint64_t len, orc_encoded **out)
{
    rb_t r = {bytes, len, 0, 0};
    orc_encoded *e = (orc_encoded *)calloc(1, sizeof *e);
    e->sample_rate = rb_u32(&r);
    e->channels = rb_u16(&r);
    e->total_samples = rb_u64(&r);
    e->n_frames = rb_u64(&r);
    const uint32_t ch = e->channels;
    if (r.err || ch == 0 || e->n_frames > len)
    {
        free(e);
        return 2;
    }
    const uint64_t nfc = e->n_frames * ch;
    e->frame_is_raw = (uint8_t *)calloc(e->n_frames ? e->n_frames : 1, 1);
    e->nnz = (uint32_t *)calloc(nfc ? nfc : 1, 4);
    e->pair_offset = (uint64_t *)calloc(nfc + 1, 8);
    e->scales = (float *)calloc(nfc ? nfc : 1, 4);
    e->raw_offset = (uint64_t *)calloc(e->n_frames + 1, 8);
    e->pairs = (orc_pair *)malloc(len + 4);
    e->raw = (int16_t *)malloc(len + 2);
    uint64_t np = 0, nr = 0;
    int rc = 0;
    for (uint64_t f = 0; f < e->n_frames && !rc; ++f)
    {
        uint64_t ncv = rb_u64(&r);
        if (ncv != 0 && ncv != ch) { rc = 3; break; }
        for (uint64_t c = 0; c < ncv; ++c)
        {
            uint64_t n = rb_u64(&r);
            if (r.err || n * 4 > len - r.pos) { rc = 2; break; }
            e->nnz[f * ch + c] = (uint32_t)n;
            rb_get(&r, e->pairs + np, n * 4);
            e->pair_offset[f * ch + c] = np;
            np += n;
        }
        if (rc) break;
        if (ncv == 0)
            for (uint32_t c = 0; c < ch; ++c)
                e->pair_offset[f * ch + c] = np;
        uint64_t ns = rb_u64(&r);
        if (ns != ncv) { rc = 3; break; }
        rb_get(&r, e->scales + f * ch, ns * 4);
        uint8_t tag = rb_u8(&r);
        e->raw_offset[f] = nr;
        if (tag == 1)
        {
            uint64_t n = rb_u64(&r);
            if (r.err || n * 2 > len - r.pos) { rc = 2; break; }
            rb_get(&r, e->raw + nr, n * 2);
            nr += n;
            e->frame_is_raw[f] = 1;
            if (ncv != 0) { rc = 3; break; }
        }
        else if (tag != 0 || ncv == 0)
            rc = 3;
        if (r.err)
            rc = 2;
    }
    e->raw_offset[e->n_frames] = nr;
    e->pair_offset[nfc] = np;
    e->encoder_delay = rb_u32(&r);
    e->padding = rb_u32(&r);
    e->original_length = rb_u64(&r);
    if (!rc && (r.err || r.pos != len))
        rc = 2;
    if (rc)
    {
        orc_encoded_free(e);
        return rc;
    }
    *out = e;
    return 0;
}

/* FNV-1a/64 (test infrastructure: table fingerprint of the pinning kit, tests/golden/dump_reference.rs) */
uint64_t orc_fnv1a64(const uint8_t *data, uint64_t len)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < len; ++i)
    {
        h ^= data[i];
        h *= 0x100000001b3ull;
    }
    return h;
}
