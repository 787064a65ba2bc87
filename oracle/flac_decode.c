/*
 * oracle/flac_decode.c -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * Independent FLAC *decoder* written from RFC 9639 (the spec src/flac.rs:1 cites),
 * standing in for the `claxon` crate the reference's tests/test_flac.rs uses to
 * read streams back.  Subset: STREAMINFO (+ skipping other metadata), fixed-size
 * or variable-size frames, CONSTANT / VERBATIM / FIXED(0..4) subframes, wasted
 * bits, partitioned Rice with 4- or 5-bit parameters including the escape code,
 * independent channels only.  Verifies every frame's CRC-8 and CRC-16 and the
 * STREAMINFO MD5.  Returns 0 on success, a distinct non-zero code per failure.
 */
#include "oracle_common.h"

#include <stdlib.h>
#include <string.h>

typedef struct
{
    const uint8_t *p;
    uint64_t len;
    uint64_t pos; /* in bits */
    int err;
} br_t;

static uint64_t br_bits(br_t *r, unsigned n)
{
    uint64_t v = 0;
    for (unsigned i = 0; i < n; ++i)
    {
        if ((r->pos >> 3) >= r->len)
        {
            r->err = 1;
            return 0;
        }
        unsigned bit = (r->p[r->pos >> 3] >> (7 - (r->pos & 7))) & 1u;
        v = (v << 1) | bit;
        r->pos++;
    }
    return v;
}

static int64_t br_sbits(br_t *r, unsigned n)
{
    if (n == 0)
        return 0;
    uint64_t v = br_bits(r, n);
    if (v & (1ull << (n - 1)))
        return (int64_t)v - ((int64_t)1 << n);
    return (int64_t)v;
}

static uint32_t br_unary(br_t *r)
{
    uint32_t z = 0;
    while (!r->err && br_bits(r, 1) == 0)
        ++z;
    return z;
}

static int decode_residual(br_t *r, int32_t *out, uint32_t bs, int order)
{
    unsigned method = (unsigned)br_bits(r, 2);
    if (method > 1)
        return 20;
    unsigned pbits = method == 0 ? 4 : 5;
    unsigned esc = method == 0 ? 15 : 31;
    unsigned po = (unsigned)br_bits(r, 4);
    uint32_t nparts = 1u << po;
    if ((bs >> po) << po != bs && po != 0)
        return 21;
    uint32_t idx = (uint32_t)order;
    for (uint32_t p = 0; p < nparts; ++p)
    {
        uint32_t cnt = bs >> po;
        if (p == 0)
        {
            if (cnt < (uint32_t)order)
                return 22;
            cnt -= (uint32_t)order;
        }
        unsigned k = (unsigned)br_bits(r, pbits);
        if (k == esc)
        {
            unsigned nb = (unsigned)br_bits(r, 5);
            for (uint32_t i = 0; i < cnt; ++i)
                out[idx++] = (int32_t)br_sbits(r, nb);
        }
        else
        {
            for (uint32_t i = 0; i < cnt; ++i)
            {
                uint32_t msb = br_unary(r);
                uint32_t lsb = k ? (uint32_t)br_bits(r, k) : 0;
                uint32_t folded = (msb << k) | lsb;
                out[idx++] = (folded & 1u) ? -(int32_t)(folded >> 1) - 1 : (int32_t)(folded >> 1);
                if (r->err)
                    return 23;
            }
        }
    }
    return r->err ? 23 : 0;
}

static int decode_subframe(br_t *r, int32_t *out, uint32_t bs, unsigned bps)
{
    if (br_bits(r, 1) != 0)
        return 30;
    unsigned type = (unsigned)br_bits(r, 6);
    unsigned wasted = 0;
    if (br_bits(r, 1))
        wasted = br_unary(r) + 1;
    bps -= wasted;
    if (type == 0)
    {
        int32_t v = (int32_t)br_sbits(r, bps);
        for (uint32_t i = 0; i < bs; ++i)
            out[i] = v;
    }
    else if (type == 1)
    {
        for (uint32_t i = 0; i < bs; ++i)
            out[i] = (int32_t)br_sbits(r, bps);
    }
    else if (type >= 8 && type <= 12)
    {
        int order = (int)type - 8;
        if ((uint32_t)order > bs)
            return 31;
        for (int i = 0; i < order; ++i)
            out[i] = (int32_t)br_sbits(r, bps);
        int rc = decode_residual(r, out, bs, order);
        if (rc)
            return rc;
        for (uint32_t i = (uint32_t)order; i < bs; ++i)
        {
            int64_t pred = 0;
            switch (order)
            {
            case 1: pred = out[i - 1]; break;
            case 2: pred = 2 * (int64_t)out[i - 1] - out[i - 2]; break;
            case 3: pred = 3 * (int64_t)out[i - 1] - 3 * (int64_t)out[i - 2] + out[i - 3]; break;
            case 4:
                pred = 4 * (int64_t)out[i - 1] - 6 * (int64_t)out[i - 2] + 4 * (int64_t)out[i - 3] -
                       out[i - 4];
                break;
            default: break;
            }
            out[i] = (int32_t)(out[i] + pred);
        }
    }
    else
        return 32; /* LPC / reserved: the reference never emits them */
    if (wasted)
        for (uint32_t i = 0; i < bs; ++i)
            out[i] = (int32_t)((uint32_t)out[i] << wasted);
    return r->err ? 33 : 0;
}

/* frame_off (nullable): receives the byte offset of every frame header, frame_off[n_frames] = len;
 * at most frame_cap entries are written. */
int orc_flac_decode_ex(const uint8_t *bytes, uint64_t len, orc_flac_info *info, uint64_t *frame_off,
                       uint64_t frame_cap)
{
    memset(info, 0, sizeof *info);
    if (len < 42 || memcmp(bytes, "fLaC", 4) != 0)
        return 1;
    br_t r = {bytes, len, 32, 0};
    int last = 0, have_si = 0;
    while (!last)
    {
        last = (int)br_bits(&r, 1);
        unsigned type = (unsigned)br_bits(&r, 7);
        uint32_t blen = (uint32_t)br_bits(&r, 24);
        if (r.err)
            return 2;
        if (type == 0)
        {
            if (blen != 34)
                return 3;
            info->min_block = (uint32_t)br_bits(&r, 16);
            info->max_block = (uint32_t)br_bits(&r, 16);
            br_bits(&r, 24);
            br_bits(&r, 24);
            info->sample_rate = (uint32_t)br_bits(&r, 20);
            info->channels = (uint32_t)br_bits(&r, 3) + 1;
            info->bits_per_sample = (uint32_t)br_bits(&r, 5) + 1;
            info->total_samples = br_bits(&r, 36);
            for (int i = 0; i < 16; ++i)
                info->md5[i] = (uint8_t)br_bits(&r, 8);
            have_si = 1;
        }
        else
            r.pos += (uint64_t)blen * 8;
        if (r.err)
            return 2;
    }
    if (!have_si)
        return 4;
    const uint32_t ch = info->channels;
    uint64_t cap = info->total_samples * ch;
    if (cap == 0)
        cap = 1;
    int32_t *outp = (int32_t *)malloc(sizeof(int32_t) * cap);
    int32_t *chan = (int32_t *)malloc(sizeof(int32_t) * 65536u * ch);
    uint64_t written = 0;
    int rc = 0;
    while ((r.pos >> 3) < len)
    {
        const uint64_t fstart = r.pos >> 3;
        if (frame_off && info->n_frames < frame_cap)
            frame_off[info->n_frames] = fstart;
        if (br_bits(&r, 14) != 0x3FFE) { rc = 10; break; }
        if (br_bits(&r, 1) != 0) { rc = 11; break; }
        br_bits(&r, 1); /* blocking strategy */
        unsigned bsb = (unsigned)br_bits(&r, 4);
        unsigned srb = (unsigned)br_bits(&r, 4);
        unsigned chb = (unsigned)br_bits(&r, 4);
        unsigned ssb = (unsigned)br_bits(&r, 3);
        if (br_bits(&r, 1) != 0) { rc = 12; break; }
        /* UTF-8 style coded number */
        unsigned b0 = (unsigned)br_bits(&r, 8);
        int extra = 0;
        if (b0 >= 0xFE) extra = 6;
        else if (b0 >= 0xFC) extra = 5;
        else if (b0 >= 0xF8) extra = 4;
        else if (b0 >= 0xF0) extra = 3;
        else if (b0 >= 0xE0) extra = 2;
        else if (b0 >= 0xC0) extra = 1;
        else if (b0 >= 0x80) { rc = 13; break; }
        uint64_t num = extra ? (b0 & (0x3Fu >> extra)) : b0;
        if (extra == 6) num = 0;
        for (int i = 0; i < extra; ++i)
        {
            unsigned b = (unsigned)br_bits(&r, 8);
            if ((b & 0xC0) != 0x80) { rc = 13; }
            num = (num << 6) | (b & 0x3F);
        }
        if (rc) break;
        if (num != info->n_frames) { rc = 14; break; } /* fixed-blocksize: frame number */
        uint32_t bs;
        switch (bsb)
        {
        case 0: rc = 15; bs = 0; break;
        case 1: bs = 192; break;
        case 2: case 3: case 4: case 5: bs = 576u << (bsb - 2); break;
        case 6: bs = (uint32_t)br_bits(&r, 8) + 1; break;
        case 7: bs = (uint32_t)br_bits(&r, 16) + 1; break;
        default: bs = 256u << (bsb - 8); break;
        }
        if (rc) break;
        if (srb == 12) br_bits(&r, 8);
        else if (srb == 13 || srb == 14) br_bits(&r, 16);
        else if (srb == 15) { rc = 16; break; }
        static const uint32_t rates[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000,
                                           32000, 44100, 48000, 96000};
        if (srb >= 1 && srb <= 11 && rates[srb] != info->sample_rate) { rc = 17; break; }
        unsigned bps = info->bits_per_sample;
        static const unsigned ss[8] = {0, 8, 12, 0, 16, 20, 24, 32};
        if (ssb != 0 && ss[ssb] != bps) { rc = 18; break; }
        if (chb >= 8) { rc = 19; break; } /* stereo decorrelation: reference never uses it */
        if (chb + 1 != ch) { rc = 19; break; }
        uint8_t c8 = orc_crc8(bytes + fstart, (r.pos >> 3) - fstart);
        if (br_bits(&r, 8) != c8) { rc = 40; break; }
        for (uint32_t c = 0; c < ch && !rc; ++c)
            rc = decode_subframe(&r, chan + (size_t)c * 65536u, bs, bps);
        if (rc) break;
        r.pos = (r.pos + 7) & ~(uint64_t)7;
        uint16_t c16 = orc_crc16(bytes + fstart, (r.pos >> 3) - fstart);
        if (br_bits(&r, 16) != c16 || r.err) { rc = 41; break; }
        if (written + (uint64_t)bs * ch > cap)
        {
            cap = (written + (uint64_t)bs * ch) * 2;
            outp = (int32_t *)realloc(outp, sizeof(int32_t) * cap);
        }
        for (uint32_t i = 0; i < bs; ++i)
            for (uint32_t c = 0; c < ch; ++c)
                outp[written++] = chan[(size_t)c * 65536u + i];
        info->n_frames++;
    }
    free(chan);
    if (frame_off && info->n_frames < frame_cap)
        frame_off[info->n_frames] = r.pos >> 3;
    info->samples = outp;
    info->n_decoded = written;
    if (rc)
        return rc;
    /* MD5 over little-endian samples of ceil(bps/8) bytes */
    unsigned bytes_ps = (info->bits_per_sample + 7) / 8;
    uint8_t *raw = (uint8_t *)malloc((size_t)(written ? written : 1) * bytes_ps);
    for (uint64_t i = 0; i < written; ++i)
        for (unsigned b = 0; b < bytes_ps; ++b)
            raw[i * bytes_ps + b] = (uint8_t)((uint32_t)outp[i] >> (8 * b));
    uint8_t dg[16];
    orc_md5(raw, written * bytes_ps, dg);
    free(raw);
    info->md5_ok = memcmp(dg, info->md5, 16) == 0;
    return 0;
}

int orc_flac_decode(const uint8_t *bytes, uint64_t len, orc_flac_info *info)
{
    return orc_flac_decode_ex(bytes, len, info, NULL, 0);
}
