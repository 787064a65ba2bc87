/*
 * oracle/flac_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * CPU restatement of the FLAC encoder of ajcm474/gapless-lossy-codec v0.5.0
 * (src/flac.rs): CRC-8/16 (:19-80), MD5 (:83-318), bit order (:340-424), UTF-8
 * number (:427-478), fixed predictor (:481-512), Rice parameter (:515-552), Rice
 * partition (:555-584), residual (:587-684), subframe (:687-745), frame
 * (:748-905), STREAMINFO (:908-944), driver (:947-1052).
 * "parity unpinned" (no reference bitstreams exist); anchored by RFC 9639
 * decodability + CRC/MD5 verification in flac_decode.c and hashlib in the tests.
 *
 * The reference writes one bit per call; only the resulting bit string matters, so
 * this writer appends whole fields MSB-first.
 */
#include "oracle_common.h"

#include <stdlib.h>
#include <string.h>

/* ----------------------------------------------------------------- CRCs */

uint8_t orc_crc8(const uint8_t *d, uint64_t len)
{
    uint8_t crc = 0;
    for (uint64_t i = 0; i < len; ++i)
    {
        crc ^= d[i];
        for (int b = 0; b < 8; ++b)
            crc = (crc & 0x80) ? (uint8_t)((crc << 1) ^ 0x07) : (uint8_t)(crc << 1);
    }
    return crc;
}

uint16_t orc_crc16(const uint8_t *d, uint64_t len)
{
    uint16_t crc = 0;
    for (uint64_t i = 0; i < len; ++i)
    {
        crc ^= (uint16_t)((uint16_t)d[i] << 8);
        for (int b = 0; b < 8; ++b)
            crc = (crc & 0x8000) ? (uint16_t)((crc << 1) ^ 0x8005) : (uint16_t)(crc << 1);
    }
    return crc;
}

/* ------------------------------------------------------------------ MD5 */
/* RFC 1321; the reference's MD5Context (:83-302) is the standard algorithm fed
 * two bytes at a time, so a one-shot digest over the same bytes is identical. */

static uint32_t rol32(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

static const uint32_t MD5_K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const int MD5_S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                              5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                              4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                              6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_block(uint32_t st[4], const uint8_t *p)
{
    uint32_t x[16];
    for (int i = 0; i < 16; ++i)
        x[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
               ((uint32_t)p[4 * i + 3] << 24);
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    for (int i = 0; i < 64; ++i)
    {
        uint32_t f;
        int g;
        if (i < 16)
        {
            f = (b & c) | (~b & d);
            g = i;
        }
        else if (i < 32)
        {
            f = (b & d) | (c & ~d);
            g = (5 * i + 1) & 15;
        }
        else if (i < 48)
        {
            f = b ^ c ^ d;
            g = (3 * i + 5) & 15;
        }
        else
        {
            f = c ^ (b | ~d);
            g = (7 * i) & 15;
        }
        uint32_t t = d;
        d = c;
        c = b;
        b = b + rol32(a + f + MD5_K[i] + x[g], MD5_S[i]);
        a = t;
    }
    st[0] += a;
    st[1] += b;
    st[2] += c;
    st[3] += d;
}

void orc_md5(const uint8_t *data, uint64_t len, uint8_t digest[16])
{
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    uint64_t full = len / 64;
    for (uint64_t i = 0; i < full; ++i)
        md5_block(st, data + 64 * i);
    uint8_t tail[128];
    uint64_t rem = len - 64 * full;
    memset(tail, 0, sizeof tail);
    if (rem)
        memcpy(tail, data + 64 * full, rem);
    tail[rem] = 0x80;
    uint64_t tl = (rem < 56) ? 64 : 128;
    uint64_t bits = len * 8;
    for (int i = 0; i < 8; ++i)
        tail[tl - 8 + i] = (uint8_t)(bits >> (8 * i));
    md5_block(st, tail);
    if (tl == 128)
        md5_block(st, tail + 64);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            digest[4 * i + j] = (uint8_t)(st[i] >> (8 * j));
}

/* ----------------------------------------------------------- bit writer */

typedef struct
{
    uint8_t *buf;
    uint64_t cap;
    uint64_t nbits;
} bw_t;

static void bw_reserve(bw_t *w, uint64_t extra_bits)
{
    uint64_t need = (w->nbits + extra_bits + 7) / 8 + 8;
    if (need > w->cap)
    {
        uint64_t nc = w->cap ? w->cap * 2 : 4096;
        while (nc < need)
            nc *= 2;
        w->buf = (uint8_t *)realloc(w->buf, nc);
        memset(w->buf + w->cap, 0, nc - w->cap);
        w->cap = nc;
    }
}

/* append the low `bits` bits of value, MSB first (src/flac.rs:340-380) */
static void bw_bits(bw_t *w, uint64_t value, unsigned bits)
{
    bw_reserve(w, bits);
    for (int b = (int)bits - 1; b >= 0; --b)
    {
        if ((value >> b) & 1u)
            w->buf[w->nbits >> 3] |= (uint8_t)(0x80u >> (w->nbits & 7));
        w->nbits++;
    }
}

static void bw_zeros(bw_t *w, uint64_t count)
{
    bw_reserve(w, count);
    w->nbits += count;
}

static void bw_align(bw_t *w) { w->nbits = (w->nbits + 7) & ~(uint64_t)7; }

/* src/flac.rs:427-478 */
static void bw_utf8(bw_t *w, uint64_t v)
{
    if (v < 0x80)
        bw_bits(w, v, 8);
    else if (v < 0x800)
    {
        bw_bits(w, 0xC0 | ((v >> 6) & 0x1F), 8);
        bw_bits(w, 0x80 | (v & 0x3F), 8);
    }
    else if (v < 0x10000)
    {
        bw_bits(w, 0xE0 | ((v >> 12) & 0x0F), 8);
        bw_bits(w, 0x80 | ((v >> 6) & 0x3F), 8);
        bw_bits(w, 0x80 | (v & 0x3F), 8);
    }
    else if (v < 0x200000)
    {
        bw_bits(w, 0xF0 | ((v >> 18) & 0x07), 8);
        bw_bits(w, 0x80 | ((v >> 12) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 6) & 0x3F), 8);
        bw_bits(w, 0x80 | (v & 0x3F), 8);
    }
    else if (v < 0x4000000)
    {
        bw_bits(w, 0xF8 | ((v >> 24) & 0x03), 8);
        bw_bits(w, 0x80 | ((v >> 18) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 12) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 6) & 0x3F), 8);
        bw_bits(w, 0x80 | (v & 0x3F), 8);
    }
    else if (v < 0x80000000ull)
    {
        bw_bits(w, 0xFC | ((v >> 30) & 0x01), 8);
        bw_bits(w, 0x80 | ((v >> 24) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 18) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 12) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 6) & 0x3F), 8);
        bw_bits(w, 0x80 | (v & 0x3F), 8);
    }
    else
    {
        bw_bits(w, 0xFE, 8);
        bw_bits(w, 0x80 | ((v >> 30) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 24) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 18) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 12) & 0x3F), 8);
        bw_bits(w, 0x80 | ((v >> 6) & 0x3F), 8);
        bw_bits(w, 0x80 | (v & 0x3F), 8);
    }
}

/* ---------------------------------------------------------- level tables */

static int predictor_order_for(int level, uint32_t bs)
{
    /* src/flac.rs:692-700 */
    switch (level)
    {
    case 0:
        return 0;
    case 1:
        return bs >= 1 ? 1 : 0;
    case 2:
        return bs >= 2 ? 2 : 0;
    case 3:
    case 4:
        return bs >= 3 ? 3 : 0;
    default:
        return bs >= 4 ? 4 : 0;
    }
}

static int partition_order_for(int level, uint32_t bs, int order)
{
    /* src/flac.rs:590-608 */
    int tz = 0;
    if (bs == 0)
        tz = 32;
    else
        while (!((bs >> tz) & 1u))
            ++tz;
    if (tz > 8)
        tz = 8;
    int cap = level == 0 ? 0 : (level <= 2 ? 2 : (level <= 5 ? 4 : 6));
    int po = cap < tz ? cap : tz;
    while (po > 0)
    {
        uint32_t ps = bs >> po;
        if (ps > (uint32_t)order && ps >= 4)
            break;
        --po;
    }
    return po;
}

static uint32_t rice_param_for(const int32_t *r, uint32_t n)
{
    /* src/flac.rs:515-552 */
    if (n == 0)
        return 0;
    uint64_t sum = 0;
    for (uint32_t i = 0; i < n; ++i)
        sum += (uint64_t)(r[i] < 0 ? -(int64_t)r[i] : (int64_t)r[i]);
    uint64_t mean = sum / n;
    if (mean == 0)
        return 0;
    uint32_t param = 0;
    uint64_t t = mean;
    while (t > 0 && param < 14)
    {
        t >>= 1;
        if (t > 0)
            ++param;
    }
    if (param > 0 && mean < (1ull << (param - 1)))
        --param;
    return param < 14 ? param : 14;
}

static void encode_subframe(bw_t *w, const int32_t *s, uint32_t bs, int level)
{
    const int order = predictor_order_for(level, bs);
    bw_bits(w, 0, 1);
    if (order == 0)
        bw_bits(w, 0x01, 6);
    else
        bw_bits(w, 0x08 | (unsigned)order, 6);
    bw_bits(w, 0, 1);
    if (order == 0)
    {
        for (uint32_t i = 0; i < bs; ++i)
            bw_bits(w, (uint64_t)(int64_t)s[i], 16);
        return;
    }
    for (int i = 0; i < order; ++i)
        bw_bits(w, (uint64_t)(int64_t)s[i], 16);
    /* residual, src/flac.rs:481-512 */
    int32_t *res = (int32_t *)malloc(sizeof(int32_t) * bs);
    for (uint32_t i = (uint32_t)order; i < bs; ++i)
    {
        int32_t pred;
        switch (order)
        {
        case 1:
            pred = s[i - 1];
            break;
        case 2:
            pred = 2 * s[i - 1] - s[i - 2];
            break;
        case 3:
            pred = 3 * s[i - 1] - 3 * s[i - 2] + s[i - 3];
            break;
        default:
            pred = 4 * s[i - 1] - 6 * s[i - 2] + 4 * s[i - 3] - s[i - 4];
            break;
        }
        res[i] = s[i] - pred;
    }
    const int po = partition_order_for(level, bs, order);
    bw_bits(w, 0, 2);
    bw_bits(w, (uint64_t)po, 4);
    const uint32_t nparts = 1u << po;
    const uint32_t dps = bs >> po;
    uint32_t idx = (uint32_t)order;
    for (uint32_t p = 0; p < nparts; ++p)
    {
        uint32_t cnt = (p == 0) ? dps - (uint32_t)order : dps;
        if (cnt == 0)
            continue;
        const int32_t *r = res + idx;
        idx += cnt;
        uint32_t k = rice_param_for(r, cnt);
        bw_bits(w, k, 4);
        for (uint32_t i = 0; i < cnt; ++i)
        {
            int32_t v = r[i];
            uint32_t folded = (v >= 0) ? ((uint32_t)v << 1) : ((((uint32_t)(-(v + 1))) << 1) | 1u);
            uint32_t msb = folded >> k;
            bw_zeros(w, msb);
            bw_bits(w, 1, 1);
            if (k)
                bw_bits(w, folded & ((1u << k) - 1u), k);
        }
    }
    free(res);
}

static void encode_frame(bw_t *w, const int16_t *smp, uint32_t ch, uint32_t rate,
                         uint32_t frame_no, uint32_t bs, int level)
{
    const uint64_t start_byte = w->nbits >> 3;
    bw_bits(w, 0x3FFE, 14);
    bw_bits(w, 0, 1);
    bw_bits(w, 0, 1);
    unsigned bsb;
    switch (bs)
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
    default: bsb = bs < 256 ? 6 : 7; break;
    }
    bw_bits(w, bsb, 4);
    unsigned srb;
    switch (rate)
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
    bw_bits(w, srb, 4);
    unsigned chb = ch == 1 ? 0 : (ch == 2 ? 1 : ch - 1);
    bw_bits(w, chb, 4);
    bw_bits(w, 4, 3); /* 16 bits per sample */
    bw_bits(w, 0, 1);
    bw_utf8(w, frame_no);
    if (bsb == 6)
        bw_bits(w, (bs - 1) & 0xFF, 8);
    else if (bsb == 7)
        bw_bits(w, (bs - 1) & 0xFFFF, 16);
    uint8_t c8 = orc_crc8(w->buf + start_byte, (w->nbits >> 3) - start_byte);
    bw_bits(w, c8, 8);

    int32_t *chan = (int32_t *)malloc(sizeof(int32_t) * bs);
    for (uint32_t c = 0; c < ch; ++c)
    {
        for (uint32_t i = 0; i < bs; ++i)
            chan[i] = smp[(uint64_t)i * ch + c];
        encode_subframe(w, chan, bs, level);
    }
    free(chan);
    bw_align(w);
    bw_reserve(w, 16);
    uint16_t c16 = orc_crc16(w->buf + start_byte, (w->nbits >> 3) - start_byte);
    bw_bits(w, c16, 16);
}

/* encode_flac_with_level, src/flac.rs:947-1052.  0 ok; 1 bad args; 4 fewer than
 * 16 samples per channel (:963-969); 5 level > 8 (:972-978). */
int orc_flac_encode(const float *pcm, uint64_t n, uint32_t rate, uint16_t channels, uint8_t level,
                    uint8_t **bytes, uint64_t *len)
{
    if (!bytes || !len || channels == 0)
        return 1;
    int16_t *s16 = (int16_t *)malloc(sizeof(int16_t) * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i)
    {
        float v = pcm[i] * 32767.0f;
        int16_t q;
        if (v != v)
            q = 0;
        else if (v <= -32768.0f)
            q = -32768;
        else if (v >= 32767.0f)
            q = 32767;
        else
            q = (int16_t)(int32_t)v;
        s16[i] = q;
    }
    const uint64_t total = n / channels;
    if (total < 16)
    {
        free(s16);
        return 4;
    }
    if (level > 8)
    {
        free(s16);
        return 5;
    }
    uint64_t bs = level <= 2 ? 1152 : 4096;
    if (bs > total)
        bs = total;
    if (bs < 16)
        bs = 16;

    bw_t w = {0, 0, 0};
    bw_bits(&w, 0x66, 8);
    bw_bits(&w, 0x4C, 8);
    bw_bits(&w, 0x61, 8);
    bw_bits(&w, 0x43, 8);
    uint8_t md5[16];
    orc_md5((const uint8_t *)s16, n * 2, md5); /* little-endian host: i16 LE bytes */
    bw_bits(&w, 1, 1);
    bw_bits(&w, 0, 7);
    bw_bits(&w, 34, 24);
    bw_bits(&w, bs & 0xFFFF, 16);
    bw_bits(&w, bs & 0xFFFF, 16);
    bw_bits(&w, 0, 24);
    bw_bits(&w, 0, 24);
    bw_bits(&w, rate, 20);
    bw_bits(&w, (uint64_t)(channels - 1), 3);
    bw_bits(&w, 15, 5);
    bw_bits(&w, total, 36);
    for (int i = 0; i < 16; ++i)
        bw_bits(&w, md5[i], 8);

    uint64_t off = 0;
    uint32_t fno = 0;
    while (off < n)
    {
        uint64_t remaining = n - off;
        uint64_t cur = remaining / channels;
        if (cur > bs)
            cur = bs;
        if (cur == 0)
            break;
        encode_frame(&w, s16 + off, channels, rate, fno, (uint32_t)cur, level);
        off += cur * channels;
        ++fno;
    }
    free(s16);
    *len = (w.nbits + 7) / 8;
    *bytes = w.buf;
    return 0;
}
