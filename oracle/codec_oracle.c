/*
 * oracle/codec_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
 *
 * CPU restatement of the codec hot path of ajcm474/gapless-lossy-codec v0.5.0:
 * src/codec.rs:15-29 (constants), :102-183 (perceptual weights / bands),
 * :188-240 (masking thresholds), :270-311 (quantizer), :326-390 (tables, MDCT,
 * IMDCT), :421-565 (encode), :595-768 (decode, overlap-add, gapless trim).
 * "parity unpinned": no reference golden vectors exist and Rust cannot be run
 * here; see the header of oracle_common.h for what pins it instead.
 *
 * Rust -> C semantic map used throughout: every f32 operation is one IEEE
 * single-precision operation (no contraction: build with -ffp-contract=off);
 * f32::cos/sin/powf -> glibc cosf/sinf/powf; sqrt -> sqrtf; round -> roundf
 * (half away from zero); max/min -> fmaxf/fminf; `as i16`/`as usize` from f32 is
 * a saturating truncation with NaN -> 0; Iterator::sum / fold run left to right.
 */
#include "oracle_common.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* Minimal dynamically scheduled parallel-for (stands in for rayon's par_iter,
 * src/codec.rs:462,620: results are index-addressed, so scheduling cannot change
 * them). */
typedef void (*pf_body)(int64_t i, void *ctx);
typedef struct
{
    pf_body body;
    void *ctx;
    int64_t n;
    int64_t next;
    pthread_mutex_t mu;
} pf_state;

static void *pf_worker(void *arg)
{
    pf_state *st = (pf_state *)arg;
    for (;;)
    {
        pthread_mutex_lock(&st->mu);
        int64_t b = st->next;
        st->next += 4;
        pthread_mutex_unlock(&st->mu);
        if (b >= st->n)
            break;
        int64_t e = b + 4 < st->n ? b + 4 : st->n;
        for (int64_t i = b; i < e; ++i)
            st->body(i, st->ctx);
    }
    return NULL;
}

static void parallel_for(int64_t n, int threads, pf_body body, void *ctx)
{
    pf_state st;
    st.body = body;
    st.ctx = ctx;
    st.n = n;
    st.next = 0;
    pthread_mutex_init(&st.mu, NULL);
    if (threads < 1)
        threads = 1;
    if (threads > 256)
        threads = 256;
    pthread_t tid[256];
    int started = 0;
    for (int t = 1; t < threads; ++t)
        if (pthread_create(&tid[started], NULL, pf_worker, &st) == 0)
            ++started;
    pf_worker(&st);
    for (int t = 0; t < started; ++t)
        pthread_join(tid[t], NULL);
    pthread_mutex_destroy(&st.mu);
}

#define N_COEF ORC_HOP
#define BLK ORC_FRAME

/* std::f32::consts::PI */
static const float PI_F32 = 3.14159274101257324219f;

/* ------------------------------------------------------------------ tables */

/* src/codec.rs:326-356.  angle = PI/n * (i + 0.5 + n/2) * (k + 0.5), evaluated
 * left to right in f32, then cosf.  window[i] = sinf(PI * (i + 0.5) / 2048).
 * norm = sqrtf(2/n). */
void orc_build_tables(float *cos_tab, float *window, float *norm)
{
    const float n = (float)N_COEF;
    const float pi_over_n = PI_F32 / n;
    const float half_n = n / 2.0f;
    for (int k = 0; k < N_COEF; ++k)
    {
        const float kk = (float)k + 0.5f;
        for (int i = 0; i < BLK; ++i)
        {
            volatile float shifted = ((float)i + 0.5f) + half_n;
            volatile float a = pi_over_n * shifted;
            volatile float angle = a * kk;
            cos_tab[(size_t)k * BLK + i] = cosf(angle);
        }
    }
    for (int i = 0; i < BLK; ++i)
    {
        volatile float num = PI_F32 * ((float)i + 0.5f);
        volatile float arg = num / (float)BLK;
        window[i] = sinf(arg);
    }
    *norm = sqrtf(2.0f / n);
}

/* src/codec.rs:102-183.  Returns the number of band EDGES written (bands has
 * edges[0]=0 ... edges[last]=1024). */
int orc_build_perceptual(uint32_t sample_rate, float *weights, int32_t *bands)
{
    const float sr = (float)sample_rate;
    for (int k = 0; k < N_COEF; ++k)
    {
        float norm_freq = (float)k / (2.0f * (float)N_COEF);
        float f = norm_freq * sr;
        float w;
        if (f < 100.0f)
            w = 0.3f + (f / 100.0f) * 0.4f;
        else if (f < 200.0f)
            w = 0.7f + ((f - 100.0f) / 100.0f) * 0.3f;
        else if (f < 5000.0f)
            w = 1.0f;
        else if (f < 10000.0f)
            w = 1.0f - ((f - 5000.0f) / 5000.0f) * 0.3f;
        else
            w = 0.7f - fminf((f - 10000.0f) / 12000.0f, 1.0f) * 0.5f;
        weights[k] = fmaxf(w, 0.2f);
    }
    /* compute_critical_bands, src/codec.rs:146-183 */
    int nb = 0;
    bands[nb++] = 0;
    const float nyq = sr / 2.0f;
    float freq = 0.0f;
    while (freq < nyq && nb < 50)
    {
        float pos = (freq / nyq) * (float)N_COEF;
        long bin = (pos >= 0.0f) ? (long)pos : 0; /* `as usize` truncates; never negative here */
        if (bin > bands[nb - 1] && bin < N_COEF)
            bands[nb++] = (int32_t)bin;
        if (freq < 500.0f)
            freq += 50.0f;
        else if (freq < 2000.0f)
            freq += 100.0f;
        else if (freq < 8000.0f)
            freq += 250.0f;
        else
            freq += 500.0f;
    }
    bands[nb++] = N_COEF;
    return nb;
}

/* 10^(NOISE_FLOOR_DB/20) with NOISE_FLOOR_DB = -48 (src/codec.rs:22,277) */
float orc_noise_floor_factor(void)
{
    volatile float e = -48.0f / 20.0f;
    return powf(10.0f, e);
}

/* ------------------------------------------------------------- transforms */

/* src/codec.rs:359-374: out[k] = (sum_i block[i]*tab[k][i]) * norm, the sum
 * strictly in ascending i, product rounded before the add. */
void orc_mdct_block(const float *cos_tab, float norm, const float *block, float *out)
{
    for (int k = 0; k < N_COEF; ++k)
    {
        const float *tb = cos_tab + (size_t)k * BLK;
        float s = 0.0f;
        for (int i = 0; i < BLK; ++i)
        {
            float p = block[i] * tb[i];
            s = s + p;
        }
        out[k] = s * norm;
    }
}

/* src/codec.rs:377-390, literal loop nest (i outer, k inner, stride-2048 walk).
 * This is the variant timed as the CPU baseline. */
void orc_imdct_block_ref_order(const float *cos_tab, float norm, const float *coeffs, float *out)
{
    for (int i = 0; i < BLK; ++i)
    {
        float s = 0.0f;
        for (int k = 0; k < N_COEF; ++k)
        {
            float p = coeffs[k] * cos_tab[(size_t)k * BLK + i];
            s = s + p;
        }
        out[i] = s * norm;
    }
}

/* Same arithmetic per output (ascending k, product then add), loops
 * interchanged and zero coefficients skipped.  Skipping is exact: a zero
 * coefficient contributes +-0, the running sum starts at +0 and can never be
 * -0 (x + y = -0 only if both are -0), so s + (+-0) == s bit for bit.
 * tests/test_oracle_codec.py checks this against the literal version. */
void orc_imdct_block(const float *cos_tab, float norm, const float *coeffs, float *out)
{
    float acc[BLK];
    for (int i = 0; i < BLK; ++i)
        acc[i] = 0.0f;
    for (int k = 0; k < N_COEF; ++k)
    {
        const float c = coeffs[k];
        if (c == 0.0f)
            continue;
        const float *tb = cos_tab + (size_t)k * BLK;
        for (int i = 0; i < BLK; ++i)
        {
            float p = c * tb[i];
            acc[i] = acc[i] + p;
        }
    }
    for (int i = 0; i < BLK; ++i)
        out[i] = acc[i] * norm;
}

/* --------------------------------------------------- masking + quantizer */

static float max_abs_1024(const float *c)
{
    float m = 0.0f;
    for (int k = 0; k < N_COEF; ++k)
        m = fmaxf(m, fabsf(c[k]));
    return m;
}

/* src/codec.rs:188-240 with quality = QUALITY_FACTOR = 0.7 */
void orc_masking_thresholds(const float *coeffs, const float *weights, const int32_t *bands,
                            int nbands_edges, float *th)
{
    for (int k = 0; k < N_COEF; ++k)
        th[k] = 0.0f;
    const float gmax = fmaxf(max_abs_1024(coeffs), 1e-10f);
    volatile float one_minus_q = 1.0f - 0.7f;
    const float cf = fmaxf(one_minus_q, 0.01f);
    for (int b = 0; b + 1 < nbands_edges; ++b)
    {
        int start = bands[b];
        int end = bands[b + 1] < N_COEF ? bands[b + 1] : N_COEF;
        if (start >= end)
            continue;
        float sumsq = 0.0f, sumw = 0.0f;
        for (int i = start; i < end; ++i)
        {
            float sq = coeffs[i] * coeffs[i];
            sumsq = sumsq + sq;
        }
        for (int i = start; i < end; ++i)
            sumw = sumw + weights[i];
        const float cnt = (float)(end - start);
        const float energy = sqrtf(sumsq / cnt);
        const float avg_w = sumw / cnt;
        const float pf = 1.0f / fmaxf(avg_w, 0.1f);
        float base = energy * 0.01f;
        base = base * cf;
        base = base * pf;
        for (int i = start; i < end; ++i)
        {
            float indiv = 1.0f / fmaxf(weights[i], 0.1f);
            float t = base * indiv;
            if (fabsf(coeffs[i]) > gmax * 0.3f)
                t = fminf(t, gmax * 0.05f);
            th[i] = t;
        }
    }
}

static int16_t f32_to_i16_sat_trunc(float v)
{
    if (v != v)
        return 0;
    if (v <= -32768.0f)
        return (int16_t)-32768;
    if (v >= 32767.0f)
        return (int16_t)32767;
    return (int16_t)(int32_t)v; /* C float->int conversion truncates toward zero */
}

/* src/codec.rs:270-311.  compute_quantization_bits_fast (:243-267) cannot
 * return 0 once abs > threshold holds, so it has no effect on the output. */
int orc_compress(const float *coeffs, float scale, const float *th, orc_pair *out)
{
    const float nf = orc_noise_floor_factor() * scale;
    int n = 0;
    for (int k = 0; k < N_COEF; ++k)
    {
        const float c = coeffs[k];
        const float a = fabsf(c);
        const float t = th[k] * scale;
        if (a > nf && a > t)
        {
            float normalized = c / scale;
            float quant = roundf(normalized * 32768.0f);
            float cl = fminf(fmaxf(quant, -32768.0f), 32767.0f);
            int16_t q = f32_to_i16_sat_trunc(cl);
            if (q != 0)
            {
                out[n].idx = (uint16_t)k;
                out[n].q = q;
                ++n;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------ encode */

typedef struct
{
    float *cos_tab;
    float window[BLK];
    float norm;
} tables_t;

static tables_t *g_tables = NULL;

static const tables_t *get_tables(void)
{
    if (!g_tables)
    {
        tables_t *t = (tables_t *)malloc(sizeof(tables_t));
        t->cos_tab = (float *)malloc(sizeof(float) * (size_t)N_COEF * BLK);
        orc_build_tables(t->cos_tab, t->window, &t->norm);
        g_tables = t;
    }
    return g_tables;
}

static uint64_t padded_len_for(uint64_t L)
{
    uint64_t v = 512 + L;
    uint64_t rem = v % N_COEF;
    if (rem)
        v += N_COEF - rem;
    return v + 512;
}

/* value of padded[c][pos], src/codec.rs:433-447, without materialising it */
static inline float padded_at(const float *pcm, uint64_t L, int ch, int c, uint64_t pos)
{
    if (pos < 512 || pos >= 512 + L)
        return 0.0f;
    return pcm[(pos - 512) * (uint64_t)ch + (uint64_t)c];
}

void orc_free(void *p) { free(p); }

void orc_encoded_free(orc_encoded *e)
{
    if (!e)
        return;
    free(e->frame_is_raw);
    free(e->nnz);
    free(e->pair_offset);
    free(e->pairs);
    free(e->scales);
    free(e->raw_offset);
    free(e->raw);
    free(e);
}

typedef struct
{
    const float *pcm;
    uint64_t L;
    int ch;
    const tables_t *T;
    const float *weights;
    const int32_t *bands;
    int nb;
    orc_encoded *e;
    orc_pair *slots;
    int16_t *rawslots;
} enc_ctx;

/* body of the per-frame closure, src/codec.rs:462-541 */
static void encode_one_frame(int64_t fi, void *vctx)
{
    enc_ctx *x = (enc_ctx *)vctx;
    const int ch = x->ch;
    const tables_t *T = x->T;
    float block[BLK], coeffs[N_COEF], th[N_COEF];
    uint64_t total_nnz_bytes = 0;
    for (int c = 0; c < ch; ++c)
    {
        const uint64_t fc = (uint64_t)fi * ch + c;
        const uint64_t start = (uint64_t)fi * N_COEF;
        int16_t *rw = x->rawslots + fc * BLK;
        for (int i = 0; i < BLK; ++i)
        {
            float v = padded_at(x->pcm, x->L, ch, c, start + i);
            float wv = v * T->window[i];
            block[i] = wv;
            /* raw candidate, src/codec.rs:498-502 */
            float s = wv * 32767.0f;
            float cl = (s != s) ? s : fminf(fmaxf(s, -32768.0f), 32767.0f);
            rw[i] = f32_to_i16_sat_trunc(cl);
        }
        orc_mdct_block(T->cos_tab, T->norm, block, coeffs);
        float scale = fmaxf(max_abs_1024(coeffs), 1e-10f);
        orc_masking_thresholds(coeffs, x->weights, x->bands, x->nb, th);
        int cnt = orc_compress(coeffs, scale, th, x->slots + fc * N_COEF);
        x->e->nnz[fc] = (uint32_t)cnt;
        x->e->scales[fc] = scale;
        total_nnz_bytes += 8 + (uint64_t)cnt * 4;
    }
    /* src/codec.rs:505-521 */
    uint64_t compressed = total_nnz_bytes + 8 + (uint64_t)ch * 4 + 64;
    uint64_t raw_size = (uint64_t)BLK * ch * 2;
    float lhs = (float)compressed;
    float rhs = (float)raw_size * 0.85f;
    x->e->frame_is_raw[fi] = (lhs >= rhs) ? 1 : 0;
}

/* Encoder::encode, src/codec.rs:421-565.  Returns 0 ok, 1 bad args, 2 input too
 * short (the reference panics for <= 512 samples per channel: slice out of range at
 * :474 because num_frames is forced to 1 at :449-452). */
int orc_encode(const float *pcm, uint64_t n, uint16_t channels, uint32_t sample_rate, int threads,
               orc_encoded **out)
{
    if (!out || channels == 0 || (n % channels) != 0)
        return 1;
    const int ch = channels;
    const uint64_t L = n / ch;
    if (L <= 512)
        return 2;
    const tables_t *T = get_tables();
    float weights[N_COEF];
    int32_t bands[64];
    const int nb = orc_build_perceptual(sample_rate, weights, bands);

    const uint64_t plen = padded_len_for(L);
    const uint64_t frames = (plen - BLK) / N_COEF + 1;
    const uint64_t nfc = frames * ch;

    orc_encoded *e = (orc_encoded *)calloc(1, sizeof(orc_encoded));
    e->sample_rate = sample_rate;
    e->channels = channels;
    e->total_samples = n;
    e->encoder_delay = 512;
    e->padding = (uint32_t)(plen - L - 512);
    e->original_length = n;
    e->n_frames = frames;
    e->frame_is_raw = (uint8_t *)calloc(frames, 1);
    e->nnz = (uint32_t *)calloc(nfc, sizeof(uint32_t));
    e->pair_offset = (uint64_t *)calloc(nfc + 1, sizeof(uint64_t));
    e->scales = (float *)calloc(nfc, sizeof(float));
    e->raw_offset = (uint64_t *)calloc(frames + 1, sizeof(uint64_t));

    /* worst-case per-frame-channel slots, compacted afterwards */
    orc_pair *slots = (orc_pair *)malloc(sizeof(orc_pair) * nfc * N_COEF);
    int16_t *rawslots = (int16_t *)malloc(sizeof(int16_t) * nfc * BLK);
    if (!slots || !rawslots)
    {
        free(slots);
        free(rawslots);
        orc_encoded_free(e);
        return 3;
    }
    enc_ctx cx = {pcm, L, ch, T, weights, bands, nb, e, slots, rawslots};
    parallel_for((int64_t)frames, threads, encode_one_frame, &cx);

    /* compact */
    uint64_t np = 0, nr = 0;
    for (uint64_t f = 0; f < frames; ++f)
    {
        e->raw_offset[f] = nr;
        if (e->frame_is_raw[f])
        {
            nr += (uint64_t)BLK * ch;
            for (int c = 0; c < ch; ++c)
            {
                e->nnz[f * ch + c] = 0;
                e->scales[f * ch + c] = 0.0f;
            }
        }
        for (int c = 0; c < ch; ++c)
        {
            e->pair_offset[f * ch + c] = np;
            np += e->nnz[f * ch + c];
        }
    }
    e->raw_offset[frames] = nr;
    e->pair_offset[nfc] = np;
    e->pairs = (orc_pair *)malloc(sizeof(orc_pair) * (np ? np : 1));
    e->raw = (int16_t *)malloc(sizeof(int16_t) * (nr ? nr : 1));
    for (uint64_t f = 0; f < frames; ++f)
    {
        if (e->frame_is_raw[f])
        {
            memcpy(e->raw + e->raw_offset[f], rawslots + f * ch * BLK,
                   sizeof(int16_t) * (size_t)BLK * ch);
        }
        else
        {
            for (int c = 0; c < ch; ++c)
            {
                uint64_t fc = f * ch + c;
                memcpy(e->pairs + e->pair_offset[fc], slots + fc * N_COEF,
                       sizeof(orc_pair) * e->nnz[fc]);
            }
        }
    }
    free(slots);
    free(rawslots);
    *out = e;
    return 0;
}

typedef struct
{
    const float *pcm;
    uint64_t L;
    int ch;
    const tables_t *T;
    float *co;
} mdct_ctx;

static void mdct_one_frame(int64_t fi, void *vctx)
{
    mdct_ctx *x = (mdct_ctx *)vctx;
    float block[BLK];
    for (int c = 0; c < x->ch; ++c)
    {
        for (int i = 0; i < BLK; ++i)
            block[i] = padded_at(x->pcm, x->L, x->ch, c, (uint64_t)fi * N_COEF + i) * x->T->window[i];
        orc_mdct_block(x->T->cos_tab, x->T->norm, block, x->co + ((uint64_t)fi * x->ch + c) * N_COEF);
    }
}

/* Dense MDCT coefficients of every frame-channel (debug aid for the GPU parity
 * tests: lets a failure be localised to the transform or to the quantizer). */
int orc_mdct_all(const float *pcm, uint64_t n, uint16_t channels, int threads, float **coeffs_out,
                 uint64_t *n_fc)
{
    if (channels == 0 || (n % channels) != 0)
        return 1;
    const int ch = channels;
    const uint64_t L = n / ch;
    if (L <= 512)
        return 2;
    const tables_t *T = get_tables();
    const uint64_t plen = padded_len_for(L);
    const uint64_t frames = (plen - BLK) / N_COEF + 1;
    float *co = (float *)malloc(sizeof(float) * frames * ch * N_COEF);
    if (threads < 1)
        threads = 1;
    mdct_ctx cx = {pcm, L, ch, T, co};
    parallel_for((int64_t)frames, threads, mdct_one_frame, &cx);
    *coeffs_out = co;
    *n_fc = frames * ch;
    return 0;
}

/* ------------------------------------------------------------ decode */

/* One frame -> per-channel blocks of 2048, src/codec.rs:622-681 */
static void decode_frame_blocks(const orc_encoded *e, uint64_t f, const tables_t *T, int literal,
                                float *blocks /* [ch][2048] */)
{
    const int ch = e->channels;
    if (e->frame_is_raw[f])
    {
        const int16_t *raw = e->raw + e->raw_offset[f];
        const uint64_t rawlen = e->raw_offset[f + 1] - e->raw_offset[f];
        for (int c = 0; c < ch; ++c)
        {
            float *b = blocks + (size_t)c * BLK;
            for (int i = 0; i < BLK; ++i)
            {
                uint64_t si = (uint64_t)i * ch + c; /* interleaved read of planar data, :635 */
                b[i] = (si < rawlen) ? ((float)raw[si] / 32767.0f) : 0.0f;
            }
        }
        return;
    }
    float coeffs[N_COEF];
    for (int c = 0; c < ch; ++c)
    {
        const uint64_t fc = f * ch + c;
        for (int k = 0; k < N_COEF; ++k)
            coeffs[k] = 0.0f;
        const float scale = fmaxf(e->scales[fc], 1e-12f);
        const orc_pair *p = e->pairs + e->pair_offset[fc];
        for (uint32_t j = 0; j < e->nnz[fc]; ++j)
        {
            if (p[j].idx < N_COEF)
            {
                float v = (float)p[j].q / 32768.0f;
                coeffs[p[j].idx] = v * scale;
            }
        }
        float *b = blocks + (size_t)c * BLK;
        if (literal)
            orc_imdct_block_ref_order(T->cos_tab, T->norm, coeffs, b);
        else
            orc_imdct_block(T->cos_tab, T->norm, coeffs, b);
        for (int i = 0; i < BLK; ++i)
            b[i] = b[i] * T->window[i];
    }
}

typedef struct
{
    const orc_encoded *e;
    const tables_t *T;
    int literal;
    float *blocks;
} dec_ctx;

static void decode_one_frame(int64_t f, void *vctx)
{
    dec_ctx *x = (dec_ctx *)vctx;
    decode_frame_blocks(x->e, (uint64_t)f, x->T, x->literal,
                        x->blocks + (size_t)f * x->e->channels * BLK);
}

/* decode_streaming concatenated (src/codec.rs:595-741): (frames+1)*1024*ch values.
 * The reference's batches of 32 / chunks of 500 only schedule; the values do not
 * depend on them, so the oracle parallelises over all frames and then performs
 * the serial overlap-add exactly as :688-705, :723-729. */
int orc_decode_untrimmed(const orc_encoded *e, int threads, int literal, float **pcm, uint64_t *n)
{
    if (!e || !pcm || !n || e->channels == 0)
        return 1;
    const tables_t *T = get_tables();
    const int ch = e->channels;
    const uint64_t frames = e->n_frames;
    const uint64_t total = (frames + 1) * N_COEF * ch;
    float *outp = (float *)malloc(sizeof(float) * (total ? total : 1));
    float *blocks = (float *)malloc(sizeof(float) * (frames ? frames : 1) * ch * BLK);
    if (!outp || !blocks)
    {
        free(outp);
        free(blocks);
        return 3;
    }
    if (threads < 1)
        threads = 1;
    dec_ctx cx = {e, T, literal, blocks};
    parallel_for((int64_t)frames, threads, decode_one_frame, &cx);

    float *overlap = (float *)calloc((size_t)ch * N_COEF, sizeof(float));
    uint64_t w = 0;
    for (uint64_t f = 0; f < frames; ++f)
    {
        const float *fb = blocks + (size_t)f * ch * BLK;
        for (int i = 0; i < N_COEF; ++i)
            for (int c = 0; c < ch; ++c)
                outp[w++] = overlap[(size_t)c * N_COEF + i] + fb[(size_t)c * BLK + i];
        for (int c = 0; c < ch; ++c)
            memcpy(overlap + (size_t)c * N_COEF, fb + (size_t)c * BLK + N_COEF,
                   sizeof(float) * N_COEF);
    }
    for (int i = 0; i < N_COEF; ++i)
        for (int c = 0; c < ch; ++c)
            outp[w++] = overlap[(size_t)c * N_COEF + i];
    free(overlap);
    free(blocks);
    *pcm = outp;
    *n = w;
    return 0;
}

/* Decoder::decode, src/codec.rs:744-768: drop `encoder_delay` interleaved values
 * (not sample frames) if longer than that, then truncate to original_length. */
int orc_decode(const orc_encoded *e, int threads, int literal, float **pcm, uint64_t *n)
{
    float *all = NULL;
    uint64_t len = 0;
    int rc = orc_decode_untrimmed(e, threads, literal, &all, &len);
    if (rc)
        return rc;
    const uint64_t delay = e->encoder_delay;
    uint64_t off = 0;
    if (len > delay)
    {
        off = delay;
        len -= delay;
    }
    if (len > e->original_length)
        len = e->original_length;
    if (off)
        memmove(all, all + off, sizeof(float) * len);
    *pcm = all;
    *n = len;
    return 0;
}
