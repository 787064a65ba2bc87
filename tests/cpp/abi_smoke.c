/* abi_smoke.c -- the C ABI of libglc_b200.so used from plain C (no Python, no C++): the calls a
 * foreign-language binding of the reference's codec/flac API would make.
 * Mirrors tests/test_simple.rs (mono 440 Hz, 2 s: exact length, SNR > -10 dB) and tests/test_flac.rs.
 * exit 0 = all checks passed, 77 = no CUDA device (the library has no CPU fallback), 1 = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "glc.h"

#define CHECK(call)                                                                   \
    do                                                                                \
    {                                                                                 \
        glc_status st_ = (call);                                                      \
        if (st_ != GLC_OK)                                                            \
        {                                                                             \
            fprintf(stderr, "%s -> %d: %s\n", #call, (int)st_, glc_last_error());     \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(void)
{
    if (glc_abi_version() != GLC_ABI_VERSION)
    {
        fprintf(stderr, "ABI mismatch: header %u, library %u\n", GLC_ABI_VERSION, glc_abi_version());
        return 1;
    }
    glc_ctx *ctx = NULL;
    glc_status st = glc_ctx_create(0, GLC_MODE_EXACT, &ctx);
    if (st == GLC_ERR_NO_DEVICE)
    {
        printf("NO DEVICE: %s\n", glc_last_error());
        return 77;
    }
    CHECK(st);

    const uint32_t sr = 44100;
    const uint64_t n = 2 * sr; /* tests/test_simple.rs: 2 s mono */
    float *x = (float *)malloc(n * sizeof(float));
    for (uint64_t i = 0; i < n; ++i)
        x[i] = 0.5f * sinf(2.0f * 3.14159265358979f * 440.0f * (float)i / (float)sr);

    glc_encoder *enc = NULL;
    glc_decoder *dec = NULL;
    CHECK(glc_encoder_new(ctx, sr, &enc));
    CHECK(glc_decoder_new(ctx, 1, sr, &dec));
    glc_encoded *e = NULL;
    CHECK(glc_encode(enc, x, n, 1, &e));
    if (e->n_frames != 86 || e->encoder_delay != 512 || e->original_length != n)
    {
        fprintf(stderr, "unexpected stream shape: frames %llu delay %u len %llu\n", (unsigned long long)e->n_frames,
                e->encoder_delay, (unsigned long long)e->original_length);
        return 1;
    }
    float *pcm = NULL;
    uint64_t m = 0;
    CHECK(glc_decode(dec, e, &pcm, &m));
    if (m != n) /* exact gapless length, tests/test_simple.rs:117 */
    {
        fprintf(stderr, "decoded %llu samples, expected %llu\n", (unsigned long long)m, (unsigned long long)n);
        return 1;
    }
    double sp = 0, np = 0; /* SNR over [1000, len-1000), tests/utils.rs:118-147 */
    for (uint64_t i = 1000; i + 1000 < n; ++i)
    {
        sp += (double)x[i] * x[i];
        np += ((double)x[i] - pcm[i]) * ((double)x[i] - pcm[i]);
    }
    const double snr = 10.0 * log10(sp / np);
    if (!(snr > -10.0))
    {
        fprintf(stderr, "SNR %.2f dB\n", snr);
        return 1;
    }

    /* streaming: chunks of exactly 500 frames, then the tail */
    glc_stream *s = NULL;
    CHECK(glc_decode_stream_open(dec, e, &s));
    uint64_t total = 0;
    int last = 0, chunks = 0;
    while (!last)
    {
        const float *c = NULL;
        uint64_t cn = 0;
        float pct = 0;
        CHECK(glc_decode_stream_next(s, &c, &cn, &last, &pct));
        total += cn;
        ++chunks;
    }
    glc_decode_stream_close(s);
    if (total != (e->n_frames + 1) * GLC_HOP_SIZE || chunks != 1)
    {
        fprintf(stderr, "streaming delivered %llu values in %d chunks\n", (unsigned long long)total, chunks);
        return 1;
    }

    /* container + FLAC */
    uint8_t *blob = NULL, *fl = NULL;
    uint64_t blen = 0, flen = 0;
    CHECK(glc_encoded_to_bincode(ctx, e, &blob, &blen));
    glc_encoded *back = NULL;
    CHECK(glc_encoded_from_bincode(ctx, blob, blen, &back));
    if (back->n_frames != e->n_frames || memcmp(back->nnz, e->nnz, e->n_frames * sizeof(uint32_t)) != 0)
    {
        fprintf(stderr, "bincode round trip differs\n");
        return 1;
    }
    CHECK(glc_flac_encode(ctx, pcm, m, sr, 1, 5, &fl, &flen));
    if (flen < 42 || memcmp(fl, "fLaC", 4) != 0)
    {
        fprintf(stderr, "bad FLAC stream\n");
        return 1;
    }
    /* the reference's two errors, in its order (src/flac.rs:963-978) */
    uint8_t *dummy = NULL;
    uint64_t dlen = 0;
    if (glc_flac_encode(ctx, x, 10, sr, 1, 9, &dummy, &dlen) != GLC_ERR_FLAC_TOO_SHORT ||
        glc_flac_encode(ctx, x, 100, sr, 1, 9, &dummy, &dlen) != GLC_ERR_FLAC_LEVEL ||
        glc_encode(enc, x, 512, 1, &back) != GLC_ERR_TOO_SHORT)
    {
        fprintf(stderr, "error codes do not mirror the reference\n");
        return 1;
    }
    glc_stats stats;
    glc_stats_get(ctx, &stats);
    printf("OK frames=%llu snr=%.2f dB glc=%llu B flac=%llu B launches(mdct)=%llu\n", (unsigned long long)e->n_frames, snr,
           (unsigned long long)blen, (unsigned long long)flen, (unsigned long long)stats.launches[GLC_K_MDCT_EXACT]);
    glc_free(ctx, fl);
    glc_free(ctx, blob);
    glc_free(ctx, pcm);
    glc_encoded_free(ctx, e);
    glc_decoder_free(dec);
    glc_encoder_free(enc);
    glc_ctx_destroy(ctx);
    free(x);
    return 0;
}
