/*
 * oracle/oracle_common.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the reference hot path of
 * ajcm474/gapless-lossy-codec v0.5.0 (src/codec.rs, src/flac.rs).  It exists to
 * CHECK the CUDA product path; nothing under gapless_lossy_codec_b200/ may
 * include, link or call it.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it.
 *
 * PARITY STATUS: "parity unpinned" at the bit level.  The reference ships no
 * golden vectors or known-answer bitstreams and there is no
 * Rust toolchain in this image, so the reference itself cannot be run.  The
 * oracle is pinned by (a) every property the reference's own tests assert
 * (ported in tests/test_oracle_*.py), (b) independent anchors: RFC 9639
 * decodability with CRC-8/CRC-16/MD5 verification, hashlib MD5, README size
 * example, (c) bit-for-bit agreement with a second restatement written separately in numpy
 * / plain Python (codec_restatement_np.py, flac_restatement_py.py, tests/test_oracle_restatements.py).
 *
 * Build flags that matter: -O2 -ffp-contract=off -fno-fast-math (Rust never
 * contracts a*b+c and never reassociates float reductions).
 */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_FRAME 2048 /* src/codec.rs:15 FRAME_SIZE */
#define ORC_HOP 1024   /* src/codec.rs:16 HOP_SIZE   */

typedef struct
{
    uint16_t idx;
    int16_t q;
} orc_pair;

/* Flat image of the reference's EncodedAudio (src/codec.rs:31-69).  Same field
 * order and meaning as glc_encoded in include/glc.h so that one ctypes
 * Structure serves both in the tests. */
typedef struct
{
    uint32_t sample_rate;     /* AudioHeader.sample_rate */
    uint16_t channels;        /* AudioHeader.channels */
    uint16_t reserved0;
    uint64_t total_samples;   /* AudioHeader.total_samples (interleaved count) */
    uint32_t encoder_delay;   /* GaplessInfo */
    uint32_t padding;
    uint64_t original_length;
    uint64_t n_frames;
    uint8_t *frame_is_raw;    /* [n_frames]  1 = raw_pcm Some(..) */
    uint32_t *nnz;            /* [n_frames*channels], 0 for raw frames */
    uint64_t *pair_offset;    /* [n_frames*channels + 1] exclusive scan of nnz */
    orc_pair *pairs;          /* [pair_offset[last]] */
    float *scales;            /* [n_frames*channels], 0 for raw frames */
    uint64_t *raw_offset;     /* [n_frames + 1] in i16 units */
    int16_t *raw;             /* planar [ch][2048] per raw frame */
} orc_encoded;

/* ---- codec ---- */
void orc_build_tables(float *cos_tab /*1024*2048*/, float *window /*2048*/, float *norm);
int orc_build_perceptual(uint32_t sample_rate, float *weights /*1024*/, int32_t *bands /*>=52*/);
float orc_noise_floor_factor(void);
void orc_mdct_block(const float *cos_tab, float norm, const float *block, float *out);
void orc_imdct_block_ref_order(const float *cos_tab, float norm, const float *coeffs, float *out);
void orc_imdct_block(const float *cos_tab, float norm, const float *coeffs, float *out);
void orc_masking_thresholds(const float *coeffs, const float *weights, const int32_t *bands,
                            int nbands_edges, float *thresholds);
int orc_compress(const float *coeffs, float scale, const float *thresholds, orc_pair *out);

int orc_encode(const float *pcm, uint64_t n, uint16_t channels, uint32_t sample_rate,
               int threads, orc_encoded **out);
int orc_decode(const orc_encoded *enc, int threads, int literal_imdct, float **pcm, uint64_t *n);
int orc_decode_untrimmed(const orc_encoded *enc, int threads, int literal_imdct, float **pcm,
                         uint64_t *n);
void orc_encoded_free(orc_encoded *e);
void orc_free(void *p);
/* dense coefficient dump for debugging parity: coeffs[frames*ch*1024] */
int orc_mdct_all(const float *pcm, uint64_t n, uint16_t channels, int threads, float **coeffs,
                 uint64_t *n_fc);

/* ---- flac ---- */
int orc_flac_encode(const float *pcm, uint64_t n, uint32_t sample_rate, uint16_t channels,
                    uint8_t level, uint8_t **bytes, uint64_t *len);
void orc_md5(const uint8_t *data, uint64_t len, uint8_t digest[16]);
uint8_t orc_crc8(const uint8_t *data, uint64_t len);
uint16_t orc_crc16(const uint8_t *data, uint64_t len);

/* ---- test-only FLAC decoder (RFC 9639 subset: verbatim/constant/fixed, Rice 4/5 bit) ---- */
typedef struct
{
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bits_per_sample;
    uint32_t min_block, max_block;
    uint64_t total_samples; /* per channel */
    uint8_t md5[16];
    uint64_t n_frames;
    int md5_ok;
    uint64_t n_decoded; /* interleaved values */
    int32_t *samples;   /* interleaved */
} orc_flac_info;
int orc_flac_decode(const uint8_t *bytes, uint64_t len, orc_flac_info *info);
int orc_flac_decode_ex(const uint8_t *bytes, uint64_t len, orc_flac_info *info, uint64_t *frame_off,
                       uint64_t frame_cap);

/* FNV-1a/64 over bytes: fingerprint of the cosine table shared with tests/golden/dump_reference.rs */
uint64_t orc_fnv1a64(const uint8_t *data, uint64_t len);

/* ---- bincode 1.3 container image (src/codec.rs:774-786) ---- */
int orc_bincode_serialize(const orc_encoded *enc, uint8_t **bytes, uint64_t *len);
int orc_bincode_deserialize(const uint8_t *bytes, uint64_t len, orc_encoded **out);

#ifdef __cplusplus
}
#endif
#endif
