// glc_tables.cpp -- host-side constant tables of the codec, built with the host libm.
//
// Parity requirement: the reference transform is "whatever its f32 table
// says".  MdctTables::new (reference src/codec.rs:326-356) evaluates the angles in f32 and calls
// f32::cos / f32::sin, i.e. the platform libm.  CUDA's cosf is a different function, so the
// tables are produced here, on the host, and uploaded; never on the device.
// Compile with -ffp-contract=off (see Makefile): one IEEE single operation per source operator.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "glc_internal.cuh"

namespace glc
{

static const float kPiF32 = 3.14159274101257324219f; // std::f32::consts::PI

// Opaque to the optimiser so that no expression is folded in higher precision.
static inline float keep(float v)
{
    volatile float t = v;
    return t;
}

void build_host_tables(HostTables *t)
{
    const float n = (float)kHop;
    t->cos_tab = (float *)malloc(sizeof(float) * (size_t)kHop * kFrame);
    const float step = keep(kPiF32 / n);      // PI / (n as f32)
    const float half = keep(n / 2.0f);        // (n as f32) / 2.0
    for (int k = 0; k < kHop; ++k)
    {
        const float kh = keep((float)k + 0.5f);
        float *row = t->cos_tab + (size_t)k * kFrame;
        for (int i = 0; i < kFrame; ++i)
        {
            // PI / n * (i + 0.5 + n/2) * (k + 0.5), left to right   (src/codec.rs:335)
            const float pos = keep(keep((float)i + 0.5f) + half);
            const float ang = keep(keep(step * pos) * kh);
            row[i] = cosf(ang);
        }
    }
    for (int i = 0; i < kFrame; ++i)
    {
        // (PI * (i + 0.5) / 2048).sin()                             (src/codec.rs:343)
        const float num = keep(kPiF32 * keep((float)i + 0.5f));
        t->window[i] = sinf(keep(num / (float)kFrame));
    }
    t->norm = sqrtf(keep(2.0f / n));                              // src/codec.rs:347
    t->noise_floor_factor = powf(10.0f, keep(-48.0f / 20.0f));    // src/codec.rs:22,277
}

void free_host_tables(HostTables *t)
{
    free(t->cos_tab);
    t->cos_tab = nullptr;
}

// PerceptualWeights::new + compute_critical_bands (src/codec.rs:102-183) and the
// signal-independent factors of compute_masking_thresholds (:218-228).
void build_host_perceptual(uint32_t sample_rate, HostPerceptual *p)
{
    memset(p, 0, sizeof *p);
    float w[kHop];
    const float sr = (float)sample_rate;
    for (int k = 0; k < kHop; ++k)
    {
        const float hz = keep(keep((float)k / keep(2.0f * (float)kHop)) * sr);
        float v;
        if (hz < 100.0f)
            v = keep(0.3f + keep(keep(hz / 100.0f) * 0.4f));
        else if (hz < 200.0f)
            v = keep(0.7f + keep(keep(keep(hz - 100.0f) / 100.0f) * 0.3f));
        else if (hz < 5000.0f)
            v = 1.0f;
        else if (hz < 10000.0f)
            v = keep(1.0f - keep(keep(keep(hz - 5000.0f) / 5000.0f) * 0.3f));
        else
            v = keep(0.7f - keep(fminf(keep(keep(hz - 10000.0f) / 12000.0f), 1.0f) * 0.5f));
        w[k] = fmaxf(v, 0.2f);
        p->inv_w[k] = keep(1.0f / fmaxf(w[k], 0.1f));
    }
    // band ladder: 50 Hz steps below 500 Hz, 100 below 2 kHz, 250 below 8 kHz, then 500;
    // at most 50 edges, then the terminating edge 1024.
    int ne = 0;
    p->band_edges[ne++] = 0;
    const float nyq = keep(sr / 2.0f);
    float f = 0.0f;
    while (f < nyq && ne < 50)
    {
        const float x = keep(keep(f / nyq) * (float)kHop);
        const long bin = x > 0.0f ? (long)x : 0; // `as usize`: truncation
        if (bin > p->band_edges[ne - 1] && bin < kHop)
            p->band_edges[ne++] = (int32_t)bin;
        const float inc = f < 500.0f ? 50.0f : (f < 2000.0f ? 100.0f : (f < 8000.0f ? 250.0f : 500.0f));
        f = keep(f + inc);
    }
    p->band_edges[ne++] = kHop;
    p->n_edges = ne;
    for (int b = 0; b + 1 < ne; ++b)
    {
        const int lo = p->band_edges[b], hi = p->band_edges[b + 1];
        float acc = 0.0f; // weights[start..end].iter().sum::<f32>()
        for (int k = lo; k < hi; ++k)
            acc = keep(acc + w[k]);
        const float cnt = (float)(hi - lo);
        p->band_cnt[b] = cnt;
        p->band_pf[b] = keep(1.0f / fmaxf(keep(acc / cnt), 0.1f));
    }
    p->cf = fmaxf(keep(1.0f - 0.7f), 0.01f);
}

// Tiled copies of the table so that one pipeline stage of a CTA is ONE contiguous 16 KiB block
// (a single bulk-copy / TMA transaction).  Layout: [n_block][stage][kKC][kBN] floats.
void tile_table_for_mdct(const float *tab, float *out)
{
    // reduction index = i (2048), output index = k (1024): element (i, k) = tab[k][i]
    const int n_blocks = kHop / kBN, stages = kFrame / kKC;
    for (int nb = 0; nb < n_blocks; ++nb)
        for (int s = 0; s < stages; ++s)
            for (int r = 0; r < kKC; ++r)
                for (int c = 0; c < kBN; ++c)
                {
                    const int i = s * kKC + r, k = nb * kBN + c;
                    out[(((size_t)nb * stages + s) * kKC + r) * kBN + c] = tab[(size_t)k * kFrame + i];
                }
}

// IMDCT: [n_block][k][kImdctBN] floats, so that the kImdctKC table rows of a pipeline stage are ONE
// contiguous block for every output block (a single bulk copy).
void tile_table_for_imdct(const float *tab, float *out)
{
    // reduction index = k (1024), output index = i (2048): element (k, i) = tab[k][i]
    const int n_blocks = kFrame / kImdctBN;
    for (int nb = 0; nb < n_blocks; ++nb)
        for (int k = 0; k < kHop; ++k)
            for (int c = 0; c < kImdctBN; ++c)
                out[((size_t)nb * kHop + k) * kImdctBN + c] = tab[(size_t)k * kFrame + nb * kImdctBN + c];
}

} // namespace glc
