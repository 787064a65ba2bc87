// glc_shard.cpp -- one call, several devices: the files of a batch are split over several contexts
// (normally one per GPU of the box) and every context runs the ordinary batch entry point on its own
// host thread.  The path shards by file with no exchange step: the reference loops over
// files (src/main.rs:546-583) and parallelises inside a file (rayon, src/codec.rs:462, 620); here the
// only inter-device step is the host-side gather of the per-file outputs back into input order.
// Host C++ only -- every device-side call goes through the C ABI of glc.h.
#include <algorithm>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "glc_internal.cuh"

#define GLC_TRY(expr)              \
    do                             \
    {                              \
        glc_status _s = (expr);    \
        if (_s != GLC_OK)          \
            return _s;             \
    } while (0)

namespace
{

// Frame count of Encoder::encode for `n` samples per channel (src/codec.rs:433-455): the planning weight.
uint64_t frames_for(uint64_t per_channel)
{
    uint64_t padded = 512 + per_channel;
    if (padded % 1024)
        padded += 1024 - padded % 1024;
    padded += 512;
    return (padded - 2048) / 1024 + 1;
}

// run fn(shard) on one thread per non-empty shard; first failure wins (status + message of that thread)
template <typename Fn>
glc_status run_shards(uint32_t n_shards, const std::vector<std::vector<uint32_t>> &members, Fn fn)
{
    std::vector<glc_status> st(n_shards, GLC_OK);
    std::vector<std::string> msg(n_shards);
    std::vector<std::thread> th;
    for (uint32_t s = 0; s < n_shards; ++s)
        if (!members[s].empty())
            th.emplace_back([&, s]() {
                st[s] = fn(s);
                if (st[s] != GLC_OK)
                    msg[s] = glc_last_error(); // thread-local in the worker
            });
    for (auto &t : th)
        t.join();
    for (uint32_t s = 0; s < n_shards; ++s)
        if (st[s] != GLC_OK)
            return glc::set_error(st[s], "shard %u: %s", s, msg[s].c_str());
    return GLC_OK;
}

std::vector<std::vector<uint32_t>> members_of(uint32_t n_files, uint32_t n_shards, const uint32_t *shard_of)
{
    std::vector<std::vector<uint32_t>> m(n_shards);
    for (uint32_t i = 0; i < n_files; ++i)
        m[shard_of[i]].push_back(i); // ascending inside a shard: outputs keep the caller's relative order
    return m;
}

} // namespace

// Longest-processing-time-first: files by descending weight (ties: lower index), each to the least
// loaded shard (ties: lower shard).  Deterministic; the same plan as shard.plan_by_file.
extern "C" glc_status glc_plan_shards(uint32_t n_files, const uint64_t *weights, uint32_t n_shards, uint32_t *shard_of)
{
    if (!weights || !shard_of || n_shards == 0)
        return glc::set_error(GLC_ERR_INVALID_ARG, "null/empty argument");
    std::vector<uint32_t> order(n_files);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weights[a] > weights[b]; });
    std::vector<uint64_t> load(n_shards, 0);
    for (uint32_t i : order)
    {
        const uint32_t s = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
        shard_of[i] = s;
        load[s] += weights[i];
    }
    return GLC_OK;
}

extern "C" glc_status glc_encode_batch_sharded(glc_encoder *const *encs, uint32_t n_shards, uint32_t n_files,
                                               const float *const *pcm, const uint64_t *n_samples,
                                               const uint16_t *channels, glc_encoded **out, uint32_t *shard_of)
{
    if (!encs || !pcm || !n_samples || !channels || !out || !shard_of || n_shards == 0 || n_files == 0)
        return glc::set_error(GLC_ERR_INVALID_ARG, "null/empty argument");
    std::vector<uint64_t> w(n_files);
    for (uint32_t i = 0; i < n_files; ++i)
    {
        // the planner's weights need valid lengths: refuse up front what glc_encode_batch would refuse later
        if (channels[i] == 0)
            return glc::set_error(GLC_ERR_INVALID_ARG, "file %u: 0 channels", i);
        if (n_samples[i] % channels[i])
            return glc::set_error(GLC_ERR_INVALID_ARG, "file %u: %llu samples is not a multiple of %u channels", i,
                                  (unsigned long long)n_samples[i], (unsigned)channels[i]);
        if (n_samples[i] / channels[i] <= 512)
            return glc::set_error(GLC_ERR_TOO_SHORT,
                                  "file %u: %llu samples per channel; the reference panics for <= 512 "
                                  "(src/codec.rs:449-452,474)",
                                  i, (unsigned long long)(n_samples[i] / channels[i]));
        w[i] = frames_for(n_samples[i] / channels[i]) * channels[i];
    }
    GLC_TRY(glc_plan_shards(n_files, w.data(), n_shards, shard_of));
    const auto members = members_of(n_files, n_shards, shard_of);
    for (uint32_t i = 0; i < n_files; ++i)
        out[i] = nullptr;
    const glc_status rs = run_shards(n_shards, members, [&](uint32_t s) -> glc_status {
        const std::vector<uint32_t> &m = members[s];
        std::vector<const float *> p(m.size());
        std::vector<uint64_t> n(m.size());
        std::vector<uint16_t> c(m.size());
        std::vector<glc_encoded *> o(m.size(), nullptr);
        for (size_t k = 0; k < m.size(); ++k)
        {
            p[k] = pcm[m[k]];
            n[k] = n_samples[m[k]];
            c[k] = channels[m[k]];
        }
        const glc_status st = glc_encode_batch(encs[s], (uint32_t)m.size(), p.data(), n.data(), c.data(), o.data());
        if (st == GLC_OK)
            for (size_t k = 0; k < m.size(); ++k)
                out[m[k]] = o[k];
        return st;
    });
    if (rs != GLC_OK) // nothing is returned on failure: release what the other shards produced
        for (uint32_t i = 0; i < n_files; ++i)
            if (out[i])
            {
                glc_encoded_free(nullptr, out[i]);
                out[i] = nullptr;
            }
    return rs;
}

extern "C" glc_status glc_decode_batch_sharded(glc_decoder *const *decs, uint32_t n_shards, uint32_t n_files,
                                               const glc_encoded *const *enc, float **pcm, uint64_t *n_samples,
                                               uint32_t *shard_of)
{
    if (!decs || !enc || !pcm || !n_samples || !shard_of || n_shards == 0 || n_files == 0)
        return glc::set_error(GLC_ERR_INVALID_ARG, "null/empty argument");
    std::vector<uint64_t> w(n_files);
    for (uint32_t i = 0; i < n_files; ++i)
    {
        if (!enc[i])
            return glc::set_error(GLC_ERR_INVALID_ARG, "stream %u is null", i);
        w[i] = enc[i]->n_frames * (uint64_t)enc[i]->channels;
    }
    GLC_TRY(glc_plan_shards(n_files, w.data(), n_shards, shard_of));
    const auto members = members_of(n_files, n_shards, shard_of);
    for (uint32_t i = 0; i < n_files; ++i)
    {
        pcm[i] = nullptr;
        n_samples[i] = 0;
    }
    const glc_status rs = run_shards(n_shards, members, [&](uint32_t s) -> glc_status {
        const std::vector<uint32_t> &m = members[s];
        std::vector<const glc_encoded *> e(m.size());
        std::vector<float *> o(m.size(), nullptr);
        std::vector<uint64_t> n(m.size(), 0);
        for (size_t k = 0; k < m.size(); ++k)
            e[k] = enc[m[k]];
        const glc_status st = glc_decode_batch(decs[s], (uint32_t)m.size(), e.data(), o.data(), n.data());
        if (st == GLC_OK)
            for (size_t k = 0; k < m.size(); ++k)
            {
                pcm[m[k]] = o[k];
                n_samples[m[k]] = n[k];
            }
        return st;
    });
    if (rs != GLC_OK)
        for (uint32_t i = 0; i < n_files; ++i)
            if (pcm[i])
            {
                glc_free(glc::decoder_ctx(decs[shard_of[i]]), pcm[i]);
                pcm[i] = nullptr;
                n_samples[i] = 0;
            }
    return rs;
}

extern "C" glc_status glc_flac_encode_batch_sharded(glc_ctx *const *ctxs, uint32_t n_shards, uint32_t n_files,
                                                    const float *const *pcm, const uint64_t *n_samples,
                                                    const uint32_t *sample_rate, const uint16_t *channels,
                                                    uint8_t level, uint8_t **bytes, uint64_t *len, uint32_t *shard_of)
{
    if (!ctxs || !pcm || !n_samples || !sample_rate || !channels || !bytes || !len || !shard_of || n_shards == 0 ||
        n_files == 0)
        return glc::set_error(GLC_ERR_INVALID_ARG, "null/empty argument");
    GLC_TRY(glc_plan_shards(n_files, n_samples, n_shards, shard_of)); // FLAC work is linear in the sample count
    const auto members = members_of(n_files, n_shards, shard_of);
    for (uint32_t i = 0; i < n_files; ++i)
    {
        bytes[i] = nullptr;
        len[i] = 0;
    }
    const glc_status rs = run_shards(n_shards, members, [&](uint32_t s) -> glc_status {
        const std::vector<uint32_t> &m = members[s];
        std::vector<const float *> p(m.size());
        std::vector<uint64_t> n(m.size()), l(m.size(), 0);
        std::vector<uint32_t> r(m.size());
        std::vector<uint16_t> c(m.size());
        std::vector<uint8_t *> b(m.size(), nullptr);
        for (size_t k = 0; k < m.size(); ++k)
        {
            p[k] = pcm[m[k]];
            n[k] = n_samples[m[k]];
            r[k] = sample_rate[m[k]];
            c[k] = channels[m[k]];
        }
        const glc_status st =
            glc_flac_encode_batch(ctxs[s], (uint32_t)m.size(), p.data(), n.data(), r.data(), c.data(), level, b.data(), l.data());
        if (st == GLC_OK)
            for (size_t k = 0; k < m.size(); ++k)
            {
                bytes[m[k]] = b[k];
                len[m[k]] = l[k];
            }
        return st;
    });
    if (rs != GLC_OK)
        for (uint32_t i = 0; i < n_files; ++i)
            if (bytes[i])
            {
                glc_free(ctxs[shard_of[i]], bytes[i]);
                bytes[i] = nullptr;
                len[i] = 0;
            }
    return rs;
}
