"""FAST-mode (FFT, true MDCT) versus the reference's transform: the measured flip histogram SURVEY.md 7.2
asks for.  Runs on a B200 (the FAST encoder is the CUDA kernel; the reference side is the CPU oracle):

    python tools/fast_report.py [--seconds 4] > profiles/r2_fast_flip_report.json

Per signal class it reports, over the frames both encoders treat as sparse:
  kept_ref / kept_fast        coefficients kept by each (of frame-channels x 1024 bins)
  kept_either                 union: the denominator of the percentages below
  membership_flips            kept by exactly one of the two
  dq_hist                     histogram of q_fast - q_ref over the union (absent = 0), clipped to +-8, ">8" beyond
  flipped_pct                 share of the union whose (index, value) differs in any way
  scale_bits_differ_pct       frame-channels whose f32 scale factor differs in any bit
  raw_decision_flips          frames whose raw/sparse decision differs
  snr_fast_enc_ref_dec_db     decoded PCM (FAST encode -> reference decode) against (reference encode -> reference decode)
  snr_ref_enc_fast_dec_db     decoded PCM (reference encode -> FAST decode) against the same
  snr_vs_input_ref_db / _fast_db   each round trip against the input itself, aligned by the codec's own shift
                              (512 values are trimmed, i.e. 512/ch sample frames, src/codec.rs:756-761)
The JSON line is what is committed under profiles/; tests/test_gpu_fast.py asserts the tolerances.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
import signals  # noqa: E402
from parity import to_oracle, to_product  # noqa: E402


def _snr(ref, got):
    ref = ref.astype(np.float64)
    err = ref - got.astype(np.float64)
    p, n = float(np.sum(ref * ref)), float(np.sum(err * err))
    return None if n == 0 else round(10 * np.log10(max(p, 1e-300) / n), 2)


def _dense(e):
    """[rows, 1024] int32 of quantised values (0 = not kept) of the sparse rows"""
    rows = e.n_frames * e.channels
    out = np.zeros((rows, 1024), np.int32)
    row_of_pair = np.repeat(np.arange(rows), e.nnz.astype(np.int64))
    ok = e.pair_idx < 1024
    out[row_of_pair[ok], e.pair_idx[ok].astype(np.int64)] = e.pair_q[ok]
    return out


def _low_passed_noise_tones(sr, ch, secs):
    total = int(secs * sr)
    out = np.empty((total, ch), np.float32)
    t = np.arange(total) / sr
    for c in range(ch):
        st = signals.lcg_u64(777 + c, total)
        wn = ((st.astype(np.float32) / np.float32(18446744073709551615.0)) - np.float32(0.5)) * np.float32(0.6)
        lp = np.convolve(wn.astype(np.float64), np.ones(16) / 16.0, mode="same")
        out[:, c] = (0.5 * lp + 0.2 * np.sin(2 * np.pi * 523.25 * t) + 0.1 * np.sin(2 * np.pi * 1318.5 * t)).astype(np.float32)
    return out.reshape(-1)


def classes(secs):
    return [
        ("sine_440", signals.sine(440, 44100, 1, secs), 1, 44100),
        ("multi_sine_x20", signals.multi_sine(44100, 2, secs), 2, 44100),
        ("sweep_100_8000", signals.sweep(100, 8000, 44100, 1, secs), 1, 44100),
        ("square_440", signals.square(440, 44100, 1, secs), 1, 44100),
        ("sawtooth_440", signals.sawtooth(440, 44100, 1, secs), 1, 44100),
        ("low_passed_noise_plus_tones", _low_passed_noise_tones(44100, 2, secs), 2, 44100),
        ("music_like_bench_signal", signals.music_like(44100, 2, secs), 2, 44100),
        ("white_noise", signals.white_noise(44100, 2, min(secs, 1.0), 12345), 2, 44100),
        ("music_like_5_1_48k", signals.music_like(48000, 6, min(secs, 2.0), seed=1000), 6, 48000),
    ]


def report_one(name, x, ch, sr, fast_ctx):
    from gapless_lossy_codec_b200 import Decoder, Encoder

    ref = oracle.encode(x, ch, sr)
    fast = to_oracle(Encoder(sr, fast_ctx).encode(x, ch))
    both_sparse = np.repeat((ref.frame_is_raw == 0) & (fast.frame_is_raw == 0), ch)
    qr, qf = _dense(ref)[both_sparse], _dense(fast)[both_sparse]
    kept_r, kept_f = qr != 0, qf != 0
    union = kept_r | kept_f
    dq = (qf - qr)[union]
    hist = {str(v): int(np.sum(dq == v)) for v in range(-8, 9)}
    hist[">8"] = int(np.sum(np.abs(dq) > 8))
    n_union = int(union.sum())
    sb = ref.scales.view(np.uint32)[both_sparse] != fast.scales.view(np.uint32)[both_sparse]
    pcm_ref = oracle.decode(ref)
    raw_flips = int(np.sum(ref.frame_is_raw != fast.frame_is_raw))
    pcm_fe = oracle.decode(fast)
    pcm_fd = Decoder(ch, sr, fast_ctx).decode(to_product(ref))
    pcm_ff = Decoder(ch, sr, fast_ctx).decode(to_product(fast))
    shift = 512 - 512 // ch * ch  # the trim removes 512 values: 512 // ch sample frames and `shift` channels
    # round trips against the input: the decoded stream is the input delayed by 512 - 512/ch sample frames
    d = 512 - 512 // ch  # sample frames
    xin = x.reshape(-1, ch)

    def vs_input(p):
        pp = p.reshape(-1, ch) if shift == 0 else None
        if pp is None:
            return None  # channels rotated by the 512-value trim (e.g. 5.1): not comparable sample for sample
        n = min(len(pp) - d, len(xin)) - 2048
        return _snr(xin[1024:n], pp[1024 + d:n + d])

    return {
        "signal": name, "sample_rate": sr, "channels": ch, "frames": int(ref.n_frames),
        "raw_frames_ref": int(ref.frame_is_raw.sum()), "raw_frames_fast": int(fast.frame_is_raw.sum()),
        "raw_decision_flips": raw_flips,
        "rows_compared": int(both_sparse.sum()),
        "kept_ref": int(kept_r.sum()), "kept_fast": int(kept_f.sum()), "kept_either": n_union,
        "kept_fraction_of_bins": round(n_union / max(1, qr.size), 4),
        "membership_flips": int(np.sum(kept_r != kept_f)),
        "flipped_pct": round(100.0 * float(np.sum(dq != 0)) / max(1, n_union), 3),
        "abs_dq_gt1_pct": round(100.0 * float(np.sum(np.abs(dq) > 1)) / max(1, n_union), 3),
        "dq_hist": hist,
        "scale_bits_differ_pct": round(100.0 * float(sb.mean()) if sb.size else 0.0, 2),
        "snr_fast_enc_ref_dec_db": _snr(pcm_ref, pcm_fe) if raw_flips == 0 else None,
        "snr_ref_enc_fast_dec_db": _snr(pcm_ref, pcm_fd),
        "snr_vs_input_ref_db": vs_input(pcm_ref), "snr_vs_input_fast_db": vs_input(pcm_ff),
        "decoded_len_equal_input": bool(len(pcm_ff) == len(x) and len(pcm_fe) == len(x)),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=4.0)
    args = ap.parse_args()
    from gapless_lossy_codec_b200.codec import Context

    fast_ctx = Context(0, mode=1)
    rows = [report_one(n, x, ch, sr, fast_ctx) for n, x, ch, sr in classes(args.seconds)]
    print(json.dumps({
        "what": "FAST (FFT true MDCT, CUDA) vs the reference transform (f32 cosine table, sequential sums; CPU oracle)",
        "why_they_differ": "the reference's table is up to 6.85e-4 (22 quantiser steps) away from the true MDCT basis "
                           "(SURVEY.md section 0 F2): index flips are a property of the reference, not of the FFT",
        "seconds_per_class": args.seconds, "classes": rows}, indent=1))
    fast_ctx.close()


if __name__ == "__main__":
    main()
