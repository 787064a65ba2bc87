"""Generates tests/golden/golden_v1.npz with the CPU oracle (run in the build container:
`python tests/golden/make_golden.py`).

The reference is a Rust crate that cannot be built or imported here and ships no golden vectors
(SURVEY.md section 4), so these fixtures are ORACLE outputs: they pin both the oracle and the CUDA
path against drift and travel to the GPU box, where /root/reference does not exist.  Inputs are
regenerated from seeds by tests/signals.py; only outputs are stored.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import signals  # noqa: E402

CODEC_CASES = {
    # name: (generator, channels, sample_rate)
    "sine440_mono": (lambda: signals.sine(440, 44100, 1, 0.25), 1, 44100),
    "music_stereo": (lambda: signals.music_like(44100, 2, 0.5), 2, 44100),
    "sweep_6ch_48k": (lambda: signals.sweep(100, 8000, 48000, 6, 0.1), 6, 48000),
    "noise_stereo_raw": (lambda: signals.white_noise(44100, 2, 0.1, 12345), 2, 44100),
}
FLAC_CASES = {
    # name: (generator, sample_rate, channels, level)
    "sine_l5": (lambda: signals.sine(440, 44100, 1, 1000 / 44100, amp=0.8)[:1000], 44100, 1, 5),
    "music_stereo_l8": (lambda: signals.music_like(44100, 2, 0.12), 44100, 2, 8),
    "noise_l0": (lambda: signals.white_noise(44100, 1, 0.05, 99), 44100, 1, 0),
    "tail_l2": (lambda: signals.sine(300, 48000, 2, 0.05)[: 2 * 1153], 48000, 2, 2),
}


def sha(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    out = {}
    cos_tab, window, norm = oracle.tables()
    out["tables/cos_sha256"] = sha(cos_tab)
    out["tables/window_sha256"] = sha(window)
    out["tables/norm_bits"] = np.array([np.float32(norm).view(np.uint32)], np.uint32)
    for name, (gen, ch, sr) in CODEC_CASES.items():
        x = gen()
        e = oracle.encode(x, ch, sr)
        pre = f"codec/{name}/"
        out[pre + "input_sha256"] = sha(x)
        out[pre + "meta"] = np.array([e.sample_rate, e.channels, e.total_samples, e.encoder_delay, e.padding,
                                      e.original_length, e.n_frames], np.uint64)
        for f in ("frame_is_raw", "nnz", "pair_offset", "pair_idx", "pair_q", "raw_offset", "raw"):
            out[pre + f] = getattr(e, f)
        out[pre + "scales_bits"] = e.scales.view(np.uint32)
        pcm = oracle.decode(e)
        out[pre + "pcm_sha256"] = sha(pcm)
        out[pre + "pcm_head_bits"] = pcm[:2048].view(np.uint32)
        out[pre + "pcm_len"] = np.array([len(pcm)], np.uint64)
        out[pre + "bincode_sha256"] = sha(np.frombuffer(oracle.bincode_serialize(e), np.uint8))
    for name, (gen, sr, ch, level) in FLAC_CASES.items():
        x = gen()
        out[f"flac/{name}/input_sha256"] = sha(x)
        out[f"flac/{name}/bytes"] = np.frombuffer(oracle.flac_encode(x, sr, ch, level), np.uint8)
    # known answers that do not depend on this repo at all
    out["kat/md5_abc"] = np.frombuffer(oracle.md5(b"abc"), np.uint8)
    out["kat/crc8_123456789"] = np.array([oracle.crc8(b"123456789")], np.uint32)
    out["kat/crc16_123456789"] = np.array([oracle.crc16(b"123456789")], np.uint32)
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
