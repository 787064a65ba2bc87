"""Two restatements of the reference's codec, written separately from the Rust source -- the C oracle
(oracle/codec_oracle.c) and a numpy one (oracle/codec_restatement_np.py) -- must agree bit for bit:
streams (flags, indices, quantised values, scale bit patterns, raw bodies, gapless metadata) and
decoded PCM.  This does not pin the oracle to the reference binary (no Rust toolchain here), but it
rules out slips of transcription in either restatement."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
import signals  # noqa: E402
from oracle import codec_restatement_np as npr  # noqa: E402


def _cases():
    rng = np.random.default_rng(42)
    return [
        ("sine_mono_44k", signals.sine(440, 44100, 1, 0.12), 1, 44100),
        ("music_noise_stereo_44k", np.concatenate([signals.music_like(44100, 2, 0.1),
                                                   signals.white_noise(44100, 2, 0.08, 9)]), 2, 44100),
        ("sweep_3ch_48k", signals.sweep(200, 9000, 48000, 3, 0.07), 3, 48000),
        ("loud_clipping_mono_8k", (rng.standard_normal(1500) * 1.5).astype(np.float32), 1, 8000),
    ]


def test_tables_agree():
    t = npr.MdctTables.get()
    cos_tab, window, norm = oracle.tables()
    assert np.array_equal(t.cos_table.view(np.uint32).ravel(), np.asarray(cos_tab, np.float32).view(np.uint32).ravel())
    assert np.array_equal(t.window.view(np.uint32), np.asarray(window, np.float32).view(np.uint32))
    assert np.float32(t.norm).view(np.uint32) == np.float32(norm).view(np.uint32) == np.uint32(0x3D3504F3)


@pytest.mark.parametrize("sr", [8000, 44100, 48000, 96000])
def test_perceptual_model_agrees(sr):
    p = npr.PerceptualWeights(1024, sr)
    weights, bands = oracle.perceptual(sr)
    assert list(bands) == p.critical_bands
    assert np.array_equal(np.asarray(weights, np.float32).view(np.uint32), p.weights.view(np.uint32))


@pytest.mark.parametrize("name,x,ch,sr", _cases(), ids=[c[0] for c in _cases()])
def test_encode_and_decode_agree(name, x, ch, sr):
    a = npr.encode(x, ch, sr)
    b = oracle.encode(x, ch, sr)
    assert (a.sample_rate, a.channels, a.total_samples) == (b.sample_rate, b.channels, b.total_samples)
    assert (a.encoder_delay, a.padding, a.original_length) == (b.encoder_delay, b.padding, b.original_length)
    assert len(a.frames) == b.n_frames
    kinds = set()
    for f, fr in enumerate(a.frames):
        assert (fr.raw_pcm is not None) == bool(b.frame_is_raw[f]), f"frame {f} raw flag"
        kinds.add(fr.raw_pcm is not None)
        if fr.raw_pcm is not None:
            assert np.array_equal(fr.raw_pcm, b.raw[int(b.raw_offset[f]):int(b.raw_offset[f + 1])]), f"frame {f} raw body"
            continue
        for c in range(ch):
            row = f * ch + c
            lo, hi = int(b.pair_offset[row]), int(b.pair_offset[row + 1])
            assert [k for k, _ in fr.sparse[c]] == b.pair_idx[lo:hi].tolist(), f"row {row} indices"
            assert [q for _, q in fr.sparse[c]] == b.pair_q[lo:hi].tolist(), f"row {row} values"
            assert np.float32(fr.scales[c]).view(np.uint32) == b.scales[row:row + 1].view(np.uint32)[0], f"row {row} scale"
    if name.startswith("music_noise"):
        assert kinds == {True, False}, "both frame kinds must occur"
    pa = npr.decode(a)
    pb = oracle.decode(b, literal_imdct=True)
    assert pa.size == pb.size == x.size
    assert np.array_equal(pa.view(np.uint32), np.asarray(pb, np.float32).view(np.uint32)), "decoded PCM bits"


# ---------------------------------------------------------------------------- FLAC

from oracle import flac_restatement_py as fpy  # noqa: E402


def _flac_cases():
    rng = np.random.default_rng(7)
    return [
        ("sine_mono_44k", signals.sine(440, 44100, 1, 0.12), 44100, 1),               # 5292 samples: 4096 + tail 1196
        ("stereo_48k", signals.sweep(100, 6000, 48000, 2, 0.06), 48000, 2),
        ("noise_loud_clipping", (rng.standard_normal(3000) * 0.9).astype(np.float32), 22050, 1),
        ("six_channels_odd_rate", (rng.standard_normal(6 * 700) * 0.2).astype(np.float32), 37000, 6),
        ("minimum_16_samples", np.linspace(-1, 1, 16, dtype=np.float32), 8000, 1),     # block size 16: 8-bit size code
        ("uncommon_block_300", signals.sine(300, 96000, 1, 300 / 96000), 96000, 1),    # block size 300: 16-bit size code
        ("silence", np.zeros(2000, np.float32), 44100, 2),
    ]


@pytest.mark.parametrize("level", range(9))
def test_flac_restatements_agree(level):
    for name, x, sr, ch in _flac_cases():
        a = fpy.encode_flac_with_level(x, sr, ch, level)
        b = oracle.flac_encode(x, sr, ch, level)
        assert a == b, (name, level, len(a), len(b), next((i for i, (p, q) in enumerate(zip(a, b)) if p != q), None))


def test_flac_crcs_agree():
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 300):
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        assert fpy.crc8(data) == oracle.crc8(data)
        assert fpy.crc16(data) == oracle.crc16(data)


# ------------------------------------------------------------------ .glc container (bincode 1.3 image)


def _bincode_py(e) -> bytes:
    """save_encoded's byte image (src/codec.rs:774-779) from the derive order of the structs (:31-69): bincode 1.3
    defaults = little-endian fixed-width integers, u64 length prefix per Vec, one tag byte per Option."""
    import struct

    ch = e.channels
    out = [struct.pack("<IHQ", e.sample_rate, ch, e.total_samples), struct.pack("<Q", e.n_frames)]
    for f in range(e.n_frames):
        if e.frame_is_raw[f]:
            body = e.raw[int(e.raw_offset[f]):int(e.raw_offset[f + 1])]
            out += [struct.pack("<Q", 0), struct.pack("<Q", 0), b"\x01", struct.pack("<Q", len(body)),
                    body.astype("<i2").tobytes()]
        else:
            out.append(struct.pack("<Q", ch))
            for c in range(ch):
                lo, hi = int(e.pair_offset[f * ch + c]), int(e.pair_offset[f * ch + c + 1])
                out.append(struct.pack("<Q", hi - lo))
                pr = np.empty(hi - lo, dtype=[("i", "<u2"), ("q", "<i2")])
                pr["i"], pr["q"] = e.pair_idx[lo:hi], e.pair_q[lo:hi]
                out.append(pr.tobytes())
            out += [struct.pack("<Q", ch), e.scales[f * ch:(f + 1) * ch].astype("<f4").tobytes(), b"\x00"]
    out.append(struct.pack("<IIQ", e.encoder_delay, e.padding, e.original_length))
    return b"".join(out)


def test_bincode_image_agrees():
    x = np.concatenate([signals.music_like(44100, 2, 0.3), signals.white_noise(44100, 2, 0.2, 4)])
    e = oracle.encode(x, 2, 44100)
    assert e.frame_is_raw.any() and not e.frame_is_raw.all()
    assert _bincode_py(e) == oracle.bincode_serialize(e)
