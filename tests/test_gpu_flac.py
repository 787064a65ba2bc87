"""GPU parity for the FLAC path: bytes from the CUDA encoder must equal the oracle's bytes, and must
decode losslessly (CRC-8, CRC-16, MD5 verified) with the independent RFC 9639 decoder."""
import hashlib

import numpy as np
import pytest

import oracle
import signals

pytestmark = pytest.mark.gpu


def _to_i16(x):
    return np.trunc(np.clip(x.astype(np.float32) * np.float32(32767.0), -32768.0, 32767.0)).astype(np.int16)


def _order_for(level, bs):
    if level == 0:
        return 0
    want = {1: 1, 2: 2, 3: 3, 4: 3}.get(level, 4)
    return want if bs >= want else 0


def _check(gpu_ctx, x, sr, ch, level, what):
    from gapless_lossy_codec_b200 import flac

    got = flac.encode_flac_with_level(x, sr, ch, level, gpu_ctx)
    ref = oracle.flac_encode(x, sr, ch, level)
    if got != ref:
        n = min(len(got), len(ref))
        first = next((i for i in range(n) if got[i] != ref[i]), n)
        raise AssertionError(f"{what}: FLAC bytes differ: len {len(got)} vs {len(ref)}, first diff at {first}")
    # Independent decode.  One reference quirk makes a stream undecodable: a tail block whose size
    # equals the predictor order has an empty first partition, and the reference then writes no Rice
    # parameter at all (src/flac.rs:632-635) although the format requires one.  Byte parity is still
    # required there; decodability is not the reference's property in that case.
    total = len(x) // ch
    bs = min(1152 if level <= 2 else 4096, total)
    tail = total % bs
    if tail and _order_for(level, tail) == tail:
        return got
    info = oracle.flac_decode(got)
    assert info["sample_rate"] == sr and info["channels"] == ch
    assert np.array_equal(info["samples"], _to_i16(x)[: total * ch].astype(np.int32)), what
    # the MD5 covers every converted sample, including a trailing partial sample frame (:1004)
    assert info["md5_ok"] == (len(x) % ch == 0)
    assert info["md5"] == hashlib.md5(_to_i16(x).tobytes()).digest()
    return got


def _noise_u32(n, seed=12345):
    """tests/test_flac.rs:79-90"""
    out = np.empty(n, np.float32)
    s = seed
    for i in range(n):
        s = (s * 1103515245 + 12345) & 0xFFFFFFFF
        out[i] = np.float32(((s >> 16) & 0x7FFF) / 32768.0) * np.float32(2.0) - np.float32(1.0)
    return out


REF_CASES = [
    ("silence", lambda: np.zeros(1000, np.float32), 44100, 1),
    ("dc", lambda: np.full(1000, 0.5, np.float32), 44100, 1),
    ("sine", lambda: signals.sine(440, 44100, 1, 0.1, amp=0.8), 44100, 1),
    ("noise", lambda: _noise_u32(8820), 44100, 1),
    ("stereo", lambda: np.stack([signals.sine(440, 44100, 1, 0.1), signals.sine(880, 44100, 1, 0.1)], 1).reshape(-1),
     44100, 2),
    ("48k", lambda: np.zeros(4800, np.float32), 48000, 1),
    ("96k", lambda: np.zeros(9600, np.float32), 96000, 1),
    ("min16", lambda: (np.arange(16, dtype=np.float32) / np.float32(16.0)) * np.float32(2.0) - np.float32(1.0), 8000, 1),
]


@pytest.mark.parametrize("name,gen,sr,ch", REF_CASES, ids=[c[0] for c in REF_CASES])
def test_reference_flac_cases(gpu_ctx, name, gen, sr, ch):
    """tests/test_flac.rs:55-133 at the default level 5"""
    _check(gpu_ctx, gen(), sr, ch, 5, name)


@pytest.mark.parametrize("level", range(9))
def test_all_levels(gpu_ctx, level):
    """tests/test_flac.rs:136-159 + a multi-block stereo input"""
    _check(gpu_ctx, signals.sine(440, 44100, 1, 1000 / 44100, amp=0.5)[:1000], 44100, 1, level, f"level{level} short")
    _check(gpu_ctx, signals.music_like(44100, 2, 0.6), 44100, 2, level, f"level{level} music")


@pytest.mark.parametrize("n", [16, 17, 19, 20, 100, 1151, 1152, 1153, 4095, 4096, 4097, 4100, 8192 + 3])
@pytest.mark.parametrize("level", [0, 2, 5, 8])
def test_ragged_tail_blocks(gpu_ctx, n, level):
    """tail block sizes (uncommon block-size codes, tiny partitions, verbatim fallback)"""
    x = signals.white_noise(44100, 1, 1.0, 99)[:n] * np.float32(2.0)
    _check(gpu_ctx, x, 44100, 1, level, f"n{n} level{level}")


def test_multichannel_and_odd_rates(gpu_ctx):
    _check(gpu_ctx, signals.music_like(48000, 6, 0.3, seed=3), 48000, 6, 8, "5.1")
    _check(gpu_ctx, signals.sine(300, 11025, 1, 0.5), 11025, 1, 5, "rate code 0")
    _check(gpu_ctx, signals.sine(300, 44100, 2, 0.2)[:-1], 44100, 2, 5, "length not a multiple of channels")


@pytest.mark.parametrize("ch", [3, 4, 5, 7, 8])
def test_channel_counts_up_to_eight(gpu_ctx, ch):
    """every channel count the frame header can carry (src/flac.rs:806-812), at the largest block size (level 8:
    the bit buffer of wide frames leaves shared memory) and at level 2 (1 152-sample blocks), ragged tails"""
    x = signals.music_like(48000, ch, 0.35, seed=70 + ch)[: (16000 + ch) * ch]
    _check(gpu_ctx, x, 48000, ch, 8, f"{ch} channels level 8")
    _check(gpu_ctx, x[: 5000 * ch], 48000, ch, 2, f"{ch} channels level 2")


def test_extreme_residuals_long_unary_runs(gpu_ctx):
    """alternating full-scale samples: |r| near 2^19, unary runs of hundreds of zeros"""
    x = np.tile(np.array([1.0, -1.0], np.float32), 5000)
    x[::7] = 0.0
    _check(gpu_ctx, x, 44100, 1, 8, "alternating")
    sparse = np.zeros(9000, np.float32)
    sparse[4500] = 1.0  # one huge residual in a partition whose mean is ~0 -> k=0, run of ~2^17 zeros
    _check(gpu_ctx, sparse, 44100, 1, 8, "impulse")


def test_errors_mirror_reference(gpu_ctx):
    from gapless_lossy_codec_b200 import GlcError, flac

    with pytest.raises(GlcError) as e:
        flac.encode_flac(np.zeros(15, np.float32), 44100, 1, gpu_ctx)
    assert e.value.status == 4 and "at least 16 samples" in e.value.message
    with pytest.raises(GlcError) as e:
        flac.encode_flac_with_level(np.zeros(100, np.float32), 44100, 1, 9, gpu_ctx)
    assert e.value.status == 5
    with pytest.raises(GlcError) as e:  # both wrong: the length check comes first (src/flac.rs:963 before :972)
        flac.encode_flac_with_level(np.zeros(10, np.float32), 44100, 1, 9, gpu_ctx)
    assert e.value.status == 4


def test_batch_and_large_lossless(gpu_ctx):
    from gapless_lossy_codec_b200 import flac

    files = [signals.music_like(44100, 2, 0.5), signals.sine(440, 96000, 1, 0.2), np.zeros(100, np.float32)]
    outs = flac.encode_flac_batch(files, [44100, 96000, 8000], [2, 1, 1], 8, gpu_ctx)
    for f, sr, ch, o in zip(files, [44100, 96000, 8000], [2, 1, 1], outs):
        assert o == oracle.flac_encode(f, sr, ch, 8)
    # BASELINE config-5 shape at reduced length: 96 kHz stereo, "24-bit" source, level 8 -> lossless
    # 16-bit round trip + MD5 (size-independent property)
    rng = np.random.default_rng(11)
    n = 96000 * 20
    t = np.arange(n) / 96000.0
    src24 = (np.sin(2 * np.pi * 441.0 * t)[:, None] * 0.4 * 8388608 + rng.integers(-2000, 2000, (n, 2))).astype(np.int32)
    x = (src24.astype(np.float32) / np.float32(8388608.0)).reshape(-1)
    out = flac.encode_flac_with_level(x, 96000, 2, 8, gpu_ctx)
    info = oracle.flac_decode(out)
    assert info["md5_ok"] and info["total_samples"] == n and info["n_frames"] == (n + 4095) // 4096
    assert np.array_equal(info["samples"], _to_i16(x).astype(np.int32))
    assert hashlib.md5(_to_i16(x).tobytes()).digest() == info["md5"]


def test_non_finite_and_out_of_range_samples(gpu_ctx):
    """`(s * 32767.0).clamp(-32768.0, 32767.0) as i16` (src/flac.rs:955-958): NaN -> 0, +-inf and anything outside
    [-1, 1] saturate, everything else truncates toward zero.  One saturating conversion instruction on the
    device; bytes must equal the oracle's."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(20000) * 0.7).astype(np.float32)
    x[::97] = np.nan
    x[5::211] = np.inf
    x[9::223] = -np.inf
    x[11::131] = 3.5
    x[13::137] = -1.0000001
    x[17::139] = np.float32(32767.4 / 32767.0)
    x[19::149] = -0.0
    _check(gpu_ctx, x, 44100, 1, 5, "non-finite mono")
    _check(gpu_ctx, x[:19998], 48000, 2, 8, "non-finite stereo")
