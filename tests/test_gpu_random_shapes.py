"""Seeded differential sweep: random shapes through the C ABI against the oracle, bit for bit.

The hand-picked cases of test_gpu_codec.py / test_gpu_flac.py follow the reference's own tests; this file walks
the product of the axes they leave between them -- channel counts 1..12, lengths from the 513-sample minimum up to
a few frames past a group or tile boundary, the sample rates the perceptual model distinguishes, signals that mix
silence, tones, full-scale noise (raw frames) and clipping -- with a fixed seed, so a failure names its case."""
import numpy as np
import pytest

import oracle
import signals
from parity import assert_encoded_equal, assert_pcm_bits_equal, to_oracle, to_product

pytestmark = pytest.mark.gpu

RATES = [8000, 11025, 16000, 22050, 32000, 44100, 48000, 88200, 96000, 192000]


def _signal(rng, n, ch):
    """n sample frames, ch channels, interleaved: segments of silence / tone / noise / clipped tone per channel"""
    out = np.zeros((n, ch), np.float32)
    for c in range(ch):
        pos = 0
        while pos < n:
            seg = int(rng.integers(200, 5000))
            kind = int(rng.integers(0, 5))
            t = np.arange(pos, min(n, pos + seg), dtype=np.float64)
            if kind == 1:
                out[pos:pos + seg, c] = (rng.uniform(0.05, 0.9) * np.sin(t * rng.uniform(0.005, 1.5))).astype(np.float32)
            elif kind == 2:
                out[pos:pos + seg, c] = rng.uniform(-1.0, 1.0, len(t)).astype(np.float32)
            elif kind == 3:
                out[pos:pos + seg, c] = np.clip(1.7 * np.sin(t * rng.uniform(0.01, 0.3)), -1.0, 1.0).astype(np.float32)
            elif kind == 4:
                out[pos:pos + seg, c] = (1e-4 * np.sin(t * 0.2)).astype(np.float32)  # far below the noise floor
            pos += seg
    return out.reshape(-1)


def _codec_cases():
    rng = np.random.default_rng(20240611)
    cases = []
    for i in range(48):
        ch = int(rng.integers(1, 13))
        n = int(rng.choice([513, 514, 1023, 1024, 1025, 1536, 2047, 2049, int(rng.integers(513, 9000)),
                            int(rng.integers(9000, 40000 // ch + 9001))]))
        cases.append((i, ch, n, int(rng.choice(RATES)), int(rng.integers(0, 2 ** 31))))
    return cases


@pytest.mark.parametrize("i,ch,n,sr,seed", _codec_cases(), ids=lambda v: str(v))
def test_random_codec_shapes_bit_exact(gpu_ctx, i, ch, n, sr, seed):
    from gapless_lossy_codec_b200 import Decoder, Encoder

    x = _signal(np.random.default_rng(seed), n, ch)
    what = f"case {i}: {ch} ch, {n} sample frames, {sr} Hz"
    ref = oracle.encode(x, ch, sr)
    enc = Encoder(sr, gpu_ctx).encode(x, ch)
    assert_encoded_equal(enc, ref, what)
    dec = Decoder(ch, sr, gpu_ctx)
    pcm = dec.decode(enc)
    assert_pcm_bits_equal(pcm, oracle.decode(ref), what + " decode")
    assert len(pcm) == len(x)
    assert_pcm_bits_equal(dec.decode_untrimmed(enc), oracle.decode(ref, trimmed=False), what + " untrimmed")


def test_random_batch_of_mixed_files(gpu_ctx):
    """the same kind of files, 24 of them with different channel counts in ONE batch call (file boundaries fall
    inside row tiles, frame groups and waves), then one decode batch"""
    from gapless_lossy_codec_b200 import Decoder, Encoder

    rng = np.random.default_rng(77)
    files, chs = [], []
    for _ in range(24):
        ch = int(rng.integers(1, 9))
        files.append(_signal(rng, int(rng.integers(513, 12000)), ch))
        chs.append(ch)
    try:
        gpu_ctx.set_tuning(0, 96)  # many small waves
        encs = Encoder(44100, gpu_ctx).encode_batch(files, chs)
        pcms = Decoder(2, 44100, gpu_ctx).decode_batch(encs)
    finally:
        gpu_ctx.set_tuning(0, 0)
    for k, (x, ch, e, p) in enumerate(zip(files, chs, encs, pcms)):
        ref = oracle.encode(x, ch, 44100)
        assert_encoded_equal(e, ref, f"file {k} ({ch} ch, {len(x) // ch} frames)")
        assert_pcm_bits_equal(p, oracle.decode(ref), f"file {k} decode")


def _flac_cases():
    rng = np.random.default_rng(424242)
    cases = []
    for i in range(48):
        ch = int(rng.integers(1, 9))
        level = int(rng.integers(0, 9))
        n = int(rng.choice([16, 17, 31, 1151, 1152, 1153, 2304, 4095, 4096, 4097, int(rng.integers(16, 3000)),
                            int(rng.integers(3000, 30000))]))
        cases.append((i, ch, n, level, int(rng.choice(RATES)), int(rng.integers(0, 2 ** 31))))
    return cases


@pytest.mark.parametrize("i,ch,n,level,sr,seed", _flac_cases(), ids=lambda v: str(v))
def test_random_flac_shapes_byte_exact(gpu_ctx, i, ch, n, level, sr, seed):
    from gapless_lossy_codec_b200 import flac

    x = _signal(np.random.default_rng(seed), n, ch)
    got = flac.encode_flac_with_level(x, sr, ch, level, gpu_ctx)
    want = oracle.flac_encode(x, sr, ch, level)
    assert got == want, f"case {i}: {ch} ch, {n} sample frames, level {level}, {sr} Hz: {len(got)} vs {len(want)} bytes"
    info = oracle.flac_decode(got)
    assert info["md5_ok"] and info["total_samples"] == n


def test_random_fast_mode_structure(gpu_ctx):
    """FAST mode on the same random files: tables consistent, raw rows empty, indices ascending, exact lengths"""
    from gapless_lossy_codec_b200 import Decoder, Encoder
    from gapless_lossy_codec_b200.codec import Context

    ctx = Context(0, mode=1)
    try:
        rng = np.random.default_rng(99)
        for k in range(24):
            ch = int(rng.integers(1, 13))
            x = _signal(rng, int(rng.integers(513, 20000)), ch)
            e = to_oracle(Encoder(48000, ctx).encode(x, ch))
            ref = oracle.encode(x, ch, 48000)
            assert e.n_frames == ref.n_frames and e.total_samples == ref.total_samples and e.padding == ref.padding
            assert np.array_equal(e.pair_offset[1:] - e.pair_offset[:-1], e.nnz), f"file {k} ({ch} ch)"
            raw_rows = np.repeat(e.frame_is_raw.astype(bool), ch)
            assert not e.nnz[raw_rows].any(), f"file {k} ({ch} ch): a raw frame carries pairs"
            assert e.raw_offset[-1] == int(e.frame_is_raw.sum()) * 2048 * ch
            for a, b in zip(e.pair_offset[:-1], e.pair_offset[1:]):
                assert np.all(np.diff(e.pair_idx[int(a):int(b)].astype(int)) > 0)
            pcm = Decoder(ch, 48000, ctx).decode(to_product(e))
            assert len(pcm) == len(x)
    finally:
        ctx.close()


def test_random_streams_through_the_other_decode_entry_points(gpu_ctx):
    """the same decode through its other doors: the chunk stream (exactly 500 frames per chunk, concatenation ==
    untrimmed stream), 16-bit PCM output (the WAV export's conversion, src/audio.rs:11-16) and decode -> FLAC on
    the device (bytes of the oracle chain), on random multi-chunk streams with raw and sparse frames"""
    from gapless_lossy_codec_b200 import Decoder

    rng = np.random.default_rng(5150)
    for k in range(6):
        ch = int(rng.integers(1, 7))
        sr = int(rng.choice([22050, 44100, 48000]))
        n = int(rng.integers(513, 40000)) if k else 1024 * 1003 // ch + 17  # case 0: more than two chunks for ch = 1
        x = _signal(rng, n, ch)
        ref = oracle.encode(x, ch, sr)
        enc = to_product(ref)
        dec = Decoder(ch, sr, gpu_ctx)
        chunks = list(dec.decode_streaming(enc))
        assert all(len(c.samples) == 500 * 1024 * ch for c in chunks[:-1]) and chunks[-1].is_last
        assert not any(c.is_last for c in chunks[:-1])
        cat = np.concatenate([c.samples for c in chunks])
        un_ref = oracle.decode(ref, trimmed=False)
        assert_pcm_bits_equal(cat, un_ref, f"stream {k} ({ch} ch, {ref.n_frames} frames) chunk concat")
        pcm_ref = oracle.decode(ref)
        want16 = np.trunc(np.clip(pcm_ref.astype(np.float32) * np.float32(32767.0), -32768.0, 32767.0)).astype(np.int16)
        assert np.array_equal(dec.decode_pcm16(enc), want16), f"stream {k}: 16-bit output"
        level = int(rng.integers(0, 9))
        assert dec.decode_to_flac(enc, level) == oracle.flac_encode(pcm_ref, sr, ch, level), f"stream {k}: decode -> FLAC level {level}"


def test_mutated_containers_are_rejected_or_decoded_like_the_reference(gpu_ctx):
    """.glc images are untrusted input (load_encoded, src/codec.rs:781-786): 600 seeded mutations of valid images
    -- truncations, byte flips, length fields overwritten with small / huge values -- must either be refused with
    an error or parse into a stream whose decode equals the oracle's decode of the same parsed arrays (the
    reference's rules for duplicate / out-of-range indices and short raw bodies, src/codec.rs:626-665).  Never a
    crash, never a huge allocation."""
    from gapless_lossy_codec_b200 import Decoder, Encoder, GlcError, encoded_from_bytes, encoded_to_bytes

    rng = np.random.default_rng(8086)
    images = []
    for ch, n in ((1, 3000), (2, 5000), (5, 2100)):
        x = _signal(rng, n, ch)
        images.append(encoded_to_bytes(Encoder(44100, gpu_ctx).encode(x, ch), gpu_ctx))
    accepted = refused = 0
    for it in range(600):
        img = bytearray(images[it % len(images)])
        kind = it % 4
        if kind == 0:
            img = img[: int(rng.integers(0, len(img)))]
        elif kind == 1:
            for _ in range(int(rng.integers(1, 5))):
                img[int(rng.integers(0, len(img)))] ^= int(rng.integers(1, 256))
        elif kind == 2:
            pos = int(rng.integers(0, max(1, len(img) - 8)))
            vals = [0, 1, 2, 7, 255, 65536, 2 ** 31, 2 ** 40, 2 ** 63 - 1, 2 ** 64 - 1]
            val = vals[int(rng.integers(0, len(vals)))]
            img[pos:pos + 8] = val.to_bytes(8, "little")
        else:
            pos = int(rng.integers(0, max(1, len(img) - 8)))
            img[pos:pos + 8] = bytes(rng.integers(0, 256, 8, dtype=np.uint8))
            img = img[: int(rng.integers(pos, len(img) + 1))]
        try:
            parsed = encoded_from_bytes(bytes(img), gpu_ctx)
        except GlcError:
            refused += 1
            continue
        try:
            pcm = Decoder(int(parsed.channels), int(parsed.sample_rate), gpu_ctx).decode(parsed)
        except GlcError:
            refused += 1
            continue
        accepted += 1
        if parsed.n_frames * max(1, int(parsed.channels)) <= 4096:  # keep the oracle's share of the test small
            assert_pcm_bits_equal(pcm, oracle.decode(to_oracle(parsed)), f"mutation {it} (kind {kind})")
    assert accepted >= 20 and refused >= 200, (accepted, refused)
