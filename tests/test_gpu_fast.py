"""FAST transform mode (GLC_MODE_FAST): FFT-based true MDCT/IMDCT fused with the quantiser.

Parity class: TOLERANCE.  The reference's transform is "whatever its f32 cosine table says"
(max 6.85e-4 away from the true MDCT basis), so a true-MDCT kernel cannot
reproduce its quantised indices; what it must reproduce is everything structural (frame counts,
gapless metadata, exact sample counts, stream layout) and the signal to within the stated
tolerances below.  EXACT mode (tests/test_gpu_codec.py) carries the bit-exact claims.

Tolerances (asserted here, reported by tools/fast_report.py):
  * decoded PCM of  FAST-encode -> reference-decode   vs  reference-encode -> reference-decode : SNR >= 30 dB
  * decoded PCM of  reference-encode -> FAST-decode    vs  reference-encode -> reference-decode : SNR >= 45 dB
  * true-MDCT check on a stationary sine: the dominant coefficient index equals numpy's float64 MDCT
    and its dequantised value is within 2e-3 relative
"""
import numpy as np
import pytest

import oracle
import signals
from parity import to_oracle, to_product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fast_ctx():
    from gapless_lossy_codec_b200.codec import Context

    ctx = Context(0, mode=1)
    yield ctx
    ctx.close()


def _snr(ref, got):
    ref = ref.astype(np.float64)
    err = ref - got.astype(np.float64)
    p, n = float(np.sum(ref * ref)), float(np.sum(err * err))
    return float("inf") if n == 0 else 10 * np.log10(max(p, 1e-300) / n)


CASES = [
    ("sine440_mono", lambda: signals.sine(440, 44100, 1, 1.0), 1, 44100),
    ("music_stereo", lambda: signals.music_like(44100, 2, 2.0), 2, 44100),
    ("sweep_6ch_48k", lambda: signals.sweep(100, 8000, 48000, 6, 0.5), 6, 48000),
    ("square_mono", lambda: signals.square(440, 44100, 1, 0.7), 1, 44100),
    ("ragged_len", lambda: signals.sine(300, 44100, 1, 0.3)[:4097], 1, 44100),
    ("three_channels", lambda: signals.music_like(44100, 3, 0.5), 3, 44100),
    ("ten_channels", lambda: signals.music_like(44100, 10, 0.4, seed=77), 10, 44100),
    # wider than one round of 8 channels AND with raw frames: the COUNT rounds park per-row tables that the
    # raw decision must clear (a race here left nnz of the last channels set; found by the wave-cut test below)
    ("ten_channels_raw_frames", lambda: signals.music_like(44100, 10, 1.2, seed=5), 10, 44100),
    ("seventeen_channels_raw_frames", lambda: signals.music_like(44100, 17, 0.9, seed=3), 17, 44100),
]
# every group geometry of the fused kernel (frames per group = max(1, 8 / ch); ch > 8: COUNT / EMIT rounds), each
# with tonal AND raw frames
CASES += [(f"{ch}_channels_raw_and_sparse", (lambda ch=ch: signals.music_like(44100, ch, 0.85, seed=40 + ch)), ch, 44100)
          for ch in (4, 5, 7, 8, 9, 16)]


@pytest.mark.parametrize("name,gen,ch,sr", CASES, ids=[c[0] for c in CASES])
def test_fast_structure_and_tolerance(fast_ctx, name, gen, ch, sr):
    from gapless_lossy_codec_b200 import Decoder, Encoder

    x = gen()
    ref = oracle.encode(x, ch, sr)
    enc = Encoder(sr, fast_ctx).encode(x, ch)
    # structure: identical by construction
    for f in ("sample_rate", "channels", "total_samples", "encoder_delay", "padding", "original_length", "n_frames"):
        assert int(getattr(enc, f)) == int(getattr(ref, f)), f
    assert np.array_equal(enc.pair_offset[1:] - enc.pair_offset[:-1], enc.nnz)
    assert all(np.all(np.diff(enc.pair_idx[int(a):int(b)].astype(int)) > 0)
               for a, b in zip(enc.pair_offset[:-1], enc.pair_offset[1:]))  # ascending indices per row
    # raw/sparse decisions may only differ on frames that sit on the decision boundary
    flips = int(np.sum(enc.frame_is_raw != ref.frame_is_raw))
    assert flips <= max(1, ref.n_frames // 50), f"{flips} raw/sparse decisions differ"
    pcm_ref = oracle.decode(ref)
    if flips == 0:
        pcm_a = oracle.decode(to_oracle(enc))  # FAST encoder, reference decoder
        assert len(pcm_a) == len(x)
        assert _snr(pcm_ref, pcm_a) >= 30.0, _snr(pcm_ref, pcm_a)
    dec = Decoder(ch, sr, fast_ctx)
    pcm_b = dec.decode(to_product(ref))  # reference encoder, FAST decoder
    assert len(pcm_b) == len(x)
    assert _snr(pcm_ref, pcm_b) >= 45.0, _snr(pcm_ref, pcm_b)
    pcm_c = dec.decode(enc)  # FAST round trip: gapless length exact
    assert len(pcm_c) == len(x)
    un = dec.decode_untrimmed(enc)
    assert len(un) == (enc.n_frames + 1) * 1024 * ch


def test_fast_is_the_true_mdct(fast_ctx):
    """stationary sine: compare the dominant coefficient of a mid-stream frame with numpy's float64 MDCT."""
    from gapless_lossy_codec_b200 import Encoder

    sr, f0 = 44100, 1000.0
    x = signals.sine(f0, sr, 1, 0.5)
    enc = Encoder(sr, fast_ctx).encode(x, 1)
    fr = 10
    a, b = int(enc.pair_offset[fr]), int(enc.pair_offset[fr + 1])
    idx, q = enc.pair_idx[a:b].astype(int), enc.pair_q[a:b].astype(np.float64)
    coef = q / 32768.0 * float(enc.scales[fr])
    # float64 reference of the same frame
    N = 1024
    padded = np.concatenate([np.zeros(512), x.astype(np.float64), np.zeros(4096)])
    blk = padded[fr * N: fr * N + 2 * N]
    w = np.sin(np.pi * (np.arange(2 * N) + 0.5) / (2 * N))
    n = np.arange(2 * N)
    k = np.arange(N)
    X = (np.cos(np.pi / N * np.outer(k + 0.5, n + 0.5 + N / 2)) @ (blk * w)) * np.sqrt(2.0 / N)
    kmax = int(np.argmax(np.abs(X)))
    assert idx[np.argmax(np.abs(coef))] == kmax
    got = coef[np.argmax(np.abs(coef))]
    assert abs(got - X[kmax]) <= 2e-3 * abs(X[kmax]), (got, X[kmax])
    # every kept coefficient is within one quantiser step (+ float error) of the true value
    step = float(enc.scales[fr]) / 32768.0
    assert np.max(np.abs(coef - X[idx])) <= 1.01 * step


def test_fast_batch_and_gapless_sum(fast_ctx):
    from gapless_lossy_codec_b200 import Decoder, Encoder

    files = [signals.sine(440, 44100, 1, 2.0), signals.sine(880, 44100, 1, 1.3), signals.music_like(44100, 2, 1.1),
             signals.white_noise(44100, 2, 0.4, 7)]
    chs = [1, 1, 2, 2]
    batch = Encoder(44100, fast_ctx).encode_batch(files, chs)
    outs = Decoder(1, 44100, fast_ctx).decode_batch(batch)
    assert [len(o) for o in outs] == [len(f) for f in files]
    # white noise must take the raw-PCM path in both modes
    assert batch[3].frame_is_raw.all()


def test_flac_and_container_are_mode_independent(fast_ctx):
    from gapless_lossy_codec_b200 import Encoder, encoded_from_bytes, encoded_to_bytes, flac

    x = signals.music_like(44100, 2, 0.4)
    assert flac.encode_flac_with_level(x, 44100, 2, 8, fast_ctx) == oracle.flac_encode(x, 44100, 2, 8)
    enc = Encoder(44100, fast_ctx).encode(x, 2)
    blob = encoded_to_bytes(enc, fast_ctx)
    assert blob == oracle.bincode_serialize(to_oracle(enc))
    back = encoded_from_bytes(blob, fast_ctx)
    assert np.array_equal(back.pair_q, enc.pair_q) and np.array_equal(back.raw, enc.raw)


@pytest.mark.parametrize("wave_rows", [8, 48, 640])
def test_fast_wave_size_does_not_change_bits(fast_ctx, wave_rows):
    """The FAST encode pipeline is cut into waves too (host-bound encodes keep two waves of compact output in a
    device ring and place each wave relative to its first offset): any cut gives the same stream, bit for bit,
    as the automatic one -- pairs, raw bodies, offsets -- on a signal with raw and sparse frames, and the decode
    of many small waves equals the decode of one."""
    from gapless_lossy_codec_b200 import Decoder, Encoder

    cases = [(signals.music_like(44100, 2, 3.0), 2, 44100), (signals.music_like(48000, 6, 1.2, seed=9), 6, 48000),
             (signals.music_like(44100, 10, 1.2, seed=5), 10, 44100)]
    for x, ch, sr in cases:
        auto = Encoder(sr, fast_ctx).encode(x, ch)
        pcm_auto = Decoder(ch, sr, fast_ctx).decode(auto)
        try:
            fast_ctx.set_tuning(0, wave_rows)
            cut = Encoder(sr, fast_ctx).encode(x, ch)
            pcm_cut = Decoder(ch, sr, fast_ctx).decode(cut)
        finally:
            fast_ctx.set_tuning(0, 0)
        a, b = to_oracle(auto), to_oracle(cut)
        assert a.n_frames == b.n_frames and np.array_equal(a.frame_is_raw, b.frame_is_raw)
        assert a.frame_is_raw.any() and not a.frame_is_raw.all(), "the case must mix raw and sparse frames"
        assert np.array_equal(a.nnz, b.nnz) and np.array_equal(a.pair_idx, b.pair_idx) and np.array_equal(a.pair_q, b.pair_q)
        assert np.array_equal(a.scales.view(np.uint32), b.scales.view(np.uint32))
        assert np.array_equal(a.raw, b.raw)
        assert np.array_equal(pcm_auto.view(np.uint32), pcm_cut.view(np.uint32))
