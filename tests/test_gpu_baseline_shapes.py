"""GPU parity at the shapes that distinguish BASELINE.json's configs 3, 4 and 5 (all against the oracle,
through the C ABI), plus streams without frames:

  config 5  FLAC frame numbers >= 0x800 and >= 0x10000 (3- and 4-byte UTF-8 forms, src/flac.rs:427-478) and
            the 96 kHz stereo level-8 shape;
  config 4  one batched call over 500 short tracks with the LCG lengths of SURVEY.md 8(d)-4;
  config 3  a five-minute 6-channel 48 kHz batch, spot windows + the channel-rotating 512-value trim;
  zero-frame streams (EncodedAudio{frames: []}): the reference pushes the all-zero final overlap and trims
            (src/codec.rs:723-765).
"""
import hashlib

import numpy as np
import pytest

import oracle
import signals
from parity import assert_encoded_equal, assert_pcm_bits_equal, to_oracle, to_product

pytestmark = pytest.mark.gpu


def _to_i16(x):
    return np.trunc(np.clip(x.astype(np.float32) * np.float32(32767.0), -32768.0, 32767.0)).astype(np.int16)


def _first_diff(a: bytes, b: bytes) -> int:
    n = min(len(a), len(b))
    av, bv = np.frombuffer(a, np.uint8, n), np.frombuffer(b, np.uint8, n)
    d = np.flatnonzero(av != bv)
    return int(d[0]) if len(d) else n


def _tone_noise(n, seed):
    """cheap deterministic mono test signal: two tones + a little LCG noise (small residuals, short codes)"""
    period = 100003  # not a multiple of any block size
    t = np.arange(period, dtype=np.float64)
    x = np.tile(0.3 * np.sin(t * 0.0123) + 0.1 * np.sin(t * 0.171 + 1.0), n // period + 1)[:n]
    h = (np.arange(n, dtype=np.uint32) + np.uint32(seed)) * np.uint32(2654435761)  # multiplicative hash, mod 2^32
    x += ((h >> np.uint32(8)).astype(np.float64) / float(1 << 24) - 0.5) * 0.002
    return x.astype(np.float32)


@pytest.mark.parametrize("n_frames", [2100, 65540], ids=["utf8_3byte_frame_numbers", "utf8_4byte_frame_numbers"])
def test_flac_long_streams_frame_number_forms(gpu_ctx, n_frames):
    """mono, level 1 (1 152-sample blocks, src/flac.rs:983-995): frame numbers run past 0x800 / 0x10000, so
    the header carries the 3- and 4-byte coded numbers and header_bytes() the matching lengths -- a wrong
    header length would shift every later frame.  Bytes equal the oracle's; the stream decodes (frame
    number sequence, CRC-8, CRC-16, MD5 verified by the independent decoder)."""
    from gapless_lossy_codec_b200 import flac

    n = n_frames * 1152 + 77  # ragged tail block
    x = _tone_noise(n, 4242)
    got = flac.encode_flac_with_level(x, 44100, 1, 1, gpu_ctx)
    ref = oracle.flac_encode(x, 44100, 1, 1)
    assert len(got) == len(ref) and got == ref, f"first difference at byte {_first_diff(got, ref)} of {len(ref)}"
    info = oracle.flac_decode(got, frame_offsets=True)
    assert info["n_frames"] == n_frames + 1 and info["md5_ok"] and info["total_samples"] == n
    assert info["md5"] == hashlib.md5(_to_i16(x).tobytes()).digest()
    # the coded number of the last frames really has the long form
    off = info["frame_off"]
    last = got[int(off[n_frames]):int(off[n_frames]) + 8]
    assert last[4] >> 4 == (0xE if n_frames < 0x10000 else 0xF)


def test_flac_config5_shape_96k_stereo_level8(gpu_ctx):
    """BASELINE config 5 at 60 s: 96 kHz stereo from a "24-bit" source (src/audio.rs:51-59 conversion),
    level 8 (4 096-sample blocks, order 4, partition order 6): bytes equal the oracle's."""
    from gapless_lossy_codec_b200 import flac

    sr, ch, secs = 96000, 2, 60
    period = signals.music_like(sr, ch, 10.0, seed=7)
    period = (np.round(period.astype(np.float64) * 8388608.0) / 8388608.0).astype(np.float32)
    x = np.tile(period, secs // 10)
    got = flac.encode_flac_with_level(x, sr, ch, 8, gpu_ctx)
    ref = oracle.flac_encode(x, sr, ch, 8)
    assert len(got) == len(ref) and got == ref, f"first difference at byte {_first_diff(got, ref)} of {len(ref)}"
    info = oracle.flac_decode(got)
    assert info["md5_ok"] and info["n_frames"] == (secs * sr + 4095) // 4096


def _lcg_lengths(n, seed=2024):
    """track length 3 + 7u seconds, u from the 64-bit LCG of tests/utils.rs:96"""
    st = signals.lcg_u64(seed, n)
    return 3.0 + 7.0 * (st.astype(np.float64) / 18446744073709551615.0)


def test_config4_batch_of_500_tracks(gpu_ctx):
    """One glc_encode_batch over 500 tracks of 3-10 s (44.1 kHz stereo), one glc_decode_batch back: 30 sampled
    tracks are bit-equal to the oracle's single-file encode and decode, every track keeps its exact sample
    count and the sum of decoded lengths equals the sum of the originals (tests/test_codec.rs:140-170)."""
    from gapless_lossy_codec_b200 import Decoder, Encoder

    sr, ch, n_tracks = 44100, 2, 500
    lens = (_lcg_lengths(n_tracks) * sr).astype(np.int64)
    bases = [signals.sine(440, sr, ch, 10.2), signals.sweep(100, 8000, sr, ch, 10.2), signals.square(330, sr, ch, 10.2),
             signals.music_like(sr, ch, 10.2, seed=5), signals.sawtooth(220, sr, ch, 10.2)]
    files = [bases[i % len(bases)][(i % 7) * 2 * 311:(i % 7) * 2 * 311 + int(lens[i]) * ch] for i in range(n_tracks)]
    assert all(len(f) == int(lens[i]) * ch for i, f in enumerate(files))
    encs = Encoder(sr, gpu_ctx).encode_batch(files, [ch] * n_tracks)
    pcm = Decoder(ch, sr, gpu_ctx).decode_batch(encs)
    assert [len(p) for p in pcm] == [len(f) for f in files]
    assert sum(len(p) for p in pcm) == sum(len(f) for f in files)
    rng = np.random.default_rng(4)
    picks = sorted(set([0, 1, n_tracks - 1] + [int(v) for v in rng.integers(0, n_tracks, 27)]))
    for i in picks:
        ref = oracle.encode(files[i], ch, sr)
        assert_encoded_equal(encs[i], ref, f"track {i} of the batch")
        assert_pcm_bits_equal(pcm[i], oracle.decode(ref), f"track {i} decode")


def _slice_frames(enc, f0, n):
    ch = enc.channels
    r0, r1 = f0 * ch, (f0 + n) * ch
    p0, p1 = int(enc.pair_offset[r0]), int(enc.pair_offset[r1])
    q0, q1 = int(enc.raw_offset[f0]), int(enc.raw_offset[f0 + n])
    return oracle.EncodedArrays(
        sample_rate=enc.sample_rate, channels=ch, total_samples=n * 1024 * ch, encoder_delay=0, padding=0,
        original_length=(n + 1) * 1024 * ch, n_frames=n, frame_is_raw=enc.frame_is_raw[f0:f0 + n].copy(),
        nnz=enc.nnz[r0:r1].copy(), pair_offset=(enc.pair_offset[r0:r1 + 1] - np.uint64(p0)).astype(np.uint64),
        pair_idx=enc.pair_idx[p0:p1].copy(), pair_q=enc.pair_q[p0:p1].copy(), scales=enc.scales[r0:r1].copy(),
        raw_offset=(enc.raw_offset[f0:f0 + n + 1] - np.uint64(q0)).astype(np.uint64), raw=enc.raw[q0:q1].copy())


def test_config3_five_minutes_of_5_1_spot_windows_and_trim(gpu_ctx):
    """BASELINE config 3 shape at five minutes: 48 kHz, 6 channels.  Windows of a few frames inside the big
    stream equal the oracle's encode of the PCM slice / decode of the sub-stream bit for bit; the trimmed
    output is the untrimmed one minus 512 interleaved VALUES (85 sample frames + 2 channels: the channel
    rotation of src/codec.rs:755-761) cut to original_length, and its head equals the oracle's trimmed decode."""
    from gapless_lossy_codec_b200 import Decoder, Encoder

    ch, sr = 6, 48000
    x = np.tile(signals.music_like(sr, ch, 10.0, seed=1000), 30)  # 300 s: 14 063 frames, 84 378 rows
    L = len(x) // ch
    enc = Encoder(sr, gpu_ctx).encode(x, ch)
    dec = Decoder(ch, sr, gpu_ctx)
    big = dec.decode_untrimmed(enc)
    assert len(big) == (enc.n_frames + 1) * 1024 * ch
    trimmed = dec.decode(enc)
    assert len(trimmed) == len(x)
    assert_pcm_bits_equal(trimmed, big[512:512 + len(x)], "trim = 512 values off the untrimmed stream")
    n = 6
    rng = np.random.default_rng(3)
    trans = np.flatnonzero(np.diff(enc.frame_is_raw.astype(np.int8)) != 0)
    starts = [0, 787, 3945, int(trans[0]) - 2, int(trans[len(trans) // 2]) - 2] + \
        [int(v) for v in rng.integers(10, enc.n_frames - n - 10, 3)]
    eo = to_oracle(enc)
    xs = x.reshape(L, ch)
    for f0 in starts:
        if f0 > 0:
            sl = xs[f0 * 1024:(f0 + n) * 1024 + 512].reshape(-1)
            ref = oracle.encode(sl, ch, sr)
            assert_encoded_equal(_slice_frames(eo, f0 + 1, n - 2), _slice_frames(ref, 1, n - 2),
                                 f"encode window at frame {f0}")
        else:
            # head of the file: frame 0 sees the 512-zero lead-in in both
            ref = oracle.encode(xs[:(n + 1) * 1024].reshape(-1), ch, sr)
            assert_encoded_equal(_slice_frames(eo, 0, n - 1), _slice_frames(ref, 0, n - 1), "encode window at the head")
        sub = _slice_frames(eo, f0, n)
        pcm = oracle.decode(sub, trimmed=False)
        assert_pcm_bits_equal(big[(f0 + 1) * 1024 * ch:(f0 + n) * 1024 * ch], pcm[1024 * ch:n * 1024 * ch],
                              f"decode window at frame {f0}")
    # the trim through the oracle's own Decoder::decode on the head of the stream
    head = _slice_frames(eo, 0, n)
    head.encoder_delay, head.original_length = 512, (n - 1) * 1024 * ch
    assert_pcm_bits_equal(trimmed[:(n - 1) * 1024 * ch], oracle.decode(head), "head of the trimmed stream")


def _empty_stream(ch, sr, original_length, delay=512):
    return oracle.EncodedArrays(
        sample_rate=sr, channels=ch, total_samples=original_length, encoder_delay=delay, padding=0,
        original_length=original_length, n_frames=0, frame_is_raw=np.zeros(0, np.uint8), nnz=np.zeros(0, np.uint32),
        pair_offset=np.zeros(1, np.uint64), pair_idx=np.zeros(0, np.uint16), pair_q=np.zeros(0, np.int16),
        scales=np.zeros(0, np.float32), raw_offset=np.zeros(1, np.uint64), raw=np.zeros(0, np.int16))


def test_streams_without_frames(gpu_ctx):
    """frames = []: decode_streaming sends only the final overlap, 1 024 zeros per channel (src/codec.rs:723-732);
    Decoder::decode drops encoder_delay values and truncates to original_length (:755-765).  Alone, and next
    to an ordinary stream in both batch orders (the empty stream owns one hop of the batch)."""
    from gapless_lossy_codec_b200 import Decoder, Encoder

    dec = Decoder(2, 44100, gpu_ctx)
    # poison the pinned pool first: stale memory must not come back as "decoded" audio
    junk = Encoder(44100, gpu_ctx).encode(signals.white_noise(44100, 2, 0.2, 9), 2)
    dec.decode(junk)
    for ch, olen in [(1, 300), (2, 5000), (2, 0), (6, 10 ** 6)]:
        z = _empty_stream(ch, 44100, olen)
        want = oracle.decode(z)
        got = dec.decode(to_product(z))
        assert_pcm_bits_equal(got, want, f"empty stream ch={ch} original_length={olen}")
        assert not got.any()
        un = dec.decode_untrimmed(to_product(z))
        assert_pcm_bits_equal(un, oracle.decode(z, trimmed=False), f"empty stream untrimmed ch={ch}")
        assert len(un) == 1024 * ch
        chunks = list(dec.decode_streaming(to_product(z)))
        assert len(chunks) == 1 and chunks[0].is_last and len(chunks[0].samples) == 1024 * ch and not chunks[0].samples.any()
    a = Encoder(44100, gpu_ctx).encode(signals.music_like(44100, 2, 0.5), 2)
    a_ref = oracle.decode(to_oracle(a))
    z = _empty_stream(2, 44100, 700)
    z_ref = oracle.decode(z)
    for order in ([z, a], [a, z], [z, z, a, z]):
        outs = dec.decode_batch([to_product(e) if isinstance(e, oracle.EncodedArrays) else e for e in order])
        for e, o in zip(order, outs):
            assert_pcm_bits_equal(o, z_ref if e is z else a_ref, "batch with empty streams")
