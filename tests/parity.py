"""Helpers shared by the GPU parity tests: compare a product EncodedAudio with the oracle's."""
from __future__ import annotations

import numpy as np


def assert_encoded_equal(gpu, ref, what=""):
    """Bit-exact comparison of every field of the flat encoded stream (integer work + f32 bit
    patterns of the scale factors)."""
    for f in ("sample_rate", "channels", "total_samples", "encoder_delay", "padding",
              "original_length", "n_frames"):
        assert int(getattr(gpu, f)) == int(getattr(ref, f)), f"{what}: field {f}"
    assert np.array_equal(gpu.frame_is_raw, ref.frame_is_raw), f"{what}: raw/sparse decisions differ"
    if not np.array_equal(gpu.nnz, ref.nnz):
        bad = np.nonzero(gpu.nnz != ref.nnz)[0]
        raise AssertionError(f"{what}: nnz differs in {len(bad)} rows, first {bad[:5]}: "
                             f"{gpu.nnz[bad[:5]]} vs {ref.nnz[bad[:5]]}")
    assert np.array_equal(gpu.pair_offset, ref.pair_offset), f"{what}: pair_offset"
    assert np.array_equal(gpu.pair_idx, ref.pair_idx), f"{what}: coefficient indices differ"
    if not np.array_equal(gpu.pair_q, ref.pair_q):
        d = np.nonzero(gpu.pair_q != ref.pair_q)[0]
        raise AssertionError(f"{what}: {len(d)} of {len(ref.pair_q)} quantized values differ "
                             f"(first at {d[:5]})")
    assert np.array_equal(gpu.scales.view(np.uint32), ref.scales.view(np.uint32)), f"{what}: scale bit patterns"
    assert np.array_equal(gpu.raw_offset, ref.raw_offset), f"{what}: raw_offset"
    assert np.array_equal(gpu.raw, ref.raw), f"{what}: raw PCM frames differ"


def to_oracle(enc):
    """product EncodedAudio -> oracle.EncodedArrays (same flat fields)."""
    import oracle

    return oracle.EncodedArrays(
        sample_rate=enc.sample_rate, channels=enc.channels, total_samples=enc.total_samples,
        encoder_delay=enc.encoder_delay, padding=enc.padding, original_length=enc.original_length,
        n_frames=enc.n_frames, frame_is_raw=enc.frame_is_raw, nnz=enc.nnz, pair_offset=enc.pair_offset,
        pair_idx=enc.pair_idx, pair_q=enc.pair_q, scales=enc.scales, raw_offset=enc.raw_offset, raw=enc.raw)


def to_product(ref):
    """oracle.EncodedArrays -> product EncodedAudio."""
    from gapless_lossy_codec_b200 import EncodedAudio

    return EncodedAudio(
        sample_rate=ref.sample_rate, channels=ref.channels, total_samples=ref.total_samples,
        encoder_delay=ref.encoder_delay, padding=ref.padding, original_length=ref.original_length,
        n_frames=ref.n_frames, frame_is_raw=ref.frame_is_raw, nnz=ref.nnz, pair_offset=ref.pair_offset,
        pair_idx=ref.pair_idx, pair_q=ref.pair_q, scales=ref.scales, raw_offset=ref.raw_offset, raw=ref.raw)


def assert_pcm_bits_equal(a, b, what=""):
    assert a.shape == b.shape, f"{what}: length {a.shape} vs {b.shape}"
    if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
        d = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
        raise AssertionError(f"{what}: {len(d)} of {len(a)} PCM values differ bitwise; first {d[:5]}, "
                             f"max abs diff {np.max(np.abs(a - b))}")
