"""Host-side mirror of the reference's `flac` module (reference src/flac.rs:947-1088)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import check
from .codec import Context, default_context, _as_f32


def encode_flac_with_level(samples, sample_rate: int, channels: int, compression_level: int,
                           ctx: Optional[Context] = None) -> bytes:
    """flac::encode_flac_with_level (src/flac.rs:947-1052).  Errors mirror the reference:
    < 16 samples per channel -> GLC_ERR_FLAC_TOO_SHORT (:963-969); level > 8 -> GLC_ERR_FLAC_LEVEL
    (:972-978)."""
    ctx = ctx or default_context()
    pcm = _as_f32(samples)
    b = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    check(ctx._lib.glc_flac_encode(ctx.handle, pcm.ctypes.data, pcm.size, int(sample_rate), int(channels),
                                   int(compression_level), C.byref(b), C.byref(n)))
    try:
        return C.string_at(b, n.value)
    finally:
        ctx._lib.glc_free(ctx.handle, b)


def encode_flac(samples, sample_rate: int, channels: int, ctx: Optional[Context] = None) -> bytes:
    """flac::encode_flac: level 5 (src/flac.rs:1055-1062)."""
    return encode_flac_with_level(samples, sample_rate, channels, 5, ctx)


def encode_flac_batch(files: Sequence, sample_rates: Sequence[int], channels: Sequence[int], level: int = 5,
                      ctx: Optional[Context] = None) -> List[bytes]:
    ctx = ctx or default_context()
    n = len(files)
    arrs = [_as_f32(f) for f in files]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    ns = (C.c_uint64 * n)(*[a.size for a in arrs])
    srs = (C.c_uint32 * n)(*[int(s) for s in sample_rates])
    chs = (C.c_uint16 * n)(*[int(c) for c in channels])
    outs = (C.POINTER(C.c_uint8) * n)()
    lens = (C.c_uint64 * n)()
    check(ctx._lib.glc_flac_encode_batch(ctx.handle, n, ptrs, ns, srs, chs, int(level), outs, lens))
    res = []
    for i in range(n):
        res.append(C.string_at(outs[i], lens[i]))
        ctx._lib.glc_free(ctx.handle, outs[i])
    return res


def export_to_flac_with_level(path, samples, sample_rate: int, channels: int, compression_level: int) -> None:
    """flac::export_to_flac_with_level (src/flac.rs:1065-1077)."""
    data = encode_flac_with_level(samples, sample_rate, channels, compression_level)
    with open(path, "wb") as f:
        f.write(data)


def export_to_flac(path, samples, sample_rate: int, channels: int) -> None:
    """flac::export_to_flac: level 5 (src/flac.rs:1080-1088)."""
    export_to_flac_with_level(path, samples, sample_rate, channels, 5)
