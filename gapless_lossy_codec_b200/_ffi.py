"""ctypes binding of libglc_b200.so (the C ABI declared in include/glc.h).

This is the binding a maintainer of the reference would write for its language (INTEGRATION.md
shows the Rust `extern "C"` equivalent).  There is deliberately no fallback: if the library is
missing or no CUDA device is usable, importing works but every call raises GlcError.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# GLC_B200_LIB overrides the library path (used to A/B kernel variants; still no fallback of any kind)
LIB_PATH = os.environ.get("GLC_B200_LIB") or os.path.join(_PKG, "libglc_b200.so")

GLC_OK = 0
STATUS_NAMES = {
    0: "GLC_OK", 1: "GLC_ERR_INVALID_ARG", 2: "GLC_ERR_TOO_SHORT", 3: "GLC_ERR_NO_MEMORY",
    4: "GLC_ERR_FLAC_TOO_SHORT", 5: "GLC_ERR_FLAC_LEVEL", 6: "GLC_ERR_NO_DEVICE", 7: "GLC_ERR_CUDA",
    8: "GLC_ERR_CORRUPT", 9: "GLC_ERR_UNSUPPORTED",
}
K_NAMES = ["mdct_exact", "quant_pack", "scan", "gather", "dequant", "imdct_exact", "ola",
           "flac_block", "flac_gather", "misc", "window_tile", "fast_encode", "fast_decode"]
K_COUNT = len(K_NAMES)


class GlcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


class Pair(C.Structure):
    _fields_ = [("idx", C.c_uint16), ("q", C.c_int16)]


class Encoded(C.Structure):
    """struct glc_encoded"""

    _fields_ = [
        ("sample_rate", C.c_uint32),
        ("channels", C.c_uint16),
        ("reserved0", C.c_uint16),
        ("total_samples", C.c_uint64),
        ("encoder_delay", C.c_uint32),
        ("padding", C.c_uint32),
        ("original_length", C.c_uint64),
        ("n_frames", C.c_uint64),
        ("frame_is_raw", C.POINTER(C.c_uint8)),
        ("nnz", C.POINTER(C.c_uint32)),
        ("pair_offset", C.POINTER(C.c_uint64)),
        ("pairs", C.POINTER(Pair)),
        ("scales", C.POINTER(C.c_float)),
        ("raw_offset", C.POINTER(C.c_uint64)),
        ("raw", C.POINTER(C.c_int16)),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("launches", C.c_uint64 * K_COUNT),
        ("kernel_ms", C.c_double * K_COUNT),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("staged_bytes", C.c_uint64),
        ("pinned_allocs", C.c_uint64),
        ("pinned_alloc_bytes", C.c_uint64),
        ("dev_allocs", C.c_uint64),
        ("dev_alloc_bytes", C.c_uint64),
    ]


# every symbol include/glc.h declares (tests/test_abi.py checks the library exports them all)
EXPORTS = [
    "glc_abi_version", "glc_last_error", "glc_device_count", "glc_ctx_create", "glc_ctx_destroy",
    "glc_ctx_set_tuning", "glc_host_alloc", "glc_host_free", "glc_free",
    "glc_encoder_new", "glc_encoder_free", "glc_encode", "glc_encode_batch", "glc_encoded_free",
    "glc_encode_i16", "glc_encode_batch_i16", "glc_encode_i32",
    "glc_decoder_new", "glc_decoder_free", "glc_decode", "glc_decode_untrimmed", "glc_decode_batch",
    "glc_decode_i16", "glc_decode_batch_i16",
    "glc_decode_stream_open", "glc_decode_stream_next", "glc_decode_stream_close",
    "glc_flac_encode", "glc_flac_encode_batch", "glc_decode_to_flac", "glc_decode_to_flac_batch",
    "glc_encoded_to_bincode", "glc_encoded_from_bincode",
    "glc_plan_shards", "glc_encode_batch_sharded", "glc_decode_batch_sharded", "glc_flac_encode_batch_sharded",
    "glc_stats_reset", "glc_stats_get", "glc_stats_enable_kernel_timing",
    "glc_dev_upload", "glc_dev_pcm_free", "glc_dev_encode", "glc_dev_decode",
    "glc_dev_encoded_download", "glc_dev_pcm_download", "glc_dev_encoded_free",
    "glc_timer_begin", "glc_timer_end", "glc_ctx_sync", "glc_flush_l2", "glc_measure_fp32_issue", "glc_dma_probe",
]

_lib = None


def load() -> C.CDLL:
    """dlopen libglc_b200.so and declare prototypes.  Raises GlcError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GlcError(6, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, u16, u8 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8
    fp = C.POINTER(C.c_float)
    pp = C.POINTER
    sig = {
        "glc_abi_version": (u32, []),
        "glc_last_error": (C.c_char_p, []),
        "glc_device_count": (C.c_int, [pp(C.c_int)]),
        "glc_ctx_create": (C.c_int, [C.c_int, C.c_int, pp(vp)]),
        "glc_ctx_destroy": (None, [vp]),
        "glc_ctx_set_tuning": (C.c_int, [vp, C.c_int, u64]),
        "glc_host_alloc": (C.c_int, [vp, C.c_size_t, pp(vp)]),
        "glc_host_free": (None, [vp, vp]),
        "glc_free": (None, [vp, vp]),
        "glc_encoder_new": (C.c_int, [vp, u32, pp(vp)]),
        "glc_encoder_free": (None, [vp]),
        "glc_encode": (C.c_int, [vp, vp, u64, u16, pp(pp(Encoded))]),
        "glc_encode_batch": (C.c_int, [vp, u32, pp(vp), pp(u64), pp(u16), pp(pp(Encoded))]),
        "glc_encoded_free": (None, [vp, pp(Encoded)]),
        "glc_encode_i16": (C.c_int, [vp, vp, u64, u16, pp(pp(Encoded))]),
        "glc_encode_batch_i16": (C.c_int, [vp, u32, pp(vp), pp(u64), pp(u16), pp(pp(Encoded))]),
        "glc_encode_i32": (C.c_int, [vp, vp, u64, u16, u32, pp(pp(Encoded))]),
        "glc_decoder_new": (C.c_int, [vp, u32, u32, pp(vp)]),
        "glc_decoder_free": (None, [vp]),
        "glc_decode": (C.c_int, [vp, pp(Encoded), pp(fp), pp(u64)]),
        "glc_decode_untrimmed": (C.c_int, [vp, pp(Encoded), pp(fp), pp(u64)]),
        "glc_decode_batch": (C.c_int, [vp, u32, pp(pp(Encoded)), pp(fp), pp(u64)]),
        "glc_decode_i16": (C.c_int, [vp, pp(Encoded), pp(pp(C.c_int16)), pp(u64)]),
        "glc_decode_batch_i16": (C.c_int, [vp, u32, pp(pp(Encoded)), pp(pp(C.c_int16)), pp(u64)]),
        "glc_decode_stream_open": (C.c_int, [vp, pp(Encoded), pp(vp)]),
        "glc_decode_stream_next": (C.c_int, [vp, pp(fp), pp(u64), pp(C.c_int), pp(C.c_float)]),
        "glc_decode_stream_close": (None, [vp]),
        "glc_flac_encode": (C.c_int, [vp, vp, u64, u32, u16, u8, pp(pp(u8)), pp(u64)]),
        "glc_flac_encode_batch": (C.c_int, [vp, u32, pp(vp), pp(u64), pp(u32), pp(u16), u8,
                                            pp(pp(u8)), pp(u64)]),
        "glc_decode_to_flac": (C.c_int, [vp, pp(Encoded), u8, pp(pp(u8)), pp(u64)]),
        "glc_decode_to_flac_batch": (C.c_int, [vp, u32, pp(pp(Encoded)), u8, pp(pp(u8)), pp(u64)]),
        "glc_encoded_to_bincode": (C.c_int, [vp, pp(Encoded), pp(pp(u8)), pp(u64)]),
        "glc_encoded_from_bincode": (C.c_int, [vp, vp, u64, pp(pp(Encoded))]),
        "glc_plan_shards": (C.c_int, [u32, pp(u64), u32, pp(u32)]),
        "glc_encode_batch_sharded": (C.c_int, [pp(vp), u32, u32, pp(vp), pp(u64), pp(u16), pp(pp(Encoded)), pp(u32)]),
        "glc_decode_batch_sharded": (C.c_int, [pp(vp), u32, u32, pp(pp(Encoded)), pp(fp), pp(u64), pp(u32)]),
        "glc_flac_encode_batch_sharded": (C.c_int, [pp(vp), u32, u32, pp(vp), pp(u64), pp(u32), pp(u16), u8,
                                                    pp(pp(u8)), pp(u64), pp(u32)]),
        "glc_stats_reset": (None, [vp]),
        "glc_stats_get": (None, [vp, pp(Stats)]),
        "glc_stats_enable_kernel_timing": (None, [vp, C.c_int]),
        "glc_dev_upload": (C.c_int, [vp, vp, u64, u16, pp(vp)]),
        "glc_dev_pcm_free": (None, [vp]),
        "glc_dev_encode": (C.c_int, [vp, vp, pp(vp)]),
        "glc_dev_decode": (C.c_int, [vp, vp, pp(vp)]),
        "glc_dev_encoded_download": (C.c_int, [vp, pp(pp(Encoded))]),
        "glc_dev_pcm_download": (C.c_int, [vp, pp(fp), pp(u64)]),
        "glc_dev_encoded_free": (None, [vp]),
        "glc_timer_begin": (C.c_int, [vp]),
        "glc_timer_end": (C.c_int, [vp, pp(C.c_float)]),
        "glc_ctx_sync": (C.c_int, [vp]),
        "glc_flush_l2": (C.c_int, [vp]),
        "glc_measure_fp32_issue": (C.c_int, [vp, C.c_int, pp(C.c_double)]),
        "glc_dma_probe": (C.c_int, [vp, u64, u64, C.c_int, pp(C.c_float)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status: int) -> None:
    if status != GLC_OK:
        raise GlcError(status, load().glc_last_error().decode("utf-8", "replace"))
