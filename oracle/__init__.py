"""ctypes bindings of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package (``gapless_lossy_codec_b200``) never does; it fails loudly when its CUDA
library is missing instead of falling back to this code.

The C sources restate /root/reference/src/codec.rs and src/flac.rs (file:line
citations are in the .c files).  Parity status: "parity unpinned" -- see
oracle_common.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libglc_oracle.so")

FRAME_SIZE = 2048
HOP_SIZE = 1024


def build(force: bool = False) -> str:
    """Compile libglc_oracle.so with the committed Makefile (gcc, no OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True, env=env)
    return _LIB_PATH


class Pair(C.Structure):
    _fields_ = [("idx", C.c_uint16), ("q", C.c_int16)]


class Encoded(C.Structure):
    """Mirror of orc_encoded (identical layout to glc_encoded in include/glc.h)."""

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


class FlacInfo(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_uint32),
        ("channels", C.c_uint32),
        ("bits_per_sample", C.c_uint32),
        ("min_block", C.c_uint32),
        ("max_block", C.c_uint32),
        ("total_samples", C.c_uint64),
        ("md5", C.c_uint8 * 16),
        ("n_frames", C.c_uint64),
        ("md5_ok", C.c_int),
        ("n_decoded", C.c_uint64),
        ("samples", C.POINTER(C.c_int32)),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.orc_build_tables.argtypes = [fp, fp, fp]
        L.orc_build_tables.restype = None
        L.orc_build_perceptual.argtypes = [C.c_uint32, fp, C.POINTER(C.c_int32)]
        L.orc_build_perceptual.restype = C.c_int
        L.orc_noise_floor_factor.restype = C.c_float
        L.orc_mdct_block.argtypes = [fp, C.c_float, fp, fp]
        L.orc_mdct_block.restype = None
        L.orc_imdct_block.argtypes = [fp, C.c_float, fp, fp]
        L.orc_imdct_block.restype = None
        L.orc_imdct_block_ref_order.argtypes = [fp, C.c_float, fp, fp]
        L.orc_imdct_block_ref_order.restype = None
        L.orc_encode.argtypes = [fp, C.c_uint64, C.c_uint16, C.c_uint32, C.c_int,
                                 C.POINTER(C.POINTER(Encoded))]
        L.orc_encode.restype = C.c_int
        L.orc_decode.argtypes = [C.POINTER(Encoded), C.c_int, C.c_int, C.POINTER(fp),
                                 C.POINTER(C.c_uint64)]
        L.orc_decode.restype = C.c_int
        L.orc_decode_untrimmed.argtypes = L.orc_decode.argtypes
        L.orc_decode_untrimmed.restype = C.c_int
        L.orc_encoded_free.argtypes = [C.POINTER(Encoded)]
        L.orc_encoded_free.restype = None
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free.restype = None
        L.orc_mdct_all.argtypes = [fp, C.c_uint64, C.c_uint16, C.c_int, C.POINTER(fp),
                                   C.POINTER(C.c_uint64)]
        L.orc_mdct_all.restype = C.c_int
        L.orc_flac_encode.argtypes = [fp, C.c_uint64, C.c_uint32, C.c_uint16, C.c_uint8,
                                      C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64)]
        L.orc_flac_encode.restype = C.c_int
        L.orc_md5.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint8)]
        L.orc_md5.restype = None
        L.orc_crc8.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_crc8.restype = C.c_uint8
        L.orc_crc16.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_crc16.restype = C.c_uint16
        L.orc_flac_decode.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(FlacInfo)]
        L.orc_flac_decode.restype = C.c_int
        L.orc_flac_decode_ex.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(FlacInfo), C.POINTER(C.c_uint64),
                                         C.c_uint64]
        L.orc_flac_decode_ex.restype = C.c_int
        L.orc_bincode_serialize.argtypes = [C.POINTER(Encoded), C.POINTER(C.POINTER(C.c_uint8)),
                                            C.POINTER(C.c_uint64)]
        L.orc_bincode_serialize.restype = C.c_int
        L.orc_bincode_deserialize.argtypes = [C.c_void_p, C.c_uint64,
                                              C.POINTER(C.POINTER(Encoded))]
        L.orc_bincode_deserialize.restype = C.c_int
        L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_fnv1a64.restype = C.c_uint64
        _lib = L
    return _lib


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


# --------------------------------------------------------------------- codec


@dataclass
class EncodedArrays:
    """numpy copy of a flat encoded stream (owning; independent of C memory)."""

    sample_rate: int
    channels: int
    total_samples: int
    encoder_delay: int
    padding: int
    original_length: int
    n_frames: int
    frame_is_raw: np.ndarray  # u8 [frames]
    nnz: np.ndarray  # u32 [frames*ch]
    pair_offset: np.ndarray  # u64 [frames*ch+1]
    pair_idx: np.ndarray  # u16 [npairs]
    pair_q: np.ndarray  # i16 [npairs]
    scales: np.ndarray  # f32 [frames*ch]
    raw_offset: np.ndarray  # u64 [frames+1]
    raw: np.ndarray  # i16

    def as_struct(self):
        """Build a ctypes Encoded that borrows this object's arrays."""
        pairs = np.empty(len(self.pair_idx), dtype=[("idx", "<u2"), ("q", "<i2")])
        pairs["idx"] = self.pair_idx
        pairs["q"] = self.pair_q
        keep = [pairs, np.ascontiguousarray(self.frame_is_raw, np.uint8),
                np.ascontiguousarray(self.nnz, np.uint32),
                np.ascontiguousarray(self.pair_offset, np.uint64),
                np.ascontiguousarray(self.scales, np.float32),
                np.ascontiguousarray(self.raw_offset, np.uint64),
                np.ascontiguousarray(self.raw, np.int16)]
        e = Encoded()
        e.sample_rate = self.sample_rate
        e.channels = self.channels
        e.total_samples = self.total_samples
        e.encoder_delay = self.encoder_delay
        e.padding = self.padding
        e.original_length = self.original_length
        e.n_frames = self.n_frames
        e.pairs = keep[0].ctypes.data_as(C.POINTER(Pair))
        e.frame_is_raw = keep[1].ctypes.data_as(C.POINTER(C.c_uint8))
        e.nnz = keep[2].ctypes.data_as(C.POINTER(C.c_uint32))
        e.pair_offset = keep[3].ctypes.data_as(C.POINTER(C.c_uint64))
        e.scales = keep[4].ctypes.data_as(C.POINTER(C.c_float))
        e.raw_offset = keep[5].ctypes.data_as(C.POINTER(C.c_uint64))
        e.raw = keep[6].ctypes.data_as(C.POINTER(C.c_int16))
        e._keep = keep
        return e


def arrays_from_struct(e) -> EncodedArrays:
    """Copy any struct with the glc_encoded/orc_encoded layout into numpy arrays."""
    frames = int(e.n_frames)
    ch = int(e.channels)
    nfc = frames * ch

    def arr(ptr, n, dt):
        if n == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(ptr, shape=(n,)).view(dt).copy() if dt is not None else None

    pair_offset = np.ctypeslib.as_array(e.pair_offset, shape=(nfc + 1,)).copy()
    raw_offset = np.ctypeslib.as_array(e.raw_offset, shape=(frames + 1,)).copy()
    npairs = int(pair_offset[-1])
    nraw = int(raw_offset[-1])
    if npairs:
        pr = np.frombuffer(
            C.string_at(C.cast(e.pairs, C.c_void_p), npairs * 4),
            dtype=[("idx", "<u2"), ("q", "<i2")],
        )
        pidx, pq = pr["idx"].copy(), pr["q"].copy()
    else:
        pidx, pq = np.zeros(0, np.uint16), np.zeros(0, np.int16)
    return EncodedArrays(
        sample_rate=int(e.sample_rate), channels=ch, total_samples=int(e.total_samples),
        encoder_delay=int(e.encoder_delay), padding=int(e.padding),
        original_length=int(e.original_length), n_frames=frames,
        frame_is_raw=np.ctypeslib.as_array(e.frame_is_raw, shape=(frames,)).copy()
        if frames else np.zeros(0, np.uint8),
        nnz=np.ctypeslib.as_array(e.nnz, shape=(nfc,)).copy() if nfc else np.zeros(0, np.uint32),
        pair_offset=pair_offset, pair_idx=pidx, pair_q=pq,
        scales=np.ctypeslib.as_array(e.scales, shape=(nfc,)).copy()
        if nfc else np.zeros(0, np.float32),
        raw_offset=raw_offset,
        raw=np.ctypeslib.as_array(e.raw, shape=(nraw,)).copy() if nraw else np.zeros(0, np.int16),
    )


class OracleError(RuntimeError):
    pass


def encode(samples: np.ndarray, channels: int, sample_rate: int, threads: int | None = None
           ) -> EncodedArrays:
    """Encoder::new(sample_rate).encode(samples, channels)  (src/codec.rs:406-565)."""
    pcm = np.ascontiguousarray(samples, np.float32)
    out = C.POINTER(Encoded)()
    rc = lib().orc_encode(_fptr(pcm), pcm.size, channels, sample_rate,
                          threads or default_threads(), C.byref(out))
    if rc:
        raise OracleError(f"orc_encode rc={rc}")
    try:
        return arrays_from_struct(out.contents)
    finally:
        lib().orc_encoded_free(out)


def decode(enc: EncodedArrays, threads: int | None = None, literal_imdct: bool = False,
           trimmed: bool = True) -> np.ndarray:
    """Decoder::decode (trimmed) or decode_streaming concatenated (untrimmed)."""
    st = enc.as_struct()
    p = C.POINTER(C.c_float)()
    n = C.c_uint64()
    fn = lib().orc_decode if trimmed else lib().orc_decode_untrimmed
    rc = fn(C.byref(st), threads or default_threads(), int(literal_imdct), C.byref(p), C.byref(n))
    if rc:
        raise OracleError(f"orc_decode rc={rc}")
    try:
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
    finally:
        lib().orc_free(p)


def mdct_all(samples: np.ndarray, channels: int, threads: int | None = None) -> np.ndarray:
    pcm = np.ascontiguousarray(samples, np.float32)
    p = C.POINTER(C.c_float)()
    n = C.c_uint64()
    rc = lib().orc_mdct_all(_fptr(pcm), pcm.size, channels, threads or default_threads(),
                            C.byref(p), C.byref(n))
    if rc:
        raise OracleError(f"orc_mdct_all rc={rc}")
    try:
        return np.ctypeslib.as_array(p, shape=(n.value, HOP_SIZE)).copy()
    finally:
        lib().orc_free(p)


def tables():
    cos_tab = np.empty((HOP_SIZE, FRAME_SIZE), np.float32)
    window = np.empty(FRAME_SIZE, np.float32)
    norm = C.c_float()
    lib().orc_build_tables(_fptr(cos_tab), _fptr(window), C.byref(norm))
    return cos_tab, window, np.float32(norm.value)


def perceptual(sample_rate: int):
    w = np.empty(HOP_SIZE, np.float32)
    b = np.zeros(64, np.int32)
    nb = lib().orc_build_perceptual(sample_rate, _fptr(w), b.ctypes.data_as(C.POINTER(C.c_int32)))
    return w, b[:nb].copy()


# ---------------------------------------------------------------------- flac


def flac_encode(samples: np.ndarray, sample_rate: int, channels: int, level: int = 5) -> bytes:
    """flac::encode_flac_with_level (src/flac.rs:947-1052)."""
    pcm = np.ascontiguousarray(samples, np.float32)
    b = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    rc = lib().orc_flac_encode(_fptr(pcm), pcm.size, sample_rate, channels, level,
                               C.byref(b), C.byref(n))
    if rc == 4:
        raise OracleError("FLAC requires at least 16 samples per channel")
    if rc == 5:
        raise OracleError("Invalid compression level")
    if rc:
        raise OracleError(f"orc_flac_encode rc={rc}")
    try:
        return C.string_at(b, n.value)
    finally:
        lib().orc_free(b)


def flac_decode(data: bytes, frame_offsets: bool = False) -> dict:
    """Independent RFC 9639 decoder (stands in for claxon in tests/test_flac.rs).  With frame_offsets the
    result also holds "frame_off": the byte offset of every frame header, then len(data)."""
    info = FlacInfo()
    buf = np.frombuffer(data, np.uint8)
    cap = len(data) // 9 + 2 if frame_offsets else 0  # a frame is at least 9 bytes
    offs = np.zeros(max(cap, 1), np.uint64)
    rc = lib().orc_flac_decode_ex(buf.ctypes.data, len(data), C.byref(info),
                                  offs.ctypes.data_as(C.POINTER(C.c_uint64)) if frame_offsets else None, cap)
    try:
        smp = (np.ctypeslib.as_array(info.samples, shape=(info.n_decoded,)).copy()
               if info.n_decoded else np.zeros(0, np.int32))
    finally:
        if info.samples:
            lib().orc_free(info.samples)
    if rc:
        raise OracleError(f"flac decode failed rc={rc}")
    out = dict(sample_rate=info.sample_rate, channels=info.channels,
               bits_per_sample=info.bits_per_sample, min_block=info.min_block,
               max_block=info.max_block, total_samples=int(info.total_samples),
               md5=bytes(info.md5), md5_ok=bool(info.md5_ok), n_frames=int(info.n_frames),
               samples=smp)
    if frame_offsets:
        out["frame_off"] = offs[: int(info.n_frames) + 1].copy()
    return out


def md5(data: bytes) -> bytes:
    out = (C.c_uint8 * 16)()
    lib().orc_md5(data, len(data), out)
    return bytes(out)


def crc8(data: bytes) -> int:
    return int(lib().orc_crc8(data, len(data)))


def crc16(data: bytes) -> int:
    return int(lib().orc_crc16(data, len(data)))


# ------------------------------------------------------------------- bincode


def bincode_serialize(enc: EncodedArrays) -> bytes:
    st = enc.as_struct()
    b = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    rc = lib().orc_bincode_serialize(C.byref(st), C.byref(b), C.byref(n))
    if rc:
        raise OracleError(f"bincode serialize rc={rc}")
    try:
        return C.string_at(b, n.value)
    finally:
        lib().orc_free(b)


def bincode_deserialize(data: bytes) -> EncodedArrays:
    out = C.POINTER(Encoded)()
    rc = lib().orc_bincode_deserialize(data, len(data), C.byref(out))
    if rc:
        raise OracleError(f"bincode deserialize rc={rc}")
    try:
        return arrays_from_struct(out.contents)
    finally:
        lib().orc_encoded_free(out)


def fnv1a64(data) -> int:
    """FNV-1a/64 of a bytes-like / numpy array (the table fingerprint of tests/golden/dump_reference.rs)."""
    a = np.ascontiguousarray(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data)
    return int(lib().orc_fnv1a64(a.ctypes.data, a.nbytes))
