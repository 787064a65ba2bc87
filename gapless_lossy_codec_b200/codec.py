"""Host-side mirror of the reference's `codec` module (reference src/codec.rs, re-exported at the
crate root by src/lib.rs:1-5): same type and method names, same argument meaning, same error
behaviour, implemented over the C ABI of libglc_b200.so.  No CPU fallback exists: every call
needs the CUDA library and a B200.

    enc = Encoder(44100)                      # Encoder::new           src/codec.rs:406
    encoded = enc.encode(samples, channels)   # Encoder::encode        :421
    dec = Decoder(channels, 44100)            # Decoder::new           :581
    pcm = dec.decode(encoded)                 # Decoder::decode        :744
    for chunk in dec.decode_streaming(encoded): ...   # decode_streaming :595
    save_encoded(encoded, path); load_encoded(path)   # :774, :781
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from ._ffi import GlcError, check

FRAME_SIZE = 2048  # src/codec.rs:15
HOP_SIZE = 1024  # src/codec.rs:16
FRAMES_PER_CHUNK = 500  # src/codec.rs:18


# ------------------------------------------------------------------ context

_ctx_lock = threading.Lock()
_ctxs: dict = {}


class Context:
    """One CUDA device (glc_ctx): streams, pinned pools, device tables."""

    def __init__(self, device: int = 0, mode: int = 0):
        """mode 0 = GLC_MODE_EXACT (bit-exact with the reference, default); 1 = GLC_MODE_FAST (FFT-based
        transform, tolerance class)."""
        self._lib = _ffi.load()
        h = C.c_void_p()
        check(self._lib.glc_ctx_create(device, int(mode), C.byref(h)))
        self.handle = h
        self.device = device
        self.mode = int(mode)

    def close(self):
        if self.handle:
            self._lib.glc_ctx_destroy(self.handle)
            self.handle = None

    def stats(self) -> dict:
        st = _ffi.Stats()
        self._lib.glc_stats_get(self.handle, C.byref(st))
        return dict(
            launches={n: int(st.launches[i]) for i, n in enumerate(_ffi.K_NAMES)},
            kernel_ms={n: float(st.kernel_ms[i]) for i, n in enumerate(_ffi.K_NAMES)},
            h2d_bytes=int(st.h2d_bytes), d2h_bytes=int(st.d2h_bytes), staged_bytes=int(st.staged_bytes),
            pinned_allocs=int(st.pinned_allocs), pinned_alloc_bytes=int(st.pinned_alloc_bytes),
            dev_allocs=int(st.dev_allocs), dev_alloc_bytes=int(st.dev_alloc_bytes))

    def stats_reset(self):
        self._lib.glc_stats_reset(self.handle)

    def enable_kernel_timing(self, on: bool):
        self._lib.glc_stats_enable_kernel_timing(self.handle, int(on))

    def set_tuning(self, reserved: int = 0, wave_rows: int = 0):
        """wave_rows = frame-channel rows per encode pipeline wave (0 = automatic)."""
        check(self._lib.glc_ctx_set_tuning(self.handle, reserved, wave_rows))

    def sync(self):
        check(self._lib.glc_ctx_sync(self.handle))

    def pinned_array(self, n: int, dtype=np.float32) -> np.ndarray:
        """numpy view of library-owned pinned host memory (freed with the context)."""
        dt = np.dtype(dtype)
        p = C.c_void_p()
        check(self._lib.glc_host_alloc(self.handle, max(n, 1) * dt.itemsize, C.byref(p)))
        buf = (C.c_byte * (max(n, 1) * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=n)


def default_context(device: int = 0) -> Context:
    with _ctx_lock:
        if device not in _ctxs:
            _ctxs[device] = Context(device)
        return _ctxs[device]


# --------------------------------------------------------------- data model


@dataclass
class AudioHeader:  # src/codec.rs:39-45
    sample_rate: int
    channels: int
    total_samples: int


@dataclass
class GaplessInfo:  # src/codec.rs:47-53
    encoder_delay: int
    padding: int
    original_length: int


@dataclass
class EncodedFrame:  # src/codec.rs:55-69
    sparse_coeffs_per_channel: List[List[Tuple[int, int]]]
    scale_factors: List[float]
    raw_pcm: Optional[List[int]]


@dataclass
class AudioChunk:  # src/codec.rs:81-85
    samples: np.ndarray
    is_last: bool


@dataclass
class Progress:  # src/codec.rs:71-79 (enum): kind in {Status, Decoding, Complete, ...}
    kind: str
    value: object


class EncodedAudio:
    """EncodedAudio (src/codec.rs:31-37) held as the flat arrays of struct glc_encoded.

    `.header`, `.gapless_info` and `.frames` give the reference's nested view; the flat numpy
    arrays (`frame_is_raw`, `nnz`, `pair_offset`, `pair_idx`, `pair_q`, `scales`, `raw_offset`,
    `raw`) are what crosses the ABI.
    """

    def __init__(self, **kw):
        self.sample_rate = int(kw["sample_rate"])
        self.channels = int(kw["channels"])
        self.total_samples = int(kw["total_samples"])
        self.encoder_delay = int(kw["encoder_delay"])
        self.padding = int(kw["padding"])
        self.original_length = int(kw["original_length"])
        self.n_frames = int(kw["n_frames"])
        self.frame_is_raw = np.ascontiguousarray(kw["frame_is_raw"], np.uint8)
        self.nnz = np.ascontiguousarray(kw["nnz"], np.uint32)
        self.pair_offset = np.ascontiguousarray(kw["pair_offset"], np.uint64)
        self.pair_idx = np.ascontiguousarray(kw["pair_idx"], np.uint16)
        self.pair_q = np.ascontiguousarray(kw["pair_q"], np.int16)
        self.scales = np.ascontiguousarray(kw["scales"], np.float32)
        self.raw_offset = np.ascontiguousarray(kw["raw_offset"], np.uint64)
        self.raw = np.ascontiguousarray(kw["raw"], np.int16)

    # -- reference-shaped views
    @property
    def header(self) -> AudioHeader:
        return AudioHeader(self.sample_rate, self.channels, self.total_samples)

    @property
    def gapless_info(self) -> GaplessInfo:
        return GaplessInfo(self.encoder_delay, self.padding, self.original_length)

    @property
    def frames(self) -> List[EncodedFrame]:
        out = []
        ch = self.channels
        for f in range(self.n_frames):
            if self.frame_is_raw[f]:
                a, b = int(self.raw_offset[f]), int(self.raw_offset[f + 1])
                out.append(EncodedFrame([], [], self.raw[a:b].tolist()))
            else:
                sp = []
                for c in range(ch):
                    a, b = int(self.pair_offset[f * ch + c]), int(self.pair_offset[f * ch + c + 1])
                    sp.append(list(zip(self.pair_idx[a:b].tolist(), self.pair_q[a:b].tolist())))
                out.append(EncodedFrame(sp, self.scales[f * ch:(f + 1) * ch].tolist(), None))
        return out

    # -- ABI plumbing
    @classmethod
    def _from_struct(cls, e) -> "EncodedAudio":
        frames, ch = int(e.n_frames), int(e.channels)
        rows = frames * ch

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)

        pair_offset = arr(e.pair_offset, rows + 1, np.uint64) if frames else np.zeros(1, np.uint64)
        raw_offset = arr(e.raw_offset, frames + 1, np.uint64) if frames else np.zeros(1, np.uint64)
        npairs, nraw = int(pair_offset[-1]), int(raw_offset[-1])
        if npairs:
            pr = np.frombuffer(C.string_at(C.cast(e.pairs, C.c_void_p), npairs * 4),
                               dtype=[("idx", "<u2"), ("q", "<i2")])
            pidx, pq = pr["idx"].copy(), pr["q"].copy()
        else:
            pidx, pq = np.zeros(0, np.uint16), np.zeros(0, np.int16)
        return cls(sample_rate=e.sample_rate, channels=ch, total_samples=e.total_samples,
                   encoder_delay=e.encoder_delay, padding=e.padding,
                   original_length=e.original_length, n_frames=frames,
                   frame_is_raw=arr(e.frame_is_raw, frames, np.uint8), nnz=arr(e.nnz, rows, np.uint32),
                   pair_offset=pair_offset, pair_idx=pidx, pair_q=pq,
                   scales=arr(e.scales, rows, np.float32), raw_offset=raw_offset,
                   raw=arr(e.raw, nraw, np.int16))

    def _as_struct(self):
        pairs = np.empty(len(self.pair_idx), dtype=[("idx", "<u2"), ("q", "<i2")])
        pairs["idx"] = self.pair_idx
        pairs["q"] = self.pair_q
        e = _ffi.Encoded()
        e.sample_rate, e.channels = self.sample_rate, self.channels
        e.total_samples = self.total_samples
        e.encoder_delay, e.padding = self.encoder_delay, self.padding
        e.original_length, e.n_frames = self.original_length, self.n_frames
        e.frame_is_raw = self.frame_is_raw.ctypes.data_as(C.POINTER(C.c_uint8))
        e.nnz = self.nnz.ctypes.data_as(C.POINTER(C.c_uint32))
        e.pair_offset = self.pair_offset.ctypes.data_as(C.POINTER(C.c_uint64))
        e.pairs = pairs.ctypes.data_as(C.POINTER(_ffi.Pair))
        e.scales = self.scales.ctypes.data_as(C.POINTER(C.c_float))
        e.raw_offset = self.raw_offset.ctypes.data_as(C.POINTER(C.c_uint64))
        e.raw = self.raw.ctypes.data_as(C.POINTER(C.c_int16))
        e._keep = (pairs, self)
        return e


def _as_f32(samples) -> np.ndarray:
    return np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)


# ------------------------------------------------------------------ Encoder


class Encoder:
    """codec::Encoder (src/codec.rs:396-566).  Reusable across files (tests/test_codec.rs:150)."""

    def __init__(self, sample_rate: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        self.sample_rate = int(sample_rate)
        h = C.c_void_p()
        check(self._lib.glc_encoder_new(self.ctx.handle, self.sample_rate, C.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.ctx.handle:
                self._lib.glc_encoder_free(self.handle)
        except Exception:
            pass

    def encode(self, samples, channels: int) -> EncodedAudio:
        """Encoder::encode.  Raises GlcError(GLC_ERR_TOO_SHORT) where the reference panics
        (<= 512 samples per channel)."""
        pcm = _as_f32(samples)
        out = C.POINTER(_ffi.Encoded)()
        check(self._lib.glc_encode(self.handle, pcm.ctypes.data, pcm.size, int(channels), C.byref(out)))
        try:
            return EncodedAudio._from_struct(out.contents)
        finally:
            self._lib.glc_encoded_free(self.ctx.handle, out)

    def encode_pcm_int(self, samples, channels: int, bits_per_sample: int = 16) -> EncodedAudio:
        """Encoder::encode over integer PCM as audio::load_wav / load_flac would convert it
        (`s as f32 / 2^(bits-1)`, src/audio.rs:51-59): int16 arrays take the 16-bit path (half the
        PCIe bytes), anything else goes through an int32 container.  Bit-identical to
        `encode(samples / 2**(bits-1))`."""
        a = np.ascontiguousarray(samples).reshape(-1)
        out = C.POINTER(_ffi.Encoded)()
        if a.dtype == np.int16 and bits_per_sample == 16:
            check(self._lib.glc_encode_i16(self.handle, a.ctypes.data, a.size, int(channels), C.byref(out)))
        else:
            a = np.ascontiguousarray(a, np.int32)
            check(self._lib.glc_encode_i32(self.handle, a.ctypes.data, a.size, int(channels), int(bits_per_sample),
                                           C.byref(out)))
        try:
            return EncodedAudio._from_struct(out.contents)
        finally:
            self._lib.glc_encoded_free(self.ctx.handle, out)

    def encode_batch(self, files: Sequence, channels: Sequence[int]) -> List[EncodedAudio]:
        """Many `encode` calls fused into one device pass (sharding unit: file)."""
        n = len(files)
        arrs = [_as_f32(f) for f in files]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        ns = (C.c_uint64 * n)(*[a.size for a in arrs])
        chs = (C.c_uint16 * n)(*[int(c) for c in channels])
        outs = (C.POINTER(_ffi.Encoded) * n)()
        check(self._lib.glc_encode_batch(self.handle, n, ptrs, ns, chs, outs))
        res = []
        try:
            for i in range(n):
                res.append(EncodedAudio._from_struct(outs[i].contents))
        finally:
            for i in range(n):
                if outs[i]:
                    self._lib.glc_encoded_free(self.ctx.handle, outs[i])
        return res

    def encode_batch_i16(self, files: Sequence, channels: Sequence[int]) -> List[EncodedAudio]:
        """`encode_batch` over 16-bit PCM (glc_encode_batch_i16): the loaders' division by 32768
        (src/audio.rs:51-59) runs on the device, 16-bit samples cross PCIe."""
        n = len(files)
        arrs = [np.ascontiguousarray(f, np.int16).reshape(-1) for f in files]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        ns = (C.c_uint64 * n)(*[a.size for a in arrs])
        chs = (C.c_uint16 * n)(*[int(c) for c in channels])
        outs = (C.POINTER(_ffi.Encoded) * n)()
        check(self._lib.glc_encode_batch_i16(self.handle, n, ptrs, ns, chs, outs))
        res = []
        try:
            for i in range(n):
                res.append(EncodedAudio._from_struct(outs[i].contents))
        finally:
            for i in range(n):
                if outs[i]:
                    self._lib.glc_encoded_free(self.ctx.handle, outs[i])
        return res


# ------------------------------------------------------------------ Decoder


class Decoder:
    """codec::Decoder (src/codec.rs:571-769).  `channels` / `sample_rate` are informational, as in
    the reference (the stream header wins, src/codec.rs:598)."""

    def __init__(self, channels: int, sample_rate: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        h = C.c_void_p()
        check(self._lib.glc_decoder_new(self.ctx.handle, int(channels), int(sample_rate), C.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.ctx.handle:
                self._lib.glc_decoder_free(self.handle)
        except Exception:
            pass

    def _take(self, p, n) -> np.ndarray:
        try:
            return (np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.float32))
        finally:
            self._lib.glc_free(self.ctx.handle, p)

    def decode(self, encoded: EncodedAudio, progress: Optional[Callable[[Progress], None]] = None
               ) -> np.ndarray:
        """Decoder::decode: all chunks concatenated, then the gapless trim (src/codec.rs:744-768)."""
        st = encoded._as_struct()
        p = C.POINTER(C.c_float)()
        n = C.c_uint64()
        if progress:
            progress(Progress("Status", f"Starting streaming decode of {encoded.n_frames} frames"))
        check(self._lib.glc_decode(self.handle, C.byref(st), C.byref(p), C.byref(n)))
        if progress:
            progress(Progress("Complete", f"Decoded {encoded.n_frames} frames"))
        return self._take(p, n.value)

    def decode_pcm16(self, encoded: EncodedAudio) -> np.ndarray:
        """The samples `glc -d file.glc` writes to its WAV (src/main.rs:95-105): Decoder::decode followed by
        audio::convert_f32_to_i16 (src/audio.rs:11-16), converted on the device so that 16-bit samples
        cross PCIe."""
        st = encoded._as_struct()
        p = C.POINTER(C.c_int16)()
        n = C.c_uint64()
        check(self._lib.glc_decode_i16(self.handle, C.byref(st), C.byref(p), C.byref(n)))
        try:
            return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int16)
        finally:
            self._lib.glc_free(self.ctx.handle, p)

    def decode_batch_pcm16(self, encoded: Sequence[EncodedAudio]) -> List[np.ndarray]:
        n = len(encoded)
        structs = [e._as_struct() for e in encoded]
        ptrs = (C.POINTER(_ffi.Encoded) * n)(*[C.pointer(s) for s in structs])
        outs = (C.POINTER(C.c_int16) * n)()
        ns = (C.c_uint64 * n)()
        check(self._lib.glc_decode_batch_i16(self.handle, n, ptrs, outs, ns))
        res = []
        for i in range(n):
            try:
                res.append(np.ctypeslib.as_array(outs[i], shape=(ns[i],)).copy() if ns[i] else np.zeros(0, np.int16))
            finally:
                self._lib.glc_free(self.ctx.handle, outs[i])
        return res

    def decode_to_flac(self, encoded: EncodedAudio, compression_level: int = 5) -> bytes:
        """`glc -d file.glc --flac-level N` (src/main.rs:55-113): decode, then FLAC-encode the decoded
        samples, with the PCM kept on the device in between."""
        st = encoded._as_struct()
        b = C.POINTER(C.c_uint8)()
        n = C.c_uint64()
        check(self._lib.glc_decode_to_flac(self.handle, C.byref(st), int(compression_level), C.byref(b), C.byref(n)))
        try:
            return C.string_at(b, n.value)
        finally:
            self._lib.glc_free(self.ctx.handle, b)

    def decode_untrimmed(self, encoded: EncodedAudio) -> np.ndarray:
        st = encoded._as_struct()
        p = C.POINTER(C.c_float)()
        n = C.c_uint64()
        check(self._lib.glc_decode_untrimmed(self.handle, C.byref(st), C.byref(p), C.byref(n)))
        return self._take(p, n.value)

    def decode_batch(self, encoded: Sequence[EncodedAudio]) -> List[np.ndarray]:
        n = len(encoded)
        structs = [e._as_struct() for e in encoded]
        ptrs = (C.POINTER(_ffi.Encoded) * n)(*[C.pointer(s) for s in structs])
        outs = (C.POINTER(C.c_float) * n)()
        ns = (C.c_uint64 * n)()
        check(self._lib.glc_decode_batch(self.handle, n, ptrs, outs, ns))
        return [self._take(outs[i], ns[i]) for i in range(n)]

    def decode_streaming(self, encoded: EncodedAudio,
                         progress: Optional[Callable[[Progress], None]] = None) -> Iterator[AudioChunk]:
        """Decoder::decode_streaming (src/codec.rs:595-741): chunks of exactly 500 frames, then the
        tail with is_last=True; Progress events Status / Decoding(%) / Complete as the reference."""
        st = encoded._as_struct()
        s = C.c_void_p()
        if progress:
            progress(Progress("Status", f"Starting streaming decode of {encoded.n_frames} frames"))
        check(self._lib.glc_decode_stream_open(self.handle, C.byref(st), C.byref(s)))
        try:
            while True:
                p = C.POINTER(C.c_float)()
                n = C.c_uint64()
                last = C.c_int()
                pct = C.c_float()
                check(self._lib.glc_decode_stream_next(s, C.byref(p), C.byref(n), C.byref(last), C.byref(pct)))
                if progress and not last.value:
                    progress(Progress("Decoding", float(pct.value)))
                data = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
                yield AudioChunk(data, bool(last.value))
                if last.value:
                    break
            if progress:
                progress(Progress("Complete", f"Decoded {encoded.n_frames} frames"))
        finally:
            self._lib.glc_decode_stream_close(s)


# ---------------------------------------------------------------- container


def encoded_to_bytes(encoded: EncodedAudio, ctx: Optional[Context] = None) -> bytes:
    ctx = ctx or default_context()
    st = encoded._as_struct()
    b = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    check(ctx._lib.glc_encoded_to_bincode(ctx.handle, C.byref(st), C.byref(b), C.byref(n)))
    try:
        return C.string_at(b, n.value)
    finally:
        ctx._lib.glc_free(ctx.handle, b)


def encoded_from_bytes(data: bytes, ctx: Optional[Context] = None) -> EncodedAudio:
    ctx = ctx or default_context()
    out = C.POINTER(_ffi.Encoded)()
    check(ctx._lib.glc_encoded_from_bincode(ctx.handle, data, len(data), C.byref(out)))
    try:
        return EncodedAudio._from_struct(out.contents)
    finally:
        ctx._lib.glc_encoded_free(ctx.handle, out)


def save_encoded(encoded: EncodedAudio, path) -> None:
    """codec::save_encoded (src/codec.rs:774-779)."""
    with open(path, "wb") as f:
        f.write(encoded_to_bytes(encoded))


def load_encoded(path) -> EncodedAudio:
    """codec::load_encoded (src/codec.rs:781-786)."""
    with open(path, "rb") as f:
        return encoded_from_bytes(f.read())
