"""oracle/codec_restatement_np.py -- TEST INFRASTRUCTURE ONLY.

A SECOND, independently written restatement of the reference's codec (ajcm474/gapless-lossy-codec
v0.5.0, src/codec.rs) in numpy, used for one thing: tests/test_oracle_restatements.py checks that it
and the C oracle (oracle/codec_oracle.c) -- two restatements written separately, from the Rust
source, in different languages and with different loop structures -- produce bit-identical
streams and PCM on small inputs.  Agreement does not pin the oracle to the reference binary (no
Rust toolchain in this image, see oracle_common.h: "parity unpinned"), but it does rule out slips of
transcription in either restatement.

How Rust's semantics are kept in numpy:
  * every value is np.float32; every product / sum is one IEEE single operation (no fused ops);
  * `s += a*b` loops and iterator `.sum::<f32>()` are left-to-right: np.cumsum over float32 (ufunc
    accumulate is a plain sequential loop), with a leading 0.0 so the chain starts like the source;
  * f32::cos / sin / powf / sqrt are the platform libm's cosf / sinf / powf / sqrtf (what Rust links
    on Linux), called through ctypes -- numpy's own float32 cos/sin are a different implementation;
  * `as usize` / `as i16` from f32 truncate toward zero (saturating); f32::round is half away from 0.
It is slow (a dense 1024 x 2048 product per frame-channel) and meant for inputs of a few frames.
"""
from __future__ import annotations

import ctypes
import ctypes.util
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("cosf", "sinf", "sqrtf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]

FRAME_SIZE = 2048          # src/codec.rs:15
HOP_SIZE = 1024            # :16
QUANTIZATION_BITS = 16     # :17
NOISE_FLOOR_DB = F(-48.0)  # :22
QUALITY_FACTOR = F(0.7)    # :23
COMPRESSION_THRESHOLD = F(0.85)  # :29
PI = F(np.pi)              # std::f32::consts::PI


def _seq_sum(v: np.ndarray) -> np.float32:
    """left-to-right f32 sum starting from 0.0"""
    if v.size == 0:
        return F(0.0)
    return np.cumsum(np.concatenate([np.zeros(1, F), v.astype(F)]), dtype=F)[-1]


class MdctTables:  # src/codec.rs:316-391
    _cache = None

    def __init__(self, n: int = HOP_SIZE):
        block = FRAME_SIZE
        table = np.empty((n, block), F)
        a0 = PI / F(n)                                 # PI / (n as f32)
        half = F(n) / F(2.0)
        ii = (np.arange(block, dtype=F) + F(0.5)) + half  # (i as f32 + 0.5 + (n as f32)/2.0), left to right
        pre = (a0 * ii).astype(F)                      # PI/n * (...)
        for k in range(n):
            ang = (pre * (F(k) + F(0.5))).astype(F)    # ... * (k as f32 + 0.5)
            table[k] = [_libm.cosf(float(x)) for x in ang]
        self.cos_table = table
        self.window = np.array([_libm.sinf(float((PI * (F(i) + F(0.5))) / F(block))) for i in range(block)], F)
        self.n = n
        self.norm = F(_libm.sqrtf(float(F(2.0) / F(n))))

    @classmethod
    def get(cls) -> "MdctTables":
        if cls._cache is None:
            cls._cache = cls()
        return cls._cache

    def mdct_block(self, block: np.ndarray) -> np.ndarray:  # :359-374
        prod = (self.cos_table * block[None, :].astype(F)).astype(F)       # block[i] * tb[i], each rounded
        s = np.cumsum(np.concatenate([np.zeros((self.n, 1), F), prod], axis=1), axis=1, dtype=F)[:, -1]
        return (s * self.norm).astype(F)

    def imdct_block(self, coeffs: np.ndarray) -> np.ndarray:  # :377-390
        prod = (self.cos_table * coeffs[:, None].astype(F)).astype(F)      # coeffs[k] * base[k*FRAME_SIZE + i]
        s = np.cumsum(np.concatenate([np.zeros((1, FRAME_SIZE), F), prod], axis=0), axis=0, dtype=F)[-1]
        return (s * self.norm).astype(F)


class PerceptualWeights:  # src/codec.rs:92-183
    def __init__(self, n: int, sample_rate: int):
        w = np.empty(n, F)
        for k in range(n):
            norm_freq = F(k) / (F(2.0) * F(n))
            f = F(norm_freq * F(sample_rate))
            if f < F(100.0):
                v = F(0.3) + F(F(f / F(100.0)) * F(0.4))
            elif f < F(200.0):
                v = F(0.7) + F(F(F(f - F(100.0)) / F(100.0)) * F(0.3))
            elif f < F(5000.0):
                v = F(1.0)
            elif f < F(10000.0):
                v = F(1.0) - F(F(F(f - F(5000.0)) / F(5000.0)) * F(0.3))
            else:
                v = F(0.7) - F(min(F(F(f - F(10000.0)) / F(12000.0)), F(1.0)) * F(0.5))
            w[k] = max(F(v), F(0.2))
        self.weights = w
        self.critical_bands = self._bands(n, sample_rate)

    @staticmethod
    def _bands(n: int, sample_rate: int) -> List[int]:
        bands = [0]
        nyquist = F(sample_rate) / F(2.0)
        freq = F(0.0)
        while freq < nyquist and len(bands) < 50:
            b = int(F(F(freq / nyquist) * F(n)))  # as usize: truncation (values are >= 0)
            if b > bands[-1] and b < n:
                bands.append(b)
            if freq < F(500.0):
                freq = F(freq + F(50.0))
            elif freq < F(2000.0):
                freq = F(freq + F(100.0))
            elif freq < F(8000.0):
                freq = F(freq + F(250.0))
            else:
                freq = F(freq + F(500.0))
        bands.append(n)
        return bands


def compute_masking_thresholds(coeffs: np.ndarray, quality: np.float32, perc: PerceptualWeights) -> np.ndarray:
    """src/codec.rs:188-240"""
    n = coeffs.size
    th = np.zeros(n, F)
    global_max = max(F(np.max(np.abs(coeffs))) if n else F(0.0), F(1e-10))
    w = perc.weights
    edges = perc.critical_bands
    for bi in range(max(len(edges) - 1, 0)):
        start, end = edges[bi], min(edges[bi + 1], n)
        if start >= end:
            continue
        cnt = F(end - start)
        sq = (coeffs[start:end] * coeffs[start:end]).astype(F)
        energy = F(_libm.sqrtf(float(F(_seq_sum(sq) / cnt))))
        avg_weight = F(_seq_sum(w[start:end]) / cnt)
        compression_factor = max(F(F(1.0) - quality), F(0.01))
        perceptual_factor = F(F(1.0) / max(avg_weight, F(0.1)))
        base = F(F(F(energy * F(0.01)) * compression_factor) * perceptual_factor)
        for i in range(start, end):
            individual = F(F(1.0) / max(w[i], F(0.1)))
            t = F(base * individual)
            if abs(coeffs[i]) > F(global_max * F(0.3)):
                t = min(t, F(global_max * F(0.05)))
            th[i] = t
    return th


def _round_half_away(x: np.float32) -> np.float32:
    return F(np.copysign(np.floor(np.abs(np.float64(x)) + 0.5), np.float64(x)))  # exact in f64 for |x| < 2^52


def compress_coefficients(coeffs: np.ndarray, scale: np.float32, th: np.ndarray, noise_floor_db: np.float32
                          ) -> List[Tuple[int, int]]:
    """src/codec.rs:270-311 (compute_quantization_bits_fast can only return 0 when abs <= threshold,
    which the caller has already excluded, so it never drops anything)"""
    noise_floor_linear = F(F(_libm.powf(10.0, float(F(noise_floor_db / F(20.0))))) * scale)
    max_q = F(1 << (QUANTIZATION_BITS - 1))
    out = []
    for k in range(coeffs.size):
        c = coeffs[k]
        a = abs(c)
        threshold = F(th[k] * scale)
        if a > noise_floor_linear and a > threshold:
            if a <= threshold:  # importance_bits == 0: unreachable, kept for the shape of the source
                continue
            normalized = F(c / scale)
            quantized = _round_half_away(F(normalized * max_q))
            q = int(min(max(quantized, F(-32768.0)), F(32767.0)))
            if q != 0:
                out.append((k, q))
    return out


@dataclass
class Frame:
    sparse: List[List[Tuple[int, int]]] = field(default_factory=list)
    scales: List[np.float32] = field(default_factory=list)
    raw_pcm: Optional[np.ndarray] = None


@dataclass
class Encoded:
    sample_rate: int
    channels: int
    total_samples: int
    frames: List[Frame]
    encoder_delay: int
    padding: int
    original_length: int


def encode(samples: np.ndarray, channels: int, sample_rate: int) -> Encoded:
    """Encoder::encode, src/codec.rs:421-565"""
    samples = np.asarray(samples, F)
    ch = channels
    tables = MdctTables.get()
    perc = PerceptualWeights(HOP_SIZE, sample_rate)
    per_chan = [samples[c::ch] for c in range(ch)]
    padded = []
    for c in range(ch):
        v = np.concatenate([np.zeros(HOP_SIZE // 2, F), per_chan[c]])
        rem = v.size % HOP_SIZE
        if rem:
            v = np.concatenate([v, np.zeros(HOP_SIZE - rem, F)])
        padded.append(np.concatenate([v, np.zeros(HOP_SIZE // 2, F)]))
    num_frames = 1 if padded[0].size < FRAME_SIZE else (padded[0].size - FRAME_SIZE) // HOP_SIZE + 1
    frames = []
    for fi in range(num_frames):
        fr = Frame()
        raw = []
        for c in range(ch):
            sl = padded[c][fi * HOP_SIZE: fi * HOP_SIZE + FRAME_SIZE]
            block = (sl * tables.window).astype(F)
            coeffs = tables.mdct_block(block)
            max_val = max(F(np.max(np.abs(coeffs))), F(1e-10))
            fr.scales.append(max_val)
            th = compute_masking_thresholds(coeffs, QUALITY_FACTOR, perc)
            fr.sparse.append(compress_coefficients(coeffs, max_val, th, NOISE_FLOOR_DB))
            v = (block * F(32767.0)).astype(F)
            raw.append(np.trunc(np.clip(v, F(-32768.0), F(32767.0))).astype(np.int16))  # planar, :498-502
        compressed = sum(8 + 4 * len(s) for s in fr.sparse) + 8 + 4 * len(fr.scales) + 64
        raw_size = FRAME_SIZE * ch * 2
        if F(compressed) >= F(F(raw_size) * COMPRESSION_THRESHOLD):
            fr = Frame([], [], np.concatenate(raw))
        frames.append(fr)
    padded_len, orig_len = padded[0].size, per_chan[0].size
    return Encoded(sample_rate, ch, samples.size, frames, HOP_SIZE // 2, padded_len - orig_len - HOP_SIZE // 2,
                   samples.size)


def decode(enc: Encoded) -> np.ndarray:
    """Decoder::decode = decode_streaming + gapless trim, src/codec.rs:595-768"""
    tables = MdctTables.get()
    ch = enc.channels
    overlap = [np.zeros(HOP_SIZE, F) for _ in range(ch)]
    out = []
    for fr in enc.frames:
        blocks = []
        if fr.raw_pcm is not None:
            for c in range(ch):
                b = np.zeros(FRAME_SIZE, F)
                idx = np.arange(FRAME_SIZE) * ch + c  # interleaved read of the planar body, :633-640
                ok = idx < fr.raw_pcm.size
                b[ok] = (fr.raw_pcm[idx[ok]].astype(F) / F(32767.0)).astype(F)
                blocks.append(b)
        else:
            for c in range(ch):
                coeffs = np.zeros(tables.n, F)
                scale = max(F(fr.scales[c]), F(1e-12))
                max_q = F(1 << (QUANTIZATION_BITS - 1))
                for (index, q) in fr.sparse[c]:
                    if index < tables.n:
                        coeffs[index] = F(F(F(q) / max_q) * scale)
                blocks.append((tables.imdct_block(coeffs) * tables.window).astype(F))
        inter = np.empty(HOP_SIZE * ch, F)
        for c in range(ch):
            inter[c::ch] = (overlap[c] + blocks[c][:HOP_SIZE]).astype(F)
            overlap[c] = blocks[c][HOP_SIZE:].copy()
        out.append(inter)
    tail = np.empty(HOP_SIZE * ch, F)
    for c in range(ch):
        tail[c::ch] = overlap[c]
    out.append(tail)
    allv = np.concatenate(out)
    if allv.size > enc.encoder_delay:
        allv = allv[enc.encoder_delay:]
    if allv.size > enc.original_length:
        allv = allv[: enc.original_length]
    return allv
