"""Waveform generators and SNR helpers: Python port of the reference's test
utilities (/root/reference/tests/utils.rs:5-173).  All arithmetic is float32 like
the Rust originals (np.sin on float32 may differ from glibc sinf in the last ulp;
that only perturbs the *inputs*, which the oracle and the CUDA path then share)."""
from __future__ import annotations

import numpy as np

F32 = np.float32
PI = F32(3.14159274101257324219)


def _t(sample_rate: int, duration: float):
    total = int(F32(sample_rate) * F32(duration))
    return total, (np.arange(total, dtype=F32) / F32(sample_rate)).astype(F32)


def _fan_out(mono: np.ndarray, channels: int) -> np.ndarray:
    return np.repeat(mono.astype(F32), channels)


def sine(freq, sample_rate, channels, duration, amp=0.5):
    """utils.rs:5-22"""
    _, t = _t(sample_rate, duration)
    ph = (F32(2.0) * PI * F32(freq)) * t
    return _fan_out(np.sin(ph.astype(F32)).astype(F32) * F32(amp), channels)


def square(freq, sample_rate, channels, duration):
    """utils.rs:25-43"""
    _, t = _t(sample_rate, duration)
    ph = ((F32(2.0) * PI * F32(freq)) * t).astype(F32)
    return _fan_out(np.where(np.sin(ph) >= 0, F32(0.3), F32(-0.3)), channels)


def sawtooth(freq, sample_rate, channels, duration):
    """utils.rs:46-64"""
    _, t = _t(sample_rate, duration)
    ph = np.fmod(((F32(2.0) * PI * F32(freq)) * t).astype(F32), F32(2.0) * PI).astype(F32)
    return _fan_out(((ph / PI) - F32(1.0)) * F32(0.3), channels)


def sweep(f0, f1, sample_rate, channels, duration):
    """utils.rs:67-86"""
    _, t = _t(sample_rate, duration)
    prog = (t / F32(duration)).astype(F32)
    f = (F32(f0) + (F32(f1) - F32(f0)) * prog).astype(F32)
    ph = (((F32(2.0) * PI) * f).astype(F32) * t).astype(F32)
    return _fan_out(np.sin(ph).astype(F32) * F32(0.3), channels)


def lcg_u64(seed: int, n: int) -> np.ndarray:
    """n successive states of state = state*1664525 + 1013904223 (mod 2^64), utils.rs:96."""
    a = np.uint64(1664525)
    c = np.uint64(1013904223)
    with np.errstate(over="ignore"):
        an = np.multiply.accumulate(np.full(n, a, np.uint64))  # a^1 .. a^n
        geo = np.concatenate([[np.uint64(1)], an[:-1]])  # a^0 .. a^(n-1)
        cs = np.add.accumulate(geo)  # sum_{j<k} a^j for k=1..n
        return an * np.uint64(seed) + c * cs


def white_noise(sample_rate, channels, duration, seed=12345, amp=0.6):
    """utils.rs:89-114: ((state as f32)/(u64::MAX as f32) - 0.5) * 0.6, channel-interleaved draws."""
    total = int(F32(sample_rate) * F32(duration)) * channels
    st = lcg_u64(seed, total)
    norm = st.astype(F32) / F32(18446744073709551615.0)
    return ((norm - F32(0.5)) * F32(amp)).astype(F32)


def snr_db(original: np.ndarray, decoded: np.ndarray) -> float:
    """utils.rs:118-147 (skips 1000 values at each end)."""
    m = min(len(original), len(decoded))
    if m < 2000:
        return 0.0
    o = original[1000:m - 1000].astype(np.float64)
    d = decoded[1000:m - 1000].astype(np.float64)
    sp = float(np.sum(o * o))
    npow = float(np.sum((o - d) ** 2))
    if npow > 0 and sp > 0:
        return 10.0 * np.log10(sp / npow)
    return float("inf") if npow == 0 else 0.0


def multi_sine(sample_rate, channels, duration, partials=20, base=100.0, step=173.0, seed=1):
    """SURVEY 8(d) config-2 style content: 20 partials at 100+173p Hz, amplitude 0.3/sqrt(20)."""
    _, t = _t(sample_rate, duration)
    out = np.zeros((len(t), channels), F32)
    for c in range(channels):
        acc = np.zeros(len(t), np.float64)
        for p in range(partials):
            f = base + step * p + 7.0 * c
            acc += np.sin(2 * np.pi * f * t.astype(np.float64) + 0.37 * p + c)
        out[:, c] = (acc * (0.3 / np.sqrt(partials))).astype(F32)
    return out.reshape(-1)


def music_like(sample_rate, channels, duration, seed=12345):
    """Deterministic mix exercising sparse AND raw frames: multi-sine + low-passed LCG noise for
    ~70 % of each second, white LCG noise (+-0.3) for the rest."""
    total = int(F32(sample_rate) * F32(duration))
    tones = multi_sine(sample_rate, channels, duration).reshape(total, channels)
    out = np.empty((total, channels), F32)
    for c in range(channels):
        st = lcg_u64(seed + 42109 * c, total)
        wn = ((st.astype(F32) / F32(18446744073709551615.0)) - F32(0.5)) * F32(0.6)
        # 8-tap moving average = cheap low-pass
        k = np.ones(8, np.float64) / 8.0
        lp = np.convolve(wn.astype(np.float64), k, mode="same").astype(F32) * F32(0.2)
        pos = (np.arange(total) % sample_rate) / sample_rate
        noisy = pos >= 0.7
        out[:, c] = np.where(noisy, wn, tones[:, c] + lp)
    return out.reshape(-1)
