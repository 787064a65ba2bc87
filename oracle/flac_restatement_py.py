"""oracle/flac_restatement_py.py -- TEST INFRASTRUCTURE ONLY.

A SECOND, independently written restatement of the reference's FLAC encoder (ajcm474/gapless-lossy-codec
v0.5.0, src/flac.rs) in plain Python, used for one thing: tests/test_oracle_restatements.py checks that
it and the C oracle (oracle/flac_oracle.c) produce byte-identical streams on small inputs at every
compression level.  (MD5 is the standard algorithm -- the reference carries its own implementation of it,
src/flac.rs:83-318 -- so hashlib stands in for it here; the C oracle's own MD5 is checked against hashlib
separately.)  Slow by design: one Python call per written bit field.
"""
from __future__ import annotations

import hashlib

import numpy as np

FLAC_SIGNATURE = b"fLaC"        # src/flac.rs:9
FRAME_SYNC_CODE = 0x3FFE        # :15
MAX_RICE_PARAM_4BIT = 14        # :12-13


def crc8(data: bytes) -> int:  # :19-51, polynomial 0x07
    crc = 0
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = ((crc << 1) ^ 0x07) & 0xFF if crc & 0x80 else (crc << 1) & 0xFF
    return crc


def crc16(data: bytes) -> int:  # :54-80, polynomial 0x8005, init 0
    crc = 0
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x8005) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


class BitWriter:  # :321-424 (MSB first)
    def __init__(self):
        self.buffer = bytearray()
        self.current = 0
        self.count = 0

    def write_bits(self, value: int, bits: int):
        value &= (1 << 64) - 1  # `as u64`: negative samples arrive sign-extended, the low `bits` bits are written
        for i in range(bits - 1, -1, -1):
            self.current = (self.current << 1) | ((value >> i) & 1)
            self.count += 1
            if self.count == 8:
                self.buffer.append(self.current)
                self.current = 0
                self.count = 0

    def write_byte(self, b: int):
        self.write_bits(b & 0xFF, 8)

    def write_unary(self, value: int):
        for _ in range(value):
            self.write_bits(0, 1)
        self.write_bits(1, 1)

    def byte_align(self):
        if self.count:
            self.buffer.append(self.current << (8 - self.count))
            self.current = 0
            self.count = 0

    def header_bytes(self, start: int) -> bytes:
        out = bytes(self.buffer[start:])
        if self.count:
            out += bytes([self.current << (8 - self.count)])
        return out

    def get_bytes(self) -> bytes:
        return self.header_bytes(0)


def write_utf8_number(w: BitWriter, v: int):  # :427-478
    if v < 0x80:
        w.write_byte(v)
    elif v < 0x800:
        w.write_byte(0xC0 | ((v >> 6) & 0x1F))
        w.write_byte(0x80 | (v & 0x3F))
    elif v < 0x10000:
        w.write_byte(0xE0 | ((v >> 12) & 0x0F))
        w.write_byte(0x80 | ((v >> 6) & 0x3F))
        w.write_byte(0x80 | (v & 0x3F))
    elif v < 0x200000:
        w.write_byte(0xF0 | ((v >> 18) & 0x07))
        for s in (12, 6, 0):
            w.write_byte(0x80 | ((v >> s) & 0x3F))
    elif v < 0x4000000:
        w.write_byte(0xF8 | ((v >> 24) & 0x03))
        for s in (18, 12, 6, 0):
            w.write_byte(0x80 | ((v >> s) & 0x3F))
    elif v < 0x80000000:
        w.write_byte(0xFC | ((v >> 30) & 0x01))
        for s in (24, 18, 12, 6, 0):
            w.write_byte(0x80 | ((v >> s) & 0x3F))
    else:
        w.write_byte(0xFE)
        for s in (30, 24, 18, 12, 6, 0):
            w.write_byte(0x80 | ((v >> s) & 0x3F))


def apply_fixed_predictor(s, order):  # :481-512
    res = []
    for i in range(len(s)):
        if i < order:
            res.append(0)
            continue
        pred = (0,
                s[i - 1] if order >= 1 else 0,
                2 * s[i - 1] - s[i - 2] if order >= 2 else 0,
                3 * s[i - 1] - 3 * s[i - 2] + s[i - 3] if order >= 3 else 0,
                4 * s[i - 1] - 6 * s[i - 2] + 4 * s[i - 3] - s[i - 4] if order >= 4 else 0)[order]
        res.append(s[i] - pred)
    return res


def calculate_rice_parameter(res):  # :515-552
    if not res:
        return 0
    mean = sum(abs(x) for x in res) // len(res)
    if mean == 0:
        return 0
    param, test = 0, mean
    while test > 0 and param < MAX_RICE_PARAM_4BIT:
        test >>= 1
        if test > 0:
            param += 1
    if param > 0 and mean < (1 << (param - 1)):
        param -= 1
    return min(param, MAX_RICE_PARAM_4BIT)


def encode_rice_partition(w: BitWriter, res, k):  # :555-584
    for x in res:
        folded = (x << 1) if x >= 0 else (((-(x + 1)) << 1) | 1)
        w.write_unary(folded >> k)
        if k > 0:
            w.write_bits(folded & ((1 << k) - 1), k)


def encode_residual(w: BitWriter, res, order, block_size, level):  # :587-684
    tz = (block_size & -block_size).bit_length() - 1
    cap = min(tz, 8)
    po = 0 if level == 0 else min(2, cap) if level <= 2 else min(4, cap) if level <= 5 else min(6, cap)
    while po > 0:
        ps = block_size >> po
        if ps > order and ps >= 4:
            break
        po -= 1
    w.write_bits(0, 2)
    w.write_bits(po, 4)
    default = block_size >> po
    idx = 0
    for p in range(1 << po):
        n = default - order if p == 0 else default
        if n == 0:
            continue
        part = res[idx:idx + n]
        idx += n
        k = calculate_rice_parameter(part)
        assert k <= MAX_RICE_PARAM_4BIT  # the escape branch (:643-672) is unreachable
        w.write_bits(k, 4)
        encode_rice_partition(w, part, k)


def encode_subframe(w: BitWriter, s, bps, level):  # :687-745
    n = len(s)
    order = (0,
             1 if n >= 1 else 0,
             2 if n >= 2 else 0,
             3 if n >= 3 else 0, 3 if n >= 3 else 0,
             4 if n >= 4 else 0, 4 if n >= 4 else 0, 4 if n >= 4 else 0, 4 if n >= 4 else 0)[level]
    w.write_bits(0, 1)
    w.write_bits(0b000001 if order == 0 else (0b001000 | order), 6)
    w.write_bits(0, 1)
    if order == 0:
        for x in s:
            w.write_bits(x, bps)
    else:
        for i in range(order):
            w.write_bits(s[i], bps)
        encode_residual(w, apply_fixed_predictor(s, order)[order:], order, n, level)


_BLOCK_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13,
                16384: 14, 32768: 15}
_RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10,
               96000: 11}


def encode_frame(w: BitWriter, samples, channels, rate, bps, frame_number, block_size, level):  # :748-905
    start = len(w.buffer)
    w.write_bits(FRAME_SYNC_CODE, 14)
    w.write_bits(0, 1)
    w.write_bits(0, 1)
    bcode = _BLOCK_CODES.get(block_size, 0b0110 if block_size < 256 else 0b0111)
    w.write_bits(bcode, 4)
    w.write_bits(_RATE_CODES.get(rate, 0), 4)
    w.write_bits(0 if channels == 1 else 1 if channels == 2 else channels - 1, 4)
    w.write_bits({8: 1, 12: 2, 16: 4, 20: 5, 24: 6}.get(bps, 0), 3)
    w.write_bits(0, 1)
    write_utf8_number(w, frame_number)
    if bcode == 0b0110:
        w.write_byte((block_size - 1) & 0xFF)
    elif bcode == 0b0111:
        w.write_bits(block_size - 1, 16)
    w.write_byte(crc8(w.header_bytes(start)))
    for c in range(channels):
        chan = [int(samples[i * channels + c]) if i * channels + c < len(samples) else 0 for i in range(block_size)]
        encode_subframe(w, chan, bps, level)
    w.byte_align()
    w.write_bits(crc16(bytes(w.buffer[start:])), 16)


def encode_flac_with_level(samples, sample_rate: int, channels: int, level: int) -> bytes:  # :947-1052
    v = np.asarray(samples, np.float32) * np.float32(32767.0)
    i16 = np.trunc(np.clip(v, np.float32(-32768.0), np.float32(32767.0))).astype(np.int16)
    total = len(i16) // channels
    if total < 16:
        raise ValueError(f"FLAC requires at least 16 samples per channel, got {total}")
    if level > 8:
        raise ValueError(f"Invalid compression level {level}, must be 0-8")
    bps = 16
    block_size = max(min(1152 if level <= 2 else 4096, total), 16)
    w = BitWriter()
    for b in FLAC_SIGNATURE:
        w.write_byte(b)
    md5 = hashlib.md5(i16.astype("<i2").tobytes()).digest()  # :305-318: every interleaved sample, little-endian
    w.write_bits(1, 1)
    w.write_bits(0, 7)
    w.write_bits(34, 24)
    w.write_bits(block_size & 0xFFFF, 16)
    w.write_bits(block_size & 0xFFFF, 16)
    w.write_bits(0, 24)
    w.write_bits(0, 24)
    w.write_bits(sample_rate, 20)
    w.write_bits(channels - 1, 3)
    w.write_bits(bps - 1, 5)
    w.write_bits(total, 36)
    for b in md5:
        w.write_byte(b)
    off, frame = 0, 0
    while off < len(i16):
        cur = min(block_size, (len(i16) - off) // channels)
        if cur == 0:
            break
        encode_frame(w, i16[off:off + cur * channels], channels, sample_rate, bps, frame, cur, level)
        off += cur * channels
        frame += 1
    return w.get_bytes()
