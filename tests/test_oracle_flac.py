"""CPU tests of the FLAC oracle (no GPU): tests/test_flac.rs ported, with the independent RFC 9639
decoder (oracle/flac_decode.c) standing in for claxon, plus external anchors (hashlib MD5, CRC
remainder property, hand-checked header bytes)."""
import hashlib

import numpy as np
import pytest

import oracle
import signals


def to_i16(x):
    return np.trunc(np.clip(x.astype(np.float32) * np.float32(32767.0), -32768.0, 32767.0)).astype(np.int16)


def check_signal(x, sr, ch, level=5):
    """tests/test_flac.rs:4-52: rate, channels, sample count equal; RMS error < 1e-4."""
    data = oracle.flac_encode(x, sr, ch, level)
    info = oracle.flac_decode(data)
    assert data[:4] == b"fLaC"
    assert info["sample_rate"] == sr and info["channels"] == ch and info["bits_per_sample"] == 16
    assert len(info["samples"]) == len(x)
    loaded = info["samples"].astype(np.float32) / np.float32(32768.0)  # audio.rs: /2^(bits-1)
    rms = float(np.sqrt(np.mean((x.astype(np.float64) - loaded) ** 2)))
    assert rms < 1e-4
    assert info["md5_ok"]
    assert np.array_equal(info["samples"], to_i16(x).astype(np.int32))  # lossless at 16 bit
    assert hashlib.md5(to_i16(x).tobytes()).digest() == info["md5"]
    return data, info


def lcg_noise(n, seed=12345):
    """tests/test_flac.rs:79-90"""
    out = np.empty(n, np.float32)
    s = seed
    for i in range(n):
        s = (s * 1103515245 + 12345) & 0xFFFFFFFF
        out[i] = np.float32(((s >> 16) & 0x7FFF) / 32768.0) * np.float32(2.0) - np.float32(1.0)
    return out


def test_flac_silence():
    check_signal(np.zeros(1000, np.float32), 44100, 1)


def test_flac_dc_offset():
    check_signal(np.full(1000, 0.5, np.float32), 44100, 1)


def test_flac_sine_wave():
    check_signal(signals.sine(440, 44100, 1, 0.1, amp=0.8), 44100, 1)


def test_flac_white_noise():
    check_signal(lcg_noise(8820), 44100, 1)


def test_flac_stereo():
    l, r = signals.sine(440, 44100, 1, 0.1), signals.sine(880, 44100, 1, 0.1)
    check_signal(np.stack([l, r], 1).reshape(-1), 44100, 2)


def test_flac_sample_rates():
    check_signal(np.zeros(4800, np.float32), 48000, 1)
    check_signal(np.zeros(9600, np.float32), 96000, 1)


def test_flac_minimum_size():
    x = (np.arange(16, dtype=np.float32) / np.float32(16.0)) * np.float32(2.0) - np.float32(1.0)
    data = oracle.flac_encode(x, 8000, 1, 5)
    info = oracle.flac_decode(data)
    assert info["md5_ok"] and len(info["samples"]) == 16 and info["min_block"] == 16
    assert np.array_equal(info["samples"], to_i16(x).astype(np.int32))


@pytest.mark.parametrize("level", range(9))
def test_flac_compression_levels(level):  # tests/test_flac.rs:136-159
    x = signals.sine(440, 44100, 1, 1.0, amp=0.5)[:1000]
    _, info = check_signal(x, 44100, 1, level)
    assert info["min_block"] == 1000  # block = min(1152 | 4096, total)


def test_levels_6_7_8_identical_and_block_sizes():
    """levels 6, 7, 8 are the same encoder; 0-2 use 1152-sample blocks, 3-8 use 4096."""
    x = signals.music_like(44100, 2, 0.5)
    outs = [oracle.flac_encode(x, 44100, 2, l) for l in range(9)]
    assert outs[6] == outs[7] == outs[8]
    for l, o in enumerate(outs):
        info = oracle.flac_decode(o)
        assert info["max_block"] == (1152 if l <= 2 else 4096) and info["md5_ok"]
    assert len(outs[0]) > len(outs[5])  # verbatim vs order-4 fixed


def test_error_cases():
    with pytest.raises(oracle.OracleError, match="at least 16"):
        oracle.flac_encode(np.zeros(15, np.float32), 44100, 1, 5)
    with pytest.raises(oracle.OracleError, match="level"):
        oracle.flac_encode(np.zeros(100, np.float32), 44100, 1, 9)
    with pytest.raises(oracle.OracleError, match="at least 16"):
        oracle.flac_encode(np.zeros(10, np.float32), 44100, 1, 9)


def test_md5_and_crc_against_external_anchors():
    rng = np.random.default_rng(1)
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 1000, 100003):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.md5(b) == hashlib.md5(b).digest()
    b = rng.integers(0, 256, 4099, dtype=np.uint8).tobytes()
    # MSB-first CRCs with zero init: appending the CRC makes the remainder zero
    c16 = oracle.crc16(b)
    assert oracle.crc16(b + bytes([c16 >> 8, c16 & 0xFF])) == 0
    c8 = oracle.crc8(b)
    assert oracle.crc8(b + bytes([c8])) == 0
    assert oracle.crc8(b"\x00") == 0 and oracle.crc8(b"\x01") == 0x07 and oracle.crc16(b"\x01") == 0x8005


def test_frame_structure_of_a_known_stream():
    """hand-checkable header bytes: 44.1 kHz mono, block 4096 -> FF F8 | C9 | 08 | frame number | crc8"""
    x = signals.sine(440, 44100, 1, 0.5)
    data = oracle.flac_encode(x, 44100, 1, 5)
    assert data[4] == 0x80 and data[5:8] == b"\x00\x00\x22"  # last-block flag + STREAMINFO, length 34
    assert int.from_bytes(data[8:10], "big") == 4096 and int.from_bytes(data[10:12], "big") == 4096
    fr = data[42:]
    assert fr[0] == 0xFF and fr[1] == 0xF8 and fr[2] == 0xC9 and fr[3] == 0x08 and fr[4] == 0x00
    assert fr[5] == oracle.crc8(fr[:5])
    assert fr[6] == 0b00011000  # subframe header: 0 | 001100 (fixed, order 4) | 0
    info = oracle.flac_decode(data)
    assert info["n_frames"] == (len(x) + 4095) // 4096


def test_decoder_rejects_corruption():
    data = bytearray(oracle.flac_encode(signals.sine(440, 44100, 1, 0.2), 44100, 1, 5))
    data[60] ^= 0x10
    with pytest.raises(oracle.OracleError):
        oracle.flac_decode(bytes(data))
