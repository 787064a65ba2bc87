"""Host-side logic of libglc_b200.so that needs no GPU: the .glc container reader / writer (hostile headers
included) and the shard planner / up-front validation of the sharded calls."""
import ctypes as C
import struct

import numpy as np
import pytest

import oracle
import signals
from gapless_lossy_codec_b200 import _ffi


def _lib():
    return _ffi.load()


def _from_bincode(blob: bytes):
    out = C.POINTER(_ffi.Encoded)()
    st = _lib().glc_encoded_from_bincode(None, blob, len(blob), C.byref(out))
    return st, out


def test_container_round_trip_matches_the_oracle_image():
    """save_encoded / load_encoded byte image (src/codec.rs:774-786): reading the oracle's image and writing it
    back gives the same bytes; the reader runs on the host only (context not needed)."""
    e = oracle.encode(signals.music_like(44100, 2, 0.4), 2, 44100)
    blob = oracle.bincode_serialize(e)
    st, out = _from_bincode(blob)
    assert st == 0
    try:
        b, n = C.POINTER(C.c_uint8)(), C.c_uint64()
        assert _lib().glc_encoded_to_bincode(None, out, C.byref(b), C.byref(n)) == 0
        assert C.string_at(b, n.value) == blob
        C.CDLL(None).free(b)  # the image is malloc'ed (glc_free would need a context)
        got = oracle.arrays_from_struct(out.contents)
        assert np.array_equal(got.pair_q, e.pair_q) and np.array_equal(got.raw, e.raw)
    finally:
        _lib().glc_encoded_free(None, out)


def _header(rate=44100, ch=2, total=0, n_frames=0):
    return struct.pack("<IHQQ", rate, ch, total, n_frames)


@pytest.mark.parametrize("blob,why", [
    (b"", "empty"),
    (_header()[:10], "truncated header"),
    (_header(ch=0) + b"\0" * 16, "zero channels"),
    (_header(n_frames=10 ** 15) + b"\0" * 16, "frame count far beyond the image size"),
    (_header(ch=65535, n_frames=60000) + b"\0" * (17 * 60000 + 16), "rows far beyond what the image can hold"),
    (_header(n_frames=1) + struct.pack("<Q", 3) + b"\0" * 40, "frame with a wrong number of channel vectors"),
], ids=lambda v: v if isinstance(v, str) else None)
def test_container_rejects_hostile_images_without_large_allocations(blob, why):
    """sizes come from an untrusted header: every count is bounded by the image size before anything is
    allocated (a 1 MB file could request hundreds of GB)."""
    st, out = _from_bincode(blob)
    assert st == 8, why  # GLC_ERR_CORRUPT
    assert not out


def test_container_accepts_a_stream_without_frames():
    blob = _header(total=1234, n_frames=0) + struct.pack("<IIQ", 512, 0, 1234)
    st, out = _from_bincode(blob)
    assert st == 0 and out.contents.n_frames == 0 and out.contents.original_length == 1234
    _lib().glc_encoded_free(None, out)


def test_sharded_calls_validate_every_file_before_planning():
    """glc_encode_batch_sharded refuses short / ragged files up front (no shard is touched, shard_of stays
    meaningless but no weight underflows), glc_plan_shards is the deterministic LPT plan."""
    L = _lib()
    n = 3
    bufs = [np.zeros(4096, np.float32), np.zeros(600, np.float32), np.zeros(4097, np.float32)]
    ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    outs = (C.POINTER(_ffi.Encoded) * n)()
    shard_of = (C.c_uint32 * n)()
    encs = (C.c_void_p * 2)(1, 1)  # never dereferenced: validation comes first

    def call(sizes, chans):
        ns = (C.c_uint64 * n)(*sizes)
        chs = (C.c_uint16 * n)(*chans)
        return L.glc_encode_batch_sharded(encs, 2, n, ptrs, ns, chs, outs, shard_of)

    assert call([4096, 600, 4096], [2, 2, 2]) == 2  # file 1: 300 samples per channel (the reference panics)
    assert b"file 1" in L.glc_last_error()
    assert call([4096, 600, 4097], [1, 1, 2]) == 1  # file 2: not a multiple of the channel count
    assert call([4096, 600, 4096], [1, 0, 1]) == 1  # zero channels
    w = (C.c_uint64 * 5)(10, 7, 7, 3, 1)
    plan = (C.c_uint32 * 5)()
    assert L.glc_plan_shards(5, w, 2, plan) == 0
    # longest first, each to the lighter shard: 10 -> s0; 7 -> s1; 7 -> s1 (14); 3 -> s0 (13); 1 -> s0 (13 < 14)
    assert list(plan) == [0, 1, 1, 0, 0]
