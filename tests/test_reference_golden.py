"""Pinning against the REAL reference crate (when its dump is present).

tests/golden/dump_reference.rs is a Cargo integration test for the reference checkout: it runs the crate's
public API (Encoder::encode, Decoder::decode, decode_streaming, save_encoded, flac::encode_flac_with_level)
over inputs it generates itself and writes inputs + outputs to a directory.  Copy that directory to
tests/golden/ref_v1/ and these tests compare, on the dumped inputs,

  * CPU (-m "not gpu"): the oracle with the crate -- .glc image byte for byte (every index, value, scale
    bit pattern, raw frame, gapless field and the bincode layout itself), decoded PCM bit for bit, chunk
    shapes of the stream, FLAC bytes at every dumped level;
  * GPU (-m gpu): the CUDA path, through the C ABI, with the same files.

Without the dump the parity claim of this repo stays "oracle-pinned only": the tests then SKIP with the
reason "parity unpinned" (this image has no cargo/rustc, so the dump cannot be produced here).
"""
import os

import numpy as np
import pytest

import oracle

REF_DIR = os.environ.get("GLC_REF_DIR") or os.path.join(os.path.dirname(__file__), "golden", "ref_v1")
HAVE = os.path.exists(os.path.join(REF_DIR, "manifest.txt"))
UNPINNED = ("parity unpinned: tests/golden/ref_v1/ is absent (produce it with tests/golden/dump_reference.rs "
            "on a machine with a Rust toolchain)")


def _manifest():
    if not HAVE:
        return []
    out = []
    with open(os.path.join(REF_DIR, "manifest.txt")) as f:
        for line in f:
            t = line.split()
            if t:
                out.append((t[0], t[1], int(t[2]), int(t[3]), [int(v) for v in t[4:]]))
    return out


MAN = _manifest()
CODEC = [m for m in MAN if m[0] == "codec"]
FLAC = [(m[1], m[2], m[3], lv) for m in MAN if m[0] == "flac" for lv in m[4]]


def _f32(name):
    return np.fromfile(os.path.join(REF_DIR, name), dtype="<f4")


def _bytes(name):
    with open(os.path.join(REF_DIR, name), "rb") as f:
        return f.read()


def _same_libm() -> bool:
    """dump host and this host must build the same cosine table (Rust's f32::cos is the platform libm)."""
    cos_tab, window, _ = oracle.tables()
    want = _bytes("table.fnv").split()
    return (f"{oracle.fnv1a64(cos_tab.view(np.uint8).reshape(-1)):016x}".encode() == want[0]
            and f"{oracle.fnv1a64(window.view(np.uint8).reshape(-1)):016x}".encode() == want[1])


def test_pinning_status():
    """Always runs: states which of the two situations this checkout is in."""
    if not HAVE:
        pytest.skip(UNPINNED)
    assert CODEC and FLAC, "tests/golden/ref_v1/manifest.txt holds no cases"
    if not _same_libm():
        pytest.skip("reference dump present, but this host's libm builds a different cosine table than the dump "
                    "host's: coefficient-level comparisons are not meaningful here")


def _chunks(name):
    return [(int(a), bool(int(b))) for a, b in (l.split() for l in _bytes(name + ".chunks.txt").decode().splitlines() if l)]


def _bits_equal(a, b, what):
    assert a.shape == b.shape, f"{what}: length {a.shape} vs {b.shape}"
    d = np.flatnonzero(a.view(np.uint32) != b.view(np.uint32))
    assert len(d) == 0, f"{what}: {len(d)} of {len(a)} values differ bitwise, first at {d[:5]}"


@pytest.mark.skipif(not HAVE, reason=UNPINNED)
@pytest.mark.parametrize("case", CODEC, ids=[c[1] for c in CODEC])
def test_oracle_equals_reference_codec(case):
    _, name, sr, ch, _ = case
    if not _same_libm():
        pytest.skip("libm of the dump host differs")
    x = _f32(name + ".in.f32")
    e = oracle.encode(x, ch, sr)
    blob = _bytes(name + ".glc")
    mine = oracle.bincode_serialize(e)
    assert mine == blob, f"{name}: .glc image differs from the reference's ({len(mine)} vs {len(blob)} bytes)"
    ref_stream = oracle.bincode_deserialize(blob)
    _bits_equal(oracle.decode(ref_stream), _f32(name + ".pcm.f32"), f"{name}: Decoder::decode")
    _bits_equal(oracle.decode(ref_stream, trimmed=False), _f32(name + ".stream.f32"), f"{name}: decode_streaming")
    n, per = ref_stream.n_frames, 1024 * ch
    want = [(500 * per, False)] * (n // 500) + [((n % 500 + 1) * per, True)]
    assert _chunks(name) == want, f"{name}: chunk shapes"


@pytest.mark.skipif(not HAVE, reason=UNPINNED)
@pytest.mark.parametrize("case", FLAC, ids=[f"{c[0]}_l{c[3]}" for c in FLAC])
def test_oracle_equals_reference_flac(case):
    name, sr, ch, level = case
    assert oracle.flac_encode(_f32(name + ".in.f32"), sr, ch, level) == _bytes(f"{name}.l{level}.flac")


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE, reason=UNPINNED)
@pytest.mark.parametrize("case", CODEC, ids=[c[1] for c in CODEC])
def test_cuda_equals_reference_codec(gpu_ctx, case):
    from gapless_lossy_codec_b200 import Decoder, Encoder, encoded_from_bytes, encoded_to_bytes

    _, name, sr, ch, _ = case
    if not _same_libm():
        pytest.skip("libm of the dump host differs")
    x = _f32(name + ".in.f32")
    blob = _bytes(name + ".glc")
    enc = Encoder(sr, gpu_ctx).encode(x, ch)
    assert encoded_to_bytes(enc, gpu_ctx) == blob, f"{name}: .glc image differs from the reference's"
    dec = Decoder(ch, sr, gpu_ctx)
    ref_stream = encoded_from_bytes(blob, gpu_ctx)
    _bits_equal(dec.decode(ref_stream), _f32(name + ".pcm.f32"), f"{name}: Decoder::decode")
    chunks = list(dec.decode_streaming(ref_stream))
    _bits_equal(np.concatenate([c.samples for c in chunks]), _f32(name + ".stream.f32"), f"{name}: decode_streaming")
    assert [(len(c.samples), c.is_last) for c in chunks] == _chunks(name)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE, reason=UNPINNED)
@pytest.mark.parametrize("case", FLAC, ids=[f"{c[0]}_l{c[3]}" for c in FLAC])
def test_cuda_equals_reference_flac(gpu_ctx, case):
    from gapless_lossy_codec_b200 import flac

    name, sr, ch, level = case
    assert flac.encode_flac_with_level(_f32(name + ".in.f32"), sr, ch, level, gpu_ctx) == _bytes(f"{name}.l{level}.flac")
