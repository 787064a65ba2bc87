"""The product path must not route through the oracle or any CPU fallback: static checks over the
package sources, and the loader must fail loudly when the CUDA library is missing."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gapless_lossy_codec_b200")


def _sources():
    for d, _, files in os.walk(PKG):
        if "build" in d.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                yield os.path.join(d, f)


def test_package_never_touches_the_oracle():
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle|#include\s+[\"<].*oracle)", re.M)
    for path in _sources():
        text = open(path, encoding="utf-8", errors="replace").read()
        assert not pat.search(text), f"{path} references oracle/"
        assert "liboracle" not in text and "libglc_oracle" not in text, path


def test_missing_library_is_a_loud_error(monkeypatch):
    from gapless_lossy_codec_b200 import _ffi

    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setattr(_ffi, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(_ffi.GlcError) as e:
        _ffi.load()
    assert e.value.status == 6 and "no CPU fallback" in e.value.message


def test_no_device_is_a_loud_error_not_a_fallback():
    """On the CPU-only build box every entry point must refuse; on a GPU box this is covered by the
    parity tests (which would fail if anything but the CUDA path produced the bytes)."""
    import ctypes as C

    from gapless_lossy_codec_b200 import GlcError, _ffi
    from gapless_lossy_codec_b200.codec import Context

    n = C.c_int()
    if _ffi.load().glc_device_count(C.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(GlcError) as e:
        Context(0)
    assert e.value.status == 6
