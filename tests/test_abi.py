"""CPU-side checks of the drop-in boundary (no compute calls without a GPU): the C-ABI library
loads, exports every symbol include/glc.h declares, and fails loudly -- never falls back -- when
there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "glc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gapless_lossy_codec_b200 import _ffi

    lib = _ffi.load()
    declared = _declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/glc.h but not exported: {missing}"
    assert sorted(_ffi.EXPORTS) == declared, "the ctypes binding must cover exactly the declared ABI"
    assert lib.glc_abi_version() == 7


def test_struct_layout_matches_header():
    from gapless_lossy_codec_b200 import _ffi

    assert C.sizeof(_ffi.Pair) == 4
    assert C.sizeof(_ffi.Encoded) == 96
    assert _ffi.Encoded.frame_is_raw.offset == 40 and _ffi.Encoded.raw.offset == 88
    import oracle

    assert C.sizeof(oracle.Encoded) == C.sizeof(_ffi.Encoded)
    for (n1, _), (n2, _) in zip(oracle.Encoded._fields_, _ffi.Encoded._fields_):
        assert n1 == n2 and getattr(oracle.Encoded, n1).offset == getattr(_ffi.Encoded, n2).offset


def _gpu_present():
    from gapless_lossy_codec_b200 import _ffi

    n = C.c_int()
    return _ffi.load().glc_device_count(C.byref(n)) == 0


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every product entry point must raise; nothing may route to the oracle."""
    if _gpu_present():
        pytest.skip("a CUDA device is present")
    import numpy as np
    from gapless_lossy_codec_b200 import Context, Encoder, GlcError, flac

    with pytest.raises(GlcError) as e:
        Context(0)
    assert e.value.status == 6 and "no CPU fallback" in e.value.message
    with pytest.raises(GlcError):
        Encoder(44100)
    with pytest.raises(GlcError):
        flac.encode_flac(np.zeros(100, np.float32), 44100, 1)


def test_product_package_never_uses_the_oracle():
    pkg = os.path.join(ROOT, "gapless_lossy_codec_b200")
    for dirpath, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d not in ("build", "__pycache__")]
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "libglc_oracle", "oracle/", "orc_"):
                    assert needle not in text, f"{f} references the oracle ({needle})"


def test_rust_shim_binds_only_exported_symbols_and_matches_integration_md():
    """The Rust shim cannot be compiled in this image (no cargo), so what can be checked is checked: every
    function its `extern "C"` block declares is exported by the library with the same number of parameters as
    include/glc.h declares, and the copies embedded in INTEGRATION.md are the files of integration/rust_shim/."""
    from gapless_lossy_codec_b200 import _ffi

    lib = _ffi.load()
    shim = os.path.join(ROOT, "integration", "rust_shim")
    codec = open(os.path.join(shim, "src", "codec.rs")).read()
    block = re.search(r'unsafe extern "C" \{(.*?)\n    \}', codec, flags=re.S).group(1)
    block = re.sub(r"//[^\n]*", "", block)
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "glc.h")).read(), flags=re.S)
    fns = re.findall(r"pub fn (glc_[a-z0-9_]+)\s*\((.*?)\)", block, flags=re.S)
    assert len(fns) >= 14
    for name, params in fns:
        assert hasattr(lib, name), f"{name}: bound by the shim, not exported"
        c_params = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, header, flags=re.S).group(1)
        n_c = 0 if c_params.strip() in ("", "void") else c_params.count(",") + 1
        n_rs = 0 if not params.strip() else params.count(",") + 1
        assert n_c == n_rs, f"{name}: {n_rs} parameters in the shim, {n_c} in include/glc.h"
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```rust\n(.*?)\n```", md, flags=re.S)
    for rel in ("build.rs", os.path.join("src", "codec.rs"), os.path.join("src", "flac.rs")):
        body = open(os.path.join(shim, rel)).read().rstrip("\n")
        assert any(b.rstrip("\n") == body for b in blocks), f"INTEGRATION.md does not embed {rel} as it is on disk"


def test_exact_kernels_never_contract_a_multiply_add():
    """Bit-exact parity rests on two roundings per multiply-add.  Read it off the SASS of
    the built library: the EXACT contractions hold no scalar FFMA, and every packed FFMA2 is one of the two exact
    forms (product + (-0.0), or addend + p * 1.0, the constant in a uniform register) in equal numbers -- ptxas has
    been seen to contract a packed multiply / add pair into one fused FFMA2, which this test would catch on a
    machine without a GPU."""
    import shutil
    import sys

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_histogram

    hist = sass_histogram.histogram()
    for kernel in ("exact_gemm_kernel<8>", "imdct_sparse_kernel"):
        h = hist[kernel]
        assert h["FFMA"] == 0, (kernel, "scalar fused multiply-add in an EXACT kernel")
        assert h["x2fused"] == 0, (kernel, "a packed multiply / add pair was contracted")
        if h["FFMA2"]:
            assert h["x2mul"] == h["x2add"] == h["FFMA2"] // 2, (kernel, dict(h))
        else:
            assert h["FMUL"] >= h["FADD"] > 0, (kernel, dict(h))  # the scalar build: FMUL + FADD pairs
        assert h["UBLKCP"] > 0 and h["SYNCS"] > 0, (kernel, "operands must arrive by TMA bulk copies on mbarriers")
