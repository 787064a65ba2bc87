import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _have_gpu() -> bool:
    try:
        import ctypes as C
        from gapless_lossy_codec_b200 import _ffi

        n = C.c_int()
        return _ffi.load().glc_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx():
    """The product context.  GPU tests FAIL (not skip) if the CUDA library cannot be used: a silent
    fallback would void the parity claims."""
    from gapless_lossy_codec_b200 import default_context

    return default_context(0)
