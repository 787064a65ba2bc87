"""The C ABI used from plain C (tests/cpp/abi_smoke.c, built by __graft_entry__.build()).  On a box
without a GPU the program must stop at glc_ctx_create with GLC_ERR_NO_DEVICE (there is no CPU
fallback); on the B200 it runs the reference's test_simple / test_flac checks through the ABI."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "abi_smoke")


def _build():
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cpp")], check=True, env=env)


def test_c_client_links_and_refuses_to_run_without_a_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    if r.returncode == 0:
        pytest.skip("a CUDA device is present: covered by the gpu test")
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_client_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert r.stdout.startswith("OK frames=86")


SUITE = os.path.join(HERE, "cpp", "reference_suite")


def test_cpp_mirror_suite_links_and_refuses_to_run_without_a_gpu():
    """include/glc.hpp (the C++ mirror of the reference's codec/flac API) compiles, links against the C
    ABI and, with no device, stops at the context with GLC_ERR_NO_DEVICE."""
    _build()
    r = subprocess.run([SUITE], capture_output=True, text=True, timeout=120)
    if r.returncode == 0:
        pytest.skip("a CUDA device is present: covered by the gpu test")
    assert r.returncode == 77, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_reference_test_suite_over_cpp_mirror_on_gpu():
    """Every test of the reference's tests/*.rs, restated over include/glc.hpp (tests/cpp/reference_suite.cpp),
    plus the mirror's output compared bit for bit with the CPU oracle."""
    _build()
    r = subprocess.run([SUITE], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-4000:], r.stderr[-2000:])
    assert "test result: ok." in r.stdout
    assert r.stdout.count("... ok") >= 50
