"""The comparer of tests/test_reference_golden.py must work the day a reference dump arrives: here a
directory with the dump's exact layout (tests/golden/dump_reference.rs) is produced by the ORACLE and the
comparer is run over it in a subprocess.  This checks the kit's plumbing (file formats, manifest, chunk
shapes, table fingerprint), not parity."""
import os
import subprocess
import sys

import numpy as np

import oracle
import signals

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_layout(d):
    man = []
    cases = [("sine_mono", signals.sine(440, 44100, 1, 0.3), 44100, 1),
             ("long_stream", signals.sine(440, 8000, 2, 66.0), 8000, 2)]  # 516 frames: two chunks
    for name, x, sr, ch in cases:
        x.astype("<f4").tofile(os.path.join(d, name + ".in.f32"))
        e = oracle.encode(x, ch, sr)
        with open(os.path.join(d, name + ".glc"), "wb") as f:
            f.write(oracle.bincode_serialize(e))
        oracle.decode(e).astype("<f4").tofile(os.path.join(d, name + ".pcm.f32"))
        oracle.decode(e, trimmed=False).astype("<f4").tofile(os.path.join(d, name + ".stream.f32"))
        per, n = 1024 * ch, e.n_frames
        with open(os.path.join(d, name + ".chunks.txt"), "w") as f:
            for _ in range(n // 500):
                f.write(f"{500 * per} 0\n")
            f.write(f"{(n % 500 + 1) * per} 1\n")
        man.append(f"codec {name} {sr} {ch}")
    x = signals.music_like(44100, 2, 0.2)
    x.astype("<f4").tofile(os.path.join(d, "fl.in.f32"))
    for lv in (0, 5, 8):
        with open(os.path.join(d, f"fl.l{lv}.flac"), "wb") as f:
            f.write(oracle.flac_encode(x, 44100, 2, lv))
    man.append("flac fl 44100 2 0 5 8")
    cos_tab, window, _ = oracle.tables()
    with open(os.path.join(d, "table.fnv"), "w") as f:
        f.write(f"{oracle.fnv1a64(cos_tab.view(np.uint8).reshape(-1)):016x} {oracle.fnv1a64(window.view(np.uint8).reshape(-1)):016x}\n")
    with open(os.path.join(d, "manifest.txt"), "w") as f:
        f.write("\n".join(man) + "\n")


def test_comparer_runs_over_a_dump_shaped_directory(tmp_path):
    _write_layout(str(tmp_path))
    env = dict(os.environ, GLC_REF_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_reference_golden.py"), "-q",
                        "-m", "not gpu", "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "6 passed" in r.stdout, r.stdout  # status + 2 codec + 3 flac


def test_fnv_known_answer():
    assert oracle.fnv1a64(b"") == 0xcbf29ce484222325 and oracle.fnv1a64(b"a") == 0xaf63dc4c8601ec8c
