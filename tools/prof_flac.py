"""Small fixed FLAC workload for ncu: level-8 encode of PROF_SECONDS of 96 kHz stereo through the C ABI."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import signals
from gapless_lossy_codec_b200 import default_context, flac

secs = float(os.environ.get("PROF_SECONDS", "120"))
ctx = default_context(0)
x = np.tile(signals.music_like(96000, 2, 10.0, seed=7), max(1, int(secs / 10)))
ctx.enable_kernel_timing(True)
for r in range(int(os.environ.get("PROF_REPS", "3"))):
    ctx.stats_reset()
    out = flac.encode_flac_with_level(x, 96000, 2, 8, ctx)
    st = ctx.stats()
    print("flac bytes", len(out), {k: round(v, 3) for k, v in st["kernel_ms"].items() if v}, flush=True)
