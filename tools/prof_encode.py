"""Small fixed workload for ncu: device-resident encode + decode of PROF_SECONDS of stereo audio."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import signals
from gapless_lossy_codec_b200 import default_context
from gapless_lossy_codec_b200.codec import Context

secs = float(os.environ.get("PROF_SECONDS", "120"))
reps = int(os.environ.get("PROF_REPS", "2"))
ctx = Context(0, 1) if os.environ.get("GLC_PROF_MODE") == "1" else default_context(0)  # 1 = FAST mode
L = ctx._lib
ctx.set_tuning(int(os.environ.get("GLC_GEMM_VARIANT", "0")), 0)
x = np.tile(signals.music_like(44100, 2, 10.0), max(1, int(secs / 10)))
enc_h = C.c_void_p(); assert L.glc_encoder_new(ctx.handle, 44100, C.byref(enc_h)) == 0
dec_h = C.c_void_p(); assert L.glc_decoder_new(ctx.handle, 2, 44100, C.byref(dec_h)) == 0
dp = C.c_void_p(); assert L.glc_dev_upload(ctx.handle, x.ctypes.data, x.size, 2, C.byref(dp)) == 0
for r in range(reps):
    de = C.c_void_p(); assert L.glc_dev_encode(enc_h, dp, C.byref(de)) == 0, L.glc_last_error()
    dq = C.c_void_p(); assert L.glc_dev_decode(dec_h, de, C.byref(dq)) == 0, L.glc_last_error()
    ctx.sync()
    L.glc_dev_pcm_free(dq); L.glc_dev_encoded_free(de)
print("prof workload done", x.size)
