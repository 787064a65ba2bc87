"""Latency of single small calls through the C ABI (BASELINE config 1: 2 s of 44.1 kHz audio)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import signals
from gapless_lossy_codec_b200 import _ffi
from gapless_lossy_codec_b200.codec import Context
ctx = Context(0); L = ctx._lib
for ch in (1, 2):
    x = signals.sine(440, 44100, ch, 2.0)
    xp = ctx.pinned_array(x.size); xp[:] = x
    enc_h, dec_h = C.c_void_p(), C.c_void_p()
    _ffi.check(L.glc_encoder_new(ctx.handle, 44100, C.byref(enc_h)))
    _ffi.check(L.glc_decoder_new(ctx.handle, ch, 44100, C.byref(dec_h)))
    te, td, tf = [], [], []
    for it in range(30):
        out = C.POINTER(_ffi.Encoded)()
        t0 = time.perf_counter()
        _ffi.check(L.glc_encode(enc_h, xp.ctypes.data, xp.size, ch, C.byref(out)))
        t1 = time.perf_counter()
        p, n = C.POINTER(C.c_float)(), C.c_uint64()
        _ffi.check(L.glc_decode(dec_h, out, C.byref(p), C.byref(n)))
        t2 = time.perf_counter()
        b, bl = C.POINTER(C.c_uint8)(), C.c_uint64()
        _ffi.check(L.glc_flac_encode(ctx.handle, xp.ctypes.data, xp.size, 44100, ch, 5, C.byref(b), C.byref(bl)))
        t3 = time.perf_counter()
        L.glc_free(ctx.handle, p); L.glc_free(ctx.handle, b); L.glc_encoded_free(ctx.handle, out)
        te.append(t1 - t0); td.append(t2 - t1); tf.append(t3 - t2)
    print(f"ch={ch}: 2 s file: encode {1e3*np.median(te[5:]):.3f} ms, decode {1e3*np.median(td[5:]):.3f} ms, flac {1e3*np.median(tf[5:]):.3f} ms "
          f"(round trip {2.0/(np.median(te[5:])+np.median(td[5:])):.0f} audio-s/s)", flush=True)

# where the FLAC call spends its time: kernel times (CUDA events) vs the whole call
for ch in (1, 2):
    x = signals.sine(440, 44100, ch, 2.0)
    xp = ctx.pinned_array(x.size); xp[:] = x
    ctx.enable_kernel_timing(True)
    for it in range(3):
        ctx.stats_reset()
        b, bl = C.POINTER(C.c_uint8)(), C.c_uint64()
        t0 = time.perf_counter()
        _ffi.check(L.glc_flac_encode(ctx.handle, xp.ctypes.data, xp.size, 44100, ch, 5, C.byref(b), C.byref(bl)))
        t1 = time.perf_counter()
        st = ctx.stats()
        L.glc_free(ctx.handle, b)
    ctx.enable_kernel_timing(False)
    print(f"flac ch={ch}: call {1e3*(t1-t0):.3f} ms; kernels {({k: round(v, 3) for k, v in st['kernel_ms'].items() if v})}", flush=True)
