import ctypes as C, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np
import bench
from gapless_lossy_codec_b200 import _ffi
from gapless_lossy_codec_b200.codec import Context
ctx = Context(0); L = ctx._lib
x = bench.synth(3600.0)
xp = ctx.pinned_array(x.size); xp[:] = x
enc_h, dec_h = C.c_void_p(), C.c_void_p()
_ffi.check(L.glc_encoder_new(ctx.handle, 44100, C.byref(enc_h)))
_ffi.check(L.glc_decoder_new(ctx.handle, 2, 44100, C.byref(dec_h)))
for it in range(4):
    out = C.POINTER(_ffi.Encoded)()
    t0=time.perf_counter()
    _ffi.check(L.glc_encode(enc_h, xp.ctypes.data, xp.size, 2, C.byref(out)))
    t1=time.perf_counter()
    p, n = C.POINTER(C.c_float)(), C.c_uint64()
    _ffi.check(L.glc_decode(dec_h, out, C.byref(p), C.byref(n)))
    t2=time.perf_counter()
    L.glc_free(ctx.handle, p); L.glc_encoded_free(ctx.handle, out)
    t3=time.perf_counter()
    print(f"iter {it}: encode {1e3*(t1-t0):.2f} ms decode {1e3*(t2-t1):.2f} ms free {1e3*(t3-t2):.2f} ms", flush=True)
