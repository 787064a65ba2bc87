"""Pageable caller memory: staging ring versus cudaHostRegister ("measure both").

    python tools/pageable_probe.py [--seconds 3600]

Times glc_encode (EXACT, 44.1 kHz stereo) of the same PCM held in (a) library-pinned memory, (b) ordinary
pageable memory (the library stages it through its pinned ring), (c) the same pageable buffer after
cudaHostRegister (the library then sees pinned memory and copies directly) -- and what the registration and
the un-registration themselves cost.  One JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import signals  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=int, default=3600)
    args = ap.parse_args()
    import torch

    from gapless_lossy_codec_b200 import _ffi
    from gapless_lossy_codec_b200.codec import Context

    sr, ch = 44100, 2
    base = signals.music_like(sr, ch, 60.0, seed=3)
    x = np.tile(base, args.seconds // 60)  # ordinary pageable memory, touched
    ctx = Context(0)
    L = ctx._lib
    enc_h = C.c_void_p()
    _ffi.check(L.glc_encoder_new(ctx.handle, sr, C.byref(enc_h)))
    xp = ctx.pinned_array(x.size, np.float32)
    xp[:] = x

    def encode_ms(arr, reps=5):
        best = 1e30
        for _ in range(reps + 1):  # the first call warms the pools
            out = C.POINTER(_ffi.Encoded)()
            t0 = time.perf_counter()
            _ffi.check(L.glc_encode(enc_h, arr.ctypes.data, arr.size, ch, C.byref(out)))
            t = time.perf_counter() - t0
            L.glc_encoded_free(ctx.handle, out)
            best = min(best, t)
        return round(best * 1e3, 2)

    res = {"what": "glc_encode of %d s of 44.1 kHz stereo f32 (%.2f GB), best of 5" % (args.seconds, x.nbytes / 1e9)}
    res["library_pinned_ms"] = encode_ms(xp)
    res["pageable_staged_ms"] = encode_ms(x)
    st = ctx.stats()
    res["staged_bytes_last_calls"] = st["staged_bytes"]
    rt = torch.cuda.cudart()
    reg, unreg = [], []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(x.ctypes.data, x.nbytes, 0)
        t1 = time.perf_counter()
        assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
        res["registered_ms"] = encode_ms(x, reps=3)
        t2 = time.perf_counter()
        rc = rt.cudaHostUnregister(x.ctypes.data)
        t3 = time.perf_counter()
        assert int(rc) == 0
        reg.append(round((t1 - t0) * 1e3, 1))
        unreg.append(round((t3 - t2) * 1e3, 1))
    res["cudaHostRegister_ms"] = reg
    res["cudaHostUnregister_ms"] = unreg
    res["register_plus_encode_plus_unregister_ms"] = round(min(reg) + res["registered_ms"] + min(unreg), 1)
    print(json.dumps(res))
    L.glc_encoder_free(enc_h)
    ctx.close()


if __name__ == "__main__":
    main()
