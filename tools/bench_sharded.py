"""One process, every visible GPU: end-to-end throughput of glc_encode_batch_sharded + glc_decode_batch_sharded
(host buffers in, host buffers out) over a batch of equal files, next to the same batch on one context.

    python tools/bench_sharded.py [--files 16] [--seconds 900] [--steps 3]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import bench
from gapless_lossy_codec_b200 import Decoder, Encoder, _ffi, shard
from gapless_lossy_codec_b200.codec import Context

ap = argparse.ArgumentParser()
ap.add_argument("--files", type=int, default=16)
ap.add_argument("--seconds", type=float, default=900.0)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()

n_dev = C.c_int()
_ffi.load().glc_device_count(C.byref(n_dev))
ctxs = [Context(d) for d in range(n_dev.value)]
x = bench.synth(args.seconds)
files = []
for i in range(args.files):  # pinned host copies, one per file
    p = ctxs[i % len(ctxs)].pinned_array(x.size)
    p[:] = x
    files.append(p)
chans = [bench.CH] * args.files
audio_s = args.files * args.seconds


def run(k):
    """raw C-ABI calls (no numpy copies of the outputs): encode sharded -> decode sharded -> free"""
    L = ctxs[0]._lib
    encs = [Encoder(bench.SR, c) for c in ctxs[:k]]
    decs = [Decoder(bench.CH, bench.SR, c) for c in ctxs[:k]]
    n = args.files
    ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in files])
    ns = (C.c_uint64 * n)(*[f.size for f in files])
    chs = (C.c_uint16 * n)(*chans)
    enc_h = (C.c_void_p * k)(*[e.handle.value for e in encs])
    dec_h = (C.c_void_p * k)(*[d.handle.value for d in decs])
    best = None
    for it in range(args.steps + 1):
        outs = (C.POINTER(_ffi.Encoded) * n)()
        where = (C.c_uint32 * n)()
        pcm = (C.POINTER(C.c_float) * n)()
        pn = (C.c_uint64 * n)()
        where_d = (C.c_uint32 * n)()
        t0 = time.perf_counter()
        _ffi.check(L.glc_encode_batch_sharded(enc_h, k, n, ptrs, ns, chs, outs, where))
        t1 = time.perf_counter()
        _ffi.check(L.glc_decode_batch_sharded(dec_h, k, n, outs, pcm, pn, where_d))
        t2 = time.perf_counter()
        assert all(pn[i] == x.size for i in range(n))
        for i in range(n):
            L.glc_free(ctxs[where_d[i]].handle, pcm[i])
            L.glc_encoded_free(ctxs[where[i]].handle, outs[i])
        if it and (best is None or t2 - t0 < best[0]):
            best = (t2 - t0, t1 - t0, t2 - t1, sorted(set(where)))
    return best


out = {"files": args.files, "seconds_per_file": args.seconds, "devices": n_dev.value,
       "note": "wall clock around the two C-ABI calls, pinned host buffers in, pinned host buffers out"}
for k in sorted({1, n_dev.value}):
    tot, te, td, used = run(k)
    out[f"contexts_{k}"] = {"roundtrip_audio_s_per_s": audio_s / tot, "encode_s": te, "decode_s": td, "shards_used": used}
print(json.dumps(out), flush=True)
