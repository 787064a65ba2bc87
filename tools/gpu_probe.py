"""First-contact GPU probe (run under gpurun): FP32 issue micro-benchmark, parity of the three
GEMM variants against the oracle on small inputs, and kernel timings on a larger batch."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle
import signals
from parity import assert_encoded_equal, assert_pcm_bits_equal, to_product
from gapless_lossy_codec_b200 import Decoder, Encoder, default_context

ctx = default_context(0)
L = ctx._lib
out = {}
for packed in (0, 1):
    t = C.c_double()
    rc = L.glc_measure_fp32_issue(ctx.handle, packed, C.byref(t))
    out[f"fp32_issue_tera_ops_packed{packed}"] = t.value
    print("fp32 issue", packed, rc, t.value, flush=True)

cases = {
    "sine_mono": (signals.sine(440, 44100, 1, 1.0), 1, 44100),
    "music_stereo": (signals.music_like(44100, 2, 2.0), 2, 44100),
    "noise_stereo": (signals.white_noise(44100, 2, 0.5), 2, 44100),
    "sweep_6ch_48k": (signals.sweep(100, 8000, 48000, 6, 0.7), 6, 48000),
}
for variant in (0, 1, 2):
    ctx.set_tuning(variant, 0)
    for name, (x, ch, sr) in cases.items():
        ref = oracle.encode(x, ch, sr)
        enc = Encoder(sr).encode(x, ch)
        try:
            assert_encoded_equal(enc, ref, f"v{variant} {name}")
            pcm_ref = oracle.decode(ref)
            pcm = Decoder(ch, sr).decode(to_product(ref))
            assert_pcm_bits_equal(pcm, pcm_ref, f"v{variant} {name} decode")
            print(f"variant {variant} {name}: PARITY OK frames={ref.n_frames} raw={int(ref.frame_is_raw.sum())} nnz={int(ref.nnz.sum())}", flush=True)
            out[f"parity_v{variant}_{name}"] = True
        except AssertionError as e:
            print(f"variant {variant} {name}: PARITY FAIL {e}", flush=True)
            out[f"parity_v{variant}_{name}"] = False

# timing: 10 minutes of stereo music-like content, device resident
dur = float(os.environ.get("PROBE_SECONDS", "600"))
x = np.tile(signals.music_like(44100, 2, 20.0), int(dur / 20))
print("timing input samples", x.size, flush=True)
enc_h = C.c_void_p(); L.glc_encoder_new(ctx.handle, 44100, C.byref(enc_h))
dec_h = C.c_void_p(); L.glc_decoder_new(ctx.handle, 2, 44100, C.byref(dec_h))
dp = C.c_void_p()
assert L.glc_dev_upload(ctx.handle, x.ctypes.data, x.size, 2, C.byref(dp)) == 0
for variant in (0, 1, 2):
    ctx.set_tuning(variant, 0)
    ctx.enable_kernel_timing(True)
    for rep in range(3):
        ctx.stats_reset()
        de = C.c_void_p()
        L.glc_timer_begin(ctx.handle)
        rc = L.glc_dev_encode(enc_h, dp, C.byref(de))
        ms = C.c_float(); L.glc_timer_end(ctx.handle, C.byref(ms))
        assert rc == 0, L.glc_last_error()
        L.glc_timer_begin(ctx.handle)
        dq = C.c_void_p()
        rc = L.glc_dev_decode(dec_h, de, C.byref(dq))
        ms2 = C.c_float(); L.glc_timer_end(ctx.handle, C.byref(ms2))
        assert rc == 0, L.glc_last_error()
        st = ctx.stats()
        L.glc_dev_pcm_free(dq); L.glc_dev_encoded_free(de)
    secs = x.size / 2 / 44100
    print(f"variant {variant}: encode {ms.value:.2f} ms ({secs/ms.value*1e3:.0f} audio-s/s) decode {ms2.value:.2f} ms "
          f"({secs/ms2.value*1e3:.0f} audio-s/s) kernels {json.dumps({k: round(v,3) for k,v in st['kernel_ms'].items() if v})}", flush=True)
    rows = (x.size // 2 + 1023) // 1024 * 2
    out[f"time_v{variant}"] = dict(enc_ms=ms.value, dec_ms=ms2.value, kernel_ms=st["kernel_ms"], rows_approx=rows)
ctx.enable_kernel_timing(False)
# e2e through the host ABI with pinned input
xp = ctx.pinned_array(x.size); xp[:] = x
for rep in range(3):
    t0 = time.perf_counter()
    o = C.POINTER(__import__("gapless_lossy_codec_b200")._ffi.Encoded)()
    rc = L.glc_encode(enc_h, xp.ctypes.data, xp.size, 2, C.byref(o)); assert rc == 0, L.glc_last_error()
    t1 = time.perf_counter()
    p = C.POINTER(C.c_float)(); n = C.c_uint64()
    rc = L.glc_decode(dec_h, o, C.byref(p), C.byref(n)); assert rc == 0, L.glc_last_error()
    t2 = time.perf_counter()
    L.glc_free(ctx.handle, p); L.glc_encoded_free(ctx.handle, o)
    print(f"e2e rep{rep}: encode {1e3*(t1-t0):.1f} ms decode {1e3*(t2-t1):.1f} ms n={n.value}", flush=True)
out["e2e_ms"] = [1e3 * (t1 - t0), 1e3 * (t2 - t1)]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
print("PROBE DONE")
