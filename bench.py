#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native gapless-lossy-codec hot path.

Metric (BASELINE.json): encode/decode audio-seconds per second.  One "step" is one pass of the hot
path over one batch: Encoder::encode followed by Decoder::decode of `--seconds` (default 3600 = the
"1 h synthetic 44.1 kHz stereo PCM on 1 B200" configuration, configs[1]) of audio PER GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU arm (C restatement of the reference)

What the single JSON line holds
  value         whole-job audio-s/s of the encode+decode round trip with the PCM ALREADY RESIDENT in
                HBM (glc_dev_encode / glc_dev_decode), timed with CUDA events on the library's compute
                stream, max over ranks.  encode_value / decode_value split it.
  e2e           the same metric through the reference-facing C ABI with HOST buffers (glc_encode /
                glc_decode: pinned host PCM in, host stream out, host PCM back), H2D and D2H inside the
                timed region.
  roofline      the dominant kernel (fused window + direct MDCT, EXACT mode).  It is FP32-issue bound by
                construction (SURVEY.md section 0, F2 / DESIGN.md section 5): `bound` says so, `peak` is the
                non-FMA FMUL+FADD issue rate measured on this GPU in the same run, and `hbm` carries
                the HBM view of the same kernel against MEASURED_PEAKS.json.
  fast_mode     side figure (EXACT runs, rank 0): the same workload in FAST mode -- the fused window + FFT-MDCT +
                quantise + pack kernel named by BASELINE.json's north_star -- with its HBM roofline.
  cpu_baseline  the oracle (C restatement of the reference; the Rust crate cannot be built in this
                image) timed on this box's host cores on a bounded prefix of the same workload.
Multi-GPU: the path shards by file / frame range with no collective: every rank
processes its own `--seconds` of audio ("weak" scaling); torch.distributed is only the barrier and
the max-over-ranks reduction of the timings.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SR = 44100
CH = 2
METRIC = "encode+decode audio-seconds per second (Encoder::encode then Decoder::decode, 44.1 kHz stereo)"
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="audio seconds per GPU per step")
    ap.add_argument("--ref-seconds", type=float, default=30.0,
                    help="audio seconds per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"],
                    help="transform mode: exact = bit-exact with the reference (default, parity-gated); "
                         "fast = FFT-based true MDCT (tolerance class, HBM-roofline showcase)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flac", action="store_true")
    ap.add_argument("--no-fast-side", action="store_true", help="skip the FAST-mode side figure of the EXACT line")
    ap.add_argument("--flac-seconds", type=float, default=600.0)
    return ap.parse_args()


def synth(seconds: float, seed: int = 12345) -> np.ndarray:
    """~70 % multi-sine + low-passed LCG noise (sparse frames), ~30 % white LCG
    noise (raw-PCM frames); a 20 s deterministic period tiled to the requested length."""
    import signals

    period = signals.music_like(SR, CH, 20.0, seed=seed)
    n = int(round(seconds * SR)) * CH
    reps = (n + period.size - 1) // period.size
    return np.tile(period, reps)[:n]


# ------------------------------------------------------------------ clocks


class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.25:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smmax.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smmax)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm


def cpu_codec_pass(x: np.ndarray, threads: int):
    """One encode + decode of the oracle, threaded like the reference (rayon over frames in encode,
    over 32-frame batches in decode) with the reference's own dense IMDCT loop (literal_imdct)."""
    import oracle

    t0 = time.perf_counter()
    enc = oracle.encode(x, CH, SR, threads=threads)
    t1 = time.perf_counter()
    pcm = oracle.decode(enc, threads=threads, literal_imdct=True)
    t2 = time.perf_counter()
    assert len(pcm) == len(x)
    return t1 - t0, t2 - t1


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path.  The reference is a Rust
    crate and this image has no cargo/rustc (oracle/_ref is empty), so the arm
    is the oracle port on all host threads; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    import oracle

    oracle.build()
    cores = os.cpu_count() or 1
    x = synth(args.ref_seconds)
    for _ in range(max(args.warmup, 0)):
        cpu_codec_pass(x[: min(x.size, 10 * SR * CH)], cores)
    te = td = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b = cpu_codec_pass(x, cores)
        te += a
        td += b
    wall = time.perf_counter() - t0
    secs = x.size / CH / SR
    value = secs * args.steps / wall
    sample = (f"first {secs:.0f} s of the 44.1 kHz stereo synthetic workload per step, {args.steps} steps; "
              "C restatement of the reference (dense reference-order IMDCT), pthreads over frames")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"encode+decode of 44.1 kHz stereo synthetic PCM, {secs:.0f} s sample per step "
                               "(bounded sample of the 1 h configuration)", "mode": "EXACT (reference arithmetic)"},
        "encode_value": secs * args.steps / te, "decode_value": secs * args.steps / td,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args, rank: int, local_rank: int, world: int):
    from gapless_lossy_codec_b200 import _ffi
    from gapless_lossy_codec_b200.codec import Context

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local_rank)
        # stdout carries exactly one JSON line: NCCL prints its version banner with printf when the first
        # communicator is created, so file descriptor 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist_mod.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        dist = dist_mod

    def barrier():
        if dist:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if not dist:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if not dist:
            return v
        import torch

        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    fast = args.mode == "fast"
    ctx = Context(local_rank, mode=1 if fast else 0)  # raises GlcError when the CUDA library / device is missing: no fallback
    L = ctx._lib
    chk = _ffi.check

    # ---- workload: this rank's shard (its own files; no data-path collective) ----
    x = synth(args.seconds, seed=12345 + 977 * rank)
    secs = x.size / CH / SR
    xp = ctx.pinned_array(x.size)
    xp[:] = x
    del x
    enc_h, dec_h = C.c_void_p(), C.c_void_p()
    chk(L.glc_encoder_new(ctx.handle, SR, C.byref(enc_h)))
    chk(L.glc_decoder_new(ctx.handle, CH, SR, C.byref(dec_h)))
    dpcm = C.c_void_p()
    chk(L.glc_dev_upload(ctx.handle, xp.ctypes.data, xp.size, CH, C.byref(dpcm)))

    # the measured non-FMA FP32 issue roof of this GPU (same run, same clocks)
    tera = C.c_double()
    chk(L.glc_measure_fp32_issue(ctx.handle, 0, C.byref(tera)))
    fp32_roof_scalar = tera.value  # 1e12 lane-ops / s, scalar FMUL -> FADD chains
    chk(L.glc_measure_fp32_issue(ctx.handle, 1, C.byref(tera)))
    fp32_roof_packed = tera.value  # the same chains as packed f32x2 instructions (what the kernels issue)
    fp32_roof = max(fp32_roof_scalar, fp32_roof_packed)

    def dev_step():
        de, dq = C.c_void_p(), C.c_void_p()
        ms_e, ms_d = C.c_float(), C.c_float()
        chk(L.glc_timer_begin(ctx.handle))
        chk(L.glc_dev_encode(enc_h, dpcm, C.byref(de)))
        chk(L.glc_timer_end(ctx.handle, C.byref(ms_e)))
        chk(L.glc_timer_begin(ctx.handle))
        chk(L.glc_dev_decode(dec_h, de, C.byref(dq)))
        chk(L.glc_timer_end(ctx.handle, C.byref(ms_d)))
        L.glc_dev_pcm_free(dq)
        L.glc_dev_encoded_free(de)
        return ms_e.value, ms_d.value

    def host_step(src=None, probe=False):
        """Encoder::encode then Decoder::decode through the C ABI with host buffers.  src: the PCM (default: the
        library-pinned copy).  probe=True additionally records per-call DMA bytes and workload statistics."""
        src = xp if src is None else src
        out = C.POINTER(_ffi.Encoded)()
        if probe:
            ctx.stats_reset()
        chk(L.glc_encode(enc_h, src.ctypes.data, src.size, CH, C.byref(out)))
        if probe:
            s_e = ctx.stats()
            ctx.stats_reset()
        p, n = C.POINTER(C.c_float)(), C.c_uint64()
        chk(L.glc_decode(dec_h, out, C.byref(p), C.byref(n)))
        got = n.value
        first = float(p[0]) if got else 0.0  # the step's result is read on the host
        e = out.contents
        host_step.pairs = int(e.pair_offset[int(e.n_frames) * int(e.channels)]) if e.n_frames else 0
        host_step.raw = int(e.raw_offset[int(e.n_frames)]) if e.n_frames else 0
        if probe:
            s_d = ctx.stats()
            host_step.bytes = {"enc_h2d": s_e["h2d_bytes"], "enc_d2h": s_e["d2h_bytes"],
                               "dec_h2d": s_d["h2d_bytes"], "dec_d2h": s_d["d2h_bytes"]}
            # measured workload statistics: raw-frame share, kept coefficients per sparse
            # frame-channel, and the size of the index-set union of the two channels of a frame (what a 2-row
            # warp of the IMDCT kernel executes) over a sample of frames
            nf, ch = int(e.n_frames), int(e.channels)
            raw = np.ctypeslib.as_array(e.frame_is_raw, (nf,))
            n_raw = int(raw.sum())
            nnz = np.ctypeslib.as_array(e.nnz, (nf * ch,))
            po = np.ctypeslib.as_array(e.pair_offset, (nf * ch + 1,))
            sparse_rows = max(1, (nf - n_raw) * ch)
            unions = []
            idx = np.ctypeslib.as_array(C.cast(e.pairs, C.POINTER(C.c_uint16)), (2 * max(host_step.pairs, 1),))[0::2]
            for f in np.flatnonzero(raw == 0)[:: max(1, (nf - n_raw) // 2000)][:2000]:
                a, b, c2 = int(po[f * ch]), int(po[f * ch + 1]), int(po[f * ch + 2]) if ch > 1 else int(po[f * ch + 1])
                unions.append(len(np.union1d(idx[a:b], idx[b:c2])))
            host_step.workload = {"raw_frame_fraction": n_raw / max(nf, 1), "mean_nnz_per_sparse_row": float(nnz.sum()) / sparse_rows,
                                  "mean_union_of_2_rows": float(np.mean(unions)) if unions else None,
                                  "union_sample_frames": len(unions)}
        L.glc_free(ctx.handle, p)
        L.glc_encoded_free(ctx.handle, out)
        return got, first

    warm = max(args.warmup, 3)  # timing rule: W >= 3
    for _ in range(warm):
        dev_step()
    ctx.sync()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if sampler else 0.0)

    # ---- timed region 1: device-resident (value) ----
    ctx.enable_kernel_timing(True)
    ctx.stats_reset()
    barrier()
    ctx.sync()
    t_region0 = time.perf_counter()
    enc_ms = dec_ms = 0.0
    for _ in range(args.steps):
        a, b = dev_step()
        enc_ms += a
        dec_ms += b
    ctx.sync()
    barrier()
    wall_dev = time.perf_counter() - t_region0
    st = ctx.stats()
    ctx.enable_kernel_timing(False)
    step_ms = max_over_ranks((enc_ms + dec_ms) / args.steps)
    enc_ms_max = max_over_ranks(enc_ms / args.steps)
    dec_ms_max = max_over_ranks(dec_ms / args.steps)
    total_secs = sum_over_ranks(secs)
    launches = sum(st["launches"].values())

    # ---- timed region 2: end to end through the C ABI with host buffers (e2e) ----
    for _ in range(max(1, min(args.warmup, 3))):
        host_step()
    ctx.stats_reset()
    barrier()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got, _first = host_step()
        assert got == xp.size, f"gapless length {got} != {xp.size}"
    ctx.sync()
    e2e_s = time.perf_counter() - t0
    barrier()
    t_region1 = time.perf_counter()
    st2 = ctx.stats()
    e2e_ms = max_over_ranks(e2e_s / args.steps * 1e3)
    launches += sum(st2["launches"].values())

    # ---- timed region 2b: the same round trip with the PCM in ORDINARY (pageable) memory, what a caller of
    #      Encoder::encode(&[f32]) hands over (src/codec.rs:421); the library stages it through pinned chunks ----
    x_pageable = np.array(xp, copy=True)  # numpy heap memory: not pinned
    host_step(x_pageable, probe=True)     # warm-up + per-call DMA bytes + workload statistics
    host_step(x_pageable)
    barrier()
    ctx.sync()
    t0 = time.perf_counter()
    n_pg = max(2, min(args.steps, 5))
    for _ in range(n_pg):
        got, _first = host_step(x_pageable)
        assert got == xp.size
    ctx.sync()
    e2e_pg_ms = max_over_ranks((time.perf_counter() - t0) / n_pg * 1e3)
    barrier()
    del x_pageable

    # ---- pure-DMA floor: the bytes of one encode call and of one decode call, up and down at once, pinned host
    #      memory <-> HBM, no kernel, every rank at the same time ----
    def dma(h2d, d2h, concurrent):
        ms = C.c_float()
        best = 1e30
        for _ in range(3):
            barrier()
            chk(L.glc_dma_probe(ctx.handle, int(h2d), int(d2h), int(concurrent), C.byref(ms)))
            best = min(best, max_over_ranks(ms.value))
        return best
    hb = host_step.bytes
    dma_enc = dma(hb["enc_h2d"], hb["enc_d2h"], 1)
    dma_dec = dma(hb["dec_h2d"], hb["dec_d2h"], 1)
    dma_h2d_only = dma(hb["enc_h2d"], 0, 0)
    dma_d2h_only = dma(0, hb["dec_d2h"], 0)
    t_region1 = time.perf_counter()

    # ---- timed region 3: the same round trip with 16-bit integer ingest (glc_encode_i16: the WAV
    #      loader's division runs on the device, half the H2D bytes) ----
    x16 = ctx.pinned_array(xp.size, np.int16)
    # half scale: the synthetic mix peaks above 1.0, which f32 carries but 16-bit PCM would clip
    np.multiply(xp, 16384.0, out=xp)  # in place: the f32 copy is not needed any more
    np.rint(xp, out=xp)
    x16[:] = xp.astype(np.int16)

    def host_step_i16():
        out = C.POINTER(_ffi.Encoded)()
        chk(L.glc_encode_i16(enc_h, x16.ctypes.data, x16.size, CH, C.byref(out)))
        p, n = C.POINTER(C.c_float)(), C.c_uint64()
        chk(L.glc_decode(dec_h, out, C.byref(p), C.byref(n)))
        got = n.value
        L.glc_free(ctx.handle, p)
        L.glc_encoded_free(ctx.handle, out)
        return got

    for _ in range(2):
        host_step_i16()
    ctx.stats_reset()
    barrier()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        assert host_step_i16() == x16.size
    ctx.sync()
    e2e16_s = time.perf_counter() - t0
    barrier()
    t_region1 = time.perf_counter()
    st3 = ctx.stats()
    e2e16_ms = max_over_ranks(e2e16_s / args.steps * 1e3)
    launches += sum(st3["launches"].values())

    # ---- timed region 4: 16-bit PCM in AND out (glc_encode_i16 + glc_decode_i16: what `glc` does between
    #      a 16-bit WAV and the WAV it writes back, src/audio.rs:11-16, 51-59; both conversions on the device) ----
    def host_step_i16_io():
        out = C.POINTER(_ffi.Encoded)()
        chk(L.glc_encode_i16(enc_h, x16.ctypes.data, x16.size, CH, C.byref(out)))
        p, n = C.POINTER(C.c_int16)(), C.c_uint64()
        chk(L.glc_decode_i16(dec_h, out, C.byref(p), C.byref(n)))
        got = n.value
        L.glc_free(ctx.handle, p)
        L.glc_encoded_free(ctx.handle, out)
        return got

    for _ in range(2):
        host_step_i16_io()
    ctx.stats_reset()
    barrier()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        assert host_step_i16_io() == x16.size
    ctx.sync()
    e2e16io_s = time.perf_counter() - t0
    barrier()
    t_region1 = time.perf_counter()
    st4 = ctx.stats()
    e2e16io_ms = max_over_ranks(e2e16io_s / args.steps * 1e3)
    launches += sum(st4["launches"].values())

    clocks = sampler.stop(t_region0, t_region1) if sampler else None

    # ---- roofline of the dominant kernel (rank 0's own launches) ----
    L_per_ch = xp.size // CH
    padded = 512 + L_per_ch
    padded += (-padded) % 1024
    padded += 512
    n_frames = (padded - 2048) // 1024 + 1
    rows = n_frames * CH
    k_ms = st["kernel_ms"]
    k_n = st["launches"]
    mdct_ms_per_step = k_ms["mdct_exact"] / args.steps
    flops = rows * 2.0 * 1024 * 2048  # non-FMA FP32 lane operations (FMUL + FADD), SURVEY.md 8(d)
    achieved_tops = flops / (mdct_ms_per_step * 1e-3) / 1e12 if mdct_ms_per_step else 0.0
    hbm_peak, peak_src = load_peaks()
    # algorithmic bytes of the MDCT stage: every new PCM sample once (4096 B per row) + the dense
    # coefficient row it hands to quantize/pack (4096 B per row)
    mdct_bytes = rows * (4096.0 + 4096.0)
    hbm_gbs = mdct_bytes / (mdct_ms_per_step * 1e-3) / 1e9 if mdct_ms_per_step else 0.0
    if fast:
        # FAST mode: the fused window + FFT-MDCT + quantise + pack kernel is HBM-bound by design.
        # Algorithmic bytes per launch: every PCM sample once (4096 B per frame-channel)
        # + the pairs it emits (4 B each) + scale and count (8 B per frame-channel).
        n_l = max(k_n["fast_encode"], 1) / args.steps
        fe_ms = k_ms["fast_encode"] / args.steps  # per step (a step is a few launches)
        fe_bytes = rows * 4096.0 + host_step.pairs * 4.0 + rows * 8.0
        fe_gbs = fe_bytes / (fe_ms * 1e-3) / 1e9
        roofline = {
            "kernel": "fast_encode_kernel (fused window + fold + 512-point FFT DCT-IV + thresholds + quantise + ordered pack)",
            "bound": "hbm", "achieved": fe_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fe_gbs / hbm_peak,
            "peak_source": peak_src,
            # profiles/r2_ncu_fast_kernels_full_summary.csv: dram read 211.9 MB + write 80.4 MB for a launch of 51 680
            # rows = 5 655 B per row against 4 792 B algorithmic (1.18x), scaled to the rows of one launch of this run
            "traffic": 5655.0 * rows / max(k_n["fast_encode"] / args.steps, 1),
            "traffic_source": "ncu capture of a 51 680-row launch (5 655 B per row), scaled to this run's rows per launch",
            "launches_per_step": k_n["fast_encode"] / args.steps,
            "ms_per_launch": fe_ms / n_l, "algorithmic_bytes_per_launch": fe_bytes / n_l,
            "kernel_ms_per_step": {k: v / args.steps for k, v in k_ms.items() if v},
        }
    else:
      roofline = {
        "kernel": "exact_gemm_kernel<MDCT> (direct-form MDCT contraction, EXACT mode; operands by TMA bulk copy, packed f32x2 multiply-then-add)",
        "bound": "fp32_issue", "achieved": achieved_tops, "peak": fp32_roof, "unit": "TFLOP/s",
        "frac": achieved_tops / fp32_roof if fp32_roof else None,
        # nominal: 148 SMs x 128 FP32 lanes x the maximum SM clock, one non-FMA operation per lane and cycle
        "peak_nominal": 148 * 128 * (clocks["sm_max_mhz"] if clocks and clocks.get("sm_max_mhz") else 1965.0) * 1e6 / 1e12,
        "frac_of_nominal": achieved_tops / (148 * 128 * (clocks["sm_max_mhz"] if clocks and clocks.get("sm_max_mhz") else 1965.0) * 1e6 / 1e12),
        "peak_scalar": fp32_roof_scalar, "peak_packed": fp32_roof_packed,
        "peak_source": "separate multiply-then-add issue micro-benchmarks on this GPU in this run (glc_measure_fp32_issue, "
                       "scalar FMUL+FADD and packed f32x2; the higher of the two); non-contracted FP32 lane-ops: tensor "
                       "cores / fused multiply-add / reordering break bit-exact parity",
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the ncu --set full capture
        # profiles/r1_ncu_exact_gemm_tma_full_summary.csv: 620 965 888 B for a launch of 51 712 rows
        # (12 008 B/row = the 8 KiB A tile it reads + the 4 KiB coefficient row it writes; the table stays
        # in L2), scaled to the rows of one launch of this run
        "traffic": 12008.0 * rows / max(k_n["mdct_exact"] / args.steps, 1),
        "traffic_source": "ncu capture of a 51 712-row launch (12 008 B per row), scaled to this run's rows per launch",
        # DRAM bytes per frame-channel of every kernel of the unfused chains (ncu --set full, one encode + decode of
        # 600 s, profiles/r2_ncu_codec_kernels_full_summary.csv) next to the algorithmic figures of SURVEY.md 8(d):
        # the encode chain moves 5.2x and the decode chain 3.8x its algorithmic bytes (windowed A tile, dense coefficient row and IMDCT blocks round-trip
        # HBM), which costs about 3 ms of the step; the step itself is bound by the FP32 issue rate of the two
        # contractions, not by these bytes
        "chain_traffic_bytes_per_frame_channel": {
            "encode": {"window_tile": 11302, "mdct_exact": 12071, "quant_pack": 5580, "scans": 6, "gather_pairs": 760,
                       "gather_raw": 2100, "total": 31819, "algorithmic": 4096 + 4 * 258 * 0.67 + 8 + 0.33 * 4096},
            "decode": {"dequant": 5814 + 4, "imdct_exact": 7339, "ola": 10376, "total": 23533,
                       "algorithmic": 4 * 258 * 0.67 + 8 + 0.33 * 4096 + 4096},
            "source": "ncu, profiles/r2_ncu_codec_kernels_full_summary.csv (51 680 rows, raw fraction 0.33, 258 pairs per sparse row)"},
        "launches_per_step": k_n["mdct_exact"] / args.steps,
        "ms_per_launch": k_ms["mdct_exact"] / max(k_n["mdct_exact"], 1),
        "hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                "peak_source": peak_src, "algorithmic_bytes_per_frame_channel": 8192},
        "kernel_ms_per_step": {k: v / args.steps for k, v in k_ms.items() if v},
      }

    line = {
        "metric": METRIC, "value": total_secs / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"batched encode+decode of {secs:.0f} s synthetic 44.1 kHz stereo PCM per GPU "
                               f"({n_frames} frames, {rows} frame-channels)",
                   "measured": host_step.workload,
                   "mode": ("FAST (FFT-based true MDCT, tolerance class -- NOT bit-exact with the reference)" if fast
                            else "EXACT (bit-exact with the reference arithmetic)"), "sharding": f"by file, {world} rank(s), no collective",
                   "l2": f"inputs larger than L2 ({xp.size * 4 / 1e6:.0f} MB PCM per step vs 126 MB)"},
        "encode_value": total_secs / (enc_ms_max * 1e-3), "decode_value": total_secs / (dec_ms_max * 1e-3),
        "e2e": {"value": total_secs / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": st2["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st2["d2h_bytes"] // args.steps,
                "ms_per_step": e2e_ms, "api": "glc_encode + glc_decode (C ABI, pinned host buffers)",
                # the same bytes as pure DMA (no kernels), every rank at once: encode call (PCM up || stream down)
                # + decode call (stream up || PCM down).  e2e cannot be faster than this on this box.
                "dma_floor_ms": dma_enc + dma_dec, "dma_floor_encode_ms": dma_enc, "dma_floor_decode_ms": dma_dec,
                "dma_floor_frac": (dma_enc + dma_dec) / e2e_ms,
                # each call can hide its kernels behind its transfers or the other way round, never both: the
                # floor of two separate calls is max(device time, DMA time) per call
                "floor_ms": max(enc_ms_max, dma_enc) + max(dec_ms_max, dma_dec),
                "floor_frac": (max(enc_ms_max, dma_enc) + max(dec_ms_max, dma_dec)) / e2e_ms,
                "dma_h2d_gbs_per_gpu": hb["enc_h2d"] / (dma_h2d_only * 1e-3) / 1e9,
                "dma_d2h_gbs_per_gpu": hb["dec_d2h"] / (dma_d2h_only * 1e-3) / 1e9,
                "bytes_per_call": hb},
        "e2e_pageable": {"value": total_secs / (e2e_pg_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_pg_ms,
                         "vs_pinned": e2e_ms / e2e_pg_ms,
                         "api": "glc_encode + glc_decode with the PCM in ordinary pageable memory (staged through a "
                                "ring of pinned chunks by host threads)"},
        "e2e_int16_ingest": {"value": total_secs / (e2e16_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e16_ms,
                             "h2d_bytes_per_step": st3["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st3["d2h_bytes"] // args.steps,
                             "api": "glc_encode_i16 + glc_decode (16-bit PCM in, f32 PCM out)"},
        "e2e_int16_io": {"value": total_secs / (e2e16io_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e16io_ms,
                         "h2d_bytes_per_step": st4["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st4["d2h_bytes"] // args.steps,
                         "api": "glc_encode_i16 + glc_decode_i16 (16-bit PCM in and out, as between two WAV files)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "clocks": clocks,
        "wall_s_device_region": wall_dev,
    }

    if rank == 0 and world == 1 and not fast and not args.no_fast_side:
        line["fast_mode"] = bench_fast_side(local_rank, None, args, rows, host_step.pairs, hbm_peak, peak_src)
    if rank == 0 and not args.no_flac:
        line["flac"] = bench_flac(ctx, args)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    L.glc_dev_pcm_free(dpcm)
    L.glc_encoder_free(enc_h)
    L.glc_decoder_free(dec_h)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def bench_fast_side(device: int, _unused, args, rows: int, pairs: int, hbm_peak: float, peak_src: str) -> dict:
    """Side figure of the default (EXACT) line: the same workload in FAST mode, i.e. the fused window +
    FFT-MDCT + quantise + pack kernel that BASELINE.json's north_star names, with its HBM roofline
    (device-resident, CUDA events, same synthetic input).  Tolerance class -- not part of `value`."""
    from gapless_lossy_codec_b200 import _ffi
    from gapless_lossy_codec_b200.codec import Context

    ctx = Context(device, mode=1)
    L = ctx._lib
    x = synth(args.seconds)
    enc_h, dec_h, dpcm = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _ffi.check(L.glc_encoder_new(ctx.handle, SR, C.byref(enc_h)))
    _ffi.check(L.glc_decoder_new(ctx.handle, CH, SR, C.byref(dec_h)))
    _ffi.check(L.glc_dev_upload(ctx.handle, x.ctypes.data, x.size, CH, C.byref(dpcm)))

    def step():
        de, dq = C.c_void_p(), C.c_void_p()
        ms_e, ms_d = C.c_float(), C.c_float()
        _ffi.check(L.glc_timer_begin(ctx.handle))
        _ffi.check(L.glc_dev_encode(enc_h, dpcm, C.byref(de)))
        _ffi.check(L.glc_timer_end(ctx.handle, C.byref(ms_e)))
        _ffi.check(L.glc_timer_begin(ctx.handle))
        _ffi.check(L.glc_dev_decode(dec_h, de, C.byref(dq)))
        _ffi.check(L.glc_timer_end(ctx.handle, C.byref(ms_d)))
        L.glc_dev_pcm_free(dq)
        L.glc_dev_encoded_free(de)
        return ms_e.value + ms_d.value

    for _ in range(3):
        step()
    ctx.enable_kernel_timing(True)
    ctx.stats_reset()
    steps = max(3, min(args.steps, 10))
    ms = sum(step() for _ in range(steps)) / steps
    st = ctx.stats()
    ctx.enable_kernel_timing(False)
    # achieved = the step's algorithmic bytes / the kernel's time per step (a step is cut into a few launches of
    # up to 151 552 frame-channels; per-launch figures are the step's divided by the launches of a step)
    n_l = max(st["launches"]["fast_encode"], 1) / steps
    fe_ms = st["kernel_ms"]["fast_encode"] / steps
    fe_bytes = rows * 4096.0 + pairs * 4.0 + rows * 8.0  # SURVEY.md 8(d); `pairs` from the EXACT stream (same order of magnitude)
    out = {
        "mode": "FAST (FFT-based true MDCT fused with the quantiser; tolerance class, NOT bit-exact)",
        "value": args.seconds / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
        "roofline": {"kernel": "fast_encode_kernel", "bound": "hbm", "achieved": fe_bytes / (fe_ms * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": fe_bytes / (fe_ms * 1e-3) / 1e9 / hbm_peak,
                     "peak_source": peak_src, "launches_per_step": n_l, "ms_per_launch": fe_ms / n_l,
                     "algorithmic_bytes_per_launch": fe_bytes / n_l,
                     # ncu (profiles/r2_ncu_fast_kernels_full_summary.csv): 5 655 B of DRAM traffic per row
                     "traffic": 5655.0 * rows / n_l,
                     # executed warp instructions per frame-channel (ncu, profiles/r2_ncu_fast_encode_*): the 512-point
                     # FFT alone is 880, i.e. at a perfect issue rate the FFT alone takes as long as moving the
                     # kernel's bytes at the full HBM rate: the kernel is instruction-bound by construction
                     "instruction_floor_note": "2 884 warp-instructions per frame-channel (ncu; FFT 867); at 100 % issue "
                                               "that is 0.29 of the HBM peak, 0.60 would need <= 1 400 in total"},
        "kernel_ms_per_step": {k: v / steps for k, v in st["kernel_ms"].items() if v},
    }
    L.glc_dev_pcm_free(dpcm)
    L.glc_encoder_free(enc_h)
    L.glc_decoder_free(dec_h)
    ctx.close()
    return out


def bench_flac(ctx, args) -> dict:
    """Secondary figure: flac::encode_flac_with_level(level 8) of 96 kHz stereo PCM through the C ABI
    (host buffers; BASELINE config 5 at `--flac-seconds`)."""
    import signals

    L = ctx._lib
    from gapless_lossy_codec_b200 import _ffi

    sr = 96000
    period = signals.music_like(sr, 2, 10.0, seed=7)
    # "24-bit" source as src/audio.rs:51-59 would load it: integer / 2^23
    period = (np.round(period.astype(np.float64) * 8388608.0) / 8388608.0).astype(np.float32)
    n = int(args.flac_seconds * sr) * 2
    x = np.tile(period, (n + period.size - 1) // period.size)[:n]
    xp = ctx.pinned_array(x.size)
    xp[:] = x
    del x

    def one():
        b, ln = C.POINTER(C.c_uint8)(), C.c_uint64()
        _ffi.check(L.glc_flac_encode(ctx.handle, xp.ctypes.data, xp.size, sr, 2, 8, C.byref(b), C.byref(ln)))
        L.glc_free(ctx.handle, b)
        return ln.value

    one()
    one()
    ctx.enable_kernel_timing(True)
    ctx.stats_reset()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        nbytes = one()
    dt = (time.perf_counter() - t0) / reps
    st = ctx.stats()
    ctx.enable_kernel_timing(False)
    hbm_peak, peak_src = load_peaks()
    k_ms = (st["kernel_ms"]["flac_block"] + st["kernel_ms"]["flac_gather"]) / reps
    # algorithmic bytes per sample: 4 B f32 in + the emitted bitstream
    alg = xp.size * 4.0 + nbytes

    # The same audio as a batch of one-minute files (glc_flac_encode_batch): the per-file MD5 chains run on one
    # host thread each, so the serial hash of a single long file no longer hides the device work.
    n_files = max(1, int(args.flac_seconds // 60))
    per = 60 * sr * 2
    batch = None
    if n_files >= 2:
        ptrs = (C.c_void_p * n_files)(*[xp.ctypes.data + 4 * per * i for i in range(n_files)])
        ns = (C.c_uint64 * n_files)(*([per] * n_files))
        srs = (C.c_uint32 * n_files)(*([sr] * n_files))
        chs = (C.c_uint16 * n_files)(*([2] * n_files))

        def many():
            outs = (C.POINTER(C.c_uint8) * n_files)()
            lens = (C.c_uint64 * n_files)()
            _ffi.check(L.glc_flac_encode_batch(ctx.handle, n_files, ptrs, ns, srs, chs, 8, outs, lens))
            tot = sum(lens[i] for i in range(n_files))
            for i in range(n_files):
                L.glc_free(ctx.handle, outs[i])
            return tot

        many()
        t0 = time.perf_counter()
        for _ in range(reps):
            many()
        dtb = (time.perf_counter() - t0) / reps
        batch = {"files": n_files, "seconds_per_file": 60, "e2e_audio_s_per_s": n_files * 60 / dtb, "e2e_ms": dtb * 1e3,
                 "api": "glc_flac_encode_batch (one host MD5 thread per file)"}
    return {"batch_of_one_minute_files": batch, "workload": f"FLAC level 8, {args.flac_seconds:.0f} s 96 kHz stereo, 16-bit (reference truncates 24-bit)",
            "e2e_audio_s_per_s": args.flac_seconds / dt, "e2e_ms": dt * 1e3, "bytes_out": int(nbytes),
            "kernel_ms": k_ms, "kernel_audio_s_per_s": args.flac_seconds / (k_ms * 1e-3) if k_ms else None,
            "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9 if k_ms else None, "peak": hbm_peak,
                         "unit": "GB/s", "frac": (alg / (k_ms * 1e-3) / 1e9) / hbm_peak if k_ms else None,
                         "peak_source": peak_src,
                         # ncu (profiles/r2_ncu_flac_kernels_full_summary.csv): flac_measure 111.4 MB + flac_emit
                         # 49.2 MB of DRAM traffic for 23.04 M samples = 6.97 B per sample
                         "traffic": 6.97 * xp.size, "algorithmic_bytes": alg}}


def cpu_baseline(args) -> dict:
    import oracle

    oracle.build()
    cores = os.cpu_count() or 1
    secs = min(args.ref_seconds * 2, args.seconds)
    x = synth(secs)
    cpu_codec_pass(x[: 5 * SR * CH], cores)
    te, td = cpu_codec_pass(x, cores)
    one = x[: 6 * SR * CH]  # single-core figure on a 6 s prefix (SURVEY.md 8d asks for both)
    te1, td1 = cpu_codec_pass(one, 1)
    return {"value": secs / (te + td), "unit": UNIT, "cores": cores, "kind": "port",
            "encode_value": secs / te, "decode_value": secs / td,
            "one_core": {"value": 6.0 / (te1 + td1), "encode_value": 6.0 / te1, "decode_value": 6.0 / td1,
                         "sample": "first 6 s, 1 thread"},
            "sample": f"first {secs:.0f} s of the same workload, one pass; C restatement of the reference "
                      f"(the Rust crate cannot be built here), {cores} pthreads over frames, dense reference-order IMDCT"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
