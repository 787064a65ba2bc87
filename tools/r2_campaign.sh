#!/bin/bash
# Round-2 measurement campaign (run under gpurun, from the repo root).  Usage: bash tools/r2_campaign.sh A|B|C
set -u
O=gpurun_out
case "${1:-A}" in
A)  # one GPU: tests, bench lines, reports
    (time python -m pytest tests -m gpu -q) > $O/r2_gputest.log 2>&1; tail -3 $O/r2_gputest.log
    python bench.py > $O/r2_bench_exact_1gpu.json 2> $O/r2_bench_exact_1gpu.err; tail -1 $O/r2_bench_exact_1gpu.err
    python bench.py --mode fast --no-cpu-baseline --no-flac > $O/r2_bench_fast_1gpu.json 2> $O/r2_bench_fast_1gpu.err
    python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
    python tools/fast_report.py --seconds 4 > $O/r2_fast_flip_report.json 2> $O/r2_fast_report.err
    python tools/bench_configs.py --configs 3,4,5 > $O/r2_configs_3_4_5_1gpu.jsonl 2> $O/r2_configs_3_4_5_1gpu.err
    python tools/latency_small.py > $O/r2_latency_small.txt 2>&1
    ;;
B)  # one GPU: ncu launch list + full captures (each after its plain run exited 0)
    CMD="python bench.py --seconds 600 --steps 2 --warmup 1 --no-cpu-baseline"
    $CMD > $O/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench_600s.csv $CMD > $O/ncu_b.log 2>&1
    export PROF_SECONDS=600 PROF_REPS=2
    python tools/prof_encode.py > $O/plain_e.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'exact_gemm|imdct_sparse|quant_pack|window_tile|dequant_tile|ola_kernel|gather|scan' -s 11 -c 11 -o $O/r2_ncu_codec_kernels python tools/prof_encode.py > $O/ncu_e.log 2>&1
    GLC_PROF_MODE=1 python tools/prof_encode.py > $O/plain_f.log 2>&1 && GLC_PROF_MODE=1 ncu --set full --clock-control none --import-source on -k regex:'fast_' -s 3 -c 3 -o $O/r2_ncu_fast_kernels python tools/prof_encode.py > $O/ncu_f.log 2>&1
    PROF_SECONDS=120 python tools/prof_flac.py > $O/plain_l.log 2>&1 && PROF_SECONDS=120 ncu --set full --clock-control none --import-source on -k regex:'flac_' -s 2 -c 2 -o $O/r2_ncu_flac_kernels python tools/prof_flac.py > $O/ncu_l.log 2>&1
    tail -2 $O/ncu_e.log $O/ncu_f.log $O/ncu_l.log
    ;;
C)  # eight GPUs: headline line + config 4
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_exact_8gpu.json 2> $O/r2_bench_exact_8gpu.err
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_configs.py --configs 4 > $O/r2_config4_8gpu.jsonl 2> $O/r2_config4_8gpu.err
    tail -c 600 $O/r2_bench_exact_8gpu.json; tail -c 1500 $O/r2_config4_8gpu.jsonl
    ;;
esac
