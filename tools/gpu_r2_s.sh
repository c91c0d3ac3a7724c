#!/bin/bash
mkdir -p gpurun_out
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_a.json 2> gpurun_out/r2s_bench.err
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_b.json 2>> gpurun_out/r2s_bench.err
LSPS_PRECISION=bf16 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_bf16.json 2>> gpurun_out/r2s_bench.err
timeout 300 python tools/step_profile.py > gpurun_out/r2s_step_profile.md 2>> gpurun_out/r2s_bench.err
cat gpurun_out/r2s_bench_a.json gpurun_out/r2s_bench_b.json gpurun_out/r2s_bench_bf16.json; grep -i "stem\|^step" gpurun_out/r2s_step_profile.md
