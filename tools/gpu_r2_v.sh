#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -4 > gpurun_out/r2v_t_kernels.log
timeout 300 python tools/microbench_conv.py > gpurun_out/r2v_microbench.md 2>&1
timeout 200 python tools/microbench_stem.py > gpurun_out/r2v_microbench_stem.md 2>&1
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_a.json 2> gpurun_out/r2v_bench.err
LSPS_PRECISION=bf16 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_bf16.json 2>> gpurun_out/r2v_bench.err
timeout 300 python tools/step_profile.py > gpurun_out/r2v_step_profile.md 2>> gpurun_out/r2v_bench.err
cat gpurun_out/r2v_t_kernels.log; cut -c1-110 gpurun_out/r2v_microbench.md | tail -8; cat gpurun_out/r2v_microbench_stem.md gpurun_out/r2v_bench_a.json gpurun_out/r2v_bench_bf16.json; head -12 gpurun_out/r2v_step_profile.md | cut -c1-120
