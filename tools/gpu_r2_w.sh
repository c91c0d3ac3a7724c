#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_kernels_gpu.py tests/test_layers_gpu.py -q -k "not teacher_forced and not free_running" 2>&1 | tail -6 > gpurun_out/r2w_tests.log
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_a.json 2> gpurun_out/r2w_bench.err
LSPS_NO_STAT_ARENA=1 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_noarena.json 2>> gpurun_out/r2w_bench.err
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_b.json 2>> gpurun_out/r2w_bench.err
LSPS_NO_STAT_ARENA=1 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_noarena_b.json 2>> gpurun_out/r2w_bench.err
cat gpurun_out/r2w_tests.log gpurun_out/r2w_bench_a.json gpurun_out/r2w_bench_noarena.json gpurun_out/r2w_bench_b.json gpurun_out/r2w_bench_noarena_b.json; tail -3 gpurun_out/r2w_bench.err
