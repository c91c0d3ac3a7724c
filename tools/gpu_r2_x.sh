#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r2x_t_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -k "golden or b64 or benchmarked" 2>&1 | tail -6 > gpurun_out/r2x_t_trainer.log
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench_a.json 2> gpurun_out/r2x_bench.err
LSPS_NO_SPLIT_FUSE=1 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench_nofuse.json 2>> gpurun_out/r2x_bench.err
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench_b.json 2>> gpurun_out/r2x_bench.err
LSPS_NO_SPLIT_FUSE=1 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench_nofuse_b.json 2>> gpurun_out/r2x_bench.err
timeout 300 python tools/step_profile.py > gpurun_out/r2x_step_profile.md 2>> gpurun_out/r2x_bench.err
cat gpurun_out/r2x_t_kernels.log gpurun_out/r2x_t_trainer.log gpurun_out/r2x_bench_a.json gpurun_out/r2x_bench_nofuse.json gpurun_out/r2x_bench_b.json gpurun_out/r2x_bench_nofuse_b.json; grep "_ex k1\|^step" gpurun_out/r2x_step_profile.md | cut -c1-110
