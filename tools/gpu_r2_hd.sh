#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/r2hd_t.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -k "golden or benchmarked or train_map or standalone or nets" 2>&1 | tail -6 >> gpurun_out/r2hd_t.log
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2hd_bench_a.json 2> gpurun_out/r2hd_bench.err
LSPS_NO_HEAD_FUSE=1 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2hd_bench_nofuse.json 2>> gpurun_out/r2hd_bench.err
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2hd_bench_b.json 2>> gpurun_out/r2hd_bench.err
LSPS_NO_HEAD_FUSE=1 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2hd_bench_nofuse_b.json 2>> gpurun_out/r2hd_bench.err
cat gpurun_out/r2hd_t.log gpurun_out/r2hd_bench_a.json gpurun_out/r2hd_bench_nofuse.json gpurun_out/r2hd_bench_b.json gpurun_out/r2hd_bench_nofuse_b.json; tail -3 gpurun_out/r2hd_bench.err
