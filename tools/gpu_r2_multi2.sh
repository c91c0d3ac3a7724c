#!/bin/bash
# 2-GPU sanity of the final build: data-parallel == single-GPU check, bench.py on 2 GPUs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -5 | tee gpurun_out/r2f_dp2_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
tail -c 900 gpurun_out/r2f_bench_n2.json; tail -3 gpurun_out/r2f_bench_n2.err
