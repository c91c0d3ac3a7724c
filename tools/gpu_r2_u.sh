#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r2u_t_kernels.log
timeout 300 python tools/microbench_conv.py > gpurun_out/r2u_microbench.md 2>&1
cat gpurun_out/r2u_t_kernels.log; cut -c1-110 gpurun_out/r2u_microbench.md | tail -8
