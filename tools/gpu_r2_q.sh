#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "stem" 2>&1 | tail -5 > gpurun_out/r2q_t_stem.log
timeout 200 python tools/microbench_stem.py > gpurun_out/r2q_microbench_stem.md 2>&1
LSPS_STEM_WG_NBUF=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "stem" 2>&1 | tail -5 > gpurun_out/r2q_t_stem_nbuf1.log
LSPS_STEM_WG_NBUF=1 timeout 200 python tools/microbench_stem.py > gpurun_out/r2q_microbench_stem_nbuf1.md 2>&1
LSPS_STEM_WG_NBUF=2 timeout 200 python tools/microbench_stem.py > gpurun_out/r2q_microbench_stem_nbuf2.md 2>&1
cat gpurun_out/r2q_t_stem.log gpurun_out/r2q_microbench_stem.md gpurun_out/r2q_t_stem_nbuf1.log gpurun_out/r2q_microbench_stem_nbuf1.md gpurun_out/r2q_microbench_stem_nbuf2.md
