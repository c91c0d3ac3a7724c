#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench_conv.py > gpurun_out/r2t_microbench_default.md 2>&1
LSPS_FORCE_CG=2 timeout 300 python tools/microbench_conv.py > gpurun_out/r2t_microbench_cg2.md 2>&1
cut -c1-150 gpurun_out/r2t_microbench_default.md; cut -c1-150 gpurun_out/r2t_microbench_cg2.md
