#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench_conv.py > gpurun_out/r2y_mb_default.md 2>&1
LSPS_KCH2_LIGHT=1 timeout 300 python tools/microbench_conv.py > gpurun_out/r2y_mb_kch2_1.md 2>&1
LSPS_KCH2_LIGHT=3 timeout 300 python tools/microbench_conv.py > gpurun_out/r2y_mb_kch2_3.md 2>&1
LSPS_KCH2_LIGHT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "conv_fwd_dgrad_wgrad" 2>&1 | tail -3
for f in gpurun_out/r2y_mb_default.md gpurun_out/r2y_mb_kch2_1.md gpurun_out/r2y_mb_kch2_3.md; do echo $f; cut -c1-100 $f | tail -8; done
