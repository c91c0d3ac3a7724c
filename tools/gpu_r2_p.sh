#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/repeat_golden.py pretrain_resx_nnyu_b1 25 > gpurun_out/r2p_repeat_resx.log 2>&1
LSPS_NO_SIDE=1 timeout 400 python tools/repeat_golden.py pretrain_resx_nnyu_b1 25 > gpurun_out/r2p_repeat_resx_noside.log 2>&1
timeout 400 python tools/repeat_golden.py pretrain_nnyu_b1 15 > gpurun_out/r2p_repeat_plain.log 2>&1
tail -30 gpurun_out/r2p_repeat_resx.log; tail -30 gpurun_out/r2p_repeat_resx_noside.log; tail -17 gpurun_out/r2p_repeat_plain.log
