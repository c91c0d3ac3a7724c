#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "grouped or instnorm" 2>&1 | tail -25 > gpurun_out/t_group_kernels.log; tail -6 gpurun_out/t_group_kernels.log
timeout 300 python tools/step_profile.py > gpurun_out/step_profile_group.md 2> gpurun_out/step_profile.err; head -14 gpurun_out/step_profile_group.md; tail -3 gpurun_out/step_profile.err
LSPS_NO_GROUP=1 timeout 300 python tools/step_profile.py 2>/dev/null | head -1
timeout 1500 python -m pytest tests/test_trainer_gpu.py -q -x 2>&1 | tail -30 > gpurun_out/t_group_trainer.log; tail -8 gpurun_out/t_group_trainer.log
