#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "instnorm" 2>&1 | tail -15 > gpurun_out/t_in2.log; tail -4 gpurun_out/t_in2.log
timeout 120 python tools/probe_in.py 2>&1 | tail -4
timeout 300 python tools/step_profile.py > gpurun_out/step_profile_in2.md 2> gpurun_out/step_profile.err; head -1 gpurun_out/step_profile_in2.md; grep instnorm gpurun_out/step_profile_in2.md
timeout 600 python -m pytest tests/test_trainer_gpu.py -q -x -k "golden" 2>&1 | tail -5
