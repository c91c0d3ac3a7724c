#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "instnorm" 2>&1 | tail -15 > gpurun_out/t_in.log; tail -5 gpurun_out/t_in.log
{ echo "## cluster kernels"; timeout 120 python tools/probe_in.py 2>&1 | tail -4; echo "## LSPS_IN_NO_CL=1"; LSPS_IN_NO_CL=1 timeout 120 python tools/probe_in.py 2>&1 | tail -4; } > gpurun_out/probe_in_cl.log 2>&1
cat gpurun_out/probe_in_cl.log
timeout 300 python tools/step_profile.py > gpurun_out/step_profile_cl.md 2> gpurun_out/step_profile.err; head -12 gpurun_out/step_profile_cl.md
LSPS_IN_NO_CL=1 timeout 300 python tools/step_profile.py 2>/dev/null | head -1
