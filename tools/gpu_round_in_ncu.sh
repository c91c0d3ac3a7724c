#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:in_cl_ -s 8 -c 4 -o gpurun_out/prof_in_cl -f python tools/probe_in.py > gpurun_out/ncu_in_cl.log 2>&1
LSPS_IN_NO_CL=1 timeout 300 ncu --set full --clock-control none -k regex:instnorm_ -s 8 -c 4 -o gpurun_out/prof_in_old -f python tools/probe_in.py > gpurun_out/ncu_in_old.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_in_cl.log
