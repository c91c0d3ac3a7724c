#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -6 gpurun_out/smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 4 -c 1 -o gpurun_out/prof_k1_v4 python tools/probe_igemm.py k1_time > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -o gpurun_out/prof_wgrad_v4 python tools/probe_igemm.py k1_time > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out/*.ncu-rep
