#!/bin/bash
# ncu --set full of the target kernels on the final build (tools/ncu_targets.py cases)
mkdir -p gpurun_out
LSPS_NCU_CASES=norm,split,conv timeout 420 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2f_prof_targets \
   python tools/ncu_targets.py > gpurun_out/r2f_ncu_targets.md 2> gpurun_out/r2f_ncu_targets.err; tail -5 gpurun_out/r2f_ncu_targets.md; tail -2 gpurun_out/r2f_ncu_targets.err
