#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_(up64|resb)" -f -o gpurun_out/r2r_two python tools/ncu_two.py > gpurun_out/r2r_ncu.log 2>&1
tail -3 gpurun_out/r2r_ncu.log
