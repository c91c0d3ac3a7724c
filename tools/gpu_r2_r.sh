#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_(fwd|wgrad)" -s 2 -c 2 -f -o gpurun_out/r2r_stem python tools/ncu_stem.py > gpurun_out/r2r_ncu.log 2>&1
tail -5 gpurun_out/r2r_ncu.log
