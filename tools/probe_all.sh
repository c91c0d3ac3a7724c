#!/bin/bash
# runs every probe case in its own process (a trapped kernel poisons the CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for c in k1_small k1_c64 k1_128 s2_64 s2_128 s2_dis2 s2_dis3 s2_dis4 dc_256 dc_128 dc_small k1_time s2_time dc_time dis4_time; do
  echo "== $c" 
  timeout 120 python tools/probe_igemm.py $c 2>&1 | tail -8
done 2>&1 | tee gpurun_out/probe_igemm.log
