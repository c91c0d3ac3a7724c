#!/bin/bash
mkdir -p gpurun_out
{
for dbg in 0 1 2 3; do
  for cg in 1 2; do
    echo "#### LSPS_DBG=$dbg LSPS_FORCE_CG=$cg"
    for c in k1_time s2_time dc_time; do LSPS_DBG=$dbg LSPS_FORCE_CG=$cg timeout 90 python tools/probe_igemm.py $c 2>&1 | grep -E "time (fwd|dgrad|wgrad)" | tr '\n' ' '; echo " [$c]"; done
  done
done
echo "#### default policy"
for c in k1_time s2_time dc_time dis4_time; do timeout 90 python tools/probe_igemm.py $c 2>&1 | tail -5; done
} > gpurun_out/probe_dbg.log 2>&1
cat gpurun_out/probe_dbg.log
