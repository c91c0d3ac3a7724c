#!/bin/bash
mkdir -p gpurun_out
{ for c in k1_small k1_c64 k1_128 s2_64 s2_128 s2_dis2 s2_dis4 dc_256 dc_128 dc_small; do timeout 90 python tools/probe_igemm.py $c 2>&1 | tail -1 | cut -c1-120; done
for kch in 1 2 3; do echo "### LSPS_KCH_SMALL=$kch"; for c in s2_time dc_time; do LSPS_KCH_SMALL=$kch timeout 90 python tools/probe_igemm.py $c 2>&1 | grep -E "FAIL|time (fwd|dgrad)" | tr '\n' ' '; echo " [$c]"; done; done
timeout 90 python tools/probe_igemm.py k1_time 2>&1 | tail -4; timeout 90 python tools/probe_igemm.py dis4_time 2>&1 | tail -4; } > gpurun_out/probe_igemm.log 2>&1
cat gpurun_out/probe_igemm.log
