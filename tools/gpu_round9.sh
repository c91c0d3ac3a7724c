#!/bin/bash
mkdir -p gpurun_out
{ for c in k1_small k1_c64 k1_128 s2_64 s2_128 dc_256 dc_128 dc_small k1_time s2_time dc_time; do timeout 90 python tools/probe_igemm.py $c 2>&1 | tail -5; done; } > gpurun_out/probe_igemm.log 2>&1
cat gpurun_out/probe_igemm.log
if grep -q "FAIL\|ERROR\|rror" gpurun_out/probe_igemm.log; then echo "PROBE FAILED - stopping"; exit 1; fi
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/t_gpu.log
tail -8 gpurun_out/t_gpu.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1300 gpurun_out/bench_n1.json
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
