#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s --durations=10 2>&1 | tail -150 > gpurun_out/t_gpu.log
tail -5 gpurun_out/t_gpu.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1500 gpurun_out/bench_n1.json
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
for c in k1_time s2_time dc_time dis4_time; do timeout 120 python tools/probe_igemm.py $c 2>&1 | tail -6; done > gpurun_out/probe_igemm.log
cat gpurun_out/probe_igemm.log
