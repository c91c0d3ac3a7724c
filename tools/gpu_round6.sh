#!/bin/bash
mkdir -p gpurun_out
{ for pf in 0 6 12 24; do echo "### LSPS_PF=$pf"; for c in k1_time s2_time dis4_time; do LSPS_PF=$pf timeout 90 python tools/probe_igemm.py $c 2>&1 | grep -E "FAIL|ERROR|time wgrad" ; done; done; } > gpurun_out/probe_pf.log 2>&1
cat gpurun_out/probe_pf.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "instnorm or conv" 2>&1 | tail -4
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 700 gpurun_out/bench_n1.json
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
