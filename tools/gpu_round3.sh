#!/bin/bash
mkdir -p gpurun_out
for c in k1_small s2_64 dc_128 s2_dis4 k1_time s2_time dc_time dis4_time; do echo "== $c"; timeout 90 python tools/probe_igemm.py $c 2>&1 | tail -6; done > gpurun_out/probe_igemm_cg2.log 2>&1
cat gpurun_out/probe_igemm_cg2.log
if grep -q "FAIL\|ERROR\|rror" gpurun_out/probe_igemm_cg2.log; then echo "PROBE FAILED - stopping"; exit 1; fi
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/t_gpu.log
tail -6 gpurun_out/t_gpu.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 900 gpurun_out/bench_n1.json
