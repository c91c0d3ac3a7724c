#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "stem or instnorm" 2>&1 | tail -25 > gpurun_out/t_stem.log
cat gpurun_out/t_stem.log
if grep -q "failed\|rror" gpurun_out/t_stem.log; then echo "KERNEL TESTS FAILED"; exit 1; fi
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/t_gpu.log
tail -8 gpurun_out/t_gpu.log
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 1300 gpurun_out/bench_n1.json
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
