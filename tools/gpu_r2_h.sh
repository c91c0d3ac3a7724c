#!/bin/bash
# round 2, call H: whole GPU suite (fused D-head + BCE in the trainer, BN / ResNeXt rows), smoke, both bench arms, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=6 -s 2>&1 | tail -150 > gpurun_out/r2h_t_gpu.log; grep -v "^  step\|adam direction\|  B=64" gpurun_out/r2h_t_gpu.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -4 gpurun_out/r2h_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -c 2600 gpurun_out/r2h_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.json 2>> gpurun_out/r2h_bench_n1.err; tail -c 500 gpurun_out/r2h_bench_ref.json
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
   --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2h_ncu_list.log 2>&1
grep -c "conv_igemm\|wgrad_kernel" gpurun_out/r2h_launches.csv
