#!/bin/bash
# final validation of HEAD: whole GPU suite, smoke, bench (both arms), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -40 > gpurun_out/t_gpu_final.log; tail -14 gpurun_out/t_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -3 gpurun_out/smoke_final.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; tail -c 2500 gpurun_out/bench_n1_final.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>> gpurun_out/bench_n1_final.err; tail -c 600 gpurun_out/bench_ref_final.json
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
   --log-file gpurun_out/launches_v6.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list_v6.log 2>&1
ls -la gpurun_out/launches_v6.csv
