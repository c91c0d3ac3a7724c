#!/bin/bash
# final validation of HEAD (round 2): whole GPU suite, smoke, bench (both arms), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=8 2>&1 | grep -v "adam direction\|^  step " | tail -60 > gpurun_out/r2f_pytest_gpu.log; tail -14 gpurun_out/r2f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -3 gpurun_out/r2f_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 2500 gpurun_out/r2f_bench_n1.json
LSPS_PRECISION=bf16 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1_bf16.json 2>> gpurun_out/r2f_bench_n1.err; tail -c 600 gpurun_out/r2f_bench_n1_bf16.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2>> gpurun_out/r2f_bench_n1.err; tail -c 600 gpurun_out/r2f_bench_ref.json
timeout 300 python tools/step_profile.py > gpurun_out/r2f_step_profile.md 2>> gpurun_out/r2f_bench_n1.err
timeout 300 python tools/microbench_conv.py > gpurun_out/r2f_microbench_conv.md 2>&1
timeout 200 python tools/microbench_stem.py > gpurun_out/r2f_microbench_stem.md 2>&1
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
   --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2f_ncu_list.log 2>&1
ls -la gpurun_out/r2f_launches.csv
