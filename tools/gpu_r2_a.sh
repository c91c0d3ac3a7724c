#!/bin/bash
# round 2, call A: baseline of HEAD before the kernel work -- full GPU suite incl. the new B=64 test, the library line
# (unmodified reference through torch/cuDNN on this GPU), ncu --set full of every kernel family but K1, both bench arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
lscpu | head -20 >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -q -m gpu --durations=6 -s 2>&1 | tail -80 > gpurun_out/r2a_t_gpu.log; tail -12 gpurun_out/r2a_t_gpu.log
timeout 900 python tools/library_line.py --steps 5 --warmup 2 > gpurun_out/r2a_library_line.json 2> gpurun_out/r2a_library_line.err; tail -c 1500 gpurun_out/r2a_library_line.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2a_prof_targets \
   python tools/ncu_targets.py > gpurun_out/r2a_ncu_targets.md 2> gpurun_out/r2a_ncu_targets.err; tail -40 gpurun_out/r2a_ncu_targets.md
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; tail -c 800 gpurun_out/r2a_bench_ref.json
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -c 1500 gpurun_out/r2a_bench_n1.json
ls -la gpurun_out/r2a_*
