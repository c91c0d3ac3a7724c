#!/bin/bash
# one GPU session: bench (N=1), ncu launch list of the same command, ncu --set full of the K1 conv + wgrad kernels
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 4 -c 2 -o gpurun_out/prof_k1 \
   python tools/probe_igemm.py k1_time > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -o gpurun_out/prof_wgrad \
   python tools/probe_igemm.py k1_time > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out
