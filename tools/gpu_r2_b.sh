#!/bin/bash
# round 2, call B: InstanceNorm statistics in the conv epilogues -- kernel tests, trainer parity, A/B bench, ncu targets
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "epilogue_statistics or dgrad_epilogue or conv_fwd_dgrad_wgrad or grouped" 2>&1 | tail -25 > gpurun_out/r2b_t_kernels.log; tail -8 gpurun_out/r2b_t_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -x -s -k "golden or benchmarked_batch or free_running" 2>&1 | tail -60 > gpurun_out/r2b_t_trainer.log; tail -30 gpurun_out/r2b_t_trainer.log
LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_light_new.json 2> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench_light_new.json
LSPS_OLD_IN=1 LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_light_old.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench_light_old.json
timeout 600 python tools/step_profile.py > gpurun_out/r2b_step_profile.md 2>> gpurun_out/r2b_bench.err; head -30 gpurun_out/r2b_step_profile.md
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2b_prof_targets \
   python tools/ncu_targets.py > gpurun_out/r2b_ncu_targets.md 2> gpurun_out/r2b_ncu_targets.err; tail -40 gpurun_out/r2b_ncu_targets.md; tail -5 gpurun_out/r2b_ncu_targets.err
ls -la gpurun_out/r2b_*
