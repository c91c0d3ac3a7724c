#!/bin/bash
# round 2, call D: fused four-phase up-sampling kernel, single-buffer split stem wgrad, in-kernel Philox noise, memset/memcpy
# instead of ATen fills/cats -- tests, A/B bench, launch list of one step, ncu of the new kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2d_t_kernels.log; tail -6 gpurun_out/r2d_t_kernels.log
timeout 1200 python -m pytest tests/test_trainer_gpu.py -q -s -k "golden or benchmarked_batch or gradients or graph or free_running" 2>&1 | tail -80 > gpurun_out/r2d_t_trainer.log; grep -v "^  step\|adam direction" gpurun_out/r2d_t_trainer.log | tail -25
for v in "mixed:" "bf16:LSPS_PRECISION=bf16" "bf16_noup64:LSPS_PRECISION=bf16 LSPS_NO_UP64=1" "mixed_noup64:LSPS_NO_UP64=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench_light_$name.json 2>> gpurun_out/r2d_bench.err; echo "$name $(cat gpurun_out/r2d_bench_light_$name.json)"
done
timeout 600 python tools/step_profile.py > gpurun_out/r2d_step_profile.md 2>> gpurun_out/r2d_bench.err; head -30 gpurun_out/r2d_step_profile.md
LSPS_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
   --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2d_ncu_list.log 2>&1
python tools/launch_table.py gpurun_out/r2d_launches.csv 437 > gpurun_out/r2d_launch_table.md 2>&1; head -45 gpurun_out/r2d_launch_table.md
LSPS_NCU_CASES=norm,split,conv timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r2d_prof_targets \
   python tools/ncu_targets.py > gpurun_out/r2d_ncu_targets.md 2> gpurun_out/r2d_ncu_targets.err; tail -40 gpurun_out/r2d_ncu_targets.md; tail -3 gpurun_out/r2d_ncu_targets.err
tail -5 gpurun_out/r2d_bench.err
