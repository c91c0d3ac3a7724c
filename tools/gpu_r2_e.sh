#!/bin/bash
# round 2, call E: ResNeXt generator (1x1 + grouped conv kinds), generator-front reuse between dis_update and gen_update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv1x1 or grouped_conv3x3 or conv_fwd_dgrad_wgrad" 2>&1 | tail -25 > gpurun_out/r2e_t_kernels.log; tail -6 gpurun_out/r2e_t_kernels.log
timeout 1200 python -m pytest tests/test_trainer_gpu.py -q -s -k "golden or benchmarked_batch or gradients" 2>&1 | tail -60 > gpurun_out/r2e_t_trainer.log; grep -v "^  step\|adam direction" gpurun_out/r2e_t_trainer.log | tail -25
for v in "mixed:" "mixed_nofront:LSPS_NO_FRONT_CACHE=1" "bf16:LSPS_PRECISION=bf16"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_light_$name.json 2>> gpurun_out/r2e_bench.err; echo "$name $(cat gpurun_out/r2e_bench_light_$name.json)"
done
tail -5 gpurun_out/r2e_bench.err
