#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_trainer_gpu.py -q -x -k "estimate3_resx" 2>&1 | tail -60 > gpurun_out/r2f_resx.log; tail -40 gpurun_out/r2f_resx.log
if grep -q "failed" gpurun_out/r2f_resx.log; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_trainer_gpu.py -q -x -k "estimate3_resx" 2>&1 | grep -v "^$" | head -80 > gpurun_out/r2f_sanitizer.log; head -60 gpurun_out/r2f_sanitizer.log
fi
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_trainer_gpu.py -q -k "golden or benchmarked_batch or gradients" 2>&1 | tail -15
for v in "mixed:" "mixed_1epi:LSPS_ONE_EPI_GROUP=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_light_$name.json 2>> gpurun_out/r2f_bench.err; echo "$name $(cat gpurun_out/r2f_bench_light_$name.json)"
done
