#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_kernels_gpu.py -q 2>&1 | tail -30 > gpurun_out/r2g_t_kernels.log; tail -25 gpurun_out/r2g_t_kernels.log
timeout 1200 python -m pytest tests/test_trainer_gpu.py -q -k "golden or benchmarked_batch or gradients" 2>&1 | tail -15
for v in "mixed:" "mixed_noup64:LSPS_NO_UP64=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_light_$name.json 2>> gpurun_out/r2g_bench.err; echo "$name $(cat gpurun_out/r2g_bench_light_$name.json)"
done
python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_n1.json 2>> gpurun_out/r2g_bench.err; tail -c 2000 gpurun_out/r2g_bench_n1.json
