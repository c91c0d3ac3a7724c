#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "wgrad_pair or conv_fwd_dgrad_wgrad" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -k "benchmarked_batch or golden" 2>&1 | tail -4
for v in "pair:" "nopair:LSPS_NO_WGRAD_PAIR=1" "pair2:" "nopair2:LSPS_NO_WGRAD_PAIR=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs LSPS_BENCH_LIGHT=1 python bench.py --steps 12 --warmup 4 2>/dev/null | tail -1 | sed "s/^/$name /"
done
