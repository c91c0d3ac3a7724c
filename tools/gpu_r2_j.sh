#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py -q 2>&1 | tail -4
timeout 1200 python -m pytest tests/test_trainer_gpu.py -q -k "golden or gradients or graph or free_running" 2>&1 | tail -4
for v in "small_bn:" "no_small_bn:LSPS_NO_SMALL_BN=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs python tools/bench_configs.py --yaml nnyu --mode estimate3 --batch 32 --steps 60 --warmup 15 2>/dev/null | tail -1 | cut -c1-260
  env $envs python tools/bench_configs.py --yaml nnyu --mode estimate3 --batch 32 --steps 60 --warmup 15 --graphs 1 2>/dev/null | tail -1 | cut -c1-260
done
python tools/bench_configs.py --yaml nnyu --mode estimate0 --batch 32 --steps 60 --warmup 15 --graphs 1 2>/dev/null | tail -1 | cut -c1-260
python tools/bench_modes.py > gpurun_out/r2j_modes.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2j_modes.json'))
for k,v in d.items(): print(k, round(v['event_ms'],3), v['launches'])"
