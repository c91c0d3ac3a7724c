#!/bin/bash
# train_map row: kind-3 conv kernels + l2 kernel, the map goldens / gradient test, and the step timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv_fwd or l2" 2>&1 | tail -25 > gpurun_out/t_map_kernels.log
tail -4 gpurun_out/t_map_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -k "map" 2>&1 | tail -40 > gpurun_out/t_map_trainer.log
tail -12 gpurun_out/t_map_trainer.log
timeout 300 python tools/bench_map.py > gpurun_out/bench_map.json 2> gpurun_out/bench_map.err
tail -30 gpurun_out/bench_map.json; tail -5 gpurun_out/bench_map.err
