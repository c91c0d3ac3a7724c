#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py -q 2>&1 | tail -5
timeout 1200 python -m pytest tests/test_trainer_gpu.py -q -s -k "golden or benchmarked_batch or teacher_forced or gradients" 2>&1 | tail -60 > gpurun_out/r2i_t_trainer.log; grep -v "^  step\|adam direction" gpurun_out/r2i_t_trainer.log | tail -22 | cut -c1-600
LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3
