#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "adam or pack" 2>&1 | tail -3 > gpurun_out/r2ad_t.log
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_layers_gpu.py -q -k "golden or layers or norm or snapshot or graph" 2>&1 | tail -4 >> gpurun_out/r2ad_t.log
timeout 200 python tools/step_profile.py > gpurun_out/r2ad_step_profile.md 2> gpurun_out/r2ad.err
cat gpurun_out/r2ad_t.log; grep "^step\|adam\|pack" gpurun_out/r2ad_step_profile.md | cut -c1-150; tail -2 gpurun_out/r2ad.err
