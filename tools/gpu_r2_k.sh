#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_drivers_gpu.py tests/test_zz_augment_gpu.py -q -x 2>&1 | tail -30
timeout 1500 python -m pytest tests/test_trainer_gpu.py -q -s -k "standalone or teacher_forced or free_running" 2>&1 | grep -v "^  step" | tail -25 | cut -c1-500
python tools/bench_side_rows.py > gpurun_out/r2k_side_rows.json 2> gpurun_out/r2k_side_rows.err; cat gpurun_out/r2k_side_rows.json; tail -3 gpurun_out/r2k_side_rows.err
timeout 600 python tools/microbench_conv.py > gpurun_out/r2k_microbench_conv.md 2> gpurun_out/r2k_microbench_conv.err; cat gpurun_out/r2k_microbench_conv.md
