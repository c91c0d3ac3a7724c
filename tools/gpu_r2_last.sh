#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_trainer_gpu.py tests/test_drivers_gpu.py -q -k "benchmarked or free_running or drivers or pipeline or map" 2>&1 | tail -4 > gpurun_out/r2last_t.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> gpurun_out/r2last_t.log
cat gpurun_out/r2last_t.log
