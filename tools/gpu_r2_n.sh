#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2n_resx.log
for i in 1 2 3; do
  timeout 200 python -m pytest tests/test_trainer_gpu.py -q -s -k "golden and resx" 2>&1 | grep "largest\|passed\|failed" >> gpurun_out/r2n_resx.log
  LSPS_ONE_EPI_GROUP=1 timeout 200 python -m pytest tests/test_trainer_gpu.py -q -s -k "golden and resx" 2>&1 | grep "largest\|passed\|failed" | sed 's/^/EG1 /' >> gpurun_out/r2n_resx.log
done
timeout 900 python -m pytest tests/test_trainer_gpu.py -q -s 2>&1 | grep -v "adam direction" | tail -40 > gpurun_out/r2n_t_trainer.log
cat gpurun_out/r2n_resx.log gpurun_out/r2n_t_trainer.log
