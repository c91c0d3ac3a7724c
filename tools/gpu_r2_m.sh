#!/bin/bash
# validation of the two-epilogue-group kernels on the split-precision convs + A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2m_t_kernels.log
timeout 900 python -m pytest tests/test_trainer_gpu.py -x -q -s -k "golden or teacher_forced or b64" 2>&1 | tail -30 > gpurun_out/r2m_t_trainer.log
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_eg2.json 2> gpurun_out/r2m_bench.err
LSPS_ONE_EPI_GROUP=1 LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_eg1.json 2>> gpurun_out/r2m_bench.err
LSPS_BENCH_LIGHT=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_eg2_b.json 2>> gpurun_out/r2m_bench.err
cat gpurun_out/r2m_t_kernels.log gpurun_out/r2m_t_trainer.log gpurun_out/r2m_bench_eg2.json gpurun_out/r2m_bench_eg1.json gpurun_out/r2m_bench_eg2_b.json
