#!/bin/bash
# round 2, call C: split-bf16 discriminator, INBWD via the stored activation, block-major multi-phase tile order, one-launch pack
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2c_t_kernels.log; tail -12 gpurun_out/r2c_t_kernels.log
timeout 1200 python -m pytest tests/test_trainer_gpu.py -q -s -k "golden or benchmarked_batch or teacher_forced or gradients or snapshot or graph or eval" 2>&1 | tail -80 > gpurun_out/r2c_t_trainer.log; grep -v "^  step" gpurun_out/r2c_t_trainer.log | tail -40
LSPS_PRECISION=bf16 timeout 900 python -m pytest tests/test_trainer_gpu.py -q -s -k "teacher_forced" 2>&1 | tail -12 > gpurun_out/r2c_t_trainer_bf16.log; grep "teacher-forced\|passed\|failed" gpurun_out/r2c_t_trainer_bf16.log
LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_light_mixed.json 2> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench_light_mixed.json
LSPS_PRECISION=bf16 LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_light_bf16.json 2>> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench_light_bf16.json
LSPS_PRECISION=bf16 LSPS_PHASE_MAJOR=1 LSPS_BENCH_LIGHT=1 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_light_bf16_phasemajor.json 2>> gpurun_out/r2c_bench.err; cat gpurun_out/r2c_bench_light_bf16_phasemajor.json
timeout 600 python tools/step_profile.py > gpurun_out/r2c_step_profile.md 2>> gpurun_out/r2c_bench.err; head -45 gpurun_out/r2c_step_profile.md
tail -5 gpurun_out/r2c_bench.err
