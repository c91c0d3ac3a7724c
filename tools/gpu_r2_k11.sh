#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "vae_step" 2>&1 | tail -5 > gpurun_out/r2k11_t.log
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_drivers_gpu.py -q -k "vae or golden or driver or pipeline or graph" 2>&1 | tail -5 >> gpurun_out/r2k11_t.log
timeout 300 python tools/bench_modes.py > gpurun_out/r2k11_modes.json 2> gpurun_out/r2k11.err
LSPS_NO_VAE_FUSED=1 timeout 300 python tools/bench_modes.py > gpurun_out/r2k11_modes_unfused.json 2>> gpurun_out/r2k11.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> gpurun_out/r2k11_t.log
cat gpurun_out/r2k11_t.log; grep -i "vae" gpurun_out/r2k11_modes.json | head; echo ---; grep -i "vae" gpurun_out/r2k11_modes_unfused.json | head; tail -3 gpurun_out/r2k11.err
