#!/bin/bash
# compute-sanitizer passes over one ResNeXt and one plain pretrain golden case (every allocation a cudaMalloc so that
# initcheck sees recycled memory as uninitialised)
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 700 compute-sanitizer --tool initcheck --print-limit 30 --log-file gpurun_out/r2o_initcheck_resx.log \
  python -m pytest tests/test_trainer_gpu.py -q -x -k "golden and pretrain_resx_nnyu_b1" > gpurun_out/r2o_initcheck_resx.out 2>&1
timeout 700 compute-sanitizer --tool initcheck --print-limit 30 --log-file gpurun_out/r2o_initcheck_plain.log \
  python -m pytest tests/test_trainer_gpu.py -q -x -k "golden and pretrain_nnyu_b2_hand" > gpurun_out/r2o_initcheck_plain.out 2>&1
timeout 700 compute-sanitizer --tool memcheck --print-limit 30 --log-file gpurun_out/r2o_memcheck_resx.log \
  python -m pytest tests/test_trainer_gpu.py -q -x -k "golden and pretrain_resx_nnyu_b1" > gpurun_out/r2o_memcheck_resx.out 2>&1
for f in gpurun_out/r2o_*.log gpurun_out/r2o_*.out; do echo "== $f"; tail -25 $f | cut -c1-250; done
