#!/bin/bash
# usage: bash tools/gpu_r2_multi.sh N   -- BASELINE configs 3 / 4 on N GPUs of one box (tools/bench_configs.py), JSON lines
N=$1
mkdir -p gpurun_out
OUT=gpurun_out/r2_multi_n$N.jsonl
: > $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) tools/bench_configs.py "$@" 2>> gpurun_out/r2_multi_n$N.err | tail -1 >> $OUT; tail -1 $OUT; }
# config 3: estimate3, nnyu, 32 per domain per rank (weak; at N=8 this IS batch 256 over 8 GPUs), eager and CUDA graphs
run --yaml nnyu --mode estimate3 --batch 32 --steps 40 --warmup 10
run --yaml nnyu --mode estimate3 --batch 32 --steps 40 --warmup 10 --graphs 1
LSPS_NO_EARLY_AR=1 run --yaml nnyu --mode estimate3 --batch 32 --steps 40 --warmup 10
# strong scaling of the same global batch 256
run --yaml nnyu --mode estimate3 --global-batch 256 --steps 40 --warmup 10
run --yaml nnyu --mode pretrain --batch 64 --steps 10 --warmup 4
# config 4: nicvl hyper-parameters, 32 per rank (batch 128 over 4 GPUs), pretrain and estimate3
run --yaml nicvl --mode pretrain --batch 32 --steps 10 --warmup 4
run --yaml nicvl --mode estimate3 --batch 32 --steps 40 --warmup 10
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -5 | tee gpurun_out/r2_dp2_test.log
fi
cat $OUT
