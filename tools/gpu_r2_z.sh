#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 300 gpurun_out/r2z_bench_n1.json; tail -2 gpurun_out/r2z_bench_n1.err
python bench.py > gpurun_out/r2z_bench_default.json 2>> gpurun_out/r2z_bench_n1.err
python -c "
import json
for f in ('gpurun_out/r2z_bench_n1.json','gpurun_out/r2z_bench_default.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d['clocks'], d['e2e']['value'])
"
