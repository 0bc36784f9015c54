#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for k in tiefree peaks; do
  timeout 300 python scripts/bench_decode.py --kind $k | cut -c1-330
done
timeout 300 python bench.py --batch 8 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r1s_bench_b8.json 2> gpurun_out/r1s_bench.err; tail -3 gpurun_out/r1s_bench.err; cat gpurun_out/r1s_bench_b8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['layers_ms'])"
