#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_detector.py -m gpu -x -q --durations=3 2>&1 | tail -12
timeout 300 python bench.py --batch 8 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r1w_bench_b8.json 2> gpurun_out/r1w_bench.err; tail -3 gpurun_out/r1w_bench.err; cat gpurun_out/r1w_bench_b8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['layers_ms'])"
