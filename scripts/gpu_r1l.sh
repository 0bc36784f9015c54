#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classify.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "pool or march" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_detector.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --batch 4 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r1l_bench_b4.json 2> gpurun_out/r1l_bench.err; tail -3 gpurun_out/r1l_bench.err; cat gpurun_out/r1l_bench_b4.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['layers_ms'])"
