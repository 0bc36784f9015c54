#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_unet.py tests/test_gpu_detector.py -m gpu -x -q 2>&1 | tail -8
for k in tiefree peaks; do
  timeout 300 python scripts/bench_decode.py --kind $k | cut -c1-330
done
timeout 300 python bench.py --batch 8 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r1q_bench_b8.json 2> gpurun_out/r1q_bench.err; tail -3 gpurun_out/r1q_bench.err; cat gpurun_out/r1q_bench_b8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])"
CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|sieve|cand_|rank_|init_state' -c 100 --csv --log-file gpurun_out/r1q_launches.csv $CMD > gpurun_out/r1q_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r1q_launches.csv
