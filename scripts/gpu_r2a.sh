#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_detector.py -m gpu -x -q 2>&1 | tail -4
for k in tiefree peaks; do
  echo -n "$k: "; timeout 300 python scripts/bench_decode.py --kind $k | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_median'], d['ms_min'], d['n_candidates'], d['flags'])"
done
timeout 300 python bench.py --batch 8 --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('1 GPU batch 8:', d['value'], d['e2e']['value'], d['roofline']['traffic'])"
CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|sieve|cand_|rank_|init_state|sort_write' -c 100 --csv --log-file gpurun_out/r2a_launches.csv $CMD > gpurun_out/r2a_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2a_launches.csv
