#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_detector.py -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
  for k in tiefree peaks; do
    echo -n "$k: "
    timeout 300 python scripts/bench_decode.py --kind $k | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_median'], d['ms_min'], d['n_candidates'], d['flags'])"
  done
done
timeout 300 python scripts/bench_decode.py --kind peaks --shape 256,512,512 --K 900 | cut -c1-200
CMD="python scripts/bench_decode.py --kind peaks --iters 2 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|sieve|cand_|rank_|init_state' -c 200 --csv --log-file gpurun_out/r1u_decode_launches.csv $CMD > gpurun_out/r1u_ncu1.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r1u_decode_launches.csv
