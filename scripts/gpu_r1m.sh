#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_classify.py -m gpu -x -q 2>&1 | tail -8
for sd in 64 256; do
for v in 0 1; do
  for k in tiefree peaks; do
    echo "sample_div $sd variant $v $k"
    CETPICK_SAMPLE_DIV=$sd CETPICK_SIEVE_VARIANT=$v timeout 300 python scripts/bench_decode.py --kind $k | cut -c1-140
    CETPICK_SAMPLE_DIV=$sd CETPICK_SIEVE_VARIANT=$v timeout 300 python scripts/bench_decode.py --kind $k | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   n_candidates', d['n_candidates'], 'flags', d['flags'])"
  done
done
done
CMD="python scripts/bench_decode.py --kind peaks --iters 2 --warmup 1"
CETPICK_SAMPLE_DIV=256 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|sieve|cand_|rank_|init_state' -c 200 --csv --log-file gpurun_out/r1m_decode_launches.csv $CMD > gpurun_out/r1m_ncu_launch.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r1m_decode_launches.csv
