#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -4
for sd in 64 256; do
for v in 0 1; do
  for k in tiefree peaks; do
    echo "sample_div $sd variant $v $k"
    CETPICK_SAMPLE_DIV=$sd CETPICK_SIEVE_VARIANT=$v timeout 300 python scripts/bench_decode.py --kind $k | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms', d['ms_median'], d['ms_min'], 'n_candidates', d['n_candidates'], 'flags', d['flags'])"
  done
done
done
CMD="python scripts/bench_decode.py --kind peaks --iters 2 --warmup 1"
CETPICK_SAMPLE_DIV=256 CETPICK_SIEVE_VARIANT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sieve_kernel' -s 1 -c 1 -o gpurun_out/r1n_sieve $CMD > gpurun_out/r1n_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r1n_sieve.ncu-rep
