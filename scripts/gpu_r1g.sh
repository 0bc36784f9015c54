#!/bin/bash
set -x
mkdir -p gpurun_out
for pdl in 0 1; do
  if [ $pdl = 0 ]; then export CETPICK_NO_PDL=1; else unset CETPICK_NO_PDL; fi
  for v in 0 1; do
    echo "pdl=$pdl variant $v"
    CETPICK_SIEVE_VARIANT=$v timeout 300 python scripts/bench_decode.py --kind peaks | cut -c1-140
  done
  CMD="python scripts/bench_decode.py --kind peaks --iters 2 --warmup 1"
  CETPICK_SIEVE_VARIANT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|sieve|cand_|rank_|init_state' -c 200 --csv --log-file gpurun_out/r1g_decode_launches_pdl$pdl.csv $CMD > gpurun_out/r1g_ncu_launch.log 2>&1
  python scripts/ncu_summary.py launches gpurun_out/r1g_decode_launches_pdl$pdl.csv
done
