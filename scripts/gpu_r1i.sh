#!/bin/bash
set -x
mkdir -p gpurun_out
export CETPICK_NO_PDL=1
for v in 0 1 2 3 4 5; do
  for k in tiefree peaks; do
    echo "variant $v $k"
    CETPICK_SIEVE_VARIANT=$v timeout 300 python scripts/bench_decode.py --kind $k | cut -c1-140
  done
done
CMD="python scripts/bench_decode.py --iters 2 --warmup 1 --kind"
for v in 0 1; do
for k in peaks tiefree; do
CETPICK_SIEVE_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sieve_kernel' -s 1 -c 1 -o gpurun_out/r1i_sieve_v${v}_$k $CMD $k > gpurun_out/r1i_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r1i_sieve_v${v}_$k.ncu-rep
done
done
