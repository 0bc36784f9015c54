#!/bin/bash
set -x
mkdir -p gpurun_out
export CETPICK_NO_PDL=1
CMD="python scripts/bench_decode.py --iters 2 --warmup 1 --kind"
for v in 0 1 4; do
CETPICK_SIEVE_VARIANT=$v CETPICK_SIEVE_NOHIT=1 timeout 600 ncu --set full --clock-control none -k regex:'sieve_kernel' -s 1 -c 1 -o gpurun_out/r1h_nohit_v$v $CMD peaks > gpurun_out/r1h_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r1h_nohit_v$v.ncu-rep
done
for k in peaks tiefree; do
CETPICK_SIEVE_VARIANT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sieve_kernel' -s 1 -c 1 -o gpurun_out/r1h_sieve_$k $CMD $k > gpurun_out/r1h_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r1h_sieve_$k.ncu-rep
done
CETPICK_SIEVE_VARIANT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scan_kernel|cand_hist' -s 19 -c 4 -o gpurun_out/r1h_sample $CMD peaks > gpurun_out/r1h_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r1h_sample.ncu-rep
