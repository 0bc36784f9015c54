#!/bin/bash
set -x
CMD="python scripts/bench_decode.py --kind peaks --iters 1 --warmup 1"
$CMD > gpurun_out/plain_d2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scan_kernel' -s 10 -c 1 -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_d_full.log 2>&1
