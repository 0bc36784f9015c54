#!/bin/bash
set -x
CMD="python bench.py --batch 1 --shape 128,512,512 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_march' -s 20 -c 10 -o gpurun_out/prof_march_r1b $CMD > gpurun_out/ncu_m_full.log 2>&1
tail -1 gpurun_out/plain_m.log | cut -c1-200
ls -la gpurun_out | tail -5
