#!/bin/bash
# Source-level stall capture of the 32->32 full-resolution marching kernel (down0.c2 / up2.c2).
#   scripts/ncu_stem.sh <tag>
TAG=${1:-rX}
mkdir -p gpurun_out
CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:'stem_tc_kernel<.bool.0>' -s 3 -c 1 -o gpurun_out/${TAG}_stem $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_stem.ncu-rep --page source --csv > gpurun_out/${TAG}_stem_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_stem.ncu-rep --page details 2>/dev/null | grep -v "^\s*$" > gpurun_out/${TAG}_stem_details.txt
ls -la gpurun_out | grep ${TAG}
tail -3 gpurun_out/${TAG}_ncu.log
