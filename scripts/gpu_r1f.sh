#!/bin/bash
# decode variants sweep + parity; then launch list and a bounded full ncu capture of the conv kernels of one forward
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -5
for v in 0 1 2 3 4 5; do
  for k in tiefree peaks; do
    echo "variant $v $k"
    CETPICK_SIEVE_VARIANT=$v timeout 300 python scripts/bench_decode.py --kind $k | cut -c1-200
  done
done
CMD="python scripts/bench_decode.py --kind peaks --iters 2 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cetpick' -c 200 --csv --log-file gpurun_out/r1f_decode_launches.csv $CMD > gpurun_out/r1f_ncu_launch.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r1f_decode_launches.csv | tee gpurun_out/r1f_decode_launches.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sieve_kernel|scan_kernel' -s 3 -c 3 -o gpurun_out/r1f_decode $CMD > gpurun_out/r1f_ncu_full.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r1f_decode.ncu-rep | tee gpurun_out/r1f_decode_full.txt
# ---- detector forward
CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r1f_plain_b1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cetpick' -c 400 --csv --log-file gpurun_out/r1f_launches.csv $CMD > gpurun_out/r1f_ncu_launch2.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r1f_launches.csv | tee gpurun_out/r1f_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_march|conv_halo|conv_up|stem|pool' -s 48 -c 24 -o gpurun_out/r1f_conv $CMD > gpurun_out/r1f_ncu_full2.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r1f_conv.ncu-rep | tee gpurun_out/r1f_conv_full.txt
ncu -i gpurun_out/r1f_conv.ncu-rep --page details > gpurun_out/r1f_conv_details.txt 2>/dev/null
ls -la gpurun_out
sz=$(stat -c %s gpurun_out/r1f_conv.ncu-rep); if [ "$sz" -gt 45000000 ]; then rm gpurun_out/r1f_conv.ncu-rep; fi
