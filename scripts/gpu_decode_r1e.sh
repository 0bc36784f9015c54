#!/bin/bash
# decode pass: parity tests, configs[2] timing, launch list and one full ncu capture of the streaming kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -15
for k in tiefree peaks; do
  timeout 300 python scripts/bench_decode.py --kind $k > gpurun_out/r1e_decode_c3_$k.json 2>> gpurun_out/r1e_decode.err
  cat gpurun_out/r1e_decode_c3_$k.json
done
timeout 300 python scripts/bench_decode.py --kind peaks --shape 256,512,512 --K 900 | tee gpurun_out/r1e_decode_c2hm_peaks.json
CMD="python scripts/bench_decode.py --kind peaks --iters 2 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1e_decode_launches.csv $CMD > gpurun_out/r1e_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sieve_kernel' -s 1 -c 1 -o gpurun_out/r1e_sieve $CMD > gpurun_out/r1e_ncu_full.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r1e_decode_launches.csv | tee gpurun_out/r1e_decode_launches.txt
python scripts/ncu_summary.py full gpurun_out/r1e_sieve.ncu-rep | tee gpurun_out/r1e_sieve_full.txt
ls -la gpurun_out | tail
