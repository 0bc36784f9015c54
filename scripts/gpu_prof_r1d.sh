#!/bin/bash
# round-1 profile pass D: gpu tests, configs[1] bench line, decode at configs[2], launch list, ncu full of the top kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest.log
tail -5 gpurun_out/r1d_pytest.log
timeout 600 python bench.py > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r1d_bench.json
timeout 300 python scripts/bench_decode.py --kind tiefree > gpurun_out/r1d_decode_c3_tiefree.json 2> gpurun_out/r1d_decode.err
timeout 300 python scripts/bench_decode.py --kind peaks > gpurun_out/r1d_decode_c3_peaks.json 2>> gpurun_out/r1d_decode.err
cat gpurun_out/r1d_decode_c3_tiefree.json gpurun_out/r1d_decode_c3_peaks.json
CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r1d_plain_b1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1d_launches.csv $CMD > gpurun_out/r1d_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_march|conv_halo|conv_up|stem|scan_kernel' -s 60 -c 40 -o gpurun_out/r1d_full $CMD > gpurun_out/r1d_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
