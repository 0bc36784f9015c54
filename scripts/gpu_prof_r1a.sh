#!/bin/bash
# round-1 profile pass A: decode at configs[2] size, launch list + full ncu of the detector at configs[0] size
set -x
mkdir -p gpurun_out
python scripts/bench_decode.py --kind tiefree > gpurun_out/decode_c3_tiefree.json 2> gpurun_out/decode_c3.err
python scripts/bench_decode.py --kind peaks > gpurun_out/decode_c3_peaks.json 2>> gpurun_out/decode_c3.err
python scripts/bench_decode.py --shape 256,512,512 --K 900 --kind peaks > gpurun_out/decode_c2hm_peaks.json 2>> gpurun_out/decode_c3.err
cat gpurun_out/decode_c3_tiefree.json gpurun_out/decode_c3_peaks.json gpurun_out/decode_c2hm_peaks.json
CMD="python bench.py --batch 1 --shape 128,512,512 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_c1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_tc|stem|pool2x2|hm_head|scan_|cand_|sort_write|init_state' -c 300 --csv --log-file gpurun_out/launches_r1a.csv $CMD > gpurun_out/ncu_c1.log 2>&1
tail -2 gpurun_out/plain_c1.log | cut -c1-600
$CMD > gpurun_out/plain_c1b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -s 30 -c 12 -o gpurun_out/prof_conv_r1a $CMD > gpurun_out/ncu_c1_full.log 2>&1
ls -la gpurun_out
