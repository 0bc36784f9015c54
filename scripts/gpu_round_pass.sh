#!/bin/bash
# Round pass: gpu tests, configs[1] bench line (+ reference arm), configs[2] decode timing, launch lists and
# bounded full ncu captures; text summaries land in gpurun_out/<tag>_* (copy the ones to keep into profiles/).
TAG=${1:-rX}
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c2_batch64.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${TAG}_bench_c2_batch64.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench_reference_arm.json
for k in tiefree peaks; do
  timeout 300 python scripts/bench_decode.py --kind $k > gpurun_out/${TAG}_decode_c3_$k.json 2>> gpurun_out/${TAG}_decode.err
  cat gpurun_out/${TAG}_decode_c3_$k.json
done
CMD="python scripts/bench_decode.py --kind peaks --iters 2 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|sieve|cand_|rank_|init_state' -c 200 --csv --log-file gpurun_out/${TAG}_decode_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/${TAG}_decode_launches.csv | tee gpurun_out/${TAG}_decode_launches.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sieve_kernel' -s 1 -c 1 -o gpurun_out/${TAG}_sieve $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
python scripts/ncu_summary.py full gpurun_out/${TAG}_sieve.ncu-rep | tee gpurun_out/${TAG}_sieve_full.txt
ncu -i gpurun_out/${TAG}_sieve.ncu-rep --page details 2>/dev/null | grep -v "^\s*$" >> gpurun_out/${TAG}_sieve_full.txt
CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/${TAG}_plain_b1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_march|conv_halo|conv_up|stem|pool' -s 46 -c 23 -o gpurun_out/${TAG}_conv $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
python scripts/ncu_summary.py full gpurun_out/${TAG}_conv.ncu-rep | tee gpurun_out/${TAG}_conv_full.txt
ncu -i gpurun_out/${TAG}_conv.ncu-rep --page details 2>/dev/null | grep -v "^\s*$" > gpurun_out/${TAG}_conv_details.txt
rm -f gpurun_out/${TAG}_conv.ncu-rep gpurun_out/${TAG}_launches.csv
ls -la gpurun_out | tail -20
