#!/bin/bash
# Round pass: gpu tests, smoke, configs[1] bench line with all legs (+ reference arm); optional probes / ncu captures.
#   scripts/gpu_round_pass.sh <tag> [probe] [ncu]
TAG=${1:-rX}
shift
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -14 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for opt in "$@"; do
  if [ "$opt" = probe ]; then
    timeout 300 python scripts/probe_mma_rate.py > gpurun_out/${TAG}_probe_mma_rate.txt 2>&1; cat gpurun_out/${TAG}_probe_mma_rate.txt
  fi
done
timeout 1200 python bench.py > gpurun_out/${TAG}_bench_c2_batch64.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_c2_batch64.json").read().strip().splitlines()[-1])
    r = j["roofline"]
    print({k: j[k] for k in ("value", "value_w1", "ms_per_step", "gpu_launches")}, "e2e", j["e2e"]["value"])
    print({k: r[k] for k in ("achieved", "frac", "frac_step", "forward_ms", "stem_uint8_input_ms", "conv_ms_all")})
    print(r["layers_ms"])
    for k in ("e2e_run", "simsiam_config3", "roofline_decode", "train_config5", "torch_cuda_baseline", "tf32_mode", "cpu_baseline", "clocks"):
        print(k, j.get(k))
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
cut -c1-300 gpurun_out/${TAG}_bench_reference_arm.json
for opt in "$@"; do
  if [ "$opt" = ncu ]; then
    CMD="python scripts/bench_decode.py --kind tiefree --iters 2 --warmup 1"
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scan_kernel|sieve|cand_|rank_|init_state|tail_' -c 200 --csv --log-file gpurun_out/${TAG}_decode_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
    python scripts/ncu_summary.py launches gpurun_out/${TAG}_decode_launches.csv | tee gpurun_out/${TAG}_decode_launches.txt
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sieve_kernel' -s 1 -c 1 -o gpurun_out/${TAG}_sieve $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
    python scripts/ncu_summary.py full gpurun_out/${TAG}_sieve.ncu-rep | tee gpurun_out/${TAG}_sieve_full.txt
    CMD="python bench.py --batch 1 --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
    timeout 300 $CMD > gpurun_out/${TAG}_plain_b1.log 2>&1 && \
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
    python scripts/ncu_summary.py launches gpurun_out/${TAG}_launches.csv | tee gpurun_out/${TAG}_launches.txt
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_march|conv_halo|conv_up|stem|pool|block_' -s 46 -c 23 -o gpurun_out/${TAG}_conv $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
    python scripts/ncu_summary.py full gpurun_out/${TAG}_conv.ncu-rep | tee gpurun_out/${TAG}_conv_full.txt
    ncu -i gpurun_out/${TAG}_conv.ncu-rep --page details 2>/dev/null | grep -v "^\s*$" > gpurun_out/${TAG}_conv_details.txt
    rm -f gpurun_out/${TAG}_conv.ncu-rep gpurun_out/${TAG}_launches.csv
  fi
done
ls -la gpurun_out | tail -12
