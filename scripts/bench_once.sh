#!/bin/bash
# one plain bench.py run with its wall time: scripts/bench_once.sh <tag>
TAG=${1:-rX}; mkdir -p gpurun_out
T0=$(date +%s)
python bench.py > gpurun_out/${TAG}_bench_c2_batch64.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$? wall=$(( $(date +%s) - T0 )) s"
python - <<PY
import json
j=json.loads(open("gpurun_out/${TAG}_bench_c2_batch64.json").read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","value_w1","ms_per_step","gpu_launches")}, "e2e", j["e2e"]["value"], "e2e_run", j["e2e_run"].get("value"), j["e2e_run"].get("tomograms"))
print("frac", j["roofline"]["frac"], "frac_step", j["roofline"]["frac_step"], "fwd", j["roofline"]["forward_ms"], "decode", j["roofline_decode"]["ms"], j["roofline_decode"]["frac"], "train", j["train_config5"].get("ours_ms_per_crop"), j["clocks"])
PY
