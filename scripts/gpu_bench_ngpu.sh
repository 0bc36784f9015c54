#!/bin/bash
# bench.py under torchrun on N GPUs of one box: scripts/gpu_bench_ngpu.sh <tag> <N> [steps] [warmup]
TAG=${1:-rX}; N=${2:-2}; K=${3:-3}; W=${4:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $K --warmup $W --no-cpu-baseline \
  > gpurun_out/${TAG}_bench_c2_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
echo "rc=$?"; python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_bench_c2_${N}gpu.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("n_gpus", "value", "value_w1", "ms_per_step")}, "e2e", j["e2e"]["value"], j["clocks"]); print("train_config5", j.get("train_config5")); print("zshard", j.get("zshard_one_tomogram"))
PY
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
