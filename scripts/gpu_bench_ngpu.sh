#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2b_bench_c2_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err; echo rc=$?
tail -3 gpurun_out/r2b_bench_${N}gpu.err | cut -c1-300
cut -c1-330 gpurun_out/r2b_bench_c2_${N}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/check_zshard_nccl.py 2>&1 | grep "z-sharded"
