#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r1k_bench_2gpu.json 2> gpurun_out/r1k_bench_2gpu.err; echo rc=$?
tail -5 gpurun_out/r1k_bench_2gpu.err; cut -c1-700 gpurun_out/r1k_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r1k_ref_2gpu.json 2> gpurun_out/r1k_ref_2gpu.err; echo rc=$?
cut -c1-600 gpurun_out/r1k_ref_2gpu.json
