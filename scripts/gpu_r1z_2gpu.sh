#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/check_zshard_nccl.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r1z_bench_c2_2gpu.json 2> gpurun_out/r1z_bench_2gpu.err; echo rc=$?
cut -c1-330 gpurun_out/r1z_bench_c2_2gpu.json
timeout 300 python bench.py --batch 8 --steps 2 --warmup 2 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('1 GPU batch 8:', d['value'], d['e2e']['value'])"
