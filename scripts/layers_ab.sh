#!/bin/bash
# quick A/B: conv + unet parity tests, then the per-layer table of a 4-tomogram bench run
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_config_sizes.py -x -q 2>&1 | tail -3
python bench.py --batch 4 --steps 2 --warmup 1 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('value',j['value'],'w1',j['value_w1'],'e2e',j['e2e']['value'],'fwd',r['forward_ms'],'frac',r['frac'])
print(r['layers_ms'])"
