for c in "2 41 300 32 2 0" "128 256 256 16 1 1" "16 512 512 16 1 1" "64 512 512 32 2 0" "1 8 100 16 1 0"; do timeout 120 python scripts/blockdiag.py $c 2>&1 | grep -v "row " | tail -12; done
python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x 2>&1 | tail -3
run() { env "$@" python bench.py --batch 4 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); l = j['roofline']['layers_ms']
print('$*', {k: v for k, v in l.items() if 'block' in k or k in ('conv:down0.c1:march', 'conv:down0.c2:march', 'conv:up2.c1:march', 'conv:up2.c2:march', 'conv:down1.c2:march')}, 'fwd', round(j['roofline']['forward_ms'], 2))"; }
run CETPICK_BLOCK=0
run CETPICK_BLOCK=1
run CETPICK_BLOCK=1 CETPICK_BLOCK_RS=5
run CETPICK_BLOCK=1 CETPICK_BLOCK_RS=12
