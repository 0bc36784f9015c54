run() { env "$@" python bench.py --batch 4 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); l = j['roofline']['layers_ms']
print('$*', {k: v for k, v in l.items() if 'block' in k or 'fhead' in k or k in ('conv:down0.c1:march', 'conv:down0.c2:march', 'conv:up2.c1:march', 'conv:up2.c2:march', 'stem')}, 'fwd', round(j['roofline']['forward_ms'], 2))"; }
run A=0
run CETPICK_BLOCK_FLAGS=12
run CETPICK_BLOCK_FLAGS=15
run CETPICK_BLOCK_LAG=6 CETPICK_BLOCK_FLAGS=15
run CETPICK_BLOCK_LAG=6 CETPICK_BLOCK_STAGES=4
run CETPICK_NO_BLOCK=1
