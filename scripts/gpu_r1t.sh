#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for n in A B C D; do
  if [ $n = A ]; then unset CETPICK_LIB; else export CETPICK_LIB=$PWD/cet_pick_b200/alt/lib$n.so; fi
  for k in tiefree peaks; do
    echo -n "variant $n $k: "
    timeout 300 python scripts/bench_decode.py --kind $k | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_median'], d['ms_min'], d['n_candidates'])"
  done
done
done
