#!/bin/bash
# repeat the GPU test-suite and the slab stress: intermittent protocol bugs show up here, not in a single pass
N=${1:-3}
for i in $(seq 1 $N); do timeout 1300 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -1; done
for i in $(seq 1 $N); do timeout 200 python scripts/stress_slab.py 40; done
