#!/bin/bash
# one GPU confined to 4 cores (the 8-GPU box has 32): handles in flight x host threads per handle
for o in 4 6 8; do for t in 1 2 3; do
  taskset -c 0-3 python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline --overlap $o --threads $t 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('overlap $o threads $t', round(d['e2e']['reads_per_s']), 'reads/s', 'util', round(d['clocks']['gpu_util_pct_mean']))"
done; done
