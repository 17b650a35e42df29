#!/bin/bash
# e2e rate of one GPU when the process is confined to 16 / 4 / 3 / 2 host cores
for c in 0-15 0-3 0-2 0-1; do
  taskset -c $c python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cores $c', round(d['e2e']['reads_per_s']), 'reads/s', 'util', round(d['clocks']['gpu_util_pct_mean']), 'threads', d['config']['host_threads_per_rank'])"
done
