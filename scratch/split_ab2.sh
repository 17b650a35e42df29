#!/bin/bash
for v in 1 0 1 0; do
SMX_S16H_SPLIT=$v python bench.py --reads 10000 --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step_rank0']; print('split=$v', 'e2e', round(d['e2e']['reads_per_s']), 'reads/s', round(d['ms_per_step']), 'ms/step util', round(d['clocks']['gpu_util_pct_mean']), 'fwd serial', round(k['dp_forward16_w1'],1))"
done
