#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 0 1; do
SMX_S16H_SPLIT=$v python bench.py --reads 6000 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step_rank0']; print('split=$v', 'fwd16h<1>', round(k['dp_forward16_w1'],1), 'ms per 6000 reads; tcups', round(d['roofline']['kernel_tcups'],3), 'e2e', round(d['e2e']['reads_per_s']), 'tb', round(k['traceback'],1))"
done
