#!/bin/bash
for t in 2 4 8 16 32; do
SMX_CHAIN_CTAS=$t python bench.py --reads 6000 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step_rank0']; print('target $t', 'chain', round(k['chain'],1), 'ms per 6000 reads; e2e', round(d['ms_per_step']))"
done
