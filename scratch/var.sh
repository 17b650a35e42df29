#!/bin/bash
for v in "$@"; do
SMX_LIB_PATH=scratch/var/$v.so python bench.py --reads 4000 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step_rank0']; print('$v', 'traceback', round(k['traceback'],1), 'ms per 4000 reads; e2e', round(d['ms_per_step']))"
done
