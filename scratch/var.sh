#!/bin/bash
for v in "$@"; do
SMX_LIB_PATH=scratch/var/$v.so python bench.py --reads 6000 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step_rank0']; print('$v', 'fwd16h<1>', round(k['dp_forward16_w1'],1), 'ms per 6000 reads; tcups', round(d['roofline']['kernel_tcups'],3), 'e2e', round(d['ms_per_step']))"
done
