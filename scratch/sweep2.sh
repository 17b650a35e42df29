#!/bin/bash
run() { env $2 python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline $3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step']), 'ms/step', round(d['e2e']['reads_per_s']), 'reads/s util', round(d['clocks']['gpu_util_pct_mean']), 'chain', round(d['kernel_ms_per_step_rank0']['chain'],1))"; }
run base A=1 ""
run inflight2 SMX_DP_INFLIGHT=2 ""
run inflight4 SMX_DP_INFLIGHT=4 ""
run inflight6 SMX_DP_INFLIGHT=6 ""
run chain8 SMX_CHAIN_CTAS=8 ""
run chain32 SMX_CHAIN_CTAS=32 ""
run base A=1 ""
