#!/bin/bash
for cfg in "6 16" "4 32" "3 40" "5 24"; do
  set -- $cfg
  SMX_TRACE_BUDGET_GB=$2 python bench.py --steps 2 --warmup 2 --overlap $1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1])
print('overlap $1 budget $2 GiB: e2e %.0f reads/s ms/step %.0f util %.0f' % (d['e2e']['reads_per_s'], d['ms_per_step'], d['clocks']['gpu_util_pct_mean']))"
done
