#!/bin/bash
# forward-DP phases admitted at a time x batches in flight, C2, one box
for cfg in "3 6" "2 6" "4 6" "6 6" "3 8" "4 8"; do
  set -- $cfg
  SMX_DP_INFLIGHT=$1 python bench.py --steps 2 --warmup 2 --overlap $2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1])
print('inflight $1 overlap $2: e2e %.0f reads/s ms/step %.0f util %.0f' % (d['e2e']['reads_per_s'], d['ms_per_step'], d['clocks']['gpu_util_pct_mean']))"
done
python bench.py --config legacy_dense --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('dense: value %.0f GCUPS e2e %.0f GCUPS %.0f reads/s' % (d['value'], d['e2e']['value'], d['e2e']['reads_per_s']), {k:round(v,1) for k,v in d['kernel_ms_per_step_rank0'].items() if v>0})"
