#!/bin/bash
# thresholds of the team classes (cells from which a problem gets 2 / 4 / 8 warps) on C2, one box
for cfg in "16 48 128" "8 32 96" "24 64 160" "12 48 128" "16 32 96"; do
  set -- $cfg
  SMX_TEAM2_MCELLS=$1 SMX_TEAM4_MCELLS=$2 SMX_TEAM8_MCELLS=$3 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); k=d['kernel_ms_per_step_rank0']
print('teams $cfg: e2e %.0f reads/s value %.0f fwd %.1f ms frac %.3f' % (d['e2e']['reads_per_s'], d['value'], k['dp_forward16_w1'], d['roofline']['frac']))"
done
python bench.py --config legacy_dense --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('dense: value %.0f GCUPS e2e %.0f GCUPS %.0f reads/s' % (d['value'], d['e2e']['value'], d['e2e']['reads_per_s']), {k:round(v,1) for k,v in d['kernel_ms_per_step_rank0'].items() if v>0})"
