#!/bin/bash
for v in 768 3000 8000 1000000 768; do
  SMX_TB_BIGWALK=$v python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); k=d['kernel_ms_per_step_rank0']
print('bigwalk $v: e2e %.0f reads/s ms/step %.0f util %.0f tb(serial) %.0f ms' % (d['e2e']['reads_per_s'], d['ms_per_step'], d['clocks']['gpu_util_pct_mean'], k['traceback']))"
done
