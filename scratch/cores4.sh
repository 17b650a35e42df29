#!/bin/bash
# one GPU confined to four host cores (the per-GPU share of the 8-GPU box): host threads per handle, handles in flight
for cfg in "2 6" "1 6" "3 6" "4 6" "2 4" "2 8"; do
  set -- $cfg
  taskset -c 0-3 python bench.py --steps 2 --warmup 2 --threads $1 --overlap $2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1])
print('4 cores, threads $1 overlap $2: e2e %.0f reads/s ms/step %.0f util %.0f' % (d['e2e']['reads_per_s'], d['ms_per_step'], d['clocks']['gpu_util_pct_mean']))"
done
SMX_PHASE_LOG=1 taskset -c 0-3 python bench.py --steps 1 --warmup 1 --threads 2 --reads 6000 --no-cpu-baseline 2> gpurun_out/phase_4cores_r02.log > /dev/null
grep -c . gpurun_out/phase_4cores_r02.log
