#!/bin/bash
# hardware work queues: 6 handles x (DP stream, high-priority stream, 3 team streams) share CUDA_DEVICE_MAX_CONNECTIONS queues (default 8)
for c in 8 32 16 32 8; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1])
print('connections $c: e2e %.0f reads/s ms/step %.0f util %.0f' % (d['e2e']['reads_per_s'], d['ms_per_step'], d['clocks']['gpu_util_pct_mean']))"
done
