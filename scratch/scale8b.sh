#!/bin/bash
N=8
for cfg in "8 2" "8 1" "6 1"; do
  set -- $cfg
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 --overlap $1 --threads $2 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N=8 overlap $1 threads $2: e2e %.0f reads/s ms/step %.0f util %s' % (e['reads_per_s'], d['ms_per_step'], d['clocks']['gpu_util_pct_mean']))"
done
