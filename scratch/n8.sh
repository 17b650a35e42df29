#!/bin/bash
N=${1:-8}
nproc; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/n$N.json 2> gpurun_out/n$N.err
tail -c 500 gpurun_out/n$N.err
python - $N <<'PY'
import json,sys
for l in open(f'gpurun_out/n{sys.argv[1]}.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d["n_gpus"], round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["config"]["host_threads_per_rank"], d["config"]["host_cores"])
PY
