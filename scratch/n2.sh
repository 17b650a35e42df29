#!/bin/bash
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/n2.json 2> gpurun_out/n2.err
tail -c 600 gpurun_out/n2.err
python - <<'PY'
import json
for l in open('gpurun_out/n2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(d["n_gpus"], round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s", d["config"]["host_threads_per_rank"], d["config"]["host_cores"])
        print({k:round(v,3) for k,v in d["stage_seconds_per_step_rank0"].items()})
        print({k:round(v,3) for k,v in d["stage_seconds_per_step_rank0_overlapped"].items()})
PY
