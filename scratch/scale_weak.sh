#!/bin/bash
# weak-scaling line on N GPUs: scratch/scale_weak.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 2> gpurun_out/scale_weak_n${N}_r02.err | grep '^{' > gpurun_out/scale_weak_n${N}_r02.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_weak_n${N}_r02.json")); e=d["e2e"]
    print("N=$N weak: value %.0f e2e %.0f GCUPS %.0f reads/s ms/step %.0f threads/rank %s cores %s util %s" % (d["value"], e["value"], e.get("reads_per_s") or 0, d["ms_per_step"], d["config"]["host_threads_per_rank"], d["config"]["host_cores"], d["clocks"]["gpu_util_pct_mean"]))
except Exception as ex: print("N=$N weak failed", ex); print(open("gpurun_out/scale_weak_n${N}_r02.err").read()[-1500:])
PY
