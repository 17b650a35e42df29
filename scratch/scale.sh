#!/bin/bash
# scaling runs on N GPUs of one box: scratch/scale.sh N [c5]
N=$1
mkdir -p gpurun_out
run() { tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" 2> gpurun_out/scale_${tag}_n${N}_r02.err | grep '^{' > gpurun_out/scale_${tag}_n${N}_r02.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${tag}_n${N}_r02.json")); e=d["e2e"]
    print("N=$N $tag: value %.0f e2e %.0f GCUPS %.0f reads/s ms/step %.0f threads/rank %s cores %s util %s" % (d["value"], e["value"], e.get("reads_per_s") or 0, d["ms_per_step"], d["config"]["host_threads_per_rank"], d["config"]["host_cores"], d["clocks"]["gpu_util_pct_mean"]))
except Exception as ex: print("N=$N $tag failed", ex); print(open("gpurun_out/scale_${tag}_n${N}_r02.err").read()[-1500:])
PY
}
run weak --steps 3 --warmup 2
run strong --steps 3 --warmup 2 --scaling strong
if [ "$2" == "c5" ]; then run c5 --config c5 --scaling strong --steps 1 --warmup 1; fi
