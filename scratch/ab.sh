#!/bin/bash
# A/B of differently compiled libsmx builds on one box: scratch/ab.sh tag1=path1 tag2=path2 ... (two rounds each)
mkdir -p gpurun_out
for round in 1 2; do
  for kv in "$@"; do
    tag=${kv%%=*}; lib=${kv#*=}
    SMX_LIB_PATH=$lib python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ab_${tag}_$round.json 2> gpurun_out/ab_${tag}_$round.err
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_${tag}_$round.json"))
k=d["kernel_ms_per_step_rank0"]
print("${tag} round $round: e2e %.0f reads/s  value %.0f GCUPS  fwd16_w1 %.1f ms  tb %.1f  chain %.1f scan %.1f  frac %.3f" % (d["e2e"]["reads_per_s"], d["value"], k["dp_forward16_w1"], k["traceback"], k["chain"], k["scan"], d["roofline"]["frac"]))
PY
  done
done
