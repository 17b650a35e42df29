#!/bin/bash
# A/B of differently compiled libsmx builds on one box: [BENCH_ARGS="..."] [ROUNDS=2] scratch/ab.sh tag1=path1 tag2=path2 ...
mkdir -p gpurun_out
ARGS=${BENCH_ARGS:---steps 2 --warmup 2}
SFX=${AB_SUFFIX:-}
for round in $(seq 1 ${ROUNDS:-2}); do
  for kv in "$@"; do
    tag=${kv%%=*}; lib=${kv#*=}
    SMX_LIB_PATH=$lib python bench.py $ARGS --no-cpu-baseline > gpurun_out/ab_${tag}${SFX}_$round.json 2> gpurun_out/ab_${tag}${SFX}_$round.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${tag}${SFX}_$round.json"))
    k=d["kernel_ms_per_step_rank0"]
    print("${tag}${SFX} round $round: e2e %.0f reads/s  value %.0f GCUPS  fwd16_w1 %.1f ms  tb %.1f  chain %.1f scan %.1f  frac %.3f util %s" % (d["e2e"]["reads_per_s"] or 0, d["value"], k["dp_forward16_w1"], k["traceback"], k["chain"], k["scan"], d["roofline"]["frac"], d["clocks"]["gpu_util_pct_mean"]))
except Exception as ex: print("${tag}${SFX} round $round: failed", ex)
PY
  done
done
