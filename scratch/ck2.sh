#!/bin/bash
# checkpointed traceback: full GPU suite, then env sweep on the default bench
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ck2_$tag.json 2> gpurun_out/ck2_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ck2_$tag.json"))
    k=d["kernel_ms_per_step_rank0"]
    print("$tag: e2e %.0f reads/s  value %.0f GCUPS  fwd16_w1 %.1f ms  tb %.1f  util %s" % (d["e2e"]["reads_per_s"] or 0, d["value"], k["dp_forward16_w1"], k["traceback"], d["clocks"]["gpu_util_pct_mean"]))
except Exception as ex: print("$tag: failed", ex)
PY
}
run side1_a SMX_CK_SIDE=1
run side0_a SMX_CK_SIDE=0
run rows513 SMX_CKPT_MINROWS=513
run rows769_cols1000 SMX_CKPT_MINROWS=769 SMX_CKPT_MINCOLS=1000
run cols2500 SMX_CKPT_MINCOLS=2500
run cols1000 SMX_CKPT_MINCOLS=1000
run side1_b SMX_CK_SIDE=1
run side0_b SMX_CK_SIDE=0
run off SMX_CKPT=0
