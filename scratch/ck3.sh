#!/bin/bash
# checkpointed traceback at its final compile-time settings: parity tests, then env sweep (team thresholds, CK thresholds)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_kernel_variants.py tests/test_pipeline_parity.py -m gpu -x -q -k "checkpointed or pipeline or golden" 2>&1 | tail -2
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ck3_$tag.json 2> gpurun_out/ck3_$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ck3_$tag.json"))
    k=d["kernel_ms_per_step_rank0"]
    print("$tag: e2e %.0f reads/s  value %.0f GCUPS  fwd16_w1 %.1f ms  tb %.1f  util %s" % (d["e2e"]["reads_per_s"] or 0, d["value"], k["dp_forward16_w1"], k["traceback"], d["clocks"]["gpu_util_pct_mean"]))
except Exception as ex: print("$tag: failed", ex)
PY
}
run base_a A=1
run teams_32_96_256 SMX_TEAM2_MCELLS=32 SMX_TEAM4_MCELLS=96 SMX_TEAM8_MCELLS=256
run teams_24_64_192 SMX_TEAM2_MCELLS=24 SMX_TEAM4_MCELLS=64 SMX_TEAM8_MCELLS=192
run noteams SMX_S16_MAXW=1
run rows513 SMX_CKPT_MINROWS=513
run rows769 SMX_CKPT_MINROWS=769
run inflight4 SMX_DP_INFLIGHT=4
run inflight2 SMX_DP_INFLIGHT=2
run base_b A=1
