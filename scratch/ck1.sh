#!/bin/bash
# checkpointed traceback: parity tests, then A/B of SMX_CKPT=0 / 1 on the default bench (same library)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_kernel_variants.py -m gpu -x -q -k "checkpointed" 2>&1 | tail -15
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "tests failed rc=$rc"; exit 1; fi
timeout 900 python -m pytest tests/test_pipeline_parity.py tests/test_dp_parity.py tests/test_fullsize_properties.py -m gpu -x -q 2>&1 | tail -3
for round in 1 2; do
for ck in 0 1; do
  SMX_CKPT=$ck timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ck_${ck}_$round.json 2> gpurun_out/ck_${ck}_$round.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ck_${ck}_$round.json"))
    k=d["kernel_ms_per_step_rank0"]
    print("ckpt=$ck round $round: e2e %.0f reads/s  value %.0f GCUPS  fwd16_w1 %.1f ms  tb %.1f  chain %.1f scan %.1f  frac %.3f util %s" % (d["e2e"]["reads_per_s"] or 0, d["value"], k["dp_forward16_w1"], k["traceback"], k["chain"], k["scan"], d["roofline"]["frac"], d["clocks"]["gpu_util_pct_mean"]))
except Exception as ex: print("ckpt=$ck round $round: failed", ex)
PY
done
done
