#!/bin/bash
run() { # label env args
  env $2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline $3 2>/dev/null | tail -1 > /tmp/line.json
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
print(sys.argv[1],":",round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; e2e GCUPS", round(d["e2e"]["value"]), "util", round(d["clocks"].get("gpu_util_pct_mean") or 0))
PY
}
run "default (gate2 ov6 nonpersistent)" "A=1" ""
run "no gate, persistent, ov6" "SMX_DP_INFLIGHT=16 SMX_PERSISTENT=1" ""
run "no gate, nonpersistent, ov6" "SMX_DP_INFLIGHT=16" ""
run "gate2 persistent ov6" "SMX_PERSISTENT=1" ""
run "gate1 ov6" "SMX_DP_INFLIGHT=1" ""
run "default ov4" "A=1" "--overlap 4"
run "default ov8 12G" "SMX_TRACE_BUDGET_GB=12" "--overlap 8"
run "default (gate2 ov6 nonpersistent)" "A=1" ""
