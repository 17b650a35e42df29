#!/bin/bash
run() { # gate overlap budget
  env SMX_DP_INFLIGHT=$1 SMX_TRACE_BUDGET_GB=$3 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --overlap $2 2>/dev/null | tail -1 > /tmp/line.json
  python - "$1" "$2" "$3" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
print("gate",sys.argv[1],"overlap",sys.argv[2],"budget",sys.argv[3],":",round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; e2e GCUPS", round(d["e2e"]["value"]), "util", round(d["clocks"].get("gpu_util_pct_mean") or 0))
PY
}
run 1 4 24
run 1 4 24
run 1 5 20
run 2 6 16
run 2 8 12
