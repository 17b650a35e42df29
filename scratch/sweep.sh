#!/bin/bash
run() { # label prefix env args
  env $3 $2 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $4 2>/dev/null | tail -1 > /tmp/line.json
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
print(sys.argv[1],":",round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; e2e GCUPS", round(d["e2e"]["value"]), "value", round(d["value"]), "util", round(d["clocks"].get("gpu_util_pct_mean") or 0), "thr", d["config"]["host_threads_per_rank"])
PY
}
run "16 cores ov6" "" "A=1" ""
run "8 cores ov6" "taskset -c 0-7" "A=1" ""
run "8 cores ov8 12G" "taskset -c 0-7" "SMX_TRACE_BUDGET_GB=12" "--overlap 8"
run "4 cores ov6" "taskset -c 0-3" "A=1" ""
run "4 cores ov8 12G" "taskset -c 0-3" "SMX_TRACE_BUDGET_GB=12" "--overlap 8"
run "4 cores ov6 thr1" "taskset -c 0-3" "A=1" "--threads 1"
