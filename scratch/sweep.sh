#!/bin/bash
run() { # label env args
  env $2 python bench.py --reads 8000 --steps 3 --warmup 2 --no-cpu-baseline $3 2>/dev/null | tail -1 > /tmp/line.json
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
print(sys.argv[1],":",round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; value", round(d["value"]), "util", round(d["clocks"].get("gpu_util_pct_mean") or 0))
PY
}
uptime
run "blocking" "A=1" ""
run "spin" "SMX_SPIN_SYNC=1" ""
run "blocking" "A=1" ""
run "spin" "SMX_SPIN_SYNC=1" ""
uptime
