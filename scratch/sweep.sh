#!/bin/bash
run() { # label env args
  env $2 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $3 2>/dev/null | tail -1 > /tmp/line.json
  python - "$1" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
print(sys.argv[1],":",round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; e2e GCUPS", round(d["e2e"]["value"]), "value", round(d["value"]), "util", round(d["clocks"].get("gpu_util_pct_mean") or 0))
PY
}
run "ramp" "A=1" ""
run "no ramp" "SMX_RAMP=0" ""
run "ramp" "A=1" ""
run "no ramp" "SMX_RAMP=0" ""
