#!/bin/bash
run() { # label, env, args
  echo "== $1"; env $2 python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline $3 2>/dev/null | tail -1 > /tmp/line.json
  python - <<'PY'
import json
d=json.loads(open('/tmp/line.json').read())
k=d["kernel_ms_per_step_rank0"]
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "e2e GCUPS", round(d["e2e"]["value"]), "dev ms", round(d["device_ms_per_step"]), "util", round(d["clocks"].get("gpu_util_pct_mean") or 0))
print("   w2", round(k["dp_forward16_w2"]), "w1", round(k["dp_forward16_w1"]), "k1", round(k["dp_forward_k1"]), "tb", round(k["traceback"]))
PY
}
run "team min 0" "SMX_TEAM_MIN_MCELLS=0" ""
run "team min 8" "SMX_TEAM_MIN_MCELLS=8" ""
run "team min 16" "SMX_TEAM_MIN_MCELLS=16" ""
run "team min 32" "SMX_TEAM_MIN_MCELLS=32" ""
run "team min 1000 (W=1 only)" "SMX_TEAM_MIN_MCELLS=1000" ""
