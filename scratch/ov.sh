#!/bin/bash
run() { # label, env, args
  echo "== $1"; env $2 python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline $3 2>/dev/null | tail -1 > /tmp/line.json
  python - <<'PY'
import json
d=json.loads(open('/tmp/line.json').read())
k=d["kernel_ms_per_step_rank0"]
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "e2e GCUPS", round(d["e2e"]["value"]), "dev ms", round(d["device_ms_per_step"]), "util", d["clocks"].get("gpu_util_pct_mean"), "W", d["clocks"].get("power_w_mean"))
print("   w2", round(k["dp_forward16_w2"]), "w1", round(k["dp_forward16_w1"]), "k4", round(k["dp_forward_k4"]), "k2", round(k["dp_forward_k2"]), "k1", round(k["dp_forward_k1"]), "tb", round(k["traceback"]))
PY
}
run "default" "A=1" ""
run "minrows 64" "SMX_S16_MINROWS=64" ""
run "minrows 32" "SMX_S16_MINROWS=32" ""
run "maxw 4" "SMX_S16_MAXW=4" ""
run "overlap 5" "A=1" "--overlap 5"
