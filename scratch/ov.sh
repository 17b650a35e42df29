#!/bin/bash
run() { # label, env, args
  echo "== $1"; env $2 python bench.py --reads 8000 --steps 2 --warmup 2 --no-cpu-baseline $3 2>/dev/null | tail -1 > /tmp/line.json
  python - <<'PY'
import json
d=json.loads(open('/tmp/line.json').read())
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "e2e GCUPS", round(d["e2e"]["value"]), "dev ms", round(d["device_ms_per_step"]))
PY
}
run "overlap4 batch1000 24G" "SMX_TRACE_BUDGET_GB=24" "--overlap 4 --batch 1000"
run "overlap4 batch1000 36G" "SMX_TRACE_BUDGET_GB=36" "--overlap 4 --batch 1000"
run "overlap6 batch500 16G" "SMX_TRACE_BUDGET_GB=16" "--overlap 6 --batch 500"
run "overlap6 batch500 20G" "SMX_TRACE_BUDGET_GB=20" "--overlap 6 --batch 500"
run "overlap8 batch250 10G" "SMX_TRACE_BUDGET_GB=10" "--overlap 8 --batch 250"
run "overlap4 batch500 20G thr8" "SMX_TRACE_BUDGET_GB=20" "--overlap 4 --batch 500 --threads 8"
run "overlap4 batch500 20G thr4" "SMX_TRACE_BUDGET_GB=20" "--overlap 4 --batch 500 --threads 4"
