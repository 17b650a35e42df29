#!/bin/bash
for gb in 12 24 44; do
export SMX_TRACE_BUDGET_GB=$gb
python bench.py --reads 4000 --steps 2 --warmup 1 --batch 1000 --overlap 2 --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
python - "$gb" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
k=d["kernel_ms_per_step_rank0"]
print("budget GB",sys.argv[1], round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "dp_all", round(k["dp_all"]), "tb", round(k["traceback"]))
PY
done
