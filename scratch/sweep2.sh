#!/bin/bash
for mw in 1 2 8; do for ov in 1 2 3; do
SMX_S16_MAXW=$mw python bench.py --reads 4000 --steps 2 --warmup 1 --batch 1000 --overlap $ov --no-cpu-baseline 2>&1 | tail -1 > /tmp/line.json
python - "$mw" "$ov" <<'PY'
import json,sys
d=json.loads(open('/tmp/line.json').read())
k=d["kernel_ms_per_step_rank0"]
print("maxw",sys.argv[1],"overlap",sys.argv[2], round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; dev GCUPS", round(d["value"]), "dp_all", round(k["dp_all"]), "s16", round(k["dp_forward16_w8"]+k["dp_forward16_w4"]+k["dp_forward16_w2"]+k["dp_forward16_w1"]))
PY
done; done
