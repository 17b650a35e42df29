#!/bin/bash
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 300 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
print(round(d["ms_per_step"]), "ms/step", round(d["e2e"]["reads_per_s"]), "reads/s; value GCUPS", round(d["value"]), "e2e GCUPS", round(d["e2e"]["value"]))
print({k:round(v,1) for k,v in d["kernel_ms_per_step_rank0"].items()})
r=d["roofline"]; print("roofline", r["kernel"][:40], r["kernel_tcups"], r["frac"], r["dpx_view"]["frac"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
PY
