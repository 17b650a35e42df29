#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py "$@" 2> gpurun_out/bench_${tag}_r02.err | grep '^{' > gpurun_out/bench_${tag}_r02.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_r02.json"))
    e=d["e2e"]; r=d.get("roofline") or {}
    print("$tag: value %.1f GCUPS e2e %.1f GCUPS ms/step %.0f %s frac %s cpu %s" % (d["value"], e["value"], d["ms_per_step"], {k:round(v) for k,v in e.items() if k.endswith("_per_s") and v}, r.get("frac"), (d.get("cpu_baseline") or {}).get("value")))
except Exception as ex: print("$tag: no json", ex)
PY
}
run c1 --config c1 --steps 3 --warmup 2
run dense --config legacy_dense --steps 2 --warmup 1 --cpu-reads 32
run c2 
