#!/bin/bash
# bench.py on every BASELINE.json config that fits one GPU (own runs for profiles/)
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py "$@" > gpurun_out/cfg_$tag.json 2> gpurun_out/cfg_$tag.err; echo "== $tag rc=$?"; tail -c 300 gpurun_out/cfg_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/cfg_$tag.json"))
    e=d["e2e"]; r=d["roofline"]
    print("$tag: value %.1f GCUPS e2e %.1f GCUPS ms/step %.0f units/s %s frac %.3f dom %s" % (d["value"], e["value"], d["ms_per_step"], {k:v for k,v in e.items() if k.endswith("_per_s")}, r["frac"] or 0, r["kernel"][:40]))
    print("   cell share:", {k[:28]: round(v,4) for k,v in r["cell_share_by_kernel_class"].items()})
    print("   cpu:", d["cpu_baseline"] and {k: d["cpu_baseline"][k] for k in ("value","units_per_s","wall_s","cores")})
except Exception as ex: print("$tag: no json", ex)
PY
}
run c2q --steps 1 --warmup 1 --reads 6000 --cpu-reads 32
run c1 --config c1 --steps 2 --warmup 1 --cpu-reads 64
run c3 --config c3 --steps 2 --warmup 1 --cpu-reads 200
run c4 --config c4 --steps 1 --warmup 1 --reads 6000 --cpu-reads 32
run dense --config legacy_dense --steps 1 --warmup 1 --cpu-reads 16
