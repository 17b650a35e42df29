#!/bin/bash
# final numbers with the checkpointed traceback: smoke, GPU suite, randomized campaign, default bench (both arms), every config
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/final3_pytest.log 2>&1; tail -2 gpurun_out/final3_pytest.log
timeout 400 python scratch/fuzz_dp.py 9000 18 > gpurun_out/fuzz_ck_r02.log 2>&1; tail -1 gpurun_out/fuzz_ck_r02.log
run() { tag=$1; shift; python bench.py "$@" 2> gpurun_out/bench_${tag}_r02.err | grep '^{' > gpurun_out/bench_${tag}_r02.json; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${tag}_r02.json"))
    e=d["e2e"]; r=d.get("roofline") or {}
    print("$tag: value %.1f GCUPS e2e %.1f GCUPS ms/step %.0f %s frac %s one-pass %s cpu %s" % (d["value"], e["value"], d["ms_per_step"], {k:round(v) for k,v in e.items() if k.endswith("_per_s") and v}, r.get("frac"), (r.get("one_pass_equivalent") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value")))
except Exception as ex: print("$tag: no json", ex)
PY
}
run c2
run c2_reference --impl reference
run c1 --config c1 --steps 3 --warmup 2
run c3 --config c3 --steps 3 --warmup 2
run c4 --config c4 --steps 2 --warmup 1
run dense --config legacy_dense --steps 2 --warmup 1 --cpu-reads 32
