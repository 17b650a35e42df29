"""Generates tests/golden/golden_v1.json.gz by running the UNMODIFIED reference Python
(/root/reference, through oracle/ref_harness.py; parasail shimmed by oracle/parasail_port.c).

Run here (the reference cannot travel):   python -m oracle.make_golden
The fixture holds inputs (plasmid maps, reads, params) and the reference's outputs per
pair-strand in the order of post_analysis_core.execute_alignment_core_loop (:68-87):
(cigar RLE as in MyResult.to_dict :512-518, score, new_query_seq_offset), plus pre_survey
distances (pre_survey_core.py:251-271) and the DP call list (mode, len_query, len_ref).
"""
from __future__ import annotations

import glob
import gzip
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from multiplexnanopore_b200 import synth  # noqa: E402  (generators only; no engine involved)
from oracle import oracle as orc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = ROOT / "tests" / "golden" / "golden_v1.json.gz"


def demo_maps():
    out = {}
    for p in sorted(glob.glob(str(rh.REF_ROOT / "resources/demo_data/my_plasmid_maps_fa/*.fa"))):
        out[Path(p).name] = "".join(l.strip() for l in open(p) if l[0] != ">").upper()
    return out


def run_case(ns, name, refs, reads, params, omit_too_long=True):
    t0 = time.time()
    log = []
    rh.load(log)
    aligners = [ns.rqa.MyOptimizedAligner(ns.mc.MySeq(r), params) for r in refs]
    results = []
    for read in reads:
        res = rh.reference_align_read_with(ns, aligners, read, omit_too_long)
        results.append([[orc.rle(c), int(s), None if o is None else int(o)] for c, s, o in res])
    print(f"{name}: {len(reads)} reads x {len(refs)} refs in {time.time() - t0:.1f}s, {len(log)} DP calls", flush=True)
    return {"name": name, "params": params, "refs": refs, "reads": reads, "omit_too_long": omit_too_long,
            "results": results, "dp_calls": [list(map(int, c)) for c in log]}


def main():
    ns = rh.load()
    assert ns is not None, "reference tree or oracle/_ref missing: run `make -C oracle` first"
    base = dict(rh.DEFAULT_PARAMS)
    cases = []
    maps = demo_maps()
    demo = list(maps.values())

    rng = np.random.default_rng(1)
    reads, _ = synth.reads_for(rng, demo, 6, topology=0)
    cases.append(run_case(ns, "c1_demo_circular", demo, reads, dict(base)))

    rng = np.random.default_rng(2)
    small = synth.plasmid_set(rng, 6, 1500, 3000, backbone=800)
    reads, _ = synth.reads_for(rng, small, 14, topology=0, frac_full=0.6, frac_partial=0.25)
    cases.append(run_case(ns, "c2_small_circular", small, reads, dict(base)))
    p2 = dict(base, gap_open_penalty=5, gap_extend_penalty=2, match_score=2, mismatch_score=-3)
    cases.append(run_case(ns, "c2_small_circular_scoring2", small, reads[:5], p2))

    rng = np.random.default_rng(5)
    reads, _ = synth.reads_for(rng, small, 8, topology=0, homopolymer_rate=3.0)
    cases.append(run_case(ns, "c2_small_circular_homopolymers", small, reads, dict(base)))

    rng = np.random.default_rng(4)
    lin = synth.plasmid_set(rng, 4, 1500, 4000, backbone=700)
    reads, _ = synth.reads_for(rng, lin, 10, topology=1)
    cases.append(run_case(ns, "c4_small_linear", lin, reads, dict(base, topology_of_dna=1)))

    # the seeds of tests/test_pipeline_parity.py::test_engine_matches_port_on_fresh_inputs: seed 31 holds a
    # my_special_DP overlap merge whose arg-max is a tie (first maximum wins, ref_query_alignment.py:136)
    for topo, seed in ((0, 31), (0, 32), (1, 33)):
        rng = np.random.default_rng(seed)
        pl = synth.plasmid_set(rng, 4, 1200, 2600, backbone=600)
        rd, _ = synth.reads_for(rng, pl, 10, topology=topo, frac_full=0.6, frac_partial=0.25, homopolymer_rate=1.0)
        cases.append(run_case(ns, f"fresh_{seed}_topology{topo}", pl, rd, dict(base, topology_of_dna=topo)))

    # round 2: wider goldens -- linear topology on the demo maps, full-size synthetic plasmids, IUPAC codes in the plasmid
    # maps (allowed since ver_0.3.1: wildcards score 0 inside parasail but are rescored as = / X), fragment-rich reads,
    # and a third scoring
    rng = np.random.default_rng(11)
    reads, _ = synth.reads_for(rng, demo, 8, topology=1)
    cases.append(run_case(ns, "c1_demo_linear", demo, reads, dict(base, topology_of_dna=1)))

    rng = np.random.default_rng(12)
    full = synth.plasmid_set(rng, 4, 5000, 9000, backbone=2500)
    reads, _ = synth.reads_for(rng, full, 10, topology=0)
    cases.append(run_case(ns, "c2_fullsize_circular", full, reads, dict(base)))

    rng = np.random.default_rng(13)
    iupac = []
    for p in synth.plasmid_set(rng, 4, 1500, 2600, backbone=700):
        a = np.frombuffer(p.encode(), dtype=np.uint8).copy()
        pos = rng.choice(a.size, size=12, replace=False)
        a[pos] = np.frombuffer(b"NRYKMSWBDHVN", dtype=np.uint8)
        iupac.append(a.tobytes().decode())
    reads, _ = synth.reads_for(rng, iupac, 12, topology=0, frac_full=0.6, frac_partial=0.3)
    reads = [r.replace("N", "A").replace("R", "G").replace("Y", "C").replace("K", "T").replace("M", "A").replace("S", "C").replace("W", "T")
             .replace("B", "C").replace("D", "G").replace("H", "A").replace("V", "G") for r in reads]   # reads are plain ACGT
    reads[0] = reads[0][:200] + "N" * 3 + reads[0][203:]                                               # ... but one carries N calls
    cases.append(run_case(ns, "iupac_maps_circular", iupac, reads, dict(base)))

    rng = np.random.default_rng(14)
    reads, _ = synth.reads_for(rng, small, 16, topology=0, frac_full=0.2, frac_partial=0.7)
    cases.append(run_case(ns, "c2_small_fragments", small, reads, dict(base)))
    p3 = dict(base, gap_open_penalty=2, gap_extend_penalty=2, match_score=1, mismatch_score=-1)
    cases.append(run_case(ns, "c2_small_circular_scoring3", small, reads[:6], p3))

    rng = np.random.default_rng(15)
    reads, _ = synth.reads_for(rng, lin, 12, topology=1, homopolymer_rate=2.0)
    cases.append(run_case(ns, "c4_small_linear_homopolymers", lin, reads, dict(base, topology_of_dna=1)))

    # pre_survey distances: demo maps (7 of the 15 pairs are the legacy known answers) + a synthetic family
    dist = {"demo_names": list(maps.keys()), "demo": {}, "family_maps": [], "family": {}}
    t0 = time.time()
    for i in range(len(demo)):
        for j in range(i + 1, len(demo)):
            dist["demo"][f"{i},{j}"] = int(rh.reference_distance(ns, demo[i], demo[j], dict(base)))
    print(f"demo distances in {time.time() - t0:.1f}s: {dist['demo']}", flush=True)
    fam = synth.config_c3(seed=3, n_maps=6, families=2, len_lo=1500, len_hi=3000)
    dist["family_maps"] = fam
    for i in range(len(fam)):
        for j in range(i + 1, len(fam)):
            dist["family"][f"{i},{j}"] = int(rh.reference_distance(ns, fam[i], fam[j], dict(base)))
    print(f"family distances: {dist['family']}", flush=True)

    OUT.parent.mkdir(parents=True, exist_ok=True)
    with gzip.open(OUT, "wt") as f:
        json.dump({"version": 1, "generator": "oracle/make_golden.py", "demo_maps": maps, "cases": cases, "distances": dist}, f)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
