"""When the reference tree is present (this container, not the GPU box) the committed goldens must be reproducible:
one case is regenerated through oracle/ref_harness.py (UNMODIFIED reference Python + reference Cython from
oracle/_ref + the parasail shim) and compared with tests/golden/golden_v1.json.gz."""
import pytest

from oracle import oracle as orc
from oracle import ref_harness as rh
from tests import golden_util


@pytest.mark.skipif(not rh.REF_ROOT.is_dir(), reason="/root/reference is not present (GPU box)")
@pytest.mark.parametrize("name,n_reads", [("c2_small_circular", 3), ("c4_small_linear", 3), ("fresh_31_topology0", 2)])
def test_golden_case_regenerates(name, n_reads):
    ns = rh.load()
    if ns is None:
        pytest.skip("oracle/_ref not built (make -C oracle)")
    case = golden_util.case(name)
    aligners = [ns.rqa.MyOptimizedAligner(ns.mc.MySeq(r), case["params"]) for r in case["refs"]]
    for read, want in list(zip(case["reads"], case["results"]))[:n_reads]:
        got = rh.reference_align_read_with(ns, aligners, read, case["omit_too_long"])
        got = [[orc.rle(c), int(s), None if o is None else int(o)] for c, s, o in got]
        assert got == want


@pytest.mark.skipif(not rh.REF_ROOT.is_dir(), reason="/root/reference is not present (GPU box)")
def test_statistics_tables_of_the_reference_load_identically():
    """consensus.SequenceBasecallQscorePDF on the reference's own NanoporeStats_*.csv files == the reference's class on them
    (P_base_calling_given_true_refseq_dict, pdf_core, q_score_distribution, version string), and one consensus case of
    tests/golden/consensus_v1.json.gz regenerates through the reference's classes (msa.py:654-733, 1749-1812)."""
    import gzip
    import json
    import tempfile
    from pathlib import Path

    import numpy as np

    from multiplexnanopore_b200 import consensus as cs
    ns = rh.load()
    if ns is None:
        pytest.skip("oracle/_ref not built (make -C oracle)")
    files = sorted((rh.REF_ROOT / "savemoney" / "modules" / "NanoporeStats_PDF").glob("*.csv"))
    assert files
    for f in files:
        want, got = ns.msa.SequenceBasecallQscorePDF(df_csv_path=f), cs.SequenceBasecallQscorePDF(df_csv_path=f)
        assert got.NanoporeStats_PDF_version == want.NanoporeStats_PDF_version
        assert got.q_score_distribution == [int(x) for x in want.q_score_distribution]
        for key, p in want.P_base_calling_given_true_refseq_dict.items():
            assert float(p) == got.P_base_calling_given_true_refseq_dict[key], (f.name, key)
            assert [float(x) for x in want.pdf_core[key]] == [float(x) for x in got.pdf_core[key]], (f.name, key)
        assert want.calc_js_divergence(want.q_score_distribution[::-1]) == got.calc_js_divergence(got.q_score_distribution[::-1])
    g = json.load(gzip.open(Path(__file__).parent / "golden" / "consensus_v1.json.gz", "rt"))
    case = next(c for c in g["cases"] if c["name"] == "iupac_reference")
    t = g["tables"][case["table"]]
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "t.csv"
        with open(p, "w") as f:
            f.write("# file_version: 9.9.9\n\t" + "\t".join(t["columns"]) + "\n")
            for idx, row in zip(range(-1, 42), t["counts"]):
                f.write(str(idx) + "\t" + "\t".join(str(x) for x in row) + "\n")
        ns.msa.MyMSA.sbq_pdf = ns.msa.SequenceBasecallQscorePDF(df_csv_path=p)
    params = dict(rh.DEFAULT_PARAMS, **case["params"])
    m = ns.msa.MyMSA.generate_msa(ns.mc.MySeq(case["ref"]), list(case["queries"]), [list(s) for s in case["q_scores"]], list(case["cigars"]), params)
    m.calculate_consensus()
    assert [m.with_prior_consensus_seq, [int(x) for x in m.with_prior_consensus_q_scores], m.with_prior_consensus_my_cigar] == case["with_prior"]
    assert [m.without_prior_consensus_seq, [int(x) for x in m.without_prior_consensus_q_scores], m.without_prior_consensus_my_cigar] == case["without_prior"]
    with_prior, without_prior = cs.consensus_params(params)
    rw, rwo = m.consensus_params(params, make_it_for_CYTHON=False)
    assert with_prior == rw and without_prior == rwo
