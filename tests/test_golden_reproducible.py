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
