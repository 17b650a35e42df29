"""Drive the UNMODIFIED reference Python (read-only, /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  The reference cannot travel to the GPU box, so this module is used
here to (1) pin the oracle restatements against the reference's own code and (2) generate the
golden fixtures under tests/golden/ (see oracle/make_golden.py).  Nothing executed on the GPU
box may import it (it returns None from `load()` when /root/reference is absent).

What is real and what is shimmed:
  real  : savemoney/modules/ref_query_alignment.py, my_classes.py, msa.py (reference Python),
          alignment_functions (the reference's Cython extension, compiled by oracle/Makefile
          from the reference's own generated C++ into oracle/_ref/)
  shim  : `parasail` -> oracle/parasail_port.c (parasail is not installable offline; see the
          header of that file: parity with the real parasail is UNPINNED),
          `Bio.Seq.Seq.reverse_complement` -> IUPAC complement table below,
          `snapgene_reader`, matplotlib*, mpl_toolkits*, spoa, PIL -> inert mocks (not on the path)
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import sys
import types
from pathlib import Path
from unittest import mock

HERE = Path(__file__).resolve().parent
REF_ROOT = Path("/root/reference")

# Biopython's ambiguous DNA complement table (Bio.Data.IUPACData.ambiguous_dna_complement)
_COMP = str.maketrans("ACGTMRWSYKVHDBNacgtmrwsykvhdbn", "TGCAKYWSRMBDHVNtgcakywsrmbdhvn")


def reverse_complement(seq: str) -> str:
    return seq.translate(_COMP)[::-1]


class _Seq:
    def __init__(self, s):
        self._s = str(s)

    def reverse_complement(self):
        return _Seq(reverse_complement(self._s))

    def __str__(self):
        return self._s


_loaded = None


def load(call_log: list | None = None):
    """Returns a namespace with .rqa (ref_query_alignment), .mc (my_classes), .msa, .af
    or None when the reference tree / compiled extension is unavailable."""
    global _loaded
    if _loaded is not None:
        _loaded.call_log_ref[0] = call_log
        return _loaded
    pkg_root = REF_ROOT / "savemoney"
    if not pkg_root.is_dir():
        return None
    ext = list((HERE / "_ref").glob("alignment_functions*.so"))
    if not ext:
        return None
    sys.path.insert(0, str(HERE.parent))
    from oracle import oracle as orc

    call_log_ref = [call_log]

    class _LogProxy(list):
        def append(self, x):  # route to whichever list the current caller supplied
            if call_log_ref[0] is not None:
                call_log_ref[0].append(x)

    sys.modules["parasail"] = orc.parasail_shim(_LogProxy())
    bio = types.ModuleType("Bio"); bio.__path__ = []
    bioseq = types.ModuleType("Bio.Seq"); bioseq.Seq = _Seq
    sys.modules.setdefault("Bio", bio); sys.modules.setdefault("Bio.Seq", bioseq)
    sg = types.ModuleType("snapgene_reader")
    sg.snapgene_file_to_dict = lambda p: (_ for _ in ()).throw(RuntimeError("snapgene_reader is shimmed"))
    sys.modules.setdefault("snapgene_reader", sg)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.backends",
                 "matplotlib.backends.backend_pdf", "matplotlib.colors", "matplotlib.ticker", "matplotlib.cm",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.inset_locator",
                 "spoa", "PIL", "PIL.Image", "PIL.ImageDraw", "PIL.ImageFont", "pysam", "pulp"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = mock.MagicMock(); m.__path__ = []
                sys.modules[name] = m
    # package skeleton without executing savemoney/__init__.py (it pulls pre_survey -> pulp)
    for name, path in (("savemoney", pkg_root), ("savemoney.modules", pkg_root / "modules"),
                       ("savemoney.modules.cython_functions", pkg_root / "modules" / "cython_functions"),
                       ("savemoney.post_analysis", pkg_root / "post_analysis"),
                       ("savemoney.pre_survey", pkg_root / "pre_survey")):
        m = types.ModuleType(name); m.__path__ = [str(path)]
        sys.modules[name] = m
    spec = importlib.util.spec_from_file_location(
        "savemoney.modules.cython_functions.alignment_functions", str(ext[0]),
        loader=importlib.machinery.ExtensionFileLoader("savemoney.modules.cython_functions.alignment_functions", str(ext[0])))
    af = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(af)
    sys.modules["savemoney.modules.cython_functions.alignment_functions"] = af
    sys.modules["savemoney.modules.cython_functions"].alignment_functions = af
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mc = importlib.import_module("savemoney.modules.my_classes")
        rqa = importlib.import_module("savemoney.modules.ref_query_alignment")
        msa = importlib.import_module("savemoney.modules.msa")
    ns = types.SimpleNamespace(rqa=rqa, mc=mc, msa=msa, af=af, call_log_ref=call_log_ref)
    _loaded = ns
    return ns


DEFAULT_PARAMS = {  # savemoney/post_analysis/post_analysis.py:14-26
    "gap_open_penalty": 3, "gap_extend_penalty": 1, "match_score": 1, "mismatch_score": -2,
    "score_threshold": 0.3, "error_rate": 1e-7, "ins_rate": 1e-7, "window": 160,
    "topology_of_dna": 0, "n_cpu": 1, "export_image_results": 0,
}


def reference_align_read(ns, ref_strs, query: str, params: dict, omit_too_long=True):
    """Exactly the loop body of post_analysis_core.execute_alignment_core_loop (:68-87),
    executed by the reference's own classes.  Returns [(my_cigar, score, new_query_seq_offset)] * 2N."""
    aligners = [ns.rqa.MyOptimizedAligner(ns.mc.MySeq(r), params) for r in ref_strs]
    return reference_align_read_with(ns, aligners, query, omit_too_long)


def reference_align_read_with(ns, aligners, query: str, omit_too_long=True):
    q = ns.mc.MySeq(query)
    q_rc = q.reverse_complement()
    out = []
    for al in aligners:
        cr = al.calc_conserved_region(q, omit_too_long=omit_too_long)
        cr_rc = al.calc_conserved_region(q_rc, omit_too_long=omit_too_long)
        for qq, c in ((q, cr), (q_rc, cr_rc)):
            r = al.execute_alignment_using_conserved_regions(qq, c) if c is not None else ns.rqa.MyResult()
            out.append((r.my_cigar, int(r.score), r.new_query_seq_offset))
    return out


def reference_distance(ns, ref_i: str, ref_j: str, params: dict) -> int:
    """pre_survey_core.RecommendedGrouping.calc_distance (:251-271) restated around the reference classes."""
    res = reference_align_read(ns, [ref_i], ref_j, params, omit_too_long=False)
    (c0, s0, _), (c1, s1, _) = res
    if s0 <= 0 and s1 <= 0:
        return len(ref_i) + len(ref_j)
    c = c0 if s0 >= s1 else c1
    return len(c) - c.count("=")
