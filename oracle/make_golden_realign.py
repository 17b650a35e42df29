"""Generates tests/golden/realign_v1.json.gz: what the reference's polish stage asks of its aligner, recorded while the
UNMODIFIED reference Python (read-only /root/reference, through oracle/ref_harness.py) polishes real multiple alignments.

For plasmids of golden alignment cases the reads assigned to them are stacked exactly like MyMSAligner.execute does
(modules/msa.py:617-639: generate_msa, set_query_seq_offset_list, add_pseudo_ref_seq_as_query) and MyMSA.polish runs its
four rounds (offsets 0, 1, 2, 0 thirds of the window).  Two things are replaced: spoa's `poa` (not installed; row N1 is
about the re-alignment, not the consensus) by a seeded stand-in that returns the reference chunk with a few edits, and
multiprocessing.Pool by an in-process map.  Recorded per exec_chunk_poa call (msa.py:1249-1355): its inputs, the
consensus, every MyAlignerBase call it made (operation, arguments, resulting my_cigar) and the my_cigar_list it handed to
generate_msa.          Run here:   python -m oracle.make_golden_realign
"""
from __future__ import annotations

import gzip
import json
import re
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_harness as rh  # noqa: E402

OUT = ROOT / "tests" / "golden" / "realign_v1.json.gz"
WINDOW = 160


def fake_poa(seqs, algorithm=1, genmsa=False, **kw):
    """stand-in for spoa.poa: the first sequence (Z + reference chunk + Z) with seeded point edits"""
    body = seqs[0][1:-1]
    rng = np.random.default_rng(zlib.crc32(body.encode()) & 0xffffffff)
    out = []
    for c in body:
        r = rng.random()
        if r < 0.01:
            out.append("ACGT"[rng.integers(4)])
        elif r < 0.015:
            out.append(c); out.append("ACGT"[rng.integers(4)])
        elif r < 0.02:
            continue
        else:
            out.append(c)
    return "Z" + ("".join(out) or body[:1]) + "Z", None


class FakePool:
    def __init__(self, processes=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def imap(self, fn, it):
        return map(fn, it)


def expand(cigar_rle: str) -> str:
    return "".join(c * int(n) for n, c in re.findall(r"(\d+)(.)", cigar_rle))


def main():
    ns = rh.load()
    assert ns is not None, "needs /root/reference and oracle/_ref"
    msa, mc = ns.msa, ns.mc
    msa.poa = fake_poa
    msa.Pool = FakePool
    g = json.load(gzip.open(ROOT / "tests" / "golden" / "golden_v1.json.gz", "rt"))
    records = []
    current = {}
    orig_exec = msa.MyMSA.exec_chunk_poa
    orig_gen = msa.MyMSA.generate_msa
    names = {"nw_trace": 0, "my_special_sg_qe_de": 1, "my_special_sg_qb_db": 2, "my_special_sw": 3, "my_special_DP": 4}
    originals = {k: getattr(ns.rqa.MyAlignerBase, k) for k in names}

    def wrap_op(name):
        def f(self, *a, **kw):
            r = originals[name](self, *a, **kw)
            if current.get("rec") is not None:
                if name == "my_special_DP":
                    q1 = kw.get("query_seq_1", a[0] if a else None); q2 = kw.get("query_seq_2", a[1] if len(a) > 1 else None)
                    ref = kw.get("ref_seq", a[2] if len(a) > 2 else None)
                    if current.get("inside_dp", 0) == 0:
                        current["rec"]["ops"].append([4, str(ref), str(q1), str(q2), r.my_cigar])
                elif current.get("inside_dp", 0) == 0:
                    ref = kw.get("ref_seq", a[0] if a else None); q = kw.get("query_seq", a[1] if len(a) > 1 else None)
                    current["rec"]["ops"].append([names[name], str(ref), str(q), "", r.my_cigar])
            return r
        return f

    def wrap_dp(self, *a, **kw):   # the two sg calls inside my_special_DP are not operations of their own
        current["inside_dp"] = current.get("inside_dp", 0) + 1
        try:
            r = originals["my_special_DP"](self, *a, **kw)
        finally:
            current["inside_dp"] -= 1
        if current.get("rec") is not None:
            q1 = kw.get("query_seq_1"); q2 = kw.get("query_seq_2"); ref = kw.get("ref_seq")
            current["rec"]["ops"].append([4, str(ref), str(q1), str(q2), r.my_cigar])
        return r

    for k in names:
        setattr(ns.rqa.MyAlignerBase, k, wrap_dp if k == "my_special_DP" else wrap_op(k))

    def exec_rec(self, ref_chunk, q_chunks, c_chunks, s_idx):
        rec = {"len_ref_seq_aligned": len(self.ref_seq_aligned), "query_seq_offset_list_aligned": [int(x) for x in self.query_seq_offset_list_aligned],
               "ref_chunk_aligned": ref_chunk, "q_chunks": list(q_chunks), "c_chunks": list(c_chunks), "s_idx": int(s_idx), "ops": []}
        current["rec"] = rec

        def gen_rec(ref_seq, query_seq_list, q_scores_list, my_cigar_list, param_dict):
            if current.get("rec") is rec and current.get("inside_dp", 0) == 0:   # my_special_DP calls generate_msa too (two rows): not this one
                rec["consensus"] = str(ref_seq)
                rec["my_cigar_list"] = list(my_cigar_list)
            return orig_gen(ref_seq, query_seq_list, q_scores_list, my_cigar_list, param_dict)
        msa.MyMSA.generate_msa = staticmethod(gen_rec)
        try:
            out = orig_exec(self, ref_chunk, q_chunks, c_chunks, s_idx)
        finally:
            msa.MyMSA.generate_msa = staticmethod(orig_gen)
            current["rec"] = None
        records.append(rec)
        return out
    msa.MyMSA.exec_chunk_poa = exec_rec

    for case_name, max_reads in (("c2_small_circular", 12), ("fresh_32_topology0", 10), ("c4_small_linear", 10)):
        case = next(c for c in g["cases"] if c["name"] == case_name)
        params = {**rh.DEFAULT_PARAMS, **case["params"], "window": WINDOW}
        topo = params["topology_of_dna"]
        for p, ref in enumerate(case["refs"]):
            qs, cigs, offs = [], [], []
            for read, row in zip(case["reads"], case["results"]):
                best = max(range(len(row)), key=lambda k: row[k][1] / len(case["refs"][k // 2]) if row[k][0] else -1e9)
                if best // 2 != p or not row[best][0] or row[best][1] / len(ref) < 0.3:
                    continue
                cig, _, off = row[best]
                seq = read if best % 2 == 0 else str(mc.MySeq(read).reverse_complement())
                qs.append(mc.MySeq.set_offset_core(seq.upper(), off)); cigs.append(expand(cig)); offs.append((off - 1) % len(seq) + 1)
            qs, cigs, offs = qs[:max_reads], cigs[:max_reads], offs[:max_reads]
            if len(qs) < 2:
                continue
            n0 = len(records)
            my_msa = msa.MyMSA.generate_msa(mc.MySeq(ref), qs, [[30] * len(q) for q in qs], cigs, params)
            my_msa.set_query_seq_offset_list(offs)
            actual_window = int(np.ceil(WINDOW / 2)) * 3
            my_msa.add_pseudo_ref_seq_as_query()
            try:
                for offset in np.array([0, 1, 2, 0]) * (actual_window // 3):
                    my_msa = my_msa.polish(offset=int(offset), window=actual_window, topology_of_dna=topo)
            except Exception as ex:   # keep what was recorded; say why the rounds stopped
                print(f"{case_name} plasmid {p}: polish stopped after {len(records) - n0} chunks: {type(ex).__name__}: {ex}")
            for r in records[n0:]:
                r["case"] = case_name; r["plasmid"] = p; r["params"] = {k: params[k] for k in ("gap_open_penalty", "gap_extend_penalty", "match_score", "mismatch_score")}
            print(f"{case_name} plasmid {p}: {len(qs)} reads, {len(records) - n0} chunks, {sum(len(r['ops']) for r in records[n0:])} operations", flush=True)
    # direct calls of the five operations on seeded chunk-sized problems (the polish rounds above never needed my_special_sw)
    for k in names:
        setattr(ns.rqa.MyAlignerBase, k, originals[k])
    from tests import util
    rng = np.random.default_rng(77)
    direct = []
    for scoring in ((3, 1, 1, -2), (5, 2, 2, -3)):
        al = ns.rqa.MyAlignerBase(dict(gap_open_penalty=scoring[0], gap_extend_penalty=scoring[1], match_score=scoring[2], mismatch_score=scoring[3]))
        for q, r in util.make_problems(rng, 220, 260, related=0.85, low_complexity=0.25):
            for kind in range(5):
                try:
                    if kind == 4:
                        cut = int(rng.integers(0, len(q) + 1))
                        q1, q2 = q[cut:], q[:cut]        # the read's physical end lies inside the chunk
                        if rng.random() < 0.1:
                            q1 = ""
                        elif rng.random() < 0.1:
                            q2 = ""
                        cig = al.my_special_DP(query_seq_1=q1, query_seq_2=q2, ref_seq=r).my_cigar
                        direct.append([4, r, q1, q2, cig, list(scoring)])
                    else:
                        fn = [al.nw_trace, al.my_special_sg_qe_de, al.my_special_sg_qb_db, al.my_special_sw][kind]
                        qq = q if kind != 3 else ("ACGT" * 3 + q[len(q) // 4: 3 * len(q) // 4 + 1] + "TTGCA")   # local: a core with foreign flanks
                        cig = fn(ref_seq=r, query_seq=qq).my_cigar
                        direct.append([kind, r, qq, "", cig, list(scoring)])
                except (IndexError, AssertionError) as ex:   # inputs on which the reference itself raises
                    direct.append([kind, r, q if kind != 4 else q1, "" if kind != 4 else q2, None, list(scoring)])
    print("direct operations:", len(direct), "of which the reference raises on", sum(1 for d in direct if d[4] is None))
    records = [r for r in records if "my_cigar_list" in r]
    kinds = np.bincount([o[0] for r in records for o in r["ops"]], minlength=5)
    print("operations by kind (nw, sg_qe_de, sg_qb_db, sw, special_DP):", kinds.tolist())
    with gzip.open(OUT, "wt") as f:
        json.dump({"generator": "oracle/make_golden_realign.py", "window": WINDOW, "records": records, "direct": direct}, f)
    print("wrote", OUT, OUT.stat().st_size, "bytes,", len(records), "chunks")


if __name__ == "__main__":
    main()
