"""Row N3, second half: the per-column Bayesian consensus of the reference's MyMSA
(savemoney/modules/msa.py:1746-1812) on the device.

`MyMSA.calculate_consensus_core` asks, for every column of the multiple alignment, the compiled
`SequenceBasecallQscorePDF.calc_consensus_error_rate` (cython_functions/alignment_functions.pyx:93-132) for the error
probability of each candidate base: 25 long-double product chains over the reads that vote in the column.  Here
`smx_consensus_run` (csrc/smx_consensus.cuh) runs those chains for all columns at once -- one CUDA thread per (column,
candidate, base), the x87 multiplication restated on integers (csrc/smx_ld80.hpp), factor table and the last seven
operations per value in the host's own long double -- and this module does what the reference's Python does around it:
the statistics file, the priors, the minimum, the mixed-base letter, the q-score, the consensus CIGAR.

    pdf = SequenceBasecallQscorePDF(df_csv_path=...)            # same file format as the reference's NanoporeStats_*.csv
    msa = generate_msa(ref_seq, query_seq_list, q_scores_list, my_cigar_list, param_dict)     # msa.py of this package
    calculate_consensus(msa, pdf)                                # sets msa.with_prior_consensus_seq / _q_scores / _my_cigar
                                                                 # and the without_prior_* triple, like msa.py:1749-1754

Results are bit-identical to the reference's (tests/golden/consensus_v1.json.gz, recorded from its own classes).  A voting
read with a character outside ATCG- or a q-score above 92 is an error here; the reference reads past the end of a vector
for those (alignment_functions.pyx:113).
"""
from __future__ import annotations

import ctypes
import re

import numpy as np

BASES = "ATCG-"   # SequenceBasecallQscorePDF.bases / MyMSA.bases (msa.py:655)
LETTER_CODE_DICT = {"ATCG": "N", "TCG": "B", "ACG": "V", "ATG": "D", "ATC": "H", "TG": "K", "AC": "M", "AG": "R", "CG": "S", "AT": "W", "TC": "Y",
                    "A": "A", "T": "T", "C": "C", "G": "G"}   # IUPAC codes keyed by their bases in the order of BASES (my_classes.py, MyCigarBase.letter_code_dict)
MAX_Q = 41        # MyFastQ.maximum_q_score_allowed: the statistics table holds the rows -1 .. 41 (msa.py:666,702)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class SequenceBasecallQscorePDF:
    """msa.py:654-733: per (true base, called base) the probability of the call and the distribution of its q-score,
    from a table of event counts (tab-separated, first line `# file_version: x.y.z`, then a header row of 25 keys `B_c`
    and the rows -1 .. 41)."""
    bases = BASES

    def __init__(self, df_csv_path=None, columns=None, counts=None):
        if df_csv_path is not None:
            with open(df_csv_path, "r") as f:
                self.NanoporeStats_PDF_version = re.match(r"^# file_version: ([0-9]+\.[0-9]+\.[0-9]+)$", f.readline().strip()).group(1)
                columns = f.readline().rstrip("\n").split("\t")[1:]
                rows = [line.rstrip("\n").split("\t") for line in f if line.strip()]
            index = [int(r[0]) for r in rows]
            counts = np.array([[int(x) for x in r[1:]] for r in rows], dtype=np.int64)
        else:
            self.NanoporeStats_PDF_version = None
            counts = np.array(counts, dtype=np.int64)
            index = list(range(-1, MAX_Q + 1))
        assert index == list(range(-1, MAX_Q + 1)) and counts.shape == (len(index), len(columns))       # msa.py:666,702
        self.columns = list(columns)
        self.counts = counts
        self.pdf_core = {}
        self.P_base_calling_given_true_refseq_dict = {}
        self.initialize_pdf()
        self.q_score_distribution = [int(x) for x in self.counts.sum(axis=1)]                              # msa.py:665-667

    def initialize_pdf(self):
        c = self.counts
        for k, name in enumerate(self.columns):            # no zero probabilities (msa.py:679-687): rows 1.. of the columns that do not end in "-"
            if name.endswith("-"):
                continue
            col = c[2:, k]
            col[col == 0] = 1
        total = {}
        for k, name in enumerate(self.columns):            # msa.py:689-695
            total[name.split("_")[0]] = total.get(name.split("_")[0], 0) + int(c[:, k].sum())
        for k, name in enumerate(self.columns):            # msa.py:697-705
            s = np.int64(c[:, k].sum())
            self.P_base_calling_given_true_refseq_dict[name] = float(s / np.int64(total[name.split("_")[0]]))
            values = c[:, k] / s
            self.pdf_core[name] = list(values)[1:] + [0.0 for _ in range(50)] + list(values)[:1]

    def calc_js_divergence(self, q_score_distribution):   # msa.py:668-676
        assert len(q_score_distribution) == len(self.q_score_distribution)
        p_dist = np.array(self.q_score_distribution[1:]) + 1
        q_dist = np.array(q_score_distribution[1:]) + 1
        p_dist = p_dist / p_dist.sum()
        q_dist = q_dist / q_dist.sum()
        r_dist = (p_dist + q_dist) / 2
        return p_dist.dot(np.log(p_dist / r_dist)) + q_dist.dot(np.log(q_dist / r_dist))

    def tables(self):
        """-> P25 [25], pdf [25, 93] float64 in the key order B_c with B, c over BASES (what smx_consensus_run takes)"""
        P = np.zeros(25, np.float64)
        pdf = np.zeros((25, 93), np.float64)
        for i, B in enumerate(BASES):
            for j, cc in enumerate(BASES):
                key = f"{B}_{cc}"
                P[i * 5 + j] = self.P_base_calling_given_true_refseq_dict[key]
                pdf[i * 5 + j, :] = self.pdf_core[key]
        return P, pdf


def select_sbq_pdf(q_score_distribution, candidates):
    """MyMSA.set_sbq_pdf (msa.py:1738-1745): of the statistics tables on offer (SequenceBasecallQscorePDF objects or csv paths,
    in the order of their minor versions) the one whose q-score distribution is closest (Jensen-Shannon) to the run's own
    (`MyFastQ.get_q_score_distribution()`: counts of the q-scores -1 .. 41)"""
    pdfs = [c if isinstance(c, SequenceBasecallQscorePDF) else SequenceBasecallQscorePDF(df_csv_path=c) for c in candidates]
    js_scores = [pdf.calc_js_divergence(q_score_distribution) for pdf in pdfs]
    return pdfs[int(np.argmin(js_scores))]


def consensus_params(param_dict):
    """msa.py:1813-1833: P_N_dict per reference letter, with and without prior (plain dicts keyed by letter).
    With prior: the bases a (possibly ambiguous) reference letter stands for share 1 - error_rate, the others -- the gap
    included -- share error_rate; a gap column expects an insertion with ins_rate.  Without prior: uniform."""
    ins_rate, error_rate = param_dict["ins_rate"], param_dict["error_rate"]
    n = len(BASES)
    assert n == 5
    with_prior, without_prior = {}, {}
    for stands_for, letter in LETTER_CODE_DICT.items():
        inside, outside = (1 - error_rate) / len(stands_for), error_rate / (n - len(stands_for))
        with_prior[letter] = {b: inside if b in stands_for else outside for b in BASES}
        without_prior[letter] = {b: 1 / n for b in BASES}
    with_prior["-"] = {b: 1 - ins_rate if b == "-" else ins_rate / 4 for b in BASES}
    without_prior["-"] = {b: 1 / n for b in BASES}
    return with_prior, without_prior


def mixed_bases(base_list):
    """msa.py:1846-1857: one candidate -> itself; several -> the IUPAC letter of the bases among them (a gap among them is dropped)"""
    if len(base_list) == 1:
        return base_list[0]
    present = set(base_list) - {"-"}
    return LETTER_CODE_DICT["".join(b for b in BASES[:-1] if b in present)]


def _column_letter(ref, call):
    """msa.py:1799-1810: consensus CIGAR letter of a column from its reference character and the consensus call"""
    if call == "-":
        return "N" if ref == "-" else "D"
    if ref == "-":
        return "I"
    return "=" if ref == call else "X"


def consensus_error_rates(msa, sbq_pdf, param_dict=None, engine=None, device_resident=False):
    """p_list of every column (with prior, without prior) [n_columns, 5] and the number of voting reads per column.
    device_resident=True: `msa` is the result of the LAST generate_msa on this engine, whose matrices are still on the
    device (no second upload; smx_consensus_run checks the dimensions)."""
    from .savemoney import default_engine
    eng = engine or default_engine()
    params = param_dict if param_dict is not None else msa.param_dict
    with_prior, without_prior = consensus_params(params)
    ref = msa.ref_seq_aligned.upper()
    letters = sorted(set(ref))
    for ch in letters:
        if ch not in with_prior:
            raise KeyError(ch)                                  # P_N_dict_dict[ref.upper()] (msa.py:1767)
    cls_of = {ch: k for k, ch in enumerate(letters)}
    col_class = np.fromiter((cls_of[ch] for ch in ref), dtype=np.int32, count=len(ref))
    pn_w = np.array([[with_prior[ch][b] for b in BASES] for ch in letters], dtype=np.float64).reshape(-1, 5)
    pn_wo = np.array([[without_prior[ch][b] for b in BASES] for ch in letters], dtype=np.float64).reshape(-1, 5)
    P, pdf = sbq_pdf.tables()
    seq = np.ascontiguousarray(msa.query_seq_matrix, dtype=np.uint8)
    qs = np.ascontiguousarray(msa.q_scores_matrix, dtype=np.int8)
    cig = np.ascontiguousarray(msa.my_cigar_matrix, dtype=np.uint8)
    n_rows, n_cols = (int(seq.shape[0]), int(seq.shape[1])) if seq.ndim == 2 else (0, len(ref))
    n_cols = len(ref)
    p_w = np.zeros((n_cols, 5), np.float64)
    p_wo = np.zeros((n_cols, 5), np.float64)
    n_ev = np.zeros(max(n_cols, 1), np.int32)
    if n_cols:
        dummy8 = np.zeros(1, np.uint8)
        mats = (None, None, None) if device_resident else (_ptr(seq) if seq.size else _ptr(dummy8), _ptr(qs) if qs.size else _ptr(dummy8.view(np.int8)),
                                                            _ptr(cig) if cig.size else _ptr(dummy8))
        rc = eng._lib.smx_consensus_run(eng._h, *mats, n_rows, n_cols, _ptr(P), _ptr(pdf), _ptr(col_class), _ptr(pn_w), _ptr(pn_wo),
                                        len(letters), _ptr(p_w), _ptr(p_wo), _ptr(n_ev))
        if rc == -7:
            raise ValueError(eng._lib.smx_last_error(eng._h).decode())
        eng._check(rc, "smx_consensus_run")
    return p_w, p_wo, n_ev[:n_cols]


def _consensus_core(ref_seq_aligned, p_rows, n_events):
    """msa.py:1755-1812 once the p_list of every column is known: the call is the candidate (or the mix of the candidates)
    with the smallest error probability p, its q-score round(-10 log10 p) capped at 50 below p = 1e-5; a column nobody votes
    in is a gap with q-score -1"""
    calls, q_scores, letters = [], [], []
    for ref, p_row, ev in zip(ref_seq_aligned, p_rows, n_events):
        if ev > 0:
            p_list = [float(x) for x in p_row]
            p = min(p_list)
            call = mixed_bases([b for b, tmp_p in zip(BASES, p_list) if tmp_p == p])
            if p >= 10 ** (-5):
                q_score = np.round(-10 * np.log10(p)).astype(int)
            elif p < 0:
                raise Exception("unknown error")
            else:
                q_score = 50
        else:
            call, q_score = "-", -1
        calls.append(call); q_scores.append(q_score); letters.append(_column_letter(ref, call))
    return "".join(calls), q_scores, "".join(letters)


def calculate_consensus(msa, sbq_pdf, param_dict=None, engine=None, device_resident=False):
    """MyMSA.calculate_consensus (msa.py:1749-1754): fills the with_prior_* and without_prior_* fields of `msa` and returns them"""
    p_w, p_wo, n_ev = consensus_error_rates(msa, sbq_pdf, param_dict, engine, device_resident)
    ref = msa.ref_seq_aligned
    msa.with_prior_consensus_seq, msa.with_prior_consensus_q_scores, msa.with_prior_consensus_my_cigar = _consensus_core(ref, p_w, n_ev)
    msa.without_prior_consensus_seq, msa.without_prior_consensus_q_scores, msa.without_prior_consensus_my_cigar = _consensus_core(ref, p_wo, n_ev)
    return ((msa.with_prior_consensus_seq, msa.with_prior_consensus_q_scores, msa.with_prior_consensus_my_cigar),
            (msa.without_prior_consensus_seq, msa.without_prior_consensus_q_scores, msa.without_prior_consensus_my_cigar))
