// Host side of smx_consensus_run (smx_consensus.cuh): the factor table and the start values in native long double before
// the kernel, the sums and 1 - 1 / sum after it -- exactly the operations of alignment_functions.pyx:111-130.  Plain C++ so
// that tests/consensus_harness.cpp can run the same code around a CPU replay of the kernel's loop.
#pragma once
#include <cstddef>
#include <vector>
#include "smx_ld80.hpp"

#define SMX_CONS_BASES 5        // "ATCG-" (SequenceBasecallQscorePDF.bases, msa.py:655)
#define SMX_CONS_QIDX 94        // 0: q-score < 0 (no pdf factor), 1 + q for q = 0..92 (pdf_core holds 93 entries, msa.py:705)

// tab[((B * 5 + b) * 5 + called) * 94 + k] = (P(b, called) * pdf(b, called, q)) / (P(B, called) * pdf(B, called, q)), k = q + 1; k = 0: P(b, called) / P(B, called)
inline void smx_cons_build_table(const double *P25, const double *pdf, std::vector<uint64_t> &tm, std::vector<int32_t> &te)
{
    const size_t NT = (size_t)25 * SMX_CONS_BASES * SMX_CONS_QIDX;
    tm.resize(NT); te.resize(NT);
    for (int B = 0; B < 5; ++B)
        for (int b = 0; b < 5; ++b)
            for (int called = 0; called < 5; ++called)
                for (int k = 0; k < SMX_CONS_QIDX; ++k) {
                    const long double PB = (long double)P25[B * 5 + called], Pb = (long double)P25[b * 5 + called];
                    long double bunshi, num;
                    if (k == 0) { bunshi = PB; num = Pb; }   // q-score < 0 (:115, :127)
                    else {
                        bunshi = PB * (long double)pdf[(size_t)(B * 5 + called) * 93 + (size_t)(k - 1)];   // :113
                        num = Pb * (long double)pdf[(size_t)(b * 5 + called) * 93 + (size_t)(k - 1)];      // :125
                    }
                    const SmxLd t = smx_ld_from_native(num / bunshi);
                    const size_t at = ((size_t)(B * 5 + b) * SMX_CONS_BASES + (size_t)called) * SMX_CONS_QIDX + (size_t)k;
                    tm[at] = t.m; te[at] = t.e;
                }
}

// init[(w * n_classes + cls) * 25 + B * 5 + b] = P_N[b] / P_N[B] (:122), w = 0 with prior, 1 without
inline void smx_cons_build_init(const double *pn_with, const double *pn_without, int n_classes, std::vector<uint64_t> &im, std::vector<int32_t> &ie)
{
    im.resize((size_t)2 * n_classes * 25); ie.resize((size_t)2 * n_classes * 25);
    for (int w = 0; w < 2; ++w)
        for (int cls = 0; cls < n_classes; ++cls)
            for (int B = 0; B < 5; ++B)
                for (int b = 0; b < 5; ++b) {
                    const double *pn = (w == 0 ? pn_with : pn_without) + (size_t)cls * 5;
                    const SmxLd t = smx_ld_from_native((long double)pn[b] / (long double)pn[B]);
                    const size_t at = ((size_t)w * n_classes + (size_t)cls) * 25 + (size_t)(B * 5 + b);
                    im[at] = t.m; ie[at] = t.e;
                }
}

SMX_LD_HD int smx_cons_base_index(unsigned c)
{
    return c == 'A' ? 0 : c == 'T' ? 1 : c == 'C' ? 2 : c == 'G' ? 3 : c == '-' ? 4 : -1;
}

// The product chains of column c for the pair (B, b) whose table slice is tab_m / tab_e [5 called][94]: the reads in row
// order, both start values at once.  This is the body of smx_consensus_kernel's thread (the table slice in shared memory
// there) and of the CPU replay in tests/consensus_harness.cpp.  Returns false if a voting read holds a character outside
// ATCG- or a q-score above 92.
SMX_LD_HD bool smx_cons_column(const uint8_t *seq, const int8_t *qs, const uint8_t *cig, int rows, int64_t cols, int64_t c, const uint64_t *tab_m,
                               const int32_t *tab_e, SmxLd &v_with, SmxLd &v_without, int &events)
{
    bool ok = true;
    events = 0;
    for (int r = 0; r < rows; ++r) {
        const size_t at = (size_t)r * (size_t)cols + (size_t)c;
        const unsigned letter = cig[at];
        if (letter == 'H' || letter == 'O') continue;   // msa.py:1762
        const int called = smx_cons_base_index(seq[at]);
        const int q = qs[at];
        ++events;
        if (called < 0 || q > 92) { ok = false; continue; }
        const int idx = called * SMX_CONS_QIDX + (q < 0 ? 0 : q + 1);
        SmxLd t;
        t.m = tab_m[idx];
        t.e = tab_e[idx];
        v_with = smx_ld_mul(v_with, t);
        v_without = smx_ld_mul(v_without, t);
    }
    return ok;
}

// products [2][cols][25] -> p lists [cols][5] per prior: bunbo_bunshi_sum over the bases in order, p = 1 - 1 / sum (:118-130),
// handed back as the double Cython returns to Python
inline void smx_cons_finish(const uint64_t *om, const int32_t *oe, size_t cols, double *p_with, double *p_without)
{
    for (int w = 0; w < 2; ++w) {
        double *out = w == 0 ? p_with : p_without;
        for (size_t c = 0; c < cols; ++c)
            for (int B = 0; B < 5; ++B) {
                long double sum = 0;
                for (int b = 0; b < 5; ++b) {
                    SmxLd v;
                    v.m = om[((size_t)w * cols + c) * 25 + (size_t)(B * 5 + b)];
                    v.e = oe[((size_t)w * cols + c) * 25 + (size_t)(B * 5 + b)];
                    sum += smx_ld_to_native(v);
                }
                out[c * 5 + (size_t)B] = (double)(1 - 1 / sum);
            }
    }
}
