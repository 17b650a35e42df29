/*
 * oracle/parasail_port.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Scalar CPU restatement of the arithmetic that SAVEMONEY delegates to the
 * third-party library `parasail` (PyPI parasail==1.3.4, requirements.txt:7 of the
 * reference; C library github.com/jeffdaily/parasail 2.6.x).  parasail is NOT vendored
 * in /root/reference and is not installable here, so this file restates its published
 * serial algorithm (src/nw_trace.c, src/sg_trace.c, src/sw_trace.c, src/cigar.c in the
 * parasail tree) as recorded in SURVEY.md Appendix A.
 *
 * PARITY STATUS: "parity unpinned" at the parasail boundary -- the reference holds no
 * golden vectors for these calls.  What pins this file:
 *   - the reference's runtime asserts (valid CIGAR, full-length SG CIGARs),
 *   - optimal-score check against an independent score-only DP (tests/),
 *   - the 7 tie-insensitive legacy pre_survey distances of the bundled demo plasmids.
 * Every tie rule that could not be verified against a real parasail is a named
 * constant below so that it can be flipped the moment a real parasail is reachable.
 *
 * Call sites in the reference that this replaces (s1 = query, s2 = reference):
 *   savemoney/modules/ref_query_alignment.py:58,343  parasail.nw_trace
 *   savemoney/modules/ref_query_alignment.py:68      parasail.sg_qe_de_trace
 *   savemoney/modules/ref_query_alignment.py:82      parasail.sg_qb_db_trace
 *   savemoney/modules/ref_query_alignment.py:35      parasail.sw_trace
 *   savemoney/modules/ref_query_alignment.py:32      parasail.matrix_create("ACGT", m, mm)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- tie rules -------------------------------------- */
/* Run-time switches, defaults = the rules the product implements.  oracle/tie_sensitivity.py
 * flips them one at a time and counts how many reference-level results (CIGAR, score, offset)
 * move: that table is what is at stake while no real parasail is reachable. */
/* gap open vs extend: parasail tests `opn > ext ? DIAG_x : extend`  => tie goes to EXTEND */
static int TIE_GAP_PREFERS_EXTEND = 1;
/* H source: DIAG if H == H_dag, else DEL (F, vertical, CIGAR 'I') if H == F, else INS (E, CIGAR 'D') */
static int TIE_H_ORDER_DIAG_F_E = 1;
/* sg_qe_de end cell, always strict '>' and ascending indices inside a pass:
 *   0 (default): last column rows 0..m-2, then last row columns 0..n-1 (the corner is the LAST
 *                candidate: a tied earlier last-row cell beats it) -- the loop structure of
 *                parasail's serial sg code (`for i < s1Len` main loop, then a separate last-row block)
 *   1          : last column rows 0..m-1 (corner included), then last row 0..n-2 (SURVEY A.6 as written:
 *                the corner beats a tied last-row cell)
 *   2          : last row first, then last column (the order of the striped SIMD kernels) */
static int TIE_SG_END_ORDER = 0;
/* sw end cell: 1 (default) smallest end_ref, then smallest end_query -- parasail's serial sw keeps `WH > score`, else
 * `WH == score && j - 1 < end_ref` (the order its striped kernels produce: first column holding the maximum, then its
 * first row); 0 = first maximum in row-major order (round-1 default of this port) */
static int TIE_SW_END_ORDER = 1;

void pp_set_tie_rules(int gap_prefers_extend, int h_order_diag_f_e, int sg_end_order, int sw_end_order)
{
    TIE_GAP_PREFERS_EXTEND = gap_prefers_extend;
    TIE_H_ORDER_DIAG_F_E = h_order_diag_f_e;
    TIE_SG_END_ORDER = sg_end_order;
    TIE_SW_END_ORDER = sw_end_order;
}

/* trace byte layout (parasail.h): low 3 bits = H source, then E and F sources */
enum { T_ZERO = 0, T_INS = 1, T_DEL = 2, T_DIAG = 4, T_DIAG_E = 8, T_INS_E = 16, T_DIAG_F = 32, T_DEL_F = 64 };

enum { MODE_NW = 0, MODE_SG_QE_DE = 1, MODE_SG_QB_DB = 2, MODE_SW = 3 };

#define NEG_INF_32 (INT32_MIN / 2)

/* parasail_matrix_create("ACGT", match, mismatch): 5x5, last row/column ('*') are 0.
 * mapper: ACGTacgt -> 0..3, every other byte -> 4. */
static inline int map_base(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

void pp_matrix_create(int match, int mismatch, int32_t out25[25])
{
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b)
            out25[a * 5 + b] = (a == 4 || b == 4) ? 0 : (a == b ? match : mismatch);
}

typedef struct {
    int32_t score;
    int32_t end_query; /* 0-based inclusive, index into s1 */
    int32_t end_ref;   /* 0-based inclusive, index into s2 */
    int32_t beg_query; /* cigar.beg_query */
    int32_t beg_ref;   /* cigar.beg_ref   */
    int64_t n_ops;     /* number of CIGAR columns written to ops_out (expanded, forward order) */
} pp_result_t;

/*
 * One pairwise alignment with full traceback.
 *   s1/len1 : query  (rows i)         s2/len2 : reference (columns j)
 *   ops_out : caller buffer of capacity >= len1 + len2, receives the expanded CIGAR
 *             ('=', 'X', 'I', 'D') in forward order.
 * Returns 0 on success, -1 on bad arguments / allocation failure.
 */
int pp_align(int mode, const char *s1, int len1, const char *s2, int len2,
             int open, int gap, int match, int mismatch,
             pp_result_t *res, char *ops_out)
{
    if (len1 <= 0 || len2 <= 0 || !res || !ops_out) return -1;
    int32_t M[25];
    pp_matrix_create(match, mismatch, M);

    const int free_beg = (mode == MODE_SG_QB_DB);
    const int free_end = (mode == MODE_SG_QE_DE);
    const int local = (mode == MODE_SW);

    uint8_t *a = (uint8_t *)malloc((size_t)len1);
    uint8_t *b = (uint8_t *)malloc((size_t)len2);
    int32_t *H = (int32_t *)malloc(sizeof(int32_t) * ((size_t)len2 + 1));
    int32_t *F = (int32_t *)malloc(sizeof(int32_t) * ((size_t)len2 + 1));
    uint8_t *HT = (uint8_t *)malloc((size_t)len1 * (size_t)len2);
    int32_t *lastcol = (int32_t *)malloc(sizeof(int32_t) * (size_t)len1);
    if (!a || !b || !H || !F || !HT || !lastcol) { free(a); free(b); free(H); free(F); free(HT); free(lastcol); return -1; }
    for (int i = 0; i < len1; ++i) a[i] = (uint8_t)map_base((unsigned char)s1[i]);
    for (int j = 0; j < len2; ++j) b[j] = (uint8_t)map_base((unsigned char)s2[j]);

    /* first row */
    H[0] = 0;
    F[0] = NEG_INF_32;
    for (int j = 1; j <= len2; ++j) {
        H[j] = (free_beg || local) ? 0 : -open - (j - 1) * gap;
        F[j] = NEG_INF_32;
    }

    int32_t best = NEG_INF_32, end_q = len1 - 1, end_r = len2 - 1;
    if (local) { best = 0; }

    for (int i = 1; i <= len1; ++i) {
        const int32_t *mrow = &M[5 * a[i - 1]];
        int32_t NH = H[0];
        int32_t WH = (free_beg || local) ? 0 : -open - (i - 1) * gap;
        int32_t E = NEG_INF_32;
        H[0] = WH;
        uint8_t *trow = HT + (size_t)(i - 1) * (size_t)len2;
        for (int j = 1; j <= len2; ++j) {
            int32_t NWH = NH;
            NH = H[j];
            int32_t F_opn = NH - open, F_ext = F[j] - gap;
            int32_t E_opn = WH - open, E_ext = E - gap;
            uint8_t t;
            if (TIE_GAP_PREFERS_EXTEND) {
                if (F_opn > F_ext) { F[j] = F_opn; t = T_DIAG_F; } else { F[j] = F_ext; t = T_DEL_F; }
                if (E_opn > E_ext) { E = E_opn; t |= T_DIAG_E; } else { E = E_ext; t |= T_INS_E; }
            } else {
                if (F_opn >= F_ext) { F[j] = F_opn; t = T_DIAG_F; } else { F[j] = F_ext; t = T_DEL_F; }
                if (E_opn >= E_ext) { E = E_opn; t |= T_DIAG_E; } else { E = E_ext; t |= T_INS_E; }
            }
            int32_t H_dag = NWH + mrow[b[j - 1]];
            int32_t Fv = F[j];
            WH = H_dag;
            if (E > WH) WH = E;
            if (Fv > WH) WH = Fv;
            if (local && WH < 0) WH = 0;
            if (local && WH == 0) {
                /* parasail sw: a cell whose value is 0 stops the walk */
                t |= T_ZERO;
            } else {
                if (WH == H_dag) t |= T_DIAG;
                else if (TIE_H_ORDER_DIAG_F_E) t |= (WH == Fv) ? T_DEL : T_INS;
                else t |= (WH == E) ? T_INS : T_DEL;
            }
            H[j] = WH;
            trow[j - 1] = t;
            if (local && (WH > best || (TIE_SW_END_ORDER == 1 && WH == best && best > 0 && j - 1 < end_r))) {
                best = WH; end_q = i - 1; end_r = j - 1;
            }
        }
        if (free_end) lastcol[i - 1] = WH;
    }
    if (free_end) {
        /* candidates: last column H(i, len2) for i = 1..len1 (kept in lastcol), last row H(len1, j) = H[j] */
        if (TIE_SG_END_ORDER == 0) {
            for (int i = 1; i < len1; ++i)
                if (lastcol[i - 1] > best) { best = lastcol[i - 1]; end_q = i - 1; end_r = len2 - 1; }
            for (int j = 1; j <= len2; ++j)
                if (H[j] > best) { best = H[j]; end_q = len1 - 1; end_r = j - 1; }
        } else if (TIE_SG_END_ORDER == 1) {
            for (int i = 1; i <= len1; ++i)
                if (lastcol[i - 1] > best) { best = lastcol[i - 1]; end_q = i - 1; end_r = len2 - 1; }
            for (int j = 1; j < len2; ++j)
                if (H[j] > best) { best = H[j]; end_q = len1 - 1; end_r = j - 1; }
        } else {
            for (int j = 1; j <= len2; ++j)
                if (H[j] > best) { best = H[j]; end_q = len1 - 1; end_r = j - 1; }
            for (int i = 1; i < len1; ++i)
                if (lastcol[i - 1] > best) { best = lastcol[i - 1]; end_q = i - 1; end_r = len2 - 1; }
        }
    } else if (!local) {
        best = H[len2]; end_q = len1 - 1; end_r = len2 - 1;
    }

    /* ---- traceback (parasail cigar_template.c); ops are produced reversed ---- */
    char *rev = (char *)malloc((size_t)len1 + (size_t)len2 + 2);
    if (!rev) { free(a); free(b); free(H); free(F); free(HT); free(lastcol); return -1; }
    int64_t n = 0;
    int i = end_q, j = end_r;
    if (free_end) {
        /* semi-global: the CIGAR also covers the free trailing part */
        if (end_q + 1 == len1) { for (int k = len2 - 1; k > end_r; --k) rev[n++] = 'D'; }
        else if (end_r + 1 == len2) { for (int k = len1 - 1; k > end_q; --k) rev[n++] = 'I'; }
    }
    int where = T_DIAG;
    while (local ? (i >= 0 && j >= 0) : (i >= 0 || j >= 0)) {
        if (i < 0) { rev[n++] = 'D'; --j; continue; }
        if (j < 0) { rev[n++] = 'I'; --i; continue; }
        uint8_t t = HT[(size_t)i * (size_t)len2 + (size_t)j];
        if (where == T_DIAG) {
            if (t & T_DIAG) { rev[n++] = (s1[i] == s2[j]) ? '=' : 'X'; --i; --j; }
            else if (t & T_INS) where = T_INS;
            else if (t & T_DEL) where = T_DEL;
            else break; /* ZERO */
        } else if (where == T_INS) {
            rev[n++] = 'D'; --j;
            where = (t & T_DIAG_E) ? T_DIAG : T_INS;
        } else {
            rev[n++] = 'I'; --i;
            where = (t & T_DIAG_F) ? T_DIAG : T_DEL;
        }
    }
    res->score = best;
    res->end_query = end_q;
    res->end_ref = end_r;
    res->beg_query = local ? i + 1 : 0;
    res->beg_ref = local ? j + 1 : 0;
    res->n_ops = n;
    for (int64_t k = 0; k < n; ++k) ops_out[k] = rev[n - 1 - k];

    free(rev); free(a); free(b); free(H); free(F); free(HT); free(lastcol);
    return 0;
}

/* Independent score-only affine DP (no trace, no tie rules): used by the tests to check
 * that pp_align's path is optimal. Textbook Gotoh, written separately on purpose. */
int32_t pp_score_only(int mode, const char *s1, int len1, const char *s2, int len2,
                      int open, int gap, int match, int mismatch)
{
    int32_t M[25];
    pp_matrix_create(match, mismatch, M);
    const int fb = (mode == MODE_SG_QB_DB), fe = (mode == MODE_SG_QE_DE), lo = (mode == MODE_SW);
    size_t W = (size_t)len2 + 1;
    int32_t *Hp = malloc(sizeof(int32_t) * W), *Hc = malloc(sizeof(int32_t) * W), *Fc = malloc(sizeof(int32_t) * W);
    int32_t best = lo ? 0 : NEG_INF_32;
    for (size_t j = 0; j < W; ++j) { Hp[j] = (j == 0 || fb || lo) ? 0 : -open - ((int)j - 1) * gap; Fc[j] = NEG_INF_32; }
    for (int i = 1; i <= len1; ++i) {
        Hc[0] = (fb || lo) ? 0 : -open - (i - 1) * gap;
        int32_t E = NEG_INF_32;
        for (int j = 1; j <= len2; ++j) {
            int32_t e1 = Hc[j - 1] - open, e2 = E - gap; E = e1 > e2 ? e1 : e2;
            int32_t f1 = Hp[j] - open, f2 = Fc[j] - gap; Fc[j] = f1 > f2 ? f1 : f2;
            int32_t h = Hp[j - 1] + M[5 * map_base((unsigned char)s1[i - 1]) + map_base((unsigned char)s2[j - 1])];
            if (E > h) h = E;
            if (Fc[j] > h) h = Fc[j];
            if (lo && h < 0) h = 0;
            Hc[j] = h;
            if (lo && h > best) best = h;
        }
        if (fe && Hc[len2] > best) best = Hc[len2];
        int32_t *t = Hp; Hp = Hc; Hc = t;
    }
    if (fe) { for (int j = 1; j <= len2; ++j) if (Hp[j] > best) best = Hp[j]; }
    else if (!lo) best = Hp[len2];
    free(Hp); free(Hc); free(Fc);
    return best;
}

/*
 * Diagonal match-count scan: restates af.k_mer_offset_analysis_2
 * (savemoney/modules/cython_functions/alignment_functions.pyx:46-64):
 *     out[k] = #{ i < n_query : ref_rep[k + i] == query[i] },  k in [0, max_k)
 * on the int64 `ord()` code vectors the reference builds
 * (ref_query_alignment.py:348-366).
 */
void pp_kmer_offset_scan(const int64_t *ref_rep, const int64_t *query, int64_t max_k, int64_t n_query, int64_t *out)
{
    for (int64_t k = 0; k < max_k; ++k) {
        int64_t s = 0;
        for (int64_t i = 0; i < n_query; ++i) s += (ref_rep[k + i] == query[i]);
        out[k] = s;
    }
}
