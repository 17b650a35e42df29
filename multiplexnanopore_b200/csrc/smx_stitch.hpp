// Rows A7, A9, A10, A11 on the host (C++): CIGAR post-processing around the DP results.
// Works on expanded column strings (one letter per alignment column, the reference's `my_cigar`,
// savemoney/modules/my_classes.py:428-429) with the alphabet = X I D S H (+ N O inside the merge).
// Restates, with file:line in savemoney/modules/ref_query_alignment.py unless noted:
//   soft_clip()        MyResult.set_soft_clipping(_core)              :562-622
//   merge_special()    MyAlignerBase.my_special_DP                    :89-163
//   aligned_columns()  msa.MyMSA.generate_msa for two rows            modules/msa.py:882-982
//   (inside merge_special) cumsum_my_cigar_scores                     :164-185
//   rotate_to_ref0()   MyResult.set_ref_seq_offset_as_0 (+ helpers)   :465-511
//   cigar_score()      MyResult.set_cigar_score                       :452-464
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace smx {

struct Scoring { int go, ge, ma, mi; };

inline char op_char(uint32_t code)
{
    switch (code) { case 1: return 'I'; case 2: return 'D'; case 4: return 'S'; case 5: return 'H'; case 7: return '='; case 8: return 'X'; default: return '?'; }
}
inline uint32_t op_code(char c)
{
    switch (c) { case 'I': return 1; case 'D': return 2; case 'S': return 4; case 'H': return 5; case '=': return 7; case 'X': return 8; default: return 15; }
}

inline void expand_runs(const uint32_t *runs, int64_t n, std::string &out)
{
    for (int64_t i = 0; i < n; ++i) out.append((size_t)(runs[i] >> 4), op_char(runs[i] & 15u));
}

inline void encode_runs(const std::string &cig, std::vector<uint32_t> &out)
{
    out.clear();
    size_t i = 0;
    while (i < cig.size()) {
        size_t j = i;
        while (j < cig.size() && cig[j] == cig[i]) ++j;
        out.push_back((uint32_t)((j - i) << 4) | op_code(cig[i]));
        i = j;
    }
}

inline int64_t count_of(const std::string &s, size_t from, size_t to, const char *letters)
{
    int64_t n = 0;
    for (size_t i = from; i < to; ++i)
        for (const char *l = letters; *l; ++l) if (s[i] == *l) { ++n; break; }
    return n;
}

// set_soft_clipping_core: keep the prefix up to the first maximum of the running affine score, turn
// the rest into S (query columns) + H (reference columns). Returns the number of aligned reference bases.
inline int64_t soft_clip(std::string &cig, const Scoring &sc)
{
    int64_t cur = 0, best = 0, best_i = -1;
    char prev = 0;
    for (size_t i = 0; i < cig.size(); ++i) {
        const char L = cig[i];
        if (L == '=') cur += sc.ma;
        else if (L == 'X') cur += sc.mi;
        else cur -= (prev == L) ? sc.ge : sc.go;  // I or D
        if (best_i < 0 || cur > best) { best = cur; best_i = (int64_t)i; }
        prev = L;
    }
    const size_t total = cig.size();
    if (best < 0) {
        const int64_t nq = (int64_t)total - count_of(cig, 0, total, "D");
        const int64_t nr = (int64_t)total - count_of(cig, 0, total, "I");
        cig.assign((size_t)nq, 'S');
        cig.append((size_t)nr, 'H');
        return 0;
    }
    const size_t keep = (size_t)best_i + 1;
    const int64_t n_s = (int64_t)(total - keep) - count_of(cig, keep, total, "D");
    const int64_t n_h = (int64_t)(total - keep) - count_of(cig, keep, total, "I");
    const int64_t aligned_ref = (int64_t)keep - count_of(cig, 0, keep, "I");
    cig.resize(keep);
    cig.append((size_t)n_s, 'S');
    cig.append((size_t)n_h, 'H');
    return aligned_ref;
}

struct SgResult {
    std::string cig;
    int64_t beg_ref = 0, end_ref = 0;
};

// my_special_sg_qe_de (:60-73): clip from the front
inline void finish_qe_de(SgResult &r, const Scoring &sc) { r.end_ref = soft_clip(r.cig, sc) - 1; }

// my_special_sg_qb_db (:74-87): clip on the reversed string
inline void finish_qb_db(SgResult &r, const Scoring &sc)
{
    std::string rev(r.cig.rbegin(), r.cig.rend());
    soft_clip(rev, sc);
    r.beg_ref = count_of(rev, 0, rev.size(), "H");
    r.cig.assign(rev.rbegin(), rev.rend());
}

// generate_msa reduced to my_cigar_list_aligned for two rows.  Each row is read as `c + "Z" + c.back()` (the sentinel
// of generate_msa and the letter Python's index -1 wraps to) without building that string; the outputs are written
// into pre-sized buffers (this runs once per matched pair-strand over ~10 k columns and was half of the host CPU time).
inline bool aligned_columns(const std::string &c1, const std::string &c2, size_t n_ref, size_t nq1, size_t nq2,
                            std::string &a1, std::string &a2)
{
    if (c1.empty() || c2.empty()) return false;
    const std::string *c[2] = {&c1, &c2};
    const size_t len[2] = {c1.size(), c2.size()};
    auto at = [&](int i, size_t x) -> char { return x < len[i] ? (*c[i])[x] : (x == len[i] ? 'Z' : c[i]->back()); };
    const size_t nq[2] = {nq1, nq2};
    size_t ci[2] = {0, 0}, qi[2] = {0, 0}, ri = 0, w = 0;
    const size_t cap = c1.size() + c2.size() + 2;  // every output column consumes a letter of at least one row
    a1.resize(cap); a2.resize(cap);
    char *out[2] = {&a1[0], &a2[0]};
    auto filler = [&](int i) -> char {
        const size_t x = ci[i];
        const char cur = at(i, x);
        const char before = x == 0 ? c[i]->back() : at(i, x - 1);  // Python index -1 wraps to the appended last letter
        return (cur == 'H' || before == 'H' || (cur == 'Z' && (*c[i])[0] == 'H')) ? 'O' : 'N';
    };
    for (;;) {
        const char col[2] = {at(0, ci[0]), at(1, ci[1])};
        const bool has_i = col[0] == 'I' || col[1] == 'I', has_s = col[0] == 'S' || col[1] == 'S';
        if (w >= cap) return false;
        if (!has_i && !has_s) {
            if (ri == n_ref) {  // the sentinel 'Z' of the reference
                if (!(col[0] == 'Z' && col[1] == 'Z' && qi[0] == nq[0] && qi[1] == nq[1])) return false;
                a1.resize(w); a2.resize(w);
                return true;
            }
            for (int i = 0; i < 2; ++i) {
                const char L = col[i];
                if (L == '=' || L == 'X') ++qi[i];
                else if (L != 'D' && L != 'H') return false;
                out[i][w] = L;
                ++ci[i];
            }
            ++ri; ++w;
        } else {
            const char want = has_s ? 'S' : 'I';
            for (int i = 0; i < 2; ++i) {
                if (col[i] == want) { out[i][w] = want; ++ci[i]; ++qi[i]; }
                else out[i][w] = filler(i);
            }
            ++w;
        }
    }
}

// my_special_DP: join the clipped sg_qe_de result of (q1, ref) with the clipped sg_qb_db result of (q2, ref)
inline bool merge_special(const SgResult &r1, const SgResult &r2, size_t n_ref, size_t nq1, size_t nq2, const Scoring &sc, std::string &out)
{
    const int64_t gap = r2.beg_ref - r1.end_ref - 1;
    if (gap > 0) {
        const size_t h1 = (size_t)count_of(r1.cig, 0, r1.cig.size(), "H");
        const size_t h2 = (size_t)count_of(r2.cig, 0, r2.cig.size(), "H");
        out.assign(r1.cig, 0, r1.cig.size() - h1);
        out.append((size_t)gap, 'H');
        out.append(r2.cig, h2, std::string::npos);
        return true;
    }
    thread_local std::string a1, a2;
    thread_local std::vector<int32_t> left, right;
    if (!aligned_columns(r1.cig, r2.cig, n_ref, nq1, nq2, a1, a2)) return false;
    const size_t L = a1.size();
    if (a2.size() != L) return false;
    // left[x] = score of the first x columns of row 1; right[x] = score of the LAST x columns of row 2 read from the
    // end (cumsum_my_cigar_scores on the reversed string: a gap run is charged its open at the column nearest the end)
    left.resize(L + 1); right.resize(L + 1);
    {
        int32_t acc = 0; char prev = 0;
        left[0] = 0;
        for (size_t i = 0; i < L; ++i) {
            const char c = a1[i];
            if (c == '=') acc += sc.ma; else if (c == 'X') acc += sc.mi; else if (c == 'D' || c == 'I') acc -= (prev == c) ? sc.ge : sc.go;
            left[i + 1] = acc; prev = c;
        }
        acc = 0; prev = 0;
        right[0] = 0;
        for (size_t i = 0; i < L; ++i) {
            const char c = a2[L - 1 - i];
            if (c == '=') acc += sc.ma; else if (c == 'X') acc += sc.mi; else if (c == 'D' || c == 'I') acc -= (prev == c) ? sc.ge : sc.go;
            right[i + 1] = acc; prev = c;
        }
    }
    size_t cut = 0;
    int64_t best = 0;
    for (size_t x = 0; x <= L; ++x) {
        const int64_t v = (int64_t)left[x] + right[L - x];
        if (x == 0 || v > best) { best = v; cut = x; }
    }
    const int64_t n_s1 = (int64_t)(L - cut) - count_of(a1, cut, L, "DNHO");
    const int64_t n_s2 = (int64_t)cut - count_of(a2, 0, cut, "DNHO");
    out.clear();
    out.reserve(L + (size_t)(n_s1 + n_s2));
    for (size_t x = 0; x < cut; ++x) if (a1[x] != 'N' && a1[x] != 'O') out.push_back(a1[x]);
    out.append((size_t)(n_s1 + n_s2), 'S');
    for (size_t x = cut; x < L; ++x) if (a2[x] != 'N' && a2[x] != 'O') out.push_back(a2[x]);
    return true;
}

// get_aligned_ref_seq_offset / get_aligned_query_seq_offset: fixed-point on the tail of the column string
inline int64_t aligned_offset(const std::string &cig, int64_t off, const char *letters)
{
    if (off == 0) return 0;
    const int64_t n = (int64_t)cig.size();
    auto tail_count = [&](int64_t a) { return count_of(cig, (size_t)(a >= n ? 0 : n - a), (size_t)n, letters); };
    int64_t a = off, prev = 0, cnt = tail_count(a);
    while (off != a - cnt) {
        a += cnt - prev;
        prev = cnt;
        cnt = tail_count(a);
    }
    return a;
}

// Python slice cig[-a:-b] (b == 0 means "to the end"); negative starts clamp at 0
inline int64_t count_dh_tail(const std::string &cig, int64_t a, int64_t b)
{
    const int64_t n = (int64_t)cig.size();
    int64_t from = n - a, to = n - b;
    if (from < 0) from = 0;
    if (to < 0) to = 0;
    if (to <= from) return 0;
    return count_of(cig, (size_t)from, (size_t)to, "DH");
}

inline int64_t rotate_to_ref0(std::string &cig, int64_t ref_off, int64_t q_off)
{
    const int64_t a_r = aligned_offset(cig, ref_off, "IS");
    const int64_t a_q = aligned_offset(cig, q_off, "DH");
    const int64_t diff = a_q - a_r;
    int64_t new_off;
    if (a_q > a_r) new_off = diff - count_dh_tail(cig, a_q, a_r);
    else if (a_q != 0) new_off = diff + count_dh_tail(cig, a_r, a_q);
    else if (a_r != 0) new_off = diff + count_dh_tail(cig, a_r, 0);
    else new_off = 0;
    const int64_t n = (int64_t)cig.size();
    if (a_r > 0) {
        const size_t from = (size_t)(a_r >= n ? 0 : n - a_r);  // cig[-a_r:] + cig[:-a_r]
        std::string rot = cig.substr(from) + cig.substr(0, from);
        cig.swap(rot);
    }
    return new_off;
}

inline int64_t cigar_score(const std::string &cig, const Scoring &sc)
{
    int64_t s = 0;
    char prev = 0;
    for (char L : cig) {
        if (L == '=') s += sc.ma;
        else if (L == 'X') s += sc.mi;
        else if (L == 'D' || L == 'I') s -= (prev == L) ? sc.ge : sc.go;
        prev = L;
    }
    return s;
}


// ---------------------------------------------------------------------------------------------
// run-length versions (uint32 len << 4 | op); adjacent runs of one letter are always merged, so a
// vector of runs is the maximal-run decomposition that set_cigar_score / MyResult.cigar iterate over
// ---------------------------------------------------------------------------------------------
typedef std::vector<uint32_t> Runs;

inline void push_run(Runs &v, uint32_t op, int64_t len)
{
    if (len <= 0) return;
    if (!v.empty() && (v.back() & 15u) == op) v.back() += (uint32_t)(len << 4);
    else v.push_back((uint32_t)(len << 4) | op);
}

inline void append_runs(Runs &v, const uint32_t *r, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) push_run(v, r[i] & 15u, (int64_t)(r[i] >> 4));
}

// append_runs for an array that is itself a maximal-run decomposition without empty runs (what the traceback kernel
// emits per problem, smx_traceback.cuh: SmxWalk::emit merges equal neighbours and drops empty runs; a clipped result
// is a contiguous slice of such an array): only the first run can merge with what is already there.
inline void append_maximal_runs(Runs &v, const uint32_t *r, int64_t n)
{
    if (n <= 0) return;
    push_run(v, r[0] & 15u, (int64_t)(r[0] >> 4));
    if (n > 1) v.insert(v.end(), r + 1, r + n);
}

inline int64_t total_columns(const Runs &v)
{
    int64_t n = 0;
    for (uint32_t r : v) n += r >> 4;
    return n;
}

// columns among the last `a` columns whose op is in `mask` (bit per op code); a beyond the length clamps
inline int64_t tail_count(const Runs &v, int64_t a, uint32_t mask)
{
    int64_t cnt = 0;
    for (size_t i = v.size(); i-- > 0 && a > 0;) {
        const int64_t len = v[i] >> 4, take = len < a ? len : a;
        if ((mask >> (v[i] & 15u)) & 1u) cnt += take;
        a -= take;
    }
    return cnt;
}

inline int64_t aligned_offset_runs(const Runs &v, int64_t off, uint32_t mask)
{
    if (off == 0) return 0;
    int64_t a = off, prev = 0, cnt = tail_count(v, a, mask);
    while (off != a - cnt) {
        a += cnt - prev;
        prev = cnt;
        cnt = tail_count(v, a, mask);
    }
    return a;
}

// set_ref_seq_offset_as_0 on runs: rotates in place, returns new_query_seq_offset.
// The reference asks its fixed-point questions about tails of the column string ("the shortest tail that holds ref_off
// reference letters", "... q_off query letters", "how many D / H columns lie in those tails"); all of them are answered by ONE
// backward scan over the runs that stops as soon as both tails are found, and the scan also leaves the run the cut falls
// into.  The suffix-sum checkpoints of the previous version cost a full pass plus a dozen searches with unpredictable branches
// on the op codes: 6.0 against 2.9 us per 1 200-run pair-strand (scratch/bench_rotate.cpp).
inline int64_t rotate_to_ref0_runs(Runs &v, int64_t ref_off, int64_t q_off)
{
    if (ref_off == 0 && q_off == 0) return 0;
    // tail = the last t columns; c0 / c1 = its I/S resp. D/H columns.  a_r = smallest t with t - c0 == ref_off, a_q = smallest
    // t with t - c1 == q_off (get_aligned_*_seq_offset); dh_r / dh_q = D/H columns among the last a_r resp. a_q columns
    // (op codes: I 1, D 2, S 4, H 5, = 7, X 8).  The loop body has no data-dependent branch: the two "found" tests flip once.
    static const int64_t kRef[16] = {0, 0, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0};   // the column holds a reference letter (not I / S)
    static const int64_t kQry[16] = {0, 1, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0};   // ... a query letter (not D / H)
    int64_t t = 0, nr = 0, nq = 0;   // tail length, its reference letters (t - c0) and query letters (t - c1)
    int64_t a_r = ref_off == 0 ? 0 : -1, a_q = q_off == 0 ? 0 : -1, dh_r = 0, dh_q = 0;
    size_t idx = v.size();   // run the cut falls into (or the first run of the tail), `part` leading columns of it stay in front
    int64_t part = 0;
    for (size_t i = v.size(); i-- > 0;) {
        const int64_t len = v[i] >> 4;
        const uint32_t op = v[i] & 15u;
        const int64_t nr2 = nr + len * kRef[op], nq2 = nq + len * kQry[op];
        if (a_r < 0 && nr2 >= ref_off) {   // a run of reference letters reaches the offset: the tail takes `need` of its columns
            const int64_t need = ref_off - nr;
            a_r = t + need; dh_r = (t - nq) + (kQry[op] ? 0 : need); idx = i; part = len - need;
            if (a_q >= 0) break;
        }
        if (a_q < 0 && nq2 >= q_off) {     // a run of query letters (none of them D / H) reaches the offset
            a_q = t + (q_off - nq); dh_q = t - nq;
            if (a_r >= 0) break;
        }
        nr = nr2; nq = nq2; t += len;
    }
    const int64_t c1 = t - nq;
    if (a_r < 0) { a_r = t; dh_r = c1; idx = 0; part = 0; }   // an offset beyond the sequence (callers never pass one): the whole string
    if (a_q < 0) { a_q = t; dh_q = c1; }
    const int64_t diff = a_q - a_r;
    int64_t new_off;
    if (a_q > a_r) new_off = diff - (dh_q - dh_r);
    else if (a_q != 0) new_off = diff + (dh_r - dh_q);
    else if (a_r != 0) new_off = diff + dh_r;
    else new_off = 0;
    // v is maximal, so after cutting only the wrap-around junction (old last run, old first run) can merge; everything else
    // moves as blocks: [tail of v[idx]] v[idx+1 ..] + v[0] v[1 .. idx-1] [head of v[idx]]
    if (a_r > 0 && (idx > 0 || part > 0)) {
        thread_local Runs tl_rot;
        Runs &r = tl_rot;
        r.clear();
        r.reserve(v.size() + 2);
        size_t start = idx;
        if (part > 0) { r.push_back((uint32_t)(((int64_t)(v[idx] >> 4) - part) << 4) | (v[idx] & 15u)); start = idx + 1; }
        r.insert(r.end(), v.begin() + (std::ptrdiff_t)start, v.end());
        if (idx > 0) {
            push_run(r, v[0] & 15u, (int64_t)(v[0] >> 4));
            r.insert(r.end(), v.begin() + 1, v.begin() + (std::ptrdiff_t)idx);
        }
        if (part > 0) push_run(r, v[idx] & 15u, part);
        v.swap(r);   // the old array becomes the next call's scratch
    }
    return new_off;
}

inline int64_t cigar_score_runs(const Runs &v, const Scoring &sc)
{
    // per op: score = per_col[op] * len + per_run[op]  ('=' 7, 'X' 8, gaps 1 / 2: -(go + ge (len - 1)); S, H, N free)
    int64_t per_col[16] = {0}, per_run[16] = {0};
    per_col[7] = sc.ma; per_col[8] = sc.mi;
    per_col[1] = per_col[2] = -sc.ge; per_run[1] = per_run[2] = -(sc.go - sc.ge);
    int64_t s = 0;
    for (uint32_t r : v) s += per_col[r & 15u] * (int64_t)(r >> 4) + per_run[r & 15u];
    return s;
}

// a semi-global result after device-side re-clipping: kept aligned runs (forward order) + counts
struct ClippedSg {
    const uint32_t *kept = nullptr;
    int64_t n_kept = 0, n_s = 0, n_h = 0;
};

// One stretch of columns of the two-row alignment of r1 and r2 against the reference (generate_msa for two rows):
// both rows hold one letter each over `len` columns.  Op codes as in the runs; FILL = the N / O filler columns.
struct MergeSeg { uint32_t o1, o2; int64_t len; };
static const uint32_t MERGE_FILL = 15u;

// score of the first t columns of a stretch of letter `op` (t <= len) when the neighbouring column on the charged side
// holds `nb`: cumsum_my_cigar_scores (:164-185) charges a gap run its open once, at its first column in reading order
inline int64_t stretch_score(uint32_t op, uint32_t nb, int64_t t, const Scoring &sc)
{
    if (t <= 0) return 0;
    switch (op) {
    case 7: return (int64_t)sc.ma * t;
    case 8: return (int64_t)sc.mi * t;
    case 1: case 2: return -((int64_t)(nb == op ? sc.ge : sc.go) + (int64_t)sc.ge * (t - 1));
    default: return 0;
    }
}

// The overlap branch of my_special_DP (:124-152) on runs.  The two-row generate_msa (modules/msa.py:882-982) only ever
// emits stretches of constant (row 1 letter, row 2 letter); over such a stretch the left prefix score of row 1 and the
// right suffix score of row 2 are linear except for the open charged at the first (row 1) resp. last (row 2) column of
// a gap run, so left + right is linear between cut 1 and cut len - 1 of a stretch and the FIRST maximum (np.argmax,
// :136) lies on one of the cuts 0, 1, len - 1 of some stretch or at the very end.  O(runs) instead of O(columns); the
// column-level version above (merge_special) is kept as the cross-check (tests/test_stitch_cpu.py).
inline bool merge_overlap_runs(const ClippedSg &r1, const ClippedSg &r2, int64_t n_ref, int64_t nq1, int64_t nq2, const Scoring &sc, Runs &out)
{
    // row 1 = kept1, S * n_s1, H * n_h1;  row 2 = H * n_h2, S * n_s2, kept2
    struct Cursor {
        const uint32_t *kept; int64_t n_kept; int64_t extra[2]; uint32_t extra_op[2]; bool extras_first;
        int64_t pos = 0, left = 0; uint32_t op = 0; bool done = false;
        int64_t n_items() const { return n_kept + 2; }
        void load()
        {
            for (;; ++pos) {
                if (pos >= n_items()) { done = true; op = 0; left = 0; return; }
                const int64_t e = extras_first ? pos : pos - n_kept;  // index into the extras, if in range
                if (e >= 0 && e < 2) { op = extra_op[e]; left = extra[e]; }
                else { const uint32_t r = kept[extras_first ? pos - 2 : pos]; op = r & 15u; left = (int64_t)(r >> 4); }
                if (left > 0) return;
            }
        }
        void take(int64_t n) { left -= n; if (left == 0) { ++pos; load(); } }
    };
    Cursor c1{r1.kept, r1.n_kept, {r1.n_s, r1.n_h}, {4u, 5u}, false};
    Cursor c2{r2.kept, r2.n_kept, {r2.n_h, r2.n_s}, {5u, 4u}, true};
    c1.load(); c2.load();
    if (c1.done || c2.done) return false;
    thread_local std::vector<MergeSeg> tl_segs;
    thread_local std::vector<int64_t> tl_suf;
    std::vector<MergeSeg> &segs = tl_segs;  // one TLS lookup per call, not per stretch
    std::vector<int64_t> &suf = tl_suf;
    segs.clear();
    segs.reserve((size_t)(r1.n_kept + r2.n_kept + 8));
    int64_t ref1 = 0, ref2 = 0, q1 = 0, q2 = 0;
    while (!c1.done || !c2.done) {
        const uint32_t a = c1.done ? 0u : c1.op, b = c2.done ? 0u : c2.op;
        const bool has_s = a == 4u || b == 4u, has_i = a == 1u || b == 1u;
        if (has_s || has_i) {
            const uint32_t want = has_s ? 4u : 1u;
            if (a == want && b == want) {
                const int64_t n = c1.left < c2.left ? c1.left : c2.left;
                segs.push_back({want, want, n}); q1 += n; q2 += n; c1.take(n); c2.take(n);
            } else if (a == want) { const int64_t n = c1.left; segs.push_back({want, MERGE_FILL, n}); q1 += n; c1.take(n); }
            else { const int64_t n = c2.left; segs.push_back({MERGE_FILL, want, n}); q2 += n; c2.take(n); }
        } else {
            if (c1.done || c2.done) return false;  // one row ran out of reference columns before the other
            const int64_t n = c1.left < c2.left ? c1.left : c2.left;
            segs.push_back({a, b, n});
            if (a == 7u || a == 8u) q1 += n;
            if (b == 7u || b == 8u) q2 += n;
            ref1 += n; ref2 += n;
            c1.take(n); c2.take(n);
        }
    }
    if (ref1 != n_ref || ref2 != n_ref || q1 != nq1 || q2 != nq2) return false;
    const size_t ns = segs.size();
    // suf[k] = score of row 2 over the stretches k.. (read from the end)
    suf.assign(ns + 1, 0);
    for (size_t k = ns; k-- > 0;) {
        const uint32_t nb = k + 1 < ns ? segs[k + 1].o2 : 0u;
        suf[k] = suf[k + 1] + stretch_score(segs[k].o2, nb, segs[k].len, sc);
    }
    // first maximum of left[x] + right[x] over the cuts x = 0 .. L
    int64_t best = suf[0], best_k = 0, best_t = 0, left = 0;
    for (size_t k = 0; k < ns; ++k) {
        const MergeSeg &g = segs[k];
        const uint32_t pb = k > 0 ? segs[k - 1].o1 : 0u, nb = k + 1 < ns ? segs[k + 1].o2 : 0u;
        const int64_t cand[3] = {0, 1, g.len - 1};
        int64_t last_t = -1;
        for (int c = 0; c < 3; ++c) {
            const int64_t t = cand[c];
            if (t <= last_t || t >= g.len) continue;
            last_t = t;
            const int64_t v = left + stretch_score(g.o1, pb, t, sc) + stretch_score(g.o2, nb, g.len - t, sc) + suf[k + 1];
            if (v > best) { best = v; best_k = (int64_t)k; best_t = t; }
        }
        left += stretch_score(g.o1, pb, g.len, sc);
    }
    if (left > best) { best = left; best_k = (int64_t)ns; best_t = 0; }  // cut = L: all of row 1
    // row 1 up to the cut, the clipped query letters of both rows as S, row 2 from the cut (fillers dropped)
    auto consumes_query = [](uint32_t op) { return op == 7u || op == 8u || op == 1u || op == 4u; };
    int64_t n_s = 0;
    for (size_t k = 0; k < ns; ++k) {
        const MergeSeg &g = segs[k];
        const int64_t before = (int64_t)k < best_k ? g.len : ((int64_t)k == best_k ? best_t : 0);
        if (consumes_query(g.o2)) n_s += before;
        if (consumes_query(g.o1)) n_s += g.len - before;
    }
    for (size_t k = 0; k < ns; ++k) {
        const MergeSeg &g = segs[k];
        const int64_t before = (int64_t)k < best_k ? g.len : ((int64_t)k == best_k ? best_t : 0);
        if (before == 0) break;
        if (g.o1 != MERGE_FILL) push_run(out, g.o1, before);
    }
    push_run(out, 4u, n_s);
    for (size_t k = (size_t)(best_k < (int64_t)ns ? best_k : (int64_t)ns); k < ns; ++k) {
        const MergeSeg &g = segs[k];
        const int64_t after = (int64_t)k == best_k ? g.len - best_t : g.len;
        if (g.o2 != MERGE_FILL) push_run(out, g.o2, after);
    }
    return true;
}

// the same branch through the column strings (the round's first implementation, kept as the cross-check)
inline bool merge_overlap_chars(const ClippedSg &r1, const ClippedSg &r2, int64_t n_ref, int64_t nq1, int64_t nq2, const Scoring &sc, Runs &out)
{
    SgResult a, b;
    expand_runs(r1.kept, r1.n_kept, a.cig);
    a.cig.append((size_t)r1.n_s, 'S');
    a.cig.append((size_t)r1.n_h, 'H');
    a.end_ref = n_ref - r1.n_h - 1;
    b.cig.assign((size_t)r2.n_h, 'H');
    b.cig.append((size_t)r2.n_s, 'S');
    expand_runs(r2.kept, r2.n_kept, b.cig);
    b.beg_ref = r2.n_h;
    std::string merged;
    if (!merge_special(a, b, (size_t)n_ref, (size_t)nq1, (size_t)nq2, sc, merged)) return false;
    size_t i = 0;
    while (i < merged.size()) {
        size_t j = i;
        while (j < merged.size() && merged[j] == merged[i]) ++j;
        push_run(out, op_code(merged[i]), (int64_t)(j - i));
        i = j;
    }
    return true;
}

// my_special_DP on re-clipped results; r1 = sg_qe_de(q1, ref) = kept1 + S*n_s1 + H*n_h1,
// r2 = sg_qb_db(q2, ref) = H*n_h2 + S*n_s2 + kept2.  Appends the merged column runs to `out`.
inline bool merge_special_runs(const ClippedSg &r1, const ClippedSg &r2, int64_t n_ref, int64_t nq1, int64_t nq2, const Scoring &sc, Runs &out,
                               bool column_strings = false)
{
    const int64_t gap = r2.n_h + r1.n_h - n_ref;  // r2.beg_ref - r1.end_ref - 1
    if (gap > 0) {
        append_maximal_runs(out, r1.kept, r1.n_kept);
        push_run(out, 4, r1.n_s);
        push_run(out, 5, gap);
        push_run(out, 4, r2.n_s);
        append_maximal_runs(out, r2.kept, r2.n_kept);
        return true;
    }
    return column_strings ? merge_overlap_chars(r1, r2, n_ref, nq1, nq2, sc, out) : merge_overlap_runs(r1, r2, n_ref, nq1, nq2, sc, out);
}

}  // namespace smx
