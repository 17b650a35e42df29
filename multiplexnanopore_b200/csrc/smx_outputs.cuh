// Row N2: what consumes the alignment stage's arrays without one Python object per pair-strand.
//
//  smx_assign_kernel   msa.QueryAssignment.__init__ / set_assignment (savemoney/modules/msa.py:31-64) as one thread per
//                      read over its 2 * N_ref scores: float64 normalisation score / len(ref) (IEEE division, the same
//                      bits as numpy's true_divide), the per-plasmid summary, and (classified_ref_seq_idx,
//                      is_reverse_compliment, assigned).
//  smx_ir_format       the `resultK(dict)` blocks of <stem>.intermediate_results.txt -- IntermediateResults.__init__ +
//                      MyTextFormat.to_text (post_analysis/post_analysis_core.py:89-128, modules/my_classes.py:43-74) over
//                      MyResult.to_dict (modules/ref_query_alignment.py:512-518) -- written straight from the run-length
//                      CIGAR arrays by host threads.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

struct SmxAssignArgs {
    const int32_t *score;     // [n_reads, 2 R]
    const uint8_t *status;    // [n_reads, 2 R]  0 = aligned, else empty MyResult()
    const int32_t *ref_len;   // [R]
    double *normalized;       // [n_reads, 2 R]  NaN where empty (msa.py:38-39)
    double *summary;          // [n_reads, R]    max over the strands, empty / negative -> 0 (msa.py:41-43)
    int32_t *assignment;      // [n_reads, 3]
    int32_t n_reads, n_refs;
    double threshold;
};

__global__ void __launch_bounds__(128) smx_assign_kernel(const SmxAssignArgs a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.n_reads) return;
    const int n2 = 2 * a.n_refs;
    const int32_t *sc = a.score + (size_t)q * n2;
    const uint8_t *st = a.status + (size_t)q * n2;
    double best = 0.0;
    int best_k = -1, n_best = 0;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int r = 0; r < a.n_refs; ++r) {
        const double len = (double)a.ref_len[r];
        double s_max = 0.0;   // "no result on either strand" and negative values both read 0 in the summary
        for (int s = 0; s < 2; ++s) {
            const int k = 2 * r + s;
            if (st[k] != 0) { a.normalized[(size_t)q * n2 + k] = nan; continue; }
            const double v = __ddiv_rn((double)sc[k], len);
            a.normalized[(size_t)q * n2 + k] = v;
            if (v > s_max) s_max = v;
            // first maximum in index order (np.ma argmax), and how many entries share it (msa.py:55-59)
            if (best_k < 0 || v > best) { best = v; best_k = k; n_best = 1; }
            else if (v == best) ++n_best;
        }
        a.summary[(size_t)q * a.n_refs + r] = s_max;
    }
    int32_t *o = a.assignment + (size_t)q * 3;
    if (best_k < 0) { o[0] = -1; o[1] = -1; o[2] = -1; }              // no result at all: undefined (msa.py:60, :62)
    else if (n_best > 1) { o[0] = -1; o[1] = -1; o[2] = 0; }            // tied maxima: neither classified nor assigned
    else { o[0] = best_k / 2; o[1] = best_k % 2; o[2] = best > a.threshold ? 1 : 0; }
}

namespace smx_out {

inline int dec_len(uint32_t v) { int n = 1; while (v >= 10) { v /= 10; ++n; } return n; }
inline char *put_dec(char *p, uint32_t v)
{
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
inline char *put_int(char *p, int64_t v)
{
    if (v < 0) { *p++ = '-'; v = -v; }
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
inline int int_len(int64_t v) { int n = v < 0 ? 1 : 0; if (v < 0) v = -v; do { ++n; v /= 10; } while (v); return n; }

static const char kOpLetter[16] = {'?', 'I', 'D', '?', 'S', 'H', '?', '=', 'X', '?', '?', '?', '?', '?', '?', '?'};
static const int32_t kNone = INT32_MIN;

// bytes of "# result<idx>(dict)\ncigar\t<RLE>\nscore\t<s>\nnew_query_seq_offset\t<o>\n\n"
inline int64_t block_len(int64_t idx, const uint32_t *runs, int64_t n_runs, bool empty, int32_t score, int32_t off)
{
    int64_t n = 8 + int_len(idx) + 7;                    // "# result" idx "(dict)\n"
    n += 6;                                              // "cigar\t"
    if (!empty) for (int64_t r = 0; r < n_runs; ++r) n += dec_len(runs[r] >> 4) + 1;
    n += 1 + 6 + int_len(empty ? 0 : score) + 1;         // "\n" "score\t" s "\n"
    n += 21 + ((empty || off == kNone) ? 4 : int_len(off)) + 2;   // "new_query_seq_offset\t" o "\n\n"
    return n;
}

inline char *block_put(char *p, int64_t idx, const uint32_t *runs, int64_t n_runs, bool empty, int32_t score, int32_t off)
{
    memcpy(p, "# result", 8); p += 8;
    p = put_int(p, idx);
    memcpy(p, "(dict)\ncigar\t", 13); p += 13;
    if (!empty) for (int64_t r = 0; r < n_runs; ++r) { p = put_dec(p, runs[r] >> 4); *p++ = kOpLetter[runs[r] & 15]; }
    memcpy(p, "\nscore\t", 7); p += 7;
    p = put_int(p, empty ? 0 : score);
    memcpy(p, "\nnew_query_seq_offset\t", 22); p += 22;
    if (empty || off == kNone) { memcpy(p, "None", 4); p += 4; }
    else p = put_int(p, off);
    *p++ = '\n'; *p++ = '\n';
    return p;
}

}  // namespace smx_out
