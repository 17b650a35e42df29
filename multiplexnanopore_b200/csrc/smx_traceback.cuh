// Second kernel of the DP path: resolves the packed direction bits written by smx_dp_forward
// into run-length CIGARs.  Restates parasail's traceback state machine (cigar_template.c,
// SURVEY.md Appendix A.4) as used through MyResult.__init__ (ref_query_alignment.py:426-433).
#pragma once
#include "smx_common.cuh"
#include "../../include/smx.h"

struct SmxTbArgs {
    const uint8_t *pool;
    const SmxProblem *probs;
    const int32_t *order;  // problems of this chunk
    int32_t n_order;
    const uint8_t *trace;
    SmxResult *results;
    uint32_t *runs;  // scratch, reversed order per problem
};

__device__ __forceinline__ uint32_t smx_trace_nibble(const uint8_t *tr, int n, int kshift, int i, int j)
{
    const int K = 1 << kshift;
    const int w = K >= 2 ? K >> 1 : 1;  // bytes per lane-step word
    const int rows_per_band = 32 << kshift;
    const int band = i / rows_per_band;
    const int rin = i - band * rows_per_band;
    const int t = rin >> kshift;
    const int k = rin & (K - 1);
    const size_t word = ((size_t)band * (size_t)(n + 31) + (size_t)(j + t)) * 32 + (size_t)t;
    const uint32_t byte = tr[word * w + (k >> 1)];
    return (byte >> ((k & 1) * 4)) & 15u;
}

// one thread per problem; runs are emitted while walking back, i.e. in reverse order
__global__ void __launch_bounds__(128) smx_traceback(const SmxTbArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_order) return;
    const int pid = a.order[t];
    const SmxProblem pr = a.probs[pid];
    SmxResult r = a.results[pid];
    const uint8_t *s1 = a.pool + pr.s1_off;
    const uint8_t *s2 = a.pool + pr.s2_off;
    const uint8_t *tr = a.trace + pr.trace_off;
    uint32_t *out = a.runs + pr.cigar_off;
    const int m = pr.m, n = pr.n;
    const bool local = pr.mode == SMX_MODE_SW;

    int nruns = 0;
    uint32_t cur_op = 0, cur_len = 0;
#define SMX_EMIT(op_, len_)                                             \
    do {                                                                \
        const uint32_t o_ = (op_);                                      \
        const uint32_t l_ = (uint32_t)(len_);                           \
        if (l_ > 0) {                                                   \
            if (o_ == cur_op) cur_len += l_;                            \
            else {                                                      \
                if (cur_len) out[nruns++] = (cur_len << 4) | cur_op;    \
                cur_op = o_; cur_len = l_;                              \
            }                                                           \
        }                                                               \
    } while (0)

    int i = r.end_q, j = r.end_r;
    if (pr.mode == SMX_MODE_SG_QE_DE) {
        // semi-global: the CIGAR also covers the free trailing part
        if (i + 1 == m) SMX_EMIT(SMX_OP_D, n - 1 - j);
        else if (j + 1 == n) SMX_EMIT(SMX_OP_I, m - 1 - i);
    }
    int where = 0;  // 0 = DIAG (H), 1 = INS (E, emits 'D'), 2 = DEL (F, emits 'I')
    while (local ? (i >= 0 && j >= 0) : (i >= 0 || j >= 0)) {
        if (i < 0) { SMX_EMIT(SMX_OP_D, j + 1); j = -1; break; }
        if (j < 0) { SMX_EMIT(SMX_OP_I, i + 1); i = -1; break; }
        const uint32_t nib = smx_trace_nibble(tr, n, pr.kshift, i, j);
        if (where == 0) {
            const uint32_t src = nib & 3u;
            if (src == SMX_SRC_DIAG) {
                SMX_EMIT(s1[i] == s2[j] ? SMX_OP_EQ : SMX_OP_X, 1);
                --i; --j;
            } else if (src == SMX_SRC_INS) where = 1;
            else if (src == SMX_SRC_DEL) where = 2;
            else break;  // ZERO: local alignment starts here
        } else if (where == 1) {
            SMX_EMIT(SMX_OP_D, 1);
            --j;
            where = (nib & SMX_BIT_EOPEN) ? 0 : 1;
        } else {
            SMX_EMIT(SMX_OP_I, 1);
            --i;
            where = (nib & SMX_BIT_FOPEN) ? 0 : 2;
        }
    }
    if (cur_len) out[nruns++] = (cur_len << 4) | cur_op;
#undef SMX_EMIT
    r.beg_q = local ? i + 1 : 0;
    r.beg_r = local ? j + 1 : 0;
    r.n_runs = nruns;
    a.results[pid] = r;
}

// gathers the per-problem reversed runs into one dense forward-order array
// dense[off[p] + x] = runs[cigar_off[p] + n_runs[p] - 1 - x]; one warp per problem.
__global__ void __launch_bounds__(256) smx_compact_runs(const SmxProblem *probs, const SmxResult *results,
                                                        const int64_t *dense_off, const uint32_t *runs,
                                                        uint32_t *dense, int32_t n)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int nr = results[w].n_runs;
    const uint32_t *src = runs + probs[w].cigar_off;
    uint32_t *dst = dense + dense_off[w];
    for (int x = lane; x < nr; x += 32) dst[x] = src[nr - 1 - x];
}
