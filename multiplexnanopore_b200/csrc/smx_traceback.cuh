// Second kernel of the DP path: resolves the packed direction bits written by smx_dp_forward
// into run-length CIGARs.  Restates parasail's traceback state machine (cigar_template.c,
// SURVEY.md Appendix A.4) as used through MyResult.__init__ (ref_query_alignment.py:426-433).
#pragma once
#include "smx_common.cuh"
#include "../../include/smx.h"

#ifndef SMX_TB_AHEAD
#define SMX_TB_AHEAD 16
#endif
#ifndef SMX_TB_PF
#define SMX_TB_PF 3
#endif

struct SmxTbArgs {
    const uint8_t *pool;
    const SmxProblem *probs;
    const int32_t *order;  // problems of this chunk
    int32_t n_order;
    int32_t n_big;  // the first n_big problems of `order` are long walks: one warp each (lane 0 walks), the rest one thread each
    const uint8_t *trace;
    SmxResult *results;
    uint32_t *runs;  // scratch, reversed order per problem
    SmxScoring sc;
    int32_t split_last;  // class-10 problems with an odd band count keep their last band as two K = 4 half-bands (smx_dp16h.cuh)
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

// Walk state shared by the emit helper.  Besides run-length encoding it evaluates, for problems flagged
// SMX_MODE_CLIP, the re-clipping of MyResult.set_soft_clipping(_core) (ref_query_alignment.py:562-622) while
// walking: the running affine score of the column string peaks at the end of an '=' run (match > 0,
// mismatch < 0, gaps > 0), so only run boundaries are candidates.
//   sg_qb_db (side="end"): the reference scores the REVERSED string = our walk order; keep the walk
//       prefix up to the first maximum.
//   sg_qe_de (side="beg"): the reference scores the forward string; walking backwards we minimise the
//       score of the suffix walked so far, ties resolved towards the forward-earliest cut ('<=').
struct SmxWalk {
    uint32_t *out;
    int nruns;
    uint32_t cur_op, cur_len;
    int clip;          // 0 none, 1 = sg_qe_de, 2 = sg_qb_db
    int go, ge, ma, mi;
    long long acc;     // score of the completed runs
    int q_used, r_used;
    long long best;    // qb_db: best prefix score; qe_de: minimal suffix score
    int have, cut_runs, cut_q, cut_r;

    __device__ __forceinline__ void candidate_qe_de()
    {
        if (!have || acc <= best) { best = acc; cut_runs = nruns; cut_q = q_used; cut_r = r_used; have = 1; }
    }
    __device__ __forceinline__ void flush()
    {
        if (!cur_len) return;
        out[nruns++] = (cur_len << 4) | cur_op;
        if (clip) {
            const int len = (int)cur_len;
            if (cur_op == SMX_OP_EQ) { acc += (long long)ma * len; q_used += len; r_used += len; }
            else if (cur_op == SMX_OP_X) { acc += (long long)mi * len; q_used += len; r_used += len; }
            else {
                acc -= go + (long long)ge * (len - 1);
                if (cur_op == SMX_OP_I) q_used += len; else r_used += len;
            }
            if (clip == 2 && cur_op == SMX_OP_EQ && (!have || acc > best)) { best = acc; cut_runs = nruns; cut_q = q_used; cut_r = r_used; have = 1; }
        }
        cur_len = 0;
    }
    __device__ __forceinline__ void emit(uint32_t op, uint32_t len)
    {
        if (!len) return;
        if (op == cur_op && cur_len) { cur_len += len; return; }
        flush();
        if (clip == 1 && op == SMX_OP_EQ) candidate_qe_de();
        cur_op = op; cur_len = len;
    }
};

// One walker per problem; runs are emitted while walking back, i.e. in reverse order.  A walk is a chain of
// dependent loads: 32 long walks in one warp would advance at the pace of the slowest load of every step and
// serialise on their divergent branches, so the long walks (m + n >= kTbBigWalk on the host) get a warp each.
__global__ void __launch_bounds__(128) smx_traceback(const SmxTbArgs a)
{
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int t;
    if (gt < 32ll * a.n_big) {
        if (threadIdx.x & 31) return;
        t = (int)(gt >> 5);
    } else {
        t = (int)(gt - 32ll * a.n_big) + a.n_big;
    }
    if (t >= a.n_order) return;
    const int pid = a.order[t];
    const SmxProblem pr = a.probs[pid];
    SmxResult r = a.results[pid];
    const uint8_t *s1 = a.pool + pr.s1_off;
    const uint8_t *s2 = a.pool + pr.s2_off;
    const uint8_t *tr = a.trace + pr.trace_off;
    const int m = pr.m, n = pr.n;
    const int mode = pr.mode & SMX_MODE_MASK;
    const bool local = mode == SMX_MODE_SW;
    const bool enc16 = pr.cls >= 10;

    SmxWalk w;
    w.out = a.runs + pr.cigar_off; w.nruns = 0; w.cur_op = 0; w.cur_len = 0;
    w.clip = (pr.mode & SMX_MODE_CLIP) ? mode : 0;
    w.go = a.sc.go; w.ge = a.sc.ge; w.ma = a.sc.match; w.mi = a.sc.mismatch;
    w.acc = 0; w.q_used = 0; w.r_used = 0; w.best = 0; w.have = 0; w.cut_runs = 0; w.cut_q = 0; w.cut_r = 0;

    int i = r.end_q, j = r.end_r;
    if (mode == SMX_MODE_SG_QE_DE) {
        // semi-global: the CIGAR also covers the free trailing part
        if (i + 1 == m) w.emit(SMX_OP_D, n - 1 - j);
        else if (j + 1 == n) w.emit(SMX_OP_I, m - 1 - i);
    }
    int where = 0;  // 0 = DIAG (H), 1 = INS (E, emits 'D'), 2 = DEL (F, emits 'I')
    // Local problems of the packed kernel carry no ZERO code in their direction bits (two bits are taken by "diagonal
    // lost" / "E beat F"): the walker tracks the value of the matrix cell it stands on instead -- H of the end cell is
    // the score, every move subtracts what the forward recurrence added -- and stops in state H on a cell worth 0,
    // which is exactly the cell parasail marks ZERO.
    const bool tracked_stop = local && enc16;
    int cur = r.score;
    // The walk is one dependent 1-byte load per step, so the loop keeps the position incrementally: `wp` points
    // at the lane-step word of the current cell (layout [band][step = j + lane][lane]), `k` is the row inside the
    // lane.  A column move lowers the step by one line, a row move changes k or, every K rows, the lane (one
    // more line and one word back).  Only a band crossing (every 32*K rows) recomputes the address.
    // Direction bits left L2 long ago: every move lowers the step index by 0..2, so the lines the walk is about
    // to touch are known without knowing the path and are prefetched SMX_TB_AHEAD lines ahead.
    // A split last band (smx_dp16h.cuh): rows >= row_base use the K = 4 layout inside that band's region; the walk starts
    // there (it starts in the last row or the last column) and switches to the problem's own layout when it leaves it.
    const int nb8 = (m + 255) >> 8;
    const bool split = a.split_last != 0 && pr.cls == 10 && (nb8 & 1);
    const int R0 = split ? (nb8 - 1) * 256 : 0;
    const uint8_t *const tr_split = tr + (long long)(nb8 - 1) * (long long)(n + 31) * 128;
    int kshift = pr.kshift, row_base = 0;
    const uint8_t *base = tr;
    if (split && i >= R0) { kshift = 2; row_base = R0; base = tr_split; }
    int K = 1 << kshift;
    int wbytes = K >= 2 ? K >> 1 : 1;
    int rows_per_band = 32 << kshift;
    long long line = 32ll * wbytes;
    long long band_bytes = (long long)(n + 31) * line;
    int lane_t = 0, k = 0;
    const uint8_t *wp = tr;
    if (i >= 0 && j >= 0) {
        const int ii = i - row_base;
        const int band = ii / rows_per_band;
        const int rin = ii - band * rows_per_band;
        lane_t = rin >> kshift;
        k = rin & (K - 1);
        wp = base + (long long)band * band_bytes + (long long)(j + lane_t) * line + (long long)lane_t * wbytes;
    }
    while (local ? (i >= 0 && j >= 0) : (i >= 0 || j >= 0)) {
        if (i < 0) { w.emit(SMX_OP_D, j + 1); j = -1; break; }
        if (j < 0) { w.emit(SMX_OP_I, i + 1); i = -1; break; }
#if SMX_TB_PF == 1
        asm volatile("prefetch.global.L2 [%0];" ::"l"(wp - SMX_TB_AHEAD * line));
#elif SMX_TB_PF == 2
        asm volatile("prefetch.global.L2 [%0];" ::"l"(wp - SMX_TB_AHEAD * line));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(wp - SMX_TB_AHEAD * line - 32));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(wp - (SMX_TB_AHEAD + 1) * line - 32));
#elif SMX_TB_PF == 3
        asm volatile("prefetch.global.L1 [%0];" ::"l"(wp - SMX_TB_AHEAD * line));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(wp - SMX_TB_AHEAD * line - 32));
#endif
        const uint32_t nib = ((uint32_t)wp[k >> 1] >> ((k & 1) * 4)) & 15u;
        int di = 0, dj = 0;
        if (where == 0) {
            if (tracked_stop && cur <= 0) break;  // ZERO: local alignment starts here
            // int32 kernels store the H source as a 2-bit code; the s16x2 kernels (classes >= 10) store
            // "diagonal lost" in bit 0 and "E beat F" in bit 1 (smx_dp16.cuh)
            const uint32_t src = enc16 ? (!(nib & 1u) ? SMX_SRC_DIAG : ((nib & 2u) ? SMX_SRC_INS : SMX_SRC_DEL)) : (nib & 3u);
            if (src == SMX_SRC_DIAG) {
                w.emit(s1[i] == s2[j] ? SMX_OP_EQ : SMX_OP_X, 1);
                if (tracked_stop) cur -= (smx_code(s1[i]) > 3 || smx_code(s2[j]) > 3) ? 0 : (s1[i] == s2[j] ? a.sc.match : a.sc.mismatch);
                di = 1; dj = 1;
            } else if (src == SMX_SRC_INS) where = 1;
            else if (src == SMX_SRC_DEL) where = 2;
            else break;  // ZERO: local alignment starts here
        } else if (where == 1) {
            w.emit(SMX_OP_D, 1);
            dj = 1;
            where = (nib & SMX_BIT_EOPEN) ? 0 : 1;
            if (tracked_stop) cur += where == 0 ? a.sc.go : a.sc.ge;  // E(i, j) = H(i, j-1) - open  or  E(i, j-1) - extend
        } else {
            w.emit(SMX_OP_I, 1);
            di = 1;
            where = (nib & SMX_BIT_FOPEN) ? 0 : 2;
            if (tracked_stop) cur += where == 0 ? a.sc.go : a.sc.ge;
        }
        if (dj) { --j; wp -= line; }
        if (di) {
            --i;
            if (k > 0) --k;
            else if (lane_t > 0) { --lane_t; k = K - 1; wp -= line + wbytes; }
            else if (i >= 0) {  // previous band: last lane, last row
                if (i < row_base) {  // leaving the split last band: back to the problem's own layout
                    kshift = pr.kshift; row_base = 0; base = tr;
                    K = 1 << kshift; wbytes = K >= 2 ? K >> 1 : 1; rows_per_band = 32 << kshift;
                    line = 32ll * wbytes; band_bytes = (long long)(n + 31) * line;
                }
                lane_t = 31; k = K - 1;
                wp = base + (long long)((i - row_base) / rows_per_band) * band_bytes + (long long)(j + 31) * line + 31ll * wbytes;
            }
        }
    }
    w.flush();
    r.beg_q = local ? i + 1 : 0;
    r.beg_r = local ? j + 1 : 0;
    r.n_runs = w.nruns; r.run_start = 0; r.n_s = 0; r.n_h = 0;
    if (w.clip == 2) {
        if (!w.have || w.best < 0) { r.n_runs = 0; r.n_s = m; r.n_h = n; }
        else { r.n_runs = w.cut_runs; r.n_s = m - w.cut_q; r.n_h = n - w.cut_r; }
    } else if (w.clip == 1) {
        // acc is now the score of the whole string; the best forward prefix scores acc - best
        if (!w.have || w.acc - w.best < 0) { r.n_runs = 0; r.n_s = m; r.n_h = n; }
        else { r.run_start = w.cut_runs; r.n_runs = w.nruns - w.cut_runs; r.n_s = w.cut_q; r.n_h = w.cut_r; }
    }
    a.results[pid] = r;
}

// gathers the per-problem reversed runs into one dense forward-order array
// dense[off[p] + x] = runs[cigar_off[p] + run_start[p] + n_runs[p] - 1 - x]; one warp per problem.
__global__ void __launch_bounds__(256) smx_compact_runs(const SmxProblem *probs, const SmxResult *results,
                                                        const int64_t *dense_off, const uint32_t *runs,
                                                        uint32_t *dense, int32_t n)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int nr = results[w].n_runs;
    const uint32_t *src = runs + probs[w].cigar_off + results[w].run_start;
    uint32_t *dst = dense + dense_off[w];
    for (int x = lane; x < nr; x += 32) dst[x] = src[nr - 1 - x];
}
