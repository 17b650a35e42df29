// Row A6 on the device: chain search over conserved regions, one CTA per pair-strand, one warp per origin.
// Same semantics as smx_chain.hpp (the host version, kept as fallback for > 256 regions and as a cross-check):
// SearchLongestPath.exec_circular / exec_linear / search_longest_path / get_1d_connection /
// get_connection_constraint and ConnectionMatrix (savemoney/modules/ref_query_alignment.py:649-709, 792-948;
// SURVEY.md Appendix C).
//
// get_1d_connection without the sort: every interval x has an "open" event and a "close" event (the
// first / second of its two endpoints in the order of (value with starts shifted by -0.1, flat index)).
// The sweep makes i -> j iff i closed before j opens and no close event lies between the first open after
// i's close and j's open.  With composite integer keys that is
//     f_i = min{ open_x : open_x > close_i },  k_i = min{ close_y : close_y > f_i },
//     i -> j  <=>  close_i < open_j < k_i.
#pragma once
#include "smx_common.cuh"

#define SMX_CHAIN_MAXN 256
#define SMX_CHAIN_WARPS 4

struct SmxChainTask {
    int64_t region_off;  // into the int4 region array (rs, re, qs, qe), reference order (k asc, start asc)
    int64_t trace_off;   // into the int32 trace output (capacity n)
    int32_t n, n_ref, n_q, pad;
};

struct SmxChainOut {
    int32_t trace_len;  // 0 = failed (the reference would raise)
    int32_t score;
};

struct SmxChainArgs {
    const SmxChainTask *tasks;
    const int4 *regions;
    int32_t *traces;
    SmxChainOut *out;
    int32_t *origin_scratch;  // 3 ints per region (score, g1, g2 of the search with that region as origin)
    int32_t topology;
    int32_t nmax;      // smem is sized for this many regions (multiple of 32)
    int32_t n_slices;  // CTAs per pair-strand: CTA s evaluates the origins r with r % (n_slices * warps) in its share
    int32_t phase;     // 0 = evaluate origins (and finish when n_slices == 1), 1 = pick the best origin and emit its trace
};

struct SmxChainScratch {
    int *rs, *re, *qs, *qe, *keep;        // shifted rows that survive the removal, and their original index
    int4 *nodej;                          // per node j, packed for the adjacency loop: {open_r, open_q, rs (INT_MIN if j is ignored), qs}
    int2 *close;                          // per node: {close_r, close_q}
    uint32_t *adj;                        // n x nw bit matrix
    int *score, *parent, *indeg, *list0, *list1, *trace;
    int nw;
};

__device__ __forceinline__ int smx_chain_scratch_ints(int nmax) { return 17 * nmax + nmax * (nmax / 32); }

// One search (search_longest_path) by one warp.  origin < 0: rows are used as they are (linear mode passes
// rows already shifted).  Returns trace length (0 on failure) with the trace (original row indices) in
// sc.trace and the score in *score_out.
__device__ int smx_chain_search(const SmxChainScratch &sc, const int4 *R, int N, int origin, bool wrap, int n_ref, int n_q,
                                int *score_out)
{
    const int lane = threadIdx.x & 31;
    // ---- A: shifted copy, wrap negatives, drop rows that span the rotated reference origin (:739-756) ----
    const int4 o = R[origin];
    int n = 0, n_init = 0, init = -1;
    for (int base = 0; base < N; base += 32) {
        const int idx = base + lane;
        int4 x = make_int4(0, -1, 0, 0);
        bool keep = false;
        if (idx < N) {
            x = R[idx];
            x.x -= o.x; x.y -= o.x; x.z -= o.z; x.w -= o.z;
            if (wrap) {
                if (x.x < 0) x.x += n_ref;
                if (x.y < 0) x.y += n_ref;
                if (x.z < 0) x.z += n_q;
                if (x.w < 0) x.w += n_q;
                keep = (x.y - x.x) >= 0;
            } else keep = true;
        }
        const unsigned bal = __ballot_sync(SMX_FULL_MASK, keep);
        const int pos = n + __popc(bal & ((1u << lane) - 1u));
        if (keep) { sc.rs[pos] = x.x; sc.re[pos] = x.y; sc.qs[pos] = x.z; sc.qe[pos] = x.w; sc.keep[pos] = idx; }
        const bool is_init = keep && x.x == 0 && x.z == 0;
        const unsigned ib = __ballot_sync(SMX_FULL_MASK, is_init);
        if (ib) { n_init += __popc(ib); init = n + __popc(bal & ((1u << (__ffs(ib) - 1)) - 1u)); }
        n += __popc(bal);
    }
    if (n_init != 1) return 0;  // assert at :878-879
    __syncwarp();
    const int nw = (n + 31) >> 5;
    // ---- B: composite event keys, packed per node so that the adjacency loop reads one 16-byte word per candidate ----
    for (int c = lane; c < n; c += 32) {
        const int a0 = (10 * sc.rs[c] - 1) * 512 + 2 * c, b0 = (10 * sc.re[c]) * 512 + 2 * c + 1;
        const int a1 = (10 * sc.qs[c] - 1) * 512 + 2 * c, b1 = (10 * sc.qe[c]) * 512 + 2 * c + 1;
        // a target j needs weight > 0 and a well-formed query interval (:891-894); INT_MIN makes `re_i < rs_j` fail
        const bool usable = (sc.re[c] - sc.rs[c]) > 0 && (sc.qe[c] - sc.qs[c]) >= 0;
        sc.nodej[c] = make_int4(min(a0, b0), min(a1, b1), usable ? sc.rs[c] : INT32_MIN, sc.qs[c]);
        sc.close[c] = make_int2(max(a0, b0), max(a1, b1));
    }
    __syncwarp();
    // ---- C: adjacency = (1-D ref | 1-D query) & strictly-after-in-both & valid nodes (:882-894, :943-948) ----
    for (int i = lane; i < ((n + 31) & ~31); i += 32) {
        if (i < n) {
            const int2 ci = sc.close[i];
            const int cr = ci.x, cq = ci.y;
            int fr = 0x7fffffff, fq = 0x7fffffff;
            for (int x = 0; x < n; ++x) {
                const int4 v = sc.nodej[x];
                if (v.x > cr && v.x < fr) fr = v.x;
                if (v.y > cq && v.y < fq) fq = v.y;
            }
            int kr = 0x7fffffff, kq = 0x7fffffff;
            for (int y = 0; y < n; ++y) {
                const int2 v = sc.close[y];
                if (v.x > fr && v.x < kr) kr = v.x;
                if (v.y > fq && v.y < kq) kq = v.y;
            }
            const int re_i = sc.re[i], qe_i = sc.qe[i];
            const bool ign_i = (re_i - sc.rs[i] < 0) || (qe_i - sc.qs[i] < 0);
            // fr == INT_MAX leaves the window (cr, kr) empty for every real key, so no separate test is needed:
            // kr stays INT_MAX only if no close key exceeds fr, and with fr == INT_MAX no open key exceeds cr
            const bool any_r = fr != 0x7fffffff, any_q = fq != 0x7fffffff;
            for (int wd = 0; wd < nw; ++wd) {
                uint32_t bits = 0;
                if (!ign_i) {
                    const int jend = min(32, n - wd * 32);
                    for (int b = 0; b < jend; ++b) {
                        const int4 v = sc.nodej[wd * 32 + b];
                        const bool one_d = (any_r && v.x > cr && v.x < kr) || (any_q && v.y > cq && v.y < kq);
                        const bool ok = one_d && re_i < v.z && qe_i < v.w;
                        bits |= (ok ? 1u : 0u) << b;
                    }
                }
                sc.adj[i * nw + wd] = bits;
            }
        }
    }
    __syncwarp();
    // ---- D: remove_unreachable_branches (:662-681) ----
    int prev_dead = -1;
    for (;;) {
        int dead_total = 0;
        for (int wd = 0; wd < nw; ++wd) {
            uint32_t acc = 0;
            for (int i = lane; i < n; i += 32) acc |= sc.adj[i * nw + wd];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc |= __shfl_xor_sync(SMX_FULL_MASK, acc, off);
            const int cnt = min(32, n - wd * 32);
            uint32_t valid = cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u);
            uint32_t dead = ~acc & valid;
            if ((init >> 5) == wd) dead &= ~(1u << (init & 31));
            dead_total += __popc(dead);
            if (lane == 0) sc.list0[wd] = (int)dead;
        }
        __syncwarp();
        if (dead_total == prev_dead) break;
        prev_dead = dead_total;
        for (int i = lane; i < n; i += 32) {
            if ((((uint32_t)sc.list0[i >> 5]) >> (i & 31)) & 1u)
                for (int wd = 0; wd < nw; ++wd) sc.adj[i * nw + wd] = 0;
        }
        __syncwarp();
    }
    // ---- E: label-correcting longest path in the reference's frontier order (:682-703) ----
    for (int j = lane; j < n; j += 32) {
        int d = 0;
        for (int i = 0; i < n; ++i) d += (sc.adj[i * nw + (j >> 5)] >> (j & 31)) & 1u;
        sc.indeg[j] = d;
        sc.score[j] = 0;
        sc.parent[j] = -1;
    }
    __syncwarp();
    if (lane == 0) { sc.score[init] = sc.re[init] - sc.rs[init]; sc.list0[0] = init; }
    __syncwarp();
    int *cur = sc.list0, *nxt = sc.list1;
    int ncur = 1;
    while (ncur > 0) {
        int nnext = 0;
        for (int c = 0; c < ncur; ++c) {
            const int i = cur[c];
            const int si = sc.score[i];
            for (int wd = 0; wd < nw; ++wd) {
                const uint32_t bits = sc.adj[i * nw + wd];
                if (!bits) continue;
                const int j = wd * 32 + lane;
                const bool has = (bits >> lane) & 1u;
                bool ready = false;
                if (has) {
                    const int t = si + (sc.re[j] - sc.rs[j]);
                    if (t > sc.score[j]) { sc.score[j] = t; sc.parent[j] = i; }
                    const int d = sc.indeg[j] - 1;
                    sc.indeg[j] = d;
                    ready = d == 0;
                }
                const unsigned rb = __ballot_sync(SMX_FULL_MASK, ready);
                if (ready) nxt[nnext + __popc(rb & ((1u << lane) - 1u))] = j;
                nnext += __popc(rb);
            }
            __syncwarp();
        }
        int *t = cur; cur = nxt; nxt = t;
        ncur = nnext;
        __syncwarp();
    }
    // ---- F: first arg-max, walk the parents ----
    int best_s = -1, best_j = 0x7fffffff;
    for (int j = lane; j < n; j += 32) { const int s = sc.score[j]; if (s > best_s) { best_s = s; best_j = j; } }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int os = __shfl_xor_sync(SMX_FULL_MASK, best_s, off), oj = __shfl_xor_sync(SMX_FULL_MASK, best_j, off);
        if (os > best_s || (os == best_s && oj < best_j)) { best_s = os; best_j = oj; }
    }
    int len = 0;
    if (lane == 0) {
        if (best_j == init || sc.parent[best_j] >= 0) {
            for (int x = best_j; x >= 0; x = sc.parent[x]) { sc.list0[len++] = x; if (len > n) { len = 0; break; } }
            for (int t = 0; t < len; ++t) sc.trace[t] = sc.keep[sc.list0[len - 1 - t]];
        }
    }
    len = __shfl_sync(SMX_FULL_MASK, len, 0);
    __syncwarp();
    *score_out = best_s;
    return len;
}

__global__ void __launch_bounds__(SMX_CHAIN_WARPS * 32) smx_chain_kernel(const SmxChainArgs a)
{
    extern __shared__ __align__(16) int chain_smem[];
    const int task_id = blockIdx.x / a.n_slices, slice = blockIdx.x % a.n_slices;
    const SmxChainTask tk = a.tasks[task_id];
    const int N = tk.n, nmax = a.nmax;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int4 *R = reinterpret_cast<int4 *>(chain_smem);      // 4 * nmax ints
    int *wbase = chain_smem + 4 * nmax + warp * smx_chain_scratch_ints(nmax);
    SmxChainScratch sc;
    sc.rs = wbase; sc.re = sc.rs + nmax; sc.qs = sc.re + nmax; sc.qe = sc.qs + nmax; sc.keep = sc.qe + nmax;
    // nodej needs 16-byte alignment: wbase is (4 + warp * (17 + nmax / 32)) * nmax ints into a 16-byte aligned array and
    // nmax is a multiple of 32, keep + nmax is therefore 16-byte aligned too
    sc.nodej = reinterpret_cast<int4 *>(sc.keep + nmax);
    sc.close = reinterpret_cast<int2 *>(sc.keep + 5 * nmax);
    sc.score = sc.keep + 7 * nmax; sc.parent = sc.score + nmax; sc.indeg = sc.parent + nmax;
    sc.list0 = sc.indeg + nmax; sc.list1 = sc.list0 + nmax; sc.trace = sc.list1 + nmax;
    sc.adj = reinterpret_cast<uint32_t *>(sc.trace + nmax);
    sc.nw = nmax / 32;
    for (int i = threadIdx.x; i < N; i += blockDim.x) R[i] = a.regions[tk.region_off + i];
    __syncthreads();
    int *trace_out = a.traces + tk.trace_off;
    int *o_score = a.origin_scratch + 3 * tk.region_off, *o_g1 = o_score + N, *o_g2 = o_g1 + N;

    if (a.topology == 1) {
        // exec_linear (:793-810): only the search with the LAST listed region as origin is returned; no wrap
        if (warp == 0 && slice == 0 && a.phase == 0) {
            int score = 0;
            const int len = smx_chain_search(sc, R, N, N - 1, false, tk.n_ref, tk.n_q, &score);
            for (int t = lane; t < len; t += 32) trace_out[t] = sc.trace[t];
            if (lane == 0) { a.out[task_id].trace_len = len; a.out[task_id].score = score; }
        }
        return;
    }
    if (a.phase == 0) {
        // exec_circular (:811-862): every region as origin; origins are dealt to (slice, warp)
        for (int r = slice * SMX_CHAIN_WARPS + warp; r < N; r += a.n_slices * SMX_CHAIN_WARPS) {
            int score = 0;
            const int len = smx_chain_search(sc, R, N, r, true, tk.n_ref, tk.n_q, &score);
            if (lane == 0) {
                if (len == 0) { o_score[r] = -1; o_g1[r] = 0; o_g2[r] = 0; }
                else {
                    // tie metrics on the path in original coordinates (:838-856)
                    int am = 0, bm = 0;
                    for (int t = 1; t < len; ++t) {
                        if (R[sc.trace[t]].z < R[sc.trace[am]].z) am = t;
                        if (R[sc.trace[t]].w > R[sc.trace[bm]].w) bm = t;
                    }
                    int d = (R[sc.trace[am]].x - R[sc.trace[bm]].y) % tk.n_ref;
                    if (d < 0) d += tk.n_ref;
                    int g = R[sc.trace[0]].z - R[sc.trace[len - 1]].w;
                    for (int t = 0; t + 1 < len; ++t) g = max(g, R[sc.trace[t + 1]].z - R[sc.trace[t]].w);
                    o_score[r] = score; o_g1[r] = d; o_g2[r] = g;
                }
            }
            __syncwarp();
        }
        if (a.n_slices > 1) return;  // phase 1 (separate launch) picks across the slices
        __threadfence_block();
        __syncthreads();
    }
    // pick: max score, then max g1, then first max g2 (:830-861); any failed origin fails the pair-strand
    __shared__ int s_best;
    if (threadIdx.x == 0) {
        int best = -1, fail = 0;
        for (int r = 0; r < N; ++r) { if (o_score[r] < 0) fail = 1; best = max(best, o_score[r]); }
        int pick = -1;
        if (!fail) {
            int top = -1;
            for (int r = 0; r < N; ++r) if (o_score[r] == best) top = max(top, o_g1[r]);
            int cnt = 0, first = -1;
            for (int r = 0; r < N; ++r) if (o_score[r] == best) { ++cnt; if (first < 0) first = r; }
            if (cnt == 1) pick = first;
            else
                for (int r = 0; r < N; ++r)
                    if (o_score[r] == best && o_g1[r] == top && (pick < 0 || o_g2[r] > o_g2[pick])) pick = r;
        }
        s_best = pick;
    }
    __syncthreads();
    if (warp == 0) {
        int len = 0, score = 0;
        if (s_best >= 0) len = smx_chain_search(sc, R, N, s_best, true, tk.n_ref, tk.n_q, &score);
        for (int t = lane; t < len; t += 32) trace_out[t] = sc.trace[t];
        if (lane == 0) { a.out[task_id].trace_len = len; a.out[task_id].score = score; }
    }
}
