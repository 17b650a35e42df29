// Row A4 on the device: threshold over the diagonal counts and conserved-run detection.
// Restates calc_conserved_region_core's selection (ref_query_alignment.py:388) and
// WindowAnalysis.exec (:966-984):
//     selected k : counts[k] > median(counts[slice]) + sigma * std(counts[slice])
//     regions    : maximal runs of ref_ext[k+i] == query[i] with length > 20 on every selected k
// One CTA per pair-strand.  median by a two-level 256-bin histogram select (exact), mean/variance
// from exact int64 sums; the float64 threshold differs from numpy's pairwise-summed value by a few
// ulps at most, so pair-strands whose threshold lies within 1e-6 of an integer are flagged and
// re-decided on the host with numpy's exact summation order (smx_pipeline.cuh).
#pragma once
#include "smx_common.cuh"

#define SMX_SEL_THREADS 256
#define SMX_SEL_MAX_K 2048   // selected diagonals kept per pair-strand (overflow -> host path)
#define SMX_RUN_MIN 20       // WindowAnalysis.consecutive_true_threshold
#define SMX_SEL_REGION_CAP 256  // regions kept per pair-strand (overflow -> host path)

struct SmxSelPair {
    int64_t ref_off, qry_off, counts_off, region_off;
    double sigma;
    int32_t nr, nq, max_k, slice_lo, slice_hi, region_cap;
};

struct SmxSelOut {
    double thr;
    int32_t n_regions;  // may exceed region_cap (then flagged)
    int32_t n_selected;
    int32_t flags;      // 1 = threshold ambiguous, 2 = selected overflow, 4 = region overflow
    int32_t region_base;  // offset of this pair-strand's regions in the compact region array
};

struct SmxRegion { int32_t k, qs, qe; };

struct SmxSelArgs {
    const uint8_t *pool;
    const SmxSelPair *pairs;
    const int32_t *counts;
    SmxRegion *regions;   // compact: every CTA reserves its slots with one atomicAdd on *region_counter
    SmxSelOut *out;
    int32_t *region_counter;
    int32_t topology;
};

// first bin b of hist[0..256) with prefix(hist, b + 1) > rank, and the rank inside that bin; warp 0 scans 8 bins per lane
__device__ __forceinline__ void smx_hist_find(const int *hist, int rank, int *out_bin, int *out_rank)
{
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    int h[8], local = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { h[b] = hist[lane * 8 + b]; local += h[b]; }
    int incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(SMX_FULL_MASK, incl, off); if (lane >= off) incl += t; }
    const int before = incl - local;  // counts in the bins of the lanes below
    const unsigned owns = __ballot_sync(SMX_FULL_MASK, incl > rank);  // lanes whose inclusive prefix passes the rank
    const int owner = owns ? __ffs(owns) - 1 : 31;
    if (lane == owner) {
        int acc = before, b = 0;
#pragma unroll
        for (; b < 8; ++b) { if (acc + h[b] > rank) break; acc += h[b]; }
        if (b == 8) { b = 7; acc -= h[7]; }  // rank beyond the total (cannot happen for rank < count)
        *out_bin = lane * 8 + b;
        *out_rank = rank - acc;
    }
}

// k-th smallest (0-based rank) of v[lo..hi) with values in [0, 65535]
__device__ int smx_select_rank(const int32_t *v, int lo, int hi, int rank, int *hist /*256*/, int *bcast)
{
    const int tid = threadIdx.x;
    hist[tid] = 0;
    __syncthreads();
    for (int x = lo + tid; x < hi; x += SMX_SEL_THREADS) atomicAdd(&hist[(v[x] >> 8) & 255], 1);
    __syncthreads();
    smx_hist_find(hist, rank, &bcast[0], &bcast[1]);
    __syncthreads();
    const int hb = bcast[0], r2 = bcast[1];
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (int x = lo + tid; x < hi; x += SMX_SEL_THREADS) { const int c = v[x]; if (((c >> 8) & 255) == hb) atomicAdd(&hist[c & 255], 1); }
    __syncthreads();
    smx_hist_find(hist, r2, &bcast[2], &bcast[3]);
    __syncthreads();
    const int res = (hb << 8) | bcast[2];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SMX_SEL_THREADS) smx_select_kernel(const SmxSelArgs a)
{
    __shared__ int hist[256];
    __shared__ int bcast[4];
    __shared__ long long red[2 * (SMX_SEL_THREADS / 32)];
    __shared__ int sel[SMX_SEL_MAX_K];
    __shared__ SmxRegion s_regs[SMX_SEL_REGION_CAP];
    __shared__ int s_base;
    __shared__ int n_sel, n_reg;
    __shared__ double thr_s;

    const SmxSelPair pr = a.pairs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *cnt = a.counts + pr.counts_off;
    const int lo = pr.slice_lo, hi = pr.slice_hi, n = hi - lo;
    int flags = 0;

    // exact sums
    long long s1 = 0, s2 = 0;
    for (int x = lo + tid; x < hi; x += SMX_SEL_THREADS) { const long long c = cnt[x]; s1 += c; s2 += c * c; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s1 += __shfl_xor_sync(SMX_FULL_MASK, s1, off); s2 += __shfl_xor_sync(SMX_FULL_MASK, s2, off); }
    if (lane == 0) { red[warp] = s1; red[SMX_SEL_THREADS / 32 + warp] = s2; }
    if (tid == 0) { n_sel = 0; n_reg = 0; }
    __syncthreads();
    // median (np.median: mean of the two middle order statistics for even n)
    const int m_lo = smx_select_rank(cnt, lo, hi, (n - 1) / 2, hist, bcast);
    const int m_hi = (n & 1) ? m_lo : smx_select_rank(cnt, lo, hi, n / 2, hist, bcast);
    if (tid == 0) {
        long long t1 = 0, t2 = 0;
        for (int w = 0; w < SMX_SEL_THREADS / 32; ++w) { t1 += red[w]; t2 += red[SMX_SEL_THREADS / 32 + w]; }
        const double med = ((double)m_lo + (double)m_hi) * 0.5;
        // n * sum(c^2) - (sum c)^2 exceeds int64 for long references with long reads (200 kb x 65 kb): 128-bit
        const unsigned __int128 vnum = (unsigned __int128)(unsigned long long)t2 * (unsigned long long)n - (unsigned __int128)(unsigned long long)t1 * (unsigned long long)t1;
        const double var = (double)vnum / ((double)n * (double)n);
        thr_s = med + pr.sigma * sqrt(var > 0.0 ? var : 0.0);
    }
    __syncthreads();
    const double thr = thr_s;
    if (fabs(thr - rint(thr)) < 1e-6) flags |= 1;
    if (pr.nq > 65535) flags |= 1;  // counts above 16 bits: the two-level histogram median does not apply, the host decides

    // selected diagonals, ascending k (ordered compaction, 256 candidates per round)
    for (int base = 0; base < pr.max_k; base += SMX_SEL_THREADS) {
        const int k = base + tid;
        const bool hit = k < pr.max_k && (double)cnt[k] > thr;
        const unsigned bal = __ballot_sync(SMX_FULL_MASK, hit);
        if (lane == 0) hist[warp] = __popc(bal);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += hist[w];
        int total = 0;
        for (int w = 0; w < SMX_SEL_THREADS / 32; ++w) total += hist[w];
        const int pos = n_sel + before + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < SMX_SEL_MAX_K) sel[pos] = k;
        __syncthreads();
        if (tid == 0) n_sel += total;
        __syncthreads();
    }
    int nsel = n_sel;
    if (nsel > SMX_SEL_MAX_K) { flags |= 2; nsel = SMX_SEL_MAX_K; }

    // conserved runs: one warp per selected diagonal
    const uint8_t *ref = a.pool + pr.ref_off;
    const uint8_t *qry = a.pool + pr.qry_off;
    const int nr = pr.nr, nq = pr.nq;
    const bool linear = a.topology == 1;
    SmxRegion *out = s_regs;
    // Lane w keeps the ballot word of bases [g + 32 w, g + 32 w + 32) of a 1024-base group; every lane then looks at
    // its own word: the run open at the word's first bit (its start comes from the nearest lower word that is not
    // all ones), and the one run of more than 20 matches that can lie strictly inside 32 bits (found by AND-shifts).
    // A run is reported by the lane in whose word it ends; bits at and beyond nq are 0, so every run ends.
    auto report = [&](int k, int rs, int re) {
        if (re - rs + 1 > SMX_RUN_MIN) {
            const int idx = atomicAdd(&n_reg, 1);
            if (idx < pr.region_cap) out[idx] = SmxRegion{k, rs, re};
        }
    };
    for (int si = warp; si < nsel; si += SMX_SEL_THREADS / 32) {
        const int k = sel[si];
        int carry = -1;  // start of the run open at the beginning of the group (warp-uniform)
        for (int g = 0; g < nq; g += 1024) {
            const int nwords = min(32, (nq - g + 31) >> 5);
            uint32_t mine = 0;
            for (int w = 0; w < nwords; ++w) {
                const int i = g + w * 32 + lane;
                bool eq = false;
                if (i < nq) {
                    if (linear) {
                        const int p = k - (nq - 1) + i;
                        if (p >= 0 && p < nr) eq = ref[p] == qry[i];
                    } else {
                        int p = k + i;
                        if (p >= nr) { p -= nr; if (p >= nr) p %= nr; }
                        eq = ref[p] == qry[i];
                    }
                }
                const unsigned m = __ballot_sync(SMX_FULL_MASK, eq);
                if (lane == w) mine = m;
            }
            const int base = g + lane * 32;
            const uint32_t nm = ~mine;
            const int c = nm ? __ffs(nm) - 1 : 32;  // ones from bit 0 up
            const int h = nm ? __clz(nm) : 32;      // ones from bit 31 down
            const unsigned allones = __ballot_sync(SMX_FULL_MASK, nm == 0u);
            // start of the run that is open when this lane's word ends (-1: none; all-ones words resolved below)
            const int open_out = h == 0 ? -1 : base + 32 - h;
            const unsigned lower = ~allones & ((1u << lane) - 1u);  // lower lanes whose word has a zero bit
            const int src = lower ? 31 - __clz(lower) : -1;
            int entering = __shfl_sync(SMX_FULL_MASK, open_out, src < 0 ? 0 : src);
            if (src < 0) entering = carry;
            // all-ones words between src and this lane: a run that was not open before them starts at the first of them
            if (entering < 0 && lane - 1 > src) entering = g + (src + 1) * 32;
            uint32_t rest = mine;
            if (entering >= 0) {
                if (c < 32) report(k, entering, base + c - 1);
                rest = c == 32 ? 0u : (mine & ~((1u << c) - 1u));
            }
            {
                const uint32_t a2 = rest & (rest >> 1), a4 = a2 & (a2 >> 2), a8 = a4 & (a4 >> 4), a16 = a8 & (a8 >> 8);
                const uint32_t a21 = a16 & (a4 >> 16) & (rest >> 20);  // bit p: bits p .. p+20 of `rest` are set
                if (a21) {
                    const int p = __ffs(a21) - 1;
                    const uint32_t inv = ~(rest >> p);
                    const int e = inv ? p + __ffs(inv) - 2 : 31;  // last bit of that run (inv == 0: all 32 bits set)
                    if (e < 31) report(k, base + p, base + e);  // a run touching bit 31 is still open: a later word ends it
                }
            }
            // run open at the end of the group
            int out31 = nm == 0u ? (entering >= 0 ? entering : base) : open_out;
            carry = __shfl_sync(SMX_FULL_MASK, out31, 31);
        }
        if (lane == 0 && carry >= 0) report(k, carry, nq - 1);  // nq a multiple of 1024 and the last base matches
    }
    __syncthreads();
    const int n_keep = min(n_reg, pr.region_cap);
    if (tid == 0) s_base = n_keep > 0 ? atomicAdd(a.region_counter, n_keep) : 0;
    __syncthreads();
    for (int x = tid; x < n_keep; x += SMX_SEL_THREADS) a.regions[s_base + x] = s_regs[x];
    if (tid == 0) {
        SmxSelOut o;
        o.thr = thr;
        o.n_regions = n_reg;
        o.n_selected = n_sel;
        if (n_reg > pr.region_cap) flags |= 4;
        o.flags = flags;
        o.region_base = s_base;
        a.out[blockIdx.x] = o;
    }
}
