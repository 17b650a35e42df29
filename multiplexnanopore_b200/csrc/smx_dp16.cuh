// Forward affine-gap DP on packed s16x2 DPX lanes: two bands of a problem per warp.
//
// Same recurrence, tie rules, direction-bit encoding and trace layout as smx_dp.cuh (so the traceback kernel is
// shared), but every 32-bit register carries TWO cells: the low half belongs to band 2P, the high half to band
// 2P+1 of the same problem, 64 columns (two chunks) behind.  One VIMNMX.S16x2 computes both maxima AND both
// "which operand won" predicates (ptxas fuses max.s16x2 + setp.eq.s16 into the predicate outputs of VIMNMX),
// VIADD.16x2 does both additions, one PRMT fetches both substitution scores (low half from the profile register
// of the low row, high half from that of the high row).  About 11 instructions per cell instead of 26.
//
// Eligibility (decided on the host, smx_api.cu): global / semi-global mode, at least two bands (m > 256), every
// intermediate value fits 16 bits with head-room, and the reference slice is pure ACGT (the 8 bytes a PRMT can
// address hold the two 4-letter profiles; a reference wildcard would need a ninth, zero, byte).  Everything
// else runs on the int32 kernels.
#pragma once
#include "smx_dp.cuh"

#define SMX16_NEG (-32000)
#ifndef SMX16_UNROLL
#define SMX16_UNROLL 2  // steps of the interior chunk loop unrolled together (4 is 7 % slower with one warp per problem: instruction cache)
#endif
#define SMX16_PRAGMA_(x) _Pragma(#x)
#define SMX16_PRAGMA(x) SMX16_PRAGMA_(x)
#define SMX16_LAG 64  // columns the high-half band trails the low-half band (two 32-column chunks)

__device__ __forceinline__ uint32_t smx16_add(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// Per-half max(a, b).  ptxas fuses max.s16x2 + the two setp.eq.s16 into one VIMNMX.S16x2 with two predicate
// outputs (p = "a won or tied"); the predicated ORs drop the direction bit IMM into the trace word of the low /
// high band when b won.  Direction-bit encoding of the s16 kernels (decoded by smx_traceback for classes >= 10):
//   bit 0: diagonal lost against max(E, F)      bit 1: E beat F (F >= E keeps it clear: DEL before INS)
//   bit 2: E opened from H (tie -> extend)      bit 3: F opened from H
// The flag is dropped with a predicated IMAD (w = one * IMM + w, `one` an opaque register holding 1): the
// integer ALU pipe of an SMSP issues one warp instruction every two cycles and carries every VIADD / VIMNMX / PRMT
// / SEL / LOP3 of this kernel, while the FMA pipe (IMAD) would otherwise idle -- measured: the kernel is bound by
// the ALU pipe, not by the issue slot.
template <uint32_t IMM>
__device__ __forceinline__ uint32_t smx16_max_tr(uint32_t a, uint32_t b, uint32_t &w_lo, uint32_t &w_hi, uint32_t one)
{
    uint32_t r;
    asm("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3;\n\t"
        "mov.b32 {r2, r3}, %3;\n\t"  // read a before %0 is written: the output may share its register
        "max.s16x2 %0, %3, %4;\n\t"
        "mov.b32 {r0, r1}, %0;\n\t"
        "setp.eq.s16 pv, r0, r2;\n\t"
        "setp.eq.s16 pu, r1, r3;\n\t"
        "@!pv mad.lo.u32 %1, %6, %5, %1;\n\t"
        "@!pu mad.lo.u32 %2, %6, %5, %2;}\n\t"
        : "=r"(r), "+r"(w_lo), "+r"(w_hi)
        : "r"(a), "r"(b), "n"(IMM), "r"(one));
    return r;
}

struct Smx16Col {  // per-column scalars threaded through the K cells of a lane
    uint32_t hd, fu, hugo, h, f, w_lo, w_hi, w_lo2, w_hi2, one;
};

template <int K, int k>
struct Smx16Cell {
    static __device__ __forceinline__ void run(uint32_t (&Hl)[K], uint32_t (&Hlgo)[K], uint32_t (&El)[K], const uint32_t (&plo)[K],
                                               const uint32_t (&phi)[K], Smx16Col &c, uint32_t sel, uint32_t nge2, uint32_t ngo2)
    {
        // the four flags of this cell land in place (bits 4k .. 4k+3) in one of two accumulators per band
        // (even / odd k), so that the predicated IMADs of neighbouring cells do not serialise
        uint32_t &n_lo = (k & 1) ? c.w_lo2 : c.w_lo;
        uint32_t &n_hi = (k & 1) ? c.w_hi2 : c.w_hi;
        const uint32_t e_ext = smx16_add(El[k], nge2);
        const uint32_t e = smx16_max_tr<(4u << (4 * k))>(e_ext, Hlgo[k], n_lo, n_hi, c.one);
        const uint32_t f_ext = smx16_add(c.fu, nge2);
        const uint32_t f = smx16_max_tr<(8u << (4 * k))>(f_ext, c.hugo, n_lo, n_hi, c.one);
        const uint32_t hdiag = smx16_add(c.hd, smx_subst_u(plo[k], phi[k], sel));
        const uint32_t t = smx16_max_tr<(2u << (4 * k))>(f, e, n_lo, n_hi, c.one);
        const uint32_t h = smx16_max_tr<(1u << (4 * k))>(hdiag, t, n_lo, n_hi, c.one);
        c.hd = Hl[k];
        Hl[k] = h;
        c.hugo = smx16_add(h, ngo2);
        Hlgo[k] = c.hugo;
        El[k] = e;
        c.fu = f;
        c.h = h;
        c.f = f;
        Smx16Cell<K, k + 1>::run(Hl, Hlgo, El, plo, phi, c, sel, nge2, ngo2);
    }
};
template <int K>
struct Smx16Cell<K, K> {
    static __device__ __forceinline__ void run(uint32_t (&)[K], uint32_t (&)[K], uint32_t (&)[K], const uint32_t (&)[K], const uint32_t (&)[K],
                                               Smx16Col &, uint32_t, uint32_t, uint32_t) {}
};

__device__ __forceinline__ uint32_t smx16_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int smx16_lo(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int smx16_hi(uint32_t v) { return (int)v >> 16; }

// One interior 32-column chunk of a band pair: every lane is strictly inside the matrix with both halves for all
// 32 steps, so there is no boundary (re)initialisation, no end-column candidate and the direction-bit stores are
// unconditional.  Lane 0 takes its inputs from the chunk's staging row in shared memory (one broadcast LDS.128
// per step) through a lane mask instead of a predicate: VIMNMX.S16x2 needs two predicate registers per
// instruction and the SM has seven, so long-lived predicates are kept out of this loop (the two bottom-line
// store predicates are re-derived from a register at their use).
template <int K, bool HAS_HI, bool TRACK>
__device__ __forceinline__ void smx16_fast_chunk(uint32_t (&Hl)[K], uint32_t (&Hlgo)[K], uint32_t (&El)[K], const uint32_t (&plo)[K],
                                                 const uint32_t (&phi)[K], uint32_t &hd, uint32_t &out_h, uint32_t &out_f, uint32_t &sel_prev,
                                                 const uint4 *tops, uint32_t lane0_mask, uint32_t nge2, uint32_t ngo2, uint32_t *tp_lo,
                                                 uint32_t *tp_hi, uint32_t *bl, uint32_t *bh, uint32_t st_flags, uint32_t tmask, int last_k,
                                                 int jbase, int &row_best, int &row_best_j, uint32_t one)
{
    SMX16_PRAGMA(unroll SMX16_UNROLL)
    for (int ss = 0; ss < 32; ++ss) {
        uint32_t hu = __shfl_up_sync(SMX_FULL_MASK, out_h, 1);
        uint32_t fu = __shfl_up_sync(SMX_FULL_MASK, out_f, 1);
        uint32_t sel = __shfl_up_sync(SMX_FULL_MASK, sel_prev, 1);
        const uint4 tv = tops[ss];
        hu = (tv.x & lane0_mask) | (hu & ~lane0_mask);
        fu = (tv.y & lane0_mask) | (fu & ~lane0_mask);
        sel = (tv.z & lane0_mask) | (sel & ~lane0_mask);
        sel_prev = sel;
        Smx16Col c;
        c.hd = hd; c.fu = fu; c.hugo = smx16_add(hu, ngo2); c.h = 0; c.f = 0; c.w_lo = 0; c.w_hi = 0; c.w_lo2 = 0; c.w_hi2 = 0; c.one = one;
        Smx16Cell<K, 0>::run(Hl, Hlgo, El, plo, phi, c, sel, nge2, ngo2);
        hd = hu;
        out_h = c.h;
        out_f = c.f;
        tp_lo[ss * 32] = c.w_lo | c.w_lo2;
        if (HAS_HI) tp_hi[ss * 32] = c.w_hi | c.w_hi2;
        uint32_t fl = st_flags;
        asm volatile("" : "+r"(fl));  // keeps the store predicates short-lived (see above)
        if (fl & 1u) __stcg(&bl[ss], __byte_perm(out_h, out_f, 0x5410));
        if (HAS_HI) { if (fl & 2u) __stcg(&bh[ss], __byte_perm(out_h, out_f, 0x7632)); }
        if (TRACK) {
            uint32_t tm = tmask;
            asm volatile("" : "+r"(tm));
            if (tm) {  // the lane that owns the last query row (semi-global end in the last row)
                uint32_t hw = Hl[0];
#pragma unroll
                for (int k = 1; k < K; ++k) if (k == last_k) hw = Hl[k];
                const int hv = (tm & 1u) ? smx16_lo(hw) : smx16_hi(hw);
                if (hv > row_best) { row_best = hv; row_best_j = jbase + ss; }
            }
        }
    }
}

// boundary lines of the s16 kernel: one uint32 per column = F << 16 | (H & 0xffff)
// W = warps that share one problem (1: every warp of the 128-thread CTA owns its own problem; 2, 4, 8: the CTA
// is one team whose warps take band pairs round-robin).  The host picks W = min(8, largest power of two <= pairs).
#ifndef SMX16_MIN_WARPS_PER_SM
#define SMX16_MIN_WARPS_PER_SM 16
#endif
template <int W>
__global__ void __launch_bounds__(W == 1 ? 128 : W * 32, SMX16_MIN_WARPS_PER_SM / (W == 1 ? 4 : W)) smx_dp_forward16(const SmxFwdArgs a)
{
    constexpr int K = 8;
    constexpr bool TEAM = W > 1;
    constexpr int LAG = SMX16_LAG;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int worker = TEAM ? blockIdx.x : ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    // the int2 line area of the worker is reused as two uint32 lines
    uint32_t *line0 = reinterpret_cast<uint32_t *>(a.boundary + (size_t)worker * 2 * (size_t)a.nmax);
    uint32_t *line1 = line0 + a.nmax;
    const int go = a.sc.go, ge = a.sc.ge;
    const uint32_t nge2 = smx16_pack(-ge, -ge), ngo2 = smx16_pack(-go, -go);
    const uint32_t lane0_mask = lane == 0 ? 0xffffffffu : 0u;
    const uint32_t one = a.one;  // 1, but unknown to ptxas: keeps the predicated flag updates on the FMA pipe (IMAD)

    __shared__ int s_ticket;
    __shared__ volatile int s_prog[W];
    __shared__ int s_red[W][6];
    // lane 0's inputs of the current 32-column chunk, one uint4 per column: {H above (lo | hi << 16), F above, PRMT
    // selector of the two reference codes, -}; read back with one broadcast LDS.128 per step
    __shared__ uint4 s_top[W == 1 ? 4 : W][32];

    for (;;) {
        int ticket = 0;
        if (TEAM) {
            __syncthreads();
            if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1);
            if (threadIdx.x < W) s_prog[threadIdx.x] = -1;
            __syncthreads();
            ticket = s_ticket;
        } else {
            if (lane == 0) ticket = atomicAdd(a.ticket, 1);
            ticket = __shfl_sync(SMX_FULL_MASK, ticket, 0);
        }
        if (ticket >= a.n_order) break;
        const int pid = a.order[ticket];
        const SmxProblem pr = a.probs[pid];
        const int m = pr.m, n = pr.n, mode = pr.mode & SMX_MODE_MASK;
        const bool free_beg = mode == 2;  // SMX_MODE_SG_QB_DB
        const uint8_t *s1 = a.pool + pr.s1_off;
        const uint8_t *s2 = a.pool + pr.s2_off;
        uint32_t *trace = reinterpret_cast<uint32_t *>(a.trace + pr.trace_off);
        const int nbands = (m + 32 * K - 1) / (32 * K);
        const int npairs = (nbands + 1) >> 1;
        const int steps = n + 31;          // steps of one band (trace layout)
        const int psteps = steps + LAG;    // steps of a band pair
        const int nchunks = (psteps + 31) >> 5;

        int col_best = SMX_NEG_INF, col_best_i = 0x7fffffff;
        int row_best = SMX_NEG_INF, row_best_j = 0;
        int corner = 0;
        const int last_band_idx = nbands - 1;
        const int last_lane = ((m - 1) % (32 * K)) / K;
        const int last_k = (m - 1) % K;

        for (int pair = wib % W; pair < npairs; pair += W) {
            const int band_lo = 2 * pair, band_hi = 2 * pair + 1;
            const bool has_hi = band_hi < nbands;
            const int i0_lo = band_lo * 32 * K + lane * K, i0_hi = band_hi * 32 * K + lane * K;
            // band b writes its bottom line to line[(b & 1) ^ 1] and reads its top from line[b & 1]
            const uint32_t *top_lo = line0;  // band_lo is even
            uint32_t *bot_lo = line1;
            const uint32_t *top_hi = line1;
            uint32_t *bot_hi = line0;
            const bool lo_is_last = band_lo == last_band_idx, hi_is_last = band_hi == last_band_idx;

            uint32_t plo[K], phi[K];
            uint32_t Hl[K], Hlgo[K], El[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int il = i0_lo + k, ih = i0_hi + k;
                plo[k] = smx_profile(il < m ? smx_code(s1[il]) : 4, a.sc.match, a.sc.mismatch);
                phi[k] = smx_profile(ih < m ? smx_code(s1[ih]) : 4, a.sc.match, a.sc.mismatch);
                Hl[k] = free_beg ? 0u : smx16_pack(-(go + il * ge), -(go + ih * ge));  // H[i][-1]
                Hlgo[k] = smx16_add(Hl[k], ngo2);
                El[k] = smx16_pack(SMX16_NEG, SMX16_NEG);
            }
            uint32_t hd = free_beg ? 0u : smx16_pack(i0_lo == 0 ? 0 : -(go + (i0_lo - 1) * ge), -(go + (i0_hi - 1) * ge));
            uint32_t out_h = 0, out_f = 0, sel_prev = 0;
            int rch1 = 0, rch2 = 0;  // reference codes of the two previous chunks (the high half reads 64 columns back)
            const bool st_lo = lane == 31 && !lo_is_last, st_hi = lane == 31 && has_hi && !hi_is_last;
            const bool track_lo = mode == 1 && lo_is_last, track_hi = mode == 1 && hi_is_last;

            for (int s0 = 0, chunk = 0; s0 < psteps; s0 += 32, ++chunk) {
                if (TEAM && pair > 0) {
                    // the low band needs columns s0..s0+31 of the previous pair's high band: that pair ran them at
                    // its steps s0+LAG .. s0+LAG+62, i.e. it must have finished chunk index chunk + 3
                    const int need = (pair - 1) * nchunks + min(chunk + 4, nchunks);
                    if (lane == 0) { while (s_prog[(pair - 1) % W] < need) __nanosleep(64); }
                    __syncwarp();
                    __threadfence_block();
                }
                __syncwarp();
                // chunk prefetch: lane l holds column s0 + l (low band inputs) and column s0 + l - LAG (high band top line)
                const int jc = s0 + lane;
                uint32_t topA = smx16_pack(0, SMX16_NEG);  // (H, F) of the row above the low band
                uint32_t topB = smx16_pack(0, SMX16_NEG);  // (H, F) of the row above the high band = bottom of the low band
                int rch0 = 0;
                if (jc < n) {
                    if (band_lo == 0) topA = smx16_pack(free_beg ? 0 : -(go + jc * ge), SMX16_NEG);
                    else topA = __ldcg(&top_lo[jc]);
                    rch0 = smx_code(s2[jc]) & 3;
                }
                const int jb = jc - LAG;
                if (jb >= 0 && jb < n) topB = __ldcg(&top_hi[jb]);
                {
                    const uint32_t c_lo = (uint32_t)rch0, c_hi = (uint32_t)rch2;
                    uint4 tv;
                    tv.x = (topA & 0xffffu) | (topB << 16);
                    tv.y = (topA >> 16) | (topB & 0xffff0000u);
                    tv.z = c_lo | ((c_lo | 8u) << 4) | ((c_hi | 4u) << 8) | ((c_hi | 12u) << 12);
                    tv.w = 0u;
                    s_top[wib][lane] = tv;
                    __syncwarp();
                }
                // ---- fast path: every lane is strictly inside the matrix with both halves for all 32 steps
                //      (1 <= j <= n-2 for the low band, 64 columns behind for the high band): no boundary
                //      (re)initialisation, no end-column candidates, unconditional direction-bit stores ----
                if (s0 >= 32 + LAG && s0 + 33 <= n) {
                    uint32_t *tp_lo = trace + ((size_t)band_lo * steps + s0) * 32 + lane;
                    uint32_t *tp_hi = trace + ((size_t)band_hi * steps + (s0 - LAG)) * 32 + lane;
                    uint32_t *bl = bot_lo + (s0 - 31);
                    uint32_t *bh = bot_hi + (s0 - 31 - LAG);
                    const uint32_t st_flags = (st_lo ? 1u : 0u) | (st_hi ? 2u : 0u);
                    if (track_lo | track_hi) {
                        const int jbase = s0 - lane - (track_lo ? 0 : LAG);
                        const uint32_t tmask = lane == last_lane ? (track_lo ? 0x0000ffffu : 0xffff0000u) : 0u;
                        if (has_hi) smx16_fast_chunk<K, true, true>(Hl, Hlgo, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, nge2, ngo2, tp_lo, tp_hi, bl, bh, st_flags, tmask, last_k, jbase, row_best, row_best_j, one);
                        else smx16_fast_chunk<K, false, true>(Hl, Hlgo, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, nge2, ngo2, tp_lo, tp_hi, bl, bh, st_flags, tmask, last_k, jbase, row_best, row_best_j, one);
                    } else {
                        if (has_hi) smx16_fast_chunk<K, true, false>(Hl, Hlgo, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, nge2, ngo2, tp_lo, tp_hi, bl, bh, st_flags, 0u, 0, 0, row_best, row_best_j, one);
                        else smx16_fast_chunk<K, false, false>(Hl, Hlgo, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, nge2, ngo2, tp_lo, tp_hi, bl, bh, st_flags, 0u, 0, 0, row_best, row_best_j, one);
                    }
                } else {
                const int send = min(32, psteps - s0);
#pragma unroll 1
                for (int ss = 0; ss < send; ++ss) {
                    const int s = s0 + ss;
                    uint32_t hu = __shfl_up_sync(SMX_FULL_MASK, out_h, 1);
                    uint32_t fu = __shfl_up_sync(SMX_FULL_MASK, out_f, 1);
                    uint32_t sel = __shfl_up_sync(SMX_FULL_MASK, sel_prev, 1);
                    const uint4 tv = s_top[wib][ss];
                    if (lane == 0) { hu = tv.x; fu = tv.y; sel = tv.z; }
                    sel_prev = sel;
                    const int j_lo = s - lane, j_hi = s - lane - LAG;
                    const bool v_lo = j_lo >= 0 && j_lo < n;
                    const bool v_hi = has_hi && j_hi >= 0 && j_hi < n;
                    // Both halves are computed on every step (no divergence); a half whose column is still left of
                    // the matrix produces values nobody reads, and its state is (re)set to the boundary column at
                    // the step where it enters the matrix.  A half right of the matrix is finished.
                    if (j_lo == 0 || j_hi == 0) {
                        const uint32_t keep = j_lo == 0 ? 0xffff0000u : 0x0000ffffu;  // the other half keeps its state
                        const int i0 = j_lo == 0 ? i0_lo : i0_hi;
                        const int sh = j_lo == 0 ? 0 : 16;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const uint32_t hv = free_beg ? 0u : ((uint32_t)(-(go + (i0 + k) * ge)) & 0xffffu);
                            Hl[k] = (Hl[k] & keep) | (hv << sh);
                            Hlgo[k] = (Hlgo[k] & keep) | ((((uint32_t)((int)(short)hv - go)) & 0xffffu) << sh);
                            El[k] = (El[k] & keep) | (((uint32_t)SMX16_NEG & 0xffffu) << sh);
                        }
                        const uint32_t hdv0 = (free_beg || i0 == 0) ? 0u : ((uint32_t)(-(go + (i0 - 1) * ge)) & 0xffffu);
                        hd = (hd & keep) | (hdv0 << sh);
                    }
                    {
                        Smx16Col c;
                        c.hd = hd; c.fu = fu; c.hugo = smx16_add(hu, ngo2); c.h = 0; c.f = 0; c.w_lo = 0; c.w_hi = 0; c.w_lo2 = 0; c.w_hi2 = 0; c.one = one;
                        Smx16Cell<K, 0>::run(Hl, Hlgo, El, plo, phi, c, sel, nge2, ngo2);
                        c.w_lo |= c.w_lo2;
                        c.w_hi |= c.w_hi2;
                        hd = hu;  // H[i0-1][j] is the diagonal input of the next column
                        out_h = c.h;
                        out_f = c.f;
                        if (v_lo) {
                            trace[((size_t)band_lo * steps + s) * 32 + lane] = c.w_lo;
                            if (lane == 31 && !lo_is_last) __stcg(&bot_lo[j_lo], (out_h & 0xffffu) | (out_f << 16));
                        }
                        if (v_hi) {
                            trace[((size_t)band_hi * steps + (s - LAG)) * 32 + lane] = c.w_hi;
                            if (lane == 31 && !hi_is_last) __stcg(&bot_hi[j_hi], (out_h >> 16) | (out_f & 0xffff0000u));
                        }
                        // end-cell candidates (rare paths, scalar)
                        if (v_lo && j_lo == n - 1) {
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                const int i = i0_lo + k, hv = smx16_lo(Hl[k]);
                                if (i < m - 1 && hv > col_best) { col_best = hv; col_best_i = i; }
                                if (i == m - 1) corner = hv;
                            }
                        }
                        if (v_hi && j_hi == n - 1) {
#pragma unroll
                            for (int k = 0; k < K; ++k) {
                                const int i = i0_hi + k, hv = smx16_hi(Hl[k]);
                                if (i < m - 1 && hv > col_best) { col_best = hv; col_best_i = i; }
                                if (i == m - 1) corner = hv;
                            }
                        }
                        if (mode == 1 && lane == last_lane) {
                            if (lo_is_last && v_lo) {
                                int hv = smx16_lo(Hl[0]);
#pragma unroll
                                for (int k = 1; k < K; ++k) if (k == last_k) hv = smx16_lo(Hl[k]);
                                if (hv > row_best) { row_best = hv; row_best_j = j_lo; }
                            }
                            if (hi_is_last && v_hi) {
                                int hv = smx16_hi(Hl[0]);
#pragma unroll
                                for (int k = 1; k < K; ++k) if (k == last_k) hv = smx16_hi(Hl[k]);
                                if (hv > row_best) { row_best = hv; row_best_j = j_hi; }
                            }
                        }
                    }
                }
                }
                rch2 = rch1;
                rch1 = rch0;
                if (TEAM) {
                    if (lane == 31) { __threadfence_block(); s_prog[wib] = pair * nchunks + chunk + 1; }
                }
            }
            __syncwarp();
        }

        // ---- end cell / score (same combination as smx_dp_forward) ----
        SmxResult r;
        r.beg_q = 0; r.beg_r = 0; r.n_runs = 0; r.run_start = 0; r.n_s = 0; r.n_h = 0; r.pad0 = 0; r.pad1 = 0; r.pad2 = 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const int ob = __shfl_xor_sync(SMX_FULL_MASK, col_best, off);
            const int oi = __shfl_xor_sync(SMX_FULL_MASK, col_best_i, off);
            if (ob > col_best || (ob == col_best && oi < col_best_i)) { col_best = ob; col_best_i = oi; }
        }
        const int owner = TEAM ? (((nbands - 1) >> 1) % W) : 0;  // warp that ran the last band
        row_best = __shfl_sync(SMX_FULL_MASK, row_best, last_lane);
        row_best_j = __shfl_sync(SMX_FULL_MASK, row_best_j, last_lane);
        corner = __shfl_sync(SMX_FULL_MASK, corner, last_lane);
        if (TEAM) {
            if (lane == 0) {
                s_red[wib][0] = col_best; s_red[wib][1] = col_best_i;
                if (wib == owner) { s_red[0][2] = row_best; s_red[0][3] = row_best_j; s_red[0][4] = corner; }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < W; ++w) {
                    const int ob = s_red[w][0], oi = s_red[w][1];
                    if (ob > col_best || (ob == col_best && oi < col_best_i)) { col_best = ob; col_best_i = oi; }
                }
                row_best = s_red[0][2]; row_best_j = s_red[0][3]; corner = s_red[0][4];
            }
        }
        if (mode == 1) {  // SMX_MODE_SG_QE_DE
            int best = SMX_NEG_INF, bi = m - 1, bj = n - 1;
            if (col_best > best) { best = col_best; bi = col_best_i; bj = n - 1; }
            if (row_best > best) { best = row_best; bi = m - 1; bj = row_best_j; }
            r.score = best; r.end_q = bi; r.end_r = bj;
        } else {
            r.score = corner; r.end_q = m - 1; r.end_r = n - 1;
        }
        if (TEAM ? (threadIdx.x == 0) : (lane == 0)) a.results[pid] = r;
    }
}
