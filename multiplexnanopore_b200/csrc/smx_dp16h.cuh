// Forward affine-gap DP on packed u16x2 lanes in the "diagonal frame": two bands of a problem per warp, five ALU
// instructions per cell pair.
//
// Same recurrence, tie rules, direction bits and trace layout as smx_dp16.cuh (the traceback kernel cannot tell the
// two apart), but the values are stored shifted by the gap-extension cost of their anti-diagonal,
//     X^(i, j) = X(i, j) + ge * (i + j)                       for X in {H, E, F},
// which turns a gap extension into a plain copy:
//     E^(i, j) = max(E^(i, j-1), Hc(i, j-1))                   Hc := H^ - (go - ge)
//     F^(i, j) = max(F^(i-1, j), Hc(i-1, j))
//     H^(i, j) = max(Hc(i-1, j-1) + s'', E^, F^)               s'' := s(q_i, r_j) + go + ge  (>= 0, in the profile)
// All three candidates of a cell are shifted by the same amount, so every comparison -- and with it every direction
// bit and every tie -- is the one the unshifted recurrence makes.  Values are kept as UNSIGNED 16-bit halves with a
// small bias, and both remaining additions only ever add a non-negative amount to, or subtract a constant from,
// a half that cannot wrap: they are ordinary 32-bit integer adds (FMA pipe on sm_100a, which would otherwise idle),
// and the integer ALU pipe -- two cycles per warp instruction, the binding resource of smx_dp16.cuh -- is left with
// four VIMNMX.U16x2 and one PRMT per cell pair.
//
// A last band without a partner band (odd band count) runs as two half-bands of 128 rows at 4 rows per lane, see the
// kernel; its direction bits then use the K = 4 layout inside that band's region of the arena.
//
// Eligibility (host, smx_api.cu: s16h_eligible): as smx_dp16.cuh, plus mismatch + 2*ge >= 0 (values never fall along
// a diagonal), match + go + ge <= 127, and the biased range bound below 65000.
#pragma once
#include "smx_dp16.cuh"
#include <type_traits>

// max.u16x2 with both "a won or tied" predicates (ptxas fuses this into one VIMNMX.U16x2 with two predicate
// outputs); when b won strictly the flag IMM is added to the direction word of the low / high band.
// FMA = true: the flag is dropped by a predicated IMAD (w = one * IMM + w, FMA pipe); false: by a predicated OR (ALU
// pipe).  The cell uses both so that neither integer pipe carries all 64 flag updates of a column step.
template <uint32_t IMM, bool FMA>
__device__ __forceinline__ uint32_t smx16h_max_tr(uint32_t a, uint32_t b, uint32_t &w_lo, uint32_t &w_hi, uint32_t one)
{
    uint32_t r;
    if (FMA) {
        asm("{.reg .pred pu, pv; .reg .u16 r0, r1, r2, r3;\n\t"
            "mov.b32 {r2, r3}, %3;\n\t"  // read a before %0 is written: the output may share its register
            "max.u16x2 %0, %3, %4;\n\t"
            "mov.b32 {r0, r1}, %0;\n\t"
            "setp.eq.u16 pv, r0, r2;\n\t"
            "setp.eq.u16 pu, r1, r3;\n\t"
            "@!pv mad.lo.u32 %1, %6, %5, %1;\n\t"
            "@!pu mad.lo.u32 %2, %6, %5, %2;}\n\t"
            : "=r"(r), "+r"(w_lo), "+r"(w_hi)
            : "r"(a), "r"(b), "n"(IMM), "r"(one));
    } else {
        asm("{.reg .pred pu, pv; .reg .u16 r0, r1, r2, r3;\n\t"
            "mov.b32 {r2, r3}, %3;\n\t"
            "max.u16x2 %0, %3, %4;\n\t"
            "mov.b32 {r0, r1}, %0;\n\t"
            "setp.eq.u16 pv, r0, r2;\n\t"
            "setp.eq.u16 pu, r1, r3;\n\t"
            "@!pv or.b32 %1, %1, %5;\n\t"
            "@!pu or.b32 %2, %2, %5;}\n\t"
            : "=r"(r), "+r"(w_lo), "+r"(w_hi)
            : "r"(a), "r"(b), "n"(IMM));
    }
    return r;
}
#ifndef SMX16H_WPC
#define SMX16H_WPC 4  // warps (= problems) per CTA of the one-warp-per-problem kernel
#endif
#ifndef SMX16H_CK_WARPS_PER_SM
// resident warps per SM the score-only (checkpointed) instantiations are built for: without the direction-word accumulators
// they fit 96 registers (C2 forward phase: 16 warps 402 ms per 20 000 reads, 20 warps 389, 24 warps / 80 registers 433)
#define SMX16H_CK_WARPS_PER_SM 20
#endif
#ifndef SMX16H_ALU_FLAGS
#define SMX16H_ALU_FLAGS 1  // how many of the four flag kinds of a cell go through the ALU pipe
#endif

struct Smx16hCol {  // per-column scalars threaded through the K cells of a lane
    uint32_t hd;    // Hc(i-1, j-1): diagonal input of the current row
    uint32_t fu;    // F^(i-1, j)
    uint32_t hcu;   // Hc(i-1, j)
    uint32_t w_lo[2], w_hi[2];  // direction-bit accumulators; K = 8: even / odd rows of one word, K = 16: rows 0-7 / 8-15
    uint32_t one;
    // LOCAL (sw_trace): the zero of the current row in the diagonal frame, its increment per row, best H seen by the lane
    uint32_t z, ge2, m;
};

// Local alignment (parasail.sw_trace, ref_query_alignment.py:35) in the diagonal frame.  H = max(0, ...) becomes
// h = max(h, Z(i, j)) with the per-cell constant Z = BIAS + ge * (i + j): along a lane's column it grows by ge per row,
// from step to step by ge.  The end cell is the maximum of H = h - Z; per lane and half only the running maximum is kept
// (one subtraction and one VIMNMX per cell pair), and a step that raises it leaves the lane's K state words and its step
// number in shared memory, from which the row is read back once per band pair.  Order among equal maxima: smallest
// reference position, then smallest query position (parasail's sw: `WH > score`, else `WH == score && j - 1 < end_ref`).
// The ZERO stop of the traceback needs no direction bit: the walker tracks the score of the cell it stands on.
struct Smx16hLocal {
    uint32_t zs;      // Z of the lane's row 0 at the current step, both halves
    uint32_t ge2;     // ge in both halves
    uint32_t lbest;   // best H of the lane in this band pair, both halves
    int s_lo, s_hi;   // step of the first time the half reached its lbest
    uint32_t *snap;   // shared: [half][k][lane] state words Hc[k] of that step, then [half][lane] zs of that step
};
template <int K>
__device__ __forceinline__ void smx16h_local_step(Smx16hLocal &L, const uint32_t (&Hc)[K], uint32_t m, int s, int lane)
{
    if (m != L.lbest) {   // some half of this lane reached a new maximum (strictly: m >= lbest in both halves)
        const bool up_lo = (m & 0xffffu) != (L.lbest & 0xffffu), up_hi = (m >> 16) != (L.lbest >> 16);
        if (up_lo) {
#pragma unroll
            for (int k = 0; k < K; ++k) L.snap[k * 32 + lane] = Hc[k];
            L.snap[2 * K * 32 + lane] = L.zs;
            L.s_lo = s;
        }
        if (up_hi) {
#pragma unroll
            for (int k = 0; k < K; ++k) L.snap[(K + k) * 32 + lane] = Hc[k];
            L.snap[2 * K * 32 + 32 + lane] = L.zs;
            L.s_hi = s;
        }
        L.lbest = m;
    }
    L.zs += L.ge2;
}

// EDGE: a chunk in which some lane's half is not inside the matrix (column < 0 or >= n).  Such a half HOLDS its state:
// `hold` has 0xffff in every half that is inside, and the new state is merged as (new & hold) | (old & ~hold) -- one
// LOP3 per state register.  A half therefore sits at its boundary-column values until it enters and at its
// column n - 1 values after it has left (the end-cell candidates are read there once per band pair).
__device__ __forceinline__ uint32_t smx16h_merge(uint32_t nv, uint32_t ov, uint32_t hold)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(hold), "r"(nv), "r"(ov));  // hold ? nv : ov, bitwise
    return r;
}

__device__ __forceinline__ uint32_t smx16h_max(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

// NOTR: score only (checkpointed problems, SMX_MODE_CKPT): the four maxima without their direction flags
template <int K, int k, bool EDGE = false, bool LOCAL = false, bool NOTR = false>
struct Smx16hCell {
    static __device__ __forceinline__ void run(uint32_t (&Hc)[K], uint32_t (&El)[K], const uint32_t (&plo)[K], const uint32_t (&phi)[K],
                                               Smx16hCol &c, uint32_t sel, uint32_t negc2, uint32_t hold = 0xffffffffu)
    {
        constexpr int A = K == 8 ? (k & 1) : (k >> 3);
        constexpr int SH = 4 * (k & 7);
        uint32_t &n_lo = c.w_lo[A];
        uint32_t &n_hi = c.w_hi[A];
        const uint32_t hdiag = c.hd + smx_subst_u(plo[k], phi[k], sel);  // no carry between the halves: s'' >= 0, no wrap
        uint32_t e, f, h;
        if (NOTR) {
            e = smx16h_max(El[k], Hc[k]);
            f = smx16h_max(c.fu, c.hcu);
            h = smx16h_max(hdiag, smx16h_max(f, e));
        } else {
            e = smx16h_max_tr<(4u << SH), (SMX16H_ALU_FLAGS < 1)>(El[k], Hc[k], n_lo, n_hi, c.one);   // extend | open
            f = smx16h_max_tr<(8u << SH), (SMX16H_ALU_FLAGS < 3)>(c.fu, c.hcu, n_lo, n_hi, c.one);
            const uint32_t t = smx16h_max_tr<(2u << SH), (SMX16H_ALU_FLAGS < 2)>(f, e, n_lo, n_hi, c.one);
            h = smx16h_max_tr<(1u << SH), (SMX16H_ALU_FLAGS < 4)>(hdiag, t, n_lo, n_hi, c.one);
        }
        if (LOCAL) {
            asm("max.u16x2 %0, %0, %1;" : "+r"(h) : "r"(c.z));            // H = max(H, 0)
            const uint32_t u = EDGE ? ((h - c.z) & hold) : (h - c.z);     // H itself (>= 0 in both halves: no borrow); a held half does not count
            asm("max.u16x2 %0, %0, %1;" : "+r"(c.m) : "r"(u));
            c.z += c.ge2;
        }
        const uint32_t hc = h + negc2;  // - (go - ge) in both halves; the low half never borrows (bias)
        c.hd = Hc[k];
        Hc[k] = EDGE ? smx16h_merge(hc, Hc[k], hold) : hc;
        El[k] = EDGE ? smx16h_merge(e, El[k], hold) : e;
        c.fu = f;
        c.hcu = hc;
        Smx16hCell<K, k + 1, EDGE, LOCAL, NOTR>::run(Hc, El, plo, phi, c, sel, negc2, hold);
    }
};
template <int K, bool EDGE, bool LOCAL, bool NOTR>
struct Smx16hCell<K, K, EDGE, LOCAL, NOTR> {
    static __device__ __forceinline__ void run(uint32_t (&)[K], uint32_t (&)[K], const uint32_t (&)[K], const uint32_t (&)[K], Smx16hCol &, uint32_t,
                                               uint32_t, uint32_t = 0) {}
};

// per-row substitution profile in the diagonal frame: byte c = s(query base, reference code c) + go + ge (0..127);
// a query wildcard / padding row scores 0 against everything
__device__ __forceinline__ uint32_t smx16h_profile(int qcode, int ma2, int mi2, int wild2)
{
    if (qcode > 3) return (uint32_t)wild2 * 0x01010101u;
    uint32_t p = (uint32_t)mi2 * 0x01010101u;
    p &= ~(0xffu << (8 * qcode));
    p |= (uint32_t)ma2 << (8 * qcode);
    return p;
}

__device__ __forceinline__ uint32_t smx16h_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ int smx16h_lo(uint32_t v) { return (int)(v & 0xffffu); }
__device__ __forceinline__ int smx16h_hi(uint32_t v) { return (int)(v >> 16); }

// One interior 32-column chunk of a band pair (see smx16_fast_chunk for the structure).  `tconst` turns the
// stored half of the last query row back into H (semi-global end in the last row): H = half - ge * j + tconst.
// SPLIT (K = 4, see the kernel): the lane's 4 nibbles per half are one 16-bit word; tp_lo / tp_hi point at the lane's
// word of step s0 in the half-band's [step][lane] region and advance by one 64-byte line per step.
template <int K, bool HAS_HI, bool TRACK, bool SPLIT, bool LOCAL = false, bool NOTR = false>
__device__ __forceinline__ void smx16h_fast_chunk(uint32_t (&Hc)[K], uint32_t (&El)[K], const uint32_t (&plo)[K], const uint32_t (&phi)[K],
                                                  uint32_t &hd, uint32_t &out_h, uint32_t &out_f, uint32_t &sel_prev, const uint4 *tops,
                                                  uint32_t lane0_mask, uint32_t negc2, uint32_t *tp_lo, uint32_t *tp_hi, uint32_t *bl, uint32_t *bh,
                                                  uint32_t st_flags, uint32_t tmask, int last_k, int jbase, int ge, int tconst, int &row_best,
                                                  int &row_best_j, uint32_t one, Smx16hLocal *L = nullptr, int s0 = 0, int lane = 0)
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
        Smx16hCol c;
        c.hd = hd; c.fu = fu; c.hcu = hu; c.w_lo[0] = 0; c.w_lo[1] = 0; c.w_hi[0] = 0; c.w_hi[1] = 0; c.one = one;
        if (LOCAL) { c.z = L->zs; c.ge2 = L->ge2; c.m = L->lbest; }
        Smx16hCell<K, 0, false, LOCAL, NOTR>::run(Hc, El, plo, phi, c, sel, negc2);
        if (LOCAL) smx16h_local_step<K>(*L, Hc, HAS_HI ? c.m : ((c.m & 0xffffu) | (L->lbest & 0xffff0000u)), s0 + ss, lane);
        hd = hu;
        out_h = c.hcu;
        out_f = c.fu;
        if (NOTR) {
        } else if (SPLIT) {
            reinterpret_cast<unsigned short *>(tp_lo)[ss * 32] = (unsigned short)c.w_lo[0];
            if (HAS_HI) reinterpret_cast<unsigned short *>(tp_hi)[ss * 32] = (unsigned short)c.w_hi[0];
        } else {
            tp_lo[ss * 32] = c.w_lo[0] | c.w_lo[1];
            if (HAS_HI) tp_hi[ss * 32] = c.w_hi[0] | c.w_hi[1];
        }
        uint32_t fl = st_flags;
        asm volatile("" : "+r"(fl));  // keeps the store predicates short-lived (VIMNMX needs the predicate registers)
        if (fl & 1u) __stcg(&bl[ss], __byte_perm(out_h, out_f, 0x5410));
        if (HAS_HI) { if (fl & 2u) __stcg(&bh[ss], __byte_perm(out_h, out_f, 0x7632)); }
        if (TRACK) {
            uint32_t tm = tmask;
            asm volatile("" : "+r"(tm));
            if (tm) {  // the lane that owns the last query row
                uint32_t hw = Hc[0];
#pragma unroll
                for (int k = 1; k < K; ++k) if (k == last_k) hw = Hc[k];
                const int jj = jbase + ss;
                const int hv = ((tm & 1u) ? smx16h_lo(hw) : smx16h_hi(hw)) - ge * jj + tconst;
                if (hv > row_best) { row_best = hv; row_best_j = jj; }
            }
        }
    }
}


// One 32-step chunk at an edge of the matrix (see Smx16hCell<EDGE>): the interior chunk plus the hold masks, per-lane
// predicates on the direction-word and boundary-line stores, and validity on the last-row tracking.  `jl0` = column
// of this lane's low half at the first step of the chunk (the high half is LAG columns behind).
template <int K, bool HAS_HI, bool SPLIT, bool LOCAL = false, bool NOTR = false>
__device__ __forceinline__ void smx16h_edge_chunk(uint32_t (&Hc)[K], uint32_t (&El)[K], const uint32_t (&plo)[K], const uint32_t (&phi)[K],
                                                  uint32_t &hd, uint32_t &out_h, uint32_t &out_f, uint32_t &sel_prev, const uint4 *tops,
                                                  uint32_t lane0_mask, uint32_t negc2, uint32_t *tp_lo, uint32_t *tp_hi, uint32_t *bl, uint32_t *bh,
                                                  uint32_t st_flags, uint32_t tmask, int last_k, int jl0, int n, int ge, int tconst, int &row_best,
                                                  int &row_best_j, uint32_t one, Smx16hLocal *L = nullptr, int s0 = 0, int lane = 0)
{
#pragma unroll 1
    for (int ss = 0; ss < 32; ++ss) {
        uint32_t hu = __shfl_up_sync(SMX_FULL_MASK, out_h, 1);
        uint32_t fu = __shfl_up_sync(SMX_FULL_MASK, out_f, 1);
        uint32_t sel = __shfl_up_sync(SMX_FULL_MASK, sel_prev, 1);
        const uint4 tv = tops[ss];
        hu = (tv.x & lane0_mask) | (hu & ~lane0_mask);
        fu = (tv.y & lane0_mask) | (fu & ~lane0_mask);
        sel = (tv.z & lane0_mask) | (sel & ~lane0_mask);
        sel_prev = sel;
        const int j_lo = jl0 + ss, j_hi = j_lo - SMX16_LAG;
        const bool v_lo = (unsigned)j_lo < (unsigned)n;
        const bool v_hi = HAS_HI && (unsigned)j_hi < (unsigned)n;
        const uint32_t hold = (v_lo ? 0x0000ffffu : 0u) | (v_hi ? 0xffff0000u : 0u);
        Smx16hCol c;
        c.hd = hd; c.fu = fu; c.hcu = hu; c.w_lo[0] = 0; c.w_lo[1] = 0; c.w_hi[0] = 0; c.w_hi[1] = 0; c.one = one;
        if (LOCAL) { c.z = L->zs; c.ge2 = L->ge2; c.m = L->lbest; }
        Smx16hCell<K, 0, true, LOCAL, NOTR>::run(Hc, El, plo, phi, c, sel, negc2, hold);
        if (LOCAL) smx16h_local_step<K>(*L, Hc, c.m, s0 + ss, lane);
        hd = smx16h_merge(hu, hd, hold);
        out_h = c.hcu;
        out_f = c.fu;
        if (NOTR) {
        } else if (SPLIT) {
            if (v_lo) reinterpret_cast<unsigned short *>(tp_lo)[ss * 32] = (unsigned short)c.w_lo[0];
            if (v_hi) reinterpret_cast<unsigned short *>(tp_hi)[ss * 32] = (unsigned short)c.w_hi[0];
        } else {
            if (v_lo) tp_lo[ss * 32] = c.w_lo[0] | c.w_lo[1];
            if (v_hi) tp_hi[ss * 32] = c.w_hi[0] | c.w_hi[1];
        }
        if ((st_flags & 1u) && v_lo) __stcg(&bl[ss], __byte_perm(out_h, out_f, 0x5410));
        if (HAS_HI) { if ((st_flags & 2u) && v_hi) __stcg(&bh[ss], __byte_perm(out_h, out_f, 0x7632)); }
        if (tmask) {  // the lane that owns the last query row (semi-global end in the last row)
            const bool lo_tr = (tmask & 1u) != 0;
            if (lo_tr ? v_lo : v_hi) {
                uint32_t hw = Hc[0];
#pragma unroll
                for (int k = 1; k < K; ++k) if (k == last_k) hw = Hc[k];
                const int jj = lo_tr ? j_lo : j_hi;
                const int hv = (lo_tr ? smx16h_lo(hw) : smx16h_hi(hw)) - ge * jj + tconst;
                if (hv > row_best) { row_best = hv; row_best_j = jj; }
            }
        }
    }
}

// boundary lines: one uint32 per column = F^ << 16 | Hc (biased unsigned halves)
// CK: the problems carry SMX_MODE_CKPT -- score-only cells, every band's bottom line and the pair checkpoints go to the
// problem's region of the arena (smx_ck_layout) instead of direction bits; no claimed boundary lines.
template <int W, bool LOCAL = false, bool CK = false>
__global__ void __launch_bounds__(W == 1 ? 32 * SMX16H_WPC : W * 32, (CK ? SMX16H_CK_WARPS_PER_SM : SMX16_MIN_WARPS_PER_SM) / (W == 1 ? SMX16H_WPC : W)) smx_dp_forward16h(const SmxFwdArgs a)
{
    static_assert(!(LOCAL && CK), "local problems keep their direction bits");
    constexpr bool TEAM = W > 1;
    constexpr bool SPLIT_OK = W == 1;  // the K = 4 code for a lone last band exists in the one-warp-per-problem kernel only
    constexpr int LAG = SMX16_LAG;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    __shared__ int s_slot[W == 1 ? SMX16H_WPC : 1];
    // claim a pair of boundary lines (see SmxFwdArgs::slot_flags); released when the worker leaves
    int worker = TEAM ? blockIdx.x : ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (a.slot_flags) {
        if (TEAM ? threadIdx.x == 0 : lane == 0) {
            unsigned smid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            int sl = (int)((smid * 61u + (unsigned)worker * 17u) % (unsigned)a.n_slots);
            while (atomicCAS(&a.slot_flags[sl], 0, 1) != 0) sl = sl + 1 == a.n_slots ? 0 : sl + 1;
            __threadfence();  // acquire: the previous owner's line stores are ordered before ours
            s_slot[TEAM ? 0 : wib] = sl;
        }
        if (TEAM) __syncthreads(); else __syncwarp();
        worker = s_slot[TEAM ? 0 : wib];
    }
    uint32_t *line0 = reinterpret_cast<uint32_t *>(a.boundary + (size_t)worker * 2 * (size_t)a.nmax);
    uint32_t *line1 = line0 + a.nmax;
    const int go = a.sc.go, ge = a.sc.ge;
    const int cc = go - ge;                                   // Hc = H^ - cc
    // stored half = value + BIAS > 0; local mode keeps the zero level Z = BIAS + ge * (i + j) non-negative for the 31 steps a
    // lane of band 0 waits left of the matrix
    const int BIAS = 3 * go + (cc < 0 ? -cc : cc) + 16 + (LOCAL ? 32 * ge : 0);
    const uint32_t negc2 = (uint32_t)(-cc) * 0x00010001u;     // 32-bit add that subtracts cc from both halves
    const int ma2 = a.sc.match + go + ge, mi2 = a.sc.mismatch + go + ge, wild2 = go + ge;
    const uint32_t lane0_mask = lane == 0 ? 0xffffffffu : 0u;
    const uint32_t one = a.one;  // 1, but unknown to ptxas: keeps the predicated flag updates on the FMA pipe (IMAD)
    const int u_edge = BIAS - 2 * go;       // Hc of row -1 / column -1 (global start): H = -(go + x ge)  ->  H^ = -go - ge
    const int u_corner = BIAS - ge - go;    // Hc(-1, -1): H = 0 -> H^ = -2 ge
    const int u_free = BIAS - ge - cc;      // free start: H = 0 at (x, -1) / (-1, x)  ->  Hc = ge * x + u_free

    __shared__ int s_ticket;
    __shared__ volatile int s_prog[W];
    __shared__ int s_red[W][6];
    __shared__ uint4 s_top[W == 1 ? SMX16H_WPC : W][32];
    __shared__ uint32_t s_snap[LOCAL ? (W == 1 ? SMX16H_WPC : W) : 1][LOCAL ? (2 * 8 + 2) * 32 : 1];  // Smx16hLocal::snap, per warp

    for (int served = 0; served < a.max_tickets; ++served) {
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
        const bool free_beg = LOCAL || mode == 2;  // SMX_MODE_SG_QB_DB; local: first row and column are 0 as well
        const uint8_t *s1 = a.pool + pr.s1_off;
        const uint8_t *s2 = a.pool + pr.s2_off;
        uint32_t *trace = reinterpret_cast<uint32_t *>(a.trace + pr.trace_off);
        const SmxCkLayout ckl = smx_ck_layout(m, n);
        uint32_t *const ck_lines = trace;                    // CK: [band][nal] bottom lines
        uint32_t *const ck_state = trace + (ckl.ck_off >> 2);  // CK: [pair][checkpoint][word][lane]
        const int nbands = (m + 255) / 256;       // bands of the trace layout: 32 lanes x 8 rows
        const int npairs = (nbands + 1) >> 1;
        const int steps = n + 31;          // steps of one band (trace layout)
        // A last band without a partner (odd band count; every problem of at most 256 rows) would leave the high halves
        // idle.  It is split into two half-bands of 128 rows -- 4 rows per lane and half, the lower half-band in the
        // high halves, 64 columns behind -- and runs the same pair code at K = 4: 92 instead of 152 instructions per
        // step.  Its direction words use the K = 4 layout inside the band's region of the arena (smx_traceback switches
        // its addressing for the rows of such a band).
        const bool split = SPLIT_OK && a.split_last != 0 && (nbands & 1);
        const int R0 = (nbands - 1) * 256;           // first row of the last band
        const bool split_hi = split && (m - R0 > 128);
        const int psteps = (nbands > 1 || split_hi) ? steps + LAG : steps;  // steps of a band pair (a single band has no trailing half)
        const int nchunks = (psteps + 31) >> 5;
        // H(i, j) = half - BIAS + cc - ge * (i + j)
        const int hconst = cc - BIAS;

        int col_best = SMX_NEG_INF, col_best_i = 0x7fffffff;
        int row_best = SMX_NEG_INF, row_best_j = 0;
        int corner = 0;
        int loc_best = 0, loc_i = m - 1, loc_j = n - 1;  // LOCAL: best H of this lane's cells, its cell (smallest j, then smallest i)
        const int last_band_idx = nbands - 1;
        const int rr_last = m - 1 - R0;              // last row inside the last band
        const int last_lane = split ? (rr_last & 127) >> 2 : rr_last >> 3;
        const int last_k = split ? rr_last & 3 : rr_last & 7;

        auto run_pair = [&](auto kc, auto splitc, const int pair) {
            constexpr int K = decltype(kc)::value;
            constexpr bool SPLIT = decltype(splitc)::value;
            const int band_lo = SPLIT ? nbands - 1 : 2 * pair, band_hi = SPLIT ? nbands - 1 : 2 * pair + 1;  // bands of the K = 8 layout
            const bool has_hi = SPLIT ? split_hi : band_hi < nbands;
            const int i0_lo = SPLIT ? R0 + lane * K : band_lo * 32 * K + lane * K;
            const int i0_hi = SPLIT ? R0 + 128 + lane * K : band_hi * 32 * K + lane * K;
            // band_lo is even: reads line0, writes line1; band_hi the other way round.  CK: band b writes line b of the problem's
            // own lines (a split last band: its lower half-band writes the otherwise unused line of the last band)
            const uint32_t *top_lo = CK ? ck_lines + (long long)(band_lo > 0 ? band_lo - 1 : 0) * ckl.nal : line0;
            uint32_t *bot_lo = CK ? ck_lines + (long long)band_lo * ckl.nal : line1;
            const uint32_t *top_hi = CK ? ck_lines + (long long)band_lo * ckl.nal : line1;
            uint32_t *bot_hi = CK ? ck_lines + (long long)band_hi * ckl.nal : line0;
            const bool lo_is_last = SPLIT ? !has_hi : band_lo == last_band_idx, hi_is_last = SPLIT ? has_hi : band_hi == last_band_idx;
            // SPLIT: the last band's region of the arena ([step][lane] 32-bit words at K = 8) holds two half-band regions of
            // 16-bit words instead, [half-band][step][lane]: the K = 4 layout of smx_traceback, every warp store one
            // contiguous 64-byte line (scattering the half-words into the K = 8 positions cost more than the split gained)
            unsigned short *const t16 = reinterpret_cast<unsigned short *>(trace + (size_t)band_lo * steps * 32);
            auto split_ptr_lo = [&](int s) { return t16 + ((long long)s * 32 + lane); };
            auto split_ptr_hi = [&](int s) { return t16 + ((long long)(steps + s - LAG) * 32 + lane); };
            uint32_t plo[K], phi[K];
            uint32_t Hc[K], El[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int il = i0_lo + k, ih = i0_hi + k;
                // rows beyond the query: wildcard rows (score 0) in the global modes; in local mode they must never tie the
                // best real cell, so they mismatch everything
                plo[k] = il < m ? smx16h_profile(smx_code(s1[il]), ma2, mi2, wild2) : (uint32_t)(LOCAL ? mi2 : wild2) * 0x01010101u;
                phi[k] = ih < m ? smx16h_profile(smx_code(s1[ih]), ma2, mi2, wild2) : (uint32_t)(LOCAL ? mi2 : wild2) * 0x01010101u;
                Hc[k] = free_beg ? smx16h_pack(ge * il + u_free, ge * ih + u_free) : smx16h_pack(u_edge, u_edge);  // column -1
                El[k] = 0u;                                                                                        // E^ = -inf
            }
            // Hc(i0 - 1, -1): diagonal input of the first row at column 0
            uint32_t hd = smx16h_pack(i0_lo == 0 ? u_corner : (free_beg ? ge * (i0_lo - 1) + u_free : u_edge),
                                      free_beg ? ge * (i0_hi - 1) + u_free : u_edge);
            uint32_t out_h = 0, out_f = 0, sel_prev = 0;
            Smx16hLocal L;
            if (LOCAL) {  // Z of the lane's row 0 at step 0: i + j = i0 - lane (low half), i0_hi - lane - LAG (high half)
                L.zs = smx16h_pack(BIAS + ge * (i0_lo - lane), BIAS + ge * (i0_hi - lane - LAG));
                L.ge2 = (uint32_t)ge * 0x00010001u;
                L.lbest = 0u; L.s_lo = 0; L.s_hi = 0;
                L.snap = s_snap[wib];
            }
            int rch1 = 0, rch2 = 0;  // reference codes of the two previous chunks (the high half reads 64 columns back)
            const bool st_lo = lane == 31 && !lo_is_last, st_hi = lane == 31 && has_hi && !hi_is_last;
            const bool track_lo = mode == 1 && lo_is_last, track_hi = mode == 1 && hi_is_last;

            for (int s0 = 0, chunk = 0; s0 < psteps; s0 += 32, ++chunk) {
                if (TEAM && pair > 0) {
                    const int need = (pair - 1) * nchunks + min(chunk + 4, nchunks);
                    if (lane == 0) { while (s_prog[(pair - 1) % W] < need) __nanosleep(64); }
                    __syncwarp();
                    __threadfence_block();
                }
                __syncwarp();
                if (CK && (chunk % SMX_CK_CHUNKS) == 0) {  // checkpoint: the loop-carried lane state at the top of this chunk
                    uint32_t *cp = ck_state + ((long long)pair * ckl.nck + chunk / SMX_CK_CHUNKS) * (SMX_CK_WORDS * 32) + lane;
#pragma unroll
                    for (int k = 0; k < K; ++k) { cp[k * 32] = Hc[k]; cp[(8 + k) * 32] = El[k]; }
                    cp[16 * 32] = hd; cp[17 * 32] = out_h; cp[18 * 32] = out_f; cp[19 * 32] = sel_prev;
                    cp[20 * 32] = (uint32_t)rch1; cp[21 * 32] = (uint32_t)rch2;
                }
                // chunk prefetch: lane l holds column s0 + l (low band inputs) and column s0 + l - LAG (high band top line)
                const int jc = s0 + lane;
                const int jb = jc - LAG;
                // row above the low band / above the high band: (Hc, F^ = -inf); band 0 reads the matrix edge
                uint32_t topA = smx16h_pack(free_beg ? ge * jc + u_free : u_edge, 0);
                uint32_t topB = smx16h_pack(u_edge, 0);
                int rch0 = 0;
                if (jc < n) {
                    if (band_lo != 0) topA = __ldcg(&top_lo[jc]);
                    rch0 = smx_code(s2[jc]) & 3;
                }
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
                if (s0 >= (has_hi ? 32 + LAG : 32) && s0 + 33 <= n) {  // a pair without a high band needs no lag (single-band problems)
                    uint32_t *tp_lo = SPLIT ? reinterpret_cast<uint32_t *>(split_ptr_lo(s0)) : trace + ((size_t)band_lo * steps + s0) * 32 + lane;
                    uint32_t *tp_hi = SPLIT ? reinterpret_cast<uint32_t *>(split_ptr_hi(s0)) : trace + ((size_t)band_hi * steps + (s0 - LAG)) * 32 + lane;
                    uint32_t *bl = bot_lo + (s0 - 31);
                    uint32_t *bh = bot_hi + (s0 - 31 - LAG);
                    const uint32_t st_flags = (st_lo ? 1u : 0u) | (st_hi ? 2u : 0u);
                    if (track_lo | track_hi) {
                        const int jbase = s0 - lane - (track_lo ? 0 : LAG);
                        const uint32_t tmask = lane == last_lane ? (track_lo ? 0x0000ffffu : 0xffff0000u) : 0u;
                        const int tconst = hconst - ge * (m - 1);
                        if (has_hi) smx16h_fast_chunk<K, true, true, SPLIT, LOCAL, CK>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, bl, bh, st_flags, tmask, last_k, jbase, ge, tconst, row_best, row_best_j, one, &L, s0, lane);
                        else smx16h_fast_chunk<K, false, true, SPLIT, LOCAL, CK>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, bl, bh, st_flags, tmask, last_k, jbase, ge, tconst, row_best, row_best_j, one, &L, s0, lane);
                    } else {
                        if (has_hi) smx16h_fast_chunk<K, true, false, SPLIT, LOCAL, CK>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, bl, bh, st_flags, 0u, 0, 0, 0, 0, row_best, row_best_j, one, &L, s0, lane);
                        else smx16h_fast_chunk<K, false, false, SPLIT, LOCAL, CK>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, bl, bh, st_flags, 0u, 0, 0, 0, 0, row_best, row_best_j, one, &L, s0, lane);
                    }
                } else {
                    // edge chunk: halves outside the matrix hold their state (smx16h_edge_chunk)
                    uint32_t *tp_lo = SPLIT ? reinterpret_cast<uint32_t *>(split_ptr_lo(s0)) : trace + ((long long)band_lo * steps + s0) * 32 + lane;
                    uint32_t *tp_hi = SPLIT ? reinterpret_cast<uint32_t *>(split_ptr_hi(s0)) : trace + ((long long)band_hi * steps + (s0 - LAG)) * 32 + lane;
                    uint32_t *bl = bot_lo + (s0 - 31);
                    uint32_t *bh = bot_hi + (s0 - 31 - LAG);
                    const uint32_t st_flags = (st_lo ? 1u : 0u) | (st_hi ? 2u : 0u);
                    const uint32_t tmask = (track_lo | track_hi) && lane == last_lane ? (track_lo ? 0x0000ffffu : 0xffff0000u) : 0u;
                    const int tconst = hconst - ge * (m - 1);
                    if (has_hi) smx16h_edge_chunk<K, true, SPLIT, LOCAL, CK>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, bl, bh, st_flags, tmask, last_k, s0 - lane, n, ge, tconst, row_best, row_best_j, one, &L, s0, lane);
                    else smx16h_edge_chunk<K, false, SPLIT, LOCAL, CK>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, bl, bh, st_flags, tmask, last_k, s0 - lane, n, ge, tconst, row_best, row_best_j, one, &L, s0, lane);
                }
                rch2 = rch1;
                rch1 = rch0;
                if (TEAM) {
                    if (lane == 31) { __threadfence_block(); s_prog[wib] = pair * nchunks + chunk + 1; }
                }
            }
            if (LOCAL) {  // this band pair's best cell per half, read back from the snapshot of the step that reached it
                __syncwarp();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int best = half ? (int)(L.lbest >> 16) : (int)(L.lbest & 0xffffu);
                    if (best <= 0 || (half && !has_hi)) continue;
                    const uint32_t zsn = L.snap[2 * K * 32 + half * 32 + lane];
                    const int z0 = half ? (int)(zsn >> 16) : (int)(zsn & 0xffffu);
                    int kk = 0;
#pragma unroll
                    for (int k = K - 1; k >= 0; --k) {  // smallest row holding the maximum
                        const uint32_t wv = L.snap[(half * K + k) * 32 + lane];
                        const int hv = (half ? (int)(wv >> 16) : (int)(wv & 0xffffu)) + cc - (z0 + k * ge);
                        if (hv == best) kk = k;
                    }
                    const int bi = (half ? i0_hi : i0_lo) + kk, bj = (half ? L.s_hi - LAG : L.s_lo) - lane;
                    if (best > loc_best || (best == loc_best && (bj < loc_j || (bj == loc_j && bi < loc_i)))) { loc_best = best; loc_i = bi; loc_j = bj; }
                }
            }
            // end-cell candidates (last column): every half has held its column n - 1 values since it left the matrix.
            // Back to H = half + hconst - ge * (i + j); rows ascending with strict '>' (the lanes are combined below).
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = i0_lo + k, hv = smx16h_lo(Hc[k]) + hconst - ge * (i + n - 1);
                if (i < m - 1 && hv > col_best) { col_best = hv; col_best_i = i; }
                if (i == m - 1) corner = hv;
            }
            if (has_hi) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int i = i0_hi + k, hv = smx16h_hi(Hc[k]) + hconst - ge * (i + n - 1);
                    if (i < m - 1 && hv > col_best) { col_best = hv; col_best_i = i; }
                    if (i == m - 1) corner = hv;
                }
            }
            __syncwarp();
        };
        for (int pair = wib % W; pair < npairs; pair += W) {
            if (SPLIT_OK && split && pair == npairs - 1) run_pair(std::integral_constant<int, 4>{}, std::true_type{}, pair);
            else run_pair(std::integral_constant<int, 8>{}, std::false_type{}, pair);
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
        if (LOCAL) {
            // best cell over the lanes (and the warps of a team): maximum, then smallest reference position, then smallest query position
            auto better = [](int b, int j, int i, int ob, int oj, int oi) { return ob > b || (ob == b && (oj < j || (oj == j && oi < i))); };
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int ob = __shfl_xor_sync(SMX_FULL_MASK, loc_best, off), oj = __shfl_xor_sync(SMX_FULL_MASK, loc_j, off), oi = __shfl_xor_sync(SMX_FULL_MASK, loc_i, off);
                if (better(loc_best, loc_j, loc_i, ob, oj, oi)) { loc_best = ob; loc_j = oj; loc_i = oi; }
            }
            if (TEAM) {
                __syncthreads();
                if (lane == 0) { s_red[wib][0] = loc_best; s_red[wib][1] = loc_j; s_red[wib][2] = loc_i; }
                __syncthreads();
                if (threadIdx.x == 0)
                    for (int w = 1; w < W; ++w)
                        if (better(loc_best, loc_j, loc_i, s_red[w][0], s_red[w][1], s_red[w][2])) { loc_best = s_red[w][0]; loc_j = s_red[w][1]; loc_i = s_red[w][2]; }
            }
            if (loc_best <= 0) { loc_best = 0; loc_i = m - 1; loc_j = n - 1; }  // parasail: score 0, end cell = last cell, empty CIGAR
            r.score = loc_best; r.end_q = loc_i; r.end_r = loc_j;
        } else if (mode == 1) {  // SMX_MODE_SG_QE_DE
            int best = SMX_NEG_INF, bi = m - 1, bj = n - 1;
            if (col_best > best) { best = col_best; bi = col_best_i; bj = n - 1; }
            if (row_best > best) { best = row_best; bi = m - 1; bj = row_best_j; }
            r.score = best; r.end_q = bi; r.end_r = bj;
        } else {
            r.score = corner; r.end_q = m - 1; r.end_r = n - 1;
        }
        if (TEAM ? (threadIdx.x == 0) : (lane == 0)) a.results[pid] = r;
    }
    if (a.slot_flags) {
        if (TEAM) __syncthreads(); else __syncwarp();
        if (TEAM ? threadIdx.x == 0 : lane == 0) { __threadfence(); atomicExch(&a.slot_flags[worker], 0); }  // release after this worker's line stores
    }
}
