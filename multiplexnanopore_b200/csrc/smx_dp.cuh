// Forward affine-gap DP (Gotoh) with packed direction bits, one warp per problem.
//
// Replaces the arithmetic of parasail.{nw,sg_qe_de,sg_qb_db,sw}_trace as called from
// savemoney/modules/ref_query_alignment.py:35,58,68,82,343 (semantics: SURVEY.md Appendix A).
//
// Parallelisation (B200: 148 SMs x 4 SMSPs, 32-wide warps):
//   - inter-task: one warp owns one (query, reference) problem; warps pull problems from a
//     size-sorted list through an atomic ticket (persistent grid = SMs x resident CTAs).
//   - intra-task: systolic anti-diagonal wavefront.  Lane t owns K consecutive query rows of the
//     current band of 32*K rows and, at step s, updates column j = s - t for those K rows.  The
//     (H, F) pair of its last row and the reference base travel to lane t+1 by warp shuffles;
//     between bands the bottom row goes through a small per-warp (H, F) line in global memory
//     (L2 resident).
//   - max-plus recurrence on the DPX path: VIMNMX / VIMNMX3 (__vimax3_s32), substitution score
//     by one PRMT from a per-row 4-byte profile register.
//   - direction bits: 4 bits per cell, the K nibbles of a lane's column are one 8/16/32-bit word;
//     words are laid out [band][step][lane] so that every warp store is one contiguous line.
#pragma once
#include "smx_common.cuh"

template <int K> struct SmxTraceWord;
template <> struct SmxTraceWord<1> { typedef uint8_t type; };
template <> struct SmxTraceWord<2> { typedef uint8_t type; };
template <> struct SmxTraceWord<4> { typedef uint16_t type; };
template <> struct SmxTraceWord<8> { typedef uint32_t type; };

// bytes of packed direction bits a problem of m x n cells needs in class K
__host__ __device__ inline int64_t smx_trace_bytes(int m, int n, int K)
{
    const int64_t bands = (m + 32 * K - 1) / (32 * K);
    const int64_t w = K >= 2 ? K / 2 : 1;
    return bands * (int64_t)(n + 31) * 32 * w;
}

// Checkpointed problems (SMX_MODE_CKPT, smx_dp16h.cuh / smx_traceback_ck.cuh): instead of direction bits the problem's
// region of the arena holds the bottom line of every band (one uint32 = F^ << 16 | Hc per column) and, for every band
// pair, the lane state at the top of every SMX_CK_CHUNKS-th 32-step chunk; the traceback recomputes the direction bits
// of the blocks (SMX_CK_CHUNKS * 32 steps) its path crosses.
#ifndef SMX_CK_CHUNKS
// 32-step chunks between two checkpoints = the block the traceback recomputes at a time.  1: an 8 KB tile per walking
// warp, 20 warps per SM (2 with 16 KB tiles and 12 warps: smx_traceback_ck 202 -> 156 ms per 20 000 reads of C2, the
// forward pass +1 % for twice the checkpoint stores)
#define SMX_CK_CHUNKS 1
#endif
#define SMX_CK_WORDS 22   // words per lane of a checkpoint: Hc[8], El[8], hd, out_h, out_f, sel_prev, rch1, rch2
struct SmxCkLayout {
    int32_t nbands, npairs, nal, nchunks, nck;
    int64_t ck_off, bytes;
};
__host__ __device__ inline SmxCkLayout smx_ck_layout(int m, int n)
{
    SmxCkLayout l;
    l.nbands = (m + 255) / 256;
    l.npairs = (l.nbands + 1) / 2;
    l.nal = (n + 31) & ~31;
    l.nchunks = (n + 31 + 64 + 31) >> 5;  // steps of a band pair: n + 31 + the 64-step lag of its high band
    l.nck = (l.nchunks + SMX_CK_CHUNKS - 1) / SMX_CK_CHUNKS;
    l.ck_off = (int64_t)l.nbands * l.nal * 4;
    l.bytes = l.ck_off + (int64_t)l.npairs * l.nck * SMX_CK_WORDS * 128;
    return l;
}
// bytes of a problem's region of the arena
__host__ __device__ inline int64_t smx_region_bytes(int m, int n, int K, int mode)
{
    return (mode & SMX_MODE_CKPT) ? smx_ck_layout(m, n).bytes : smx_trace_bytes(m, n, K);
}

// per-row substitution profile: byte c = score(query base, reference code c), c = 0..3; code 4 -> 0
__device__ __forceinline__ uint32_t smx_profile(int qcode, int match, int mismatch)
{
    if (qcode > 3) return 0u;
    const uint32_t mm = (uint32_t)(mismatch & 0xff), ma = (uint32_t)(match & 0xff);
    uint32_t p = mm * 0x01010101u;
    p &= ~(0xffu << (8 * qcode));
    p |= ma << (8 * qcode);
    return p;
}

// PRMT selector that picks byte `rcode` of the profile (or of the zero register for the
// wildcard) and sign-extends it to 32 bits.
__device__ __forceinline__ uint32_t smx_selector(int rcode)
{
    const uint32_t idx = (uint32_t)rcode;  // 0..3 -> profile bytes, 4 -> first byte of the zero register
    return idx | ((idx | 8u) << 4) | ((idx | 8u) << 8) | ((idx | 8u) << 12);
}

// prmt.b32 in its default mode honours bit 3 of every selector nibble (replicate the sign of the
// selected byte); the __byte_perm intrinsic masks that bit away, hence the inline PTX.
__device__ __forceinline__ int smx_subst(uint32_t profile, uint32_t sel)
{
    int r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(profile), "r"(0u), "r"(sel));
    return r;
}

// both halves at once: low 16 bits from the profile of the low row (bytes 0-3), high 16 bits from the profile
// of the high row (bytes 4-7); selector built per column by smx_dp_forward16
__device__ __forceinline__ uint32_t smx_subst_u(uint32_t profile_lo, uint32_t profile_hi, uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(profile_lo), "r"(profile_hi), "r"(sel));
    return r;
}

struct SmxFwdArgs {
    const uint8_t *pool;
    const SmxProblem *probs;
    const int32_t *order;  // problem indices of this class, largest first
    int32_t n_order;
    int32_t *ticket;
    uint8_t *trace;
    int2 *boundary;  // per worker (warp, or CTA in team mode): 2 lines of nmax (H, F) pairs
    int32_t nmax;
    SmxResult *results;
    SmxScoring sc;
    uint32_t one;  // always 1 (a kernel parameter so that ptxas cannot fold the IMAD it feeds, smx_dp16.cuh)
    // smx_dp16h.cuh only: non-persistent launch.  Every worker (warp, or CTA in team mode) takes `max_tickets` problems
    // and leaves, so that blocks retire all the time and short kernels of other streams (scan / select / chain of the
    // next batch, on a high-priority stream) get SM slots while a long forward launch is running.  Boundary lines are
    // then claimed from `n_slots` line pairs through `slot_flags` (0 = free) instead of being indexed by the worker.
    int32_t max_tickets;
    int32_t n_slots;
    int32_t *slot_flags;
    int32_t split_last;  // smx_dp16h.cuh, W = 1: run a last band without a partner as two half-bands of 128 rows (K = 4)
};

#define SMX_TEAM_WARPS 8  // warps of a CTA that share one large problem (bands interleaved)

// WORKER = one warp (TEAM == false: a warp owns a problem and walks its bands one after the other) or
// one CTA of SMX_TEAM_WARPS warps (TEAM == true: band b belongs to warp b % W and runs two 32-column
// chunks behind band b-1; the (H, F) bottom line goes through L2, progress through shared memory).
template <int K, bool LOCAL, bool TEAM>
__global__ void __launch_bounds__(TEAM ? SMX_TEAM_WARPS * 32 : 128) smx_dp_forward(const SmxFwdArgs a)
{
    typedef typename SmxTraceWord<K>::type TW;
    constexpr int W = TEAM ? SMX_TEAM_WARPS : 1;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;  // warp in block
    const int worker = TEAM ? blockIdx.x : ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int2 *line0 = a.boundary + (size_t)worker * 2 * (size_t)a.nmax;
    int2 *line1 = line0 + a.nmax;
    const int go = a.sc.go, ge = a.sc.ge;

    __shared__ int s_ticket;
    __shared__ volatile int s_prog[SMX_TEAM_WARPS];      // band * nchunks + chunks done, per warp
    __shared__ int s_red[SMX_TEAM_WARPS][6];

    for (;;) {
        int ticket = 0;
        if (TEAM) {
            __syncthreads();  // previous problem fully retired (s_red / s_prog reuse)
            if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1);
            if (threadIdx.x < SMX_TEAM_WARPS) s_prog[threadIdx.x] = -1;
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
        const bool free_beg = LOCAL || mode == 2;  // SMX_MODE_SG_QB_DB
        const uint8_t *s1 = a.pool + pr.s1_off;
        const uint8_t *s2 = a.pool + pr.s2_off;
        TW *trace = reinterpret_cast<TW *>(a.trace + pr.trace_off);
        const int nbands = (m + 32 * K - 1) / (32 * K);
        const int steps = n + 31;
        const int nchunks = (steps + 31) >> 5;

        // end-cell candidates: last column (rows ascending) then last row (columns ascending)
        int col_best = SMX_NEG_INF, col_best_i = 0x7fffffff;
        int row_best = SMX_NEG_INF, row_best_j = 0;
        int corner = 0;
        int loc_best = 0, loc_i = m - 1, loc_j = n - 1;  // local mode: first maximum in row-major order
        const int last_lane = ((m - 1) % (32 * K)) / K;
        const int last_k = (m - 1) % K;

        for (int band = wib % W; band < nbands; band += W) {
            const int i0 = band * 32 * K + lane * K;
            const int2 *top_line = (band & 1) ? line1 : line0;  // written by the previous band
            int2 *bot_line = (band & 1) ? line0 : line1;
            const bool last_band = band == nbands - 1;

            uint32_t prof[K];
            int Hl[K], Hlgo[K], El[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = i0 + k;
                const int qc = i < m ? smx_code(s1[i]) : 4;
                prof[k] = smx_profile(qc, a.sc.match, a.sc.mismatch);
                Hl[k] = free_beg ? 0 : -(go + i * ge);  // H[i][-1]
                Hlgo[k] = Hl[k] - go;
                El[k] = SMX_NEG_INF;
            }
            int hd = (i0 == 0 || free_beg) ? 0 : -(go + (i0 - 1) * ge);  // H[i0-1][-1]
            int out_h = 0, out_f = 0, rc_prev = 4;

            for (int s0 = 0, chunk = 0; s0 < steps; s0 += 32, ++chunk) {
                if (TEAM && band > 0) {
                    // columns s0..s0+31 of the previous band are complete once it finished chunk + 1
                    const int need = (band - 1) * nchunks + min(chunk + 2, nchunks);
                    if (lane == 0) { while (s_prog[(band - 1) % W] < need) __nanosleep(64); }
                    __syncwarp();
                    __threadfence_block();
                }
                // chunk prefetch for lane 0's inputs: top boundary (H, F) and reference codes of columns s0..s0+31
                const int jc = s0 + lane;
                int2 top = make_int2(0, SMX_NEG_INF);
                int rch = 4;
                if (jc < n) {
                    if (band == 0) top.x = free_beg ? 0 : -(go + jc * ge);
                    else top = __ldcg(&top_line[jc]);
                    rch = smx_code(s2[jc]);
                }
                const int send = min(32, steps - s0);
#pragma unroll 1
                for (int ss = 0; ss < send; ++ss) {
                    const int s = s0 + ss;
                    int hu = __shfl_up_sync(SMX_FULL_MASK, out_h, 1);
                    int fu = __shfl_up_sync(SMX_FULL_MASK, out_f, 1);
                    int rc = __shfl_up_sync(SMX_FULL_MASK, rc_prev, 1);
                    const int th = __shfl_sync(SMX_FULL_MASK, top.x, ss);
                    const int tf = __shfl_sync(SMX_FULL_MASK, top.y, ss);
                    const int tr = __shfl_sync(SMX_FULL_MASK, rch, ss);
                    if (lane == 0) { hu = th; fu = tf; rc = tr; }
                    rc_prev = rc;
                    const int j = s - lane;
                    if (j >= 0 && j < n) {
                        const uint32_t sel = smx_selector(rc);
                        const int hu_in = hu;
                        int hugo = hu - go;
                        uint32_t word = 0;
                        int h = 0, f = 0;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const int e_ext = El[k] - ge;
                            const int e_opn = Hlgo[k];
                            const int e = max(e_ext, e_opn);
                            const int f_ext = fu - ge;
                            const int f_opn = hugo;
                            f = max(f_ext, f_opn);
                            const int hdiag = hd + smx_subst(prof[k], sel);
                            // NOTE: the H source is decided on the unclamped maximum.  Writing it as
                            // max(max3(..), 0) followed by `h == f` makes ptxas 12.9 fuse the clamp into
                            // VIMNMX.RELU with a predicate output whose sense is inverted (observed on
                            // sm_100a: INS/DEL swapped); for h > 0 both forms are equivalent.
                            const int h3 = __vimax3_s32(hdiag, e, f);
                            uint32_t nib = (h3 == hdiag) ? SMX_SRC_DIAG : ((h3 == f) ? SMX_SRC_DEL : SMX_SRC_INS);
                            h = h3;
                            if (LOCAL) {
                                if (h3 <= 0) { h = 0; nib = SMX_SRC_ZERO; }
                            }
                            if (e_opn > e_ext) nib |= SMX_BIT_EOPEN;  // tie -> extend
                            if (f_opn > f_ext) nib |= SMX_BIT_FOPEN;
                            word |= nib << (4 * k);
                            if (LOCAL) {
                                const int i = i0 + k;
                                if (i < m && (h > loc_best || (h == loc_best && h > 0 && (j < loc_j || (j == loc_j && i < loc_i))))) {
                                    loc_best = h; loc_i = i; loc_j = j;
                                }
                            }
                            hd = Hl[k];
                            Hl[k] = h;
                            hugo = h - go;
                            Hlgo[k] = hugo;
                            El[k] = e;
                            fu = f;
                        }
                        hd = hu_in;  // H[i0-1][j] is the diagonal input of the next column
                        out_h = h;
                        out_f = f;
                        trace[((size_t)band * steps + s) * 32 + lane] = (TW)word;
                        if (lane == 31 && !last_band) __stcg(&bot_line[j], make_int2(h, f));
                        if (!LOCAL) {
                            if (j == n - 1) {
#pragma unroll
                                for (int k = 0; k < K; ++k) {
                                    const int i = i0 + k;
                                    if (i < m - 1 && Hl[k] > col_best) { col_best = Hl[k]; col_best_i = i; }
                                    if (i == m - 1) corner = Hl[k];
                                }
                            }
                            if (mode == 1 && last_band && lane == last_lane) {
                                int hv = Hl[0];
#pragma unroll
                                for (int k = 1; k < K; ++k) if (k == last_k) hv = Hl[k];
                                if (hv > row_best) { row_best = hv; row_best_j = j; }
                            }
                        }
                    }
                }
                if (TEAM && !last_band) {
                    // lane 31 wrote the bottom line of this chunk: publish after a block-scope fence
                    if (lane == 31) { __threadfence_block(); s_prog[wib] = band * nchunks + chunk + 1; }
                }
            }
            __syncwarp();
        }

        // ---- end cell / score -------------------------------------------------------------
        SmxResult r;
        r.beg_q = 0; r.beg_r = 0; r.n_runs = 0; r.run_start = 0; r.n_s = 0; r.n_h = 0; r.pad0 = 0; r.pad1 = 0; r.pad2 = 0;
        if (LOCAL) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int ob = __shfl_xor_sync(SMX_FULL_MASK, loc_best, off);
                const int oi = __shfl_xor_sync(SMX_FULL_MASK, loc_i, off);
                const int oj = __shfl_xor_sync(SMX_FULL_MASK, loc_j, off);
                if (ob > loc_best || (ob == loc_best && ob > 0 && (oj < loc_j || (oj == loc_j && oi < loc_i)))) {
                    loc_best = ob; loc_i = oi; loc_j = oj;
                }
            }
            if (TEAM) {
                if (lane == 0) { s_red[wib][0] = loc_best; s_red[wib][1] = loc_i; s_red[wib][2] = loc_j; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    for (int w = 1; w < W; ++w) {
                        const int ob = s_red[w][0], oi = s_red[w][1], oj = s_red[w][2];
                        if (ob > loc_best || (ob == loc_best && ob > 0 && (oj < loc_j || (oj == loc_j && oi < loc_i)))) {
                            loc_best = ob; loc_i = oi; loc_j = oj;
                        }
                    }
                }
            }
            r.score = loc_best; r.end_q = loc_i; r.end_r = loc_j;
        } else {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const int ob = __shfl_xor_sync(SMX_FULL_MASK, col_best, off);
                const int oi = __shfl_xor_sync(SMX_FULL_MASK, col_best_i, off);
                if (ob > col_best || (ob == col_best && oi < col_best_i)) { col_best = ob; col_best_i = oi; }
            }
            const int owner = TEAM ? ((nbands - 1) % W) : 0;  // warp that ran the last band
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
        }
        if (TEAM ? (threadIdx.x == 0) : (lane == 0)) a.results[pid] = r;
    }
}
