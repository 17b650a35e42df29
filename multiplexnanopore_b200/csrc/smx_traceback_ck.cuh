// Traceback of the checkpointed problems (SMX_MODE_CKPT): the forward pass (smx_dp_forward16h<W, false, true>) kept no
// direction bits, only every band's bottom line and the lane state of every band pair at the top of every
// SMX_CK_CHUNKS-th chunk.  One warp per problem: for the cell the walk stands on it restores the checkpoint of the block
// (SMX_CK_CHUNKS * 32 steps) that holds the cell's step, re-runs the chunks of that block up to the cell with the SAME chunk
// code the full-trace forward kernel runs (smx16h_fast_chunk / smx16h_edge_chunk: identical comparisons, identical direction
// bits), the direction words going to an 8 KB tile in shared memory, and lane 0 walks inside the tile until the path leaves
// the block or the band pair.  The walk itself is the state machine of smx_traceback (parasail's cigar_template.c through
// MyResult.__init__, ref_query_alignment.py:426-433), including the device-side re-clipping.
//
// A path crosses a band pair (512 rows) in about 543 steps of the skewed frame out of the n + 95 the pair has: for a
// 2 500 x 4 000 problem about 14 % of the cells are computed a second time.  The kernel is latency-bound (one warp per
// problem, lane 0 alone while walking), so it is built for residency: 96 registers, five CTAs of four warps per SM.
#pragma once
#include "smx_dp16h.cuh"
#include "smx_traceback.cuh"

#define SMX_TBCK_WARPS 4                                      // problems per CTA
#ifndef SMX_TBCK_MINBLOCKS
#define SMX_TBCK_MINBLOCKS 5                                  // CTAs per SM the register budget is set for (6: 80 registers, spills, 1.7x slower)
#endif
#define SMX_TBCK_TILE_WORDS (2 * SMX_CK_CHUNKS * 32 * 32)     // [half][step in block][lane] 32-bit words (K = 4: 16-bit words, half used)

struct SmxTbCkArgs {
    const uint8_t *pool;
    const SmxProblem *probs;
    const int32_t *order;  // the checkpointed problems of this chunk
    int32_t n_order;
    const uint8_t *trace;
    SmxResult *results;
    uint32_t *runs;
    SmxScoring sc;
    int32_t split_last;
    uint32_t one;
};

__global__ void __launch_bounds__(32 * SMX_TBCK_WARPS, SMX_TBCK_MINBLOCKS) smx_traceback_ck(const SmxTbCkArgs a)
{
    constexpr int LAG = SMX16_LAG;
    constexpr int BLK = SMX_CK_CHUNKS * 32;  // steps of a block
    extern __shared__ uint32_t s_dyn[];
    __shared__ uint4 s_top[SMX_TBCK_WARPS][32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= a.n_order) return;  // whole warps
    uint32_t *const tile = s_dyn + wib * SMX_TBCK_TILE_WORDS;
    const int pid = a.order[t];
    const SmxProblem pr = a.probs[pid];
    SmxResult r = a.results[pid];
    const uint8_t *s1 = a.pool + pr.s1_off;
    const uint8_t *s2 = a.pool + pr.s2_off;
    const int m = pr.m, n = pr.n;
    const int mode = pr.mode & SMX_MODE_MASK;
    const bool free_beg = mode == 2;  // SMX_MODE_SG_QB_DB

    // the forward kernel's constants (smx_dp_forward16h)
    const int go = a.sc.go, ge = a.sc.ge;
    const int cc = go - ge;
    const int BIAS = 3 * go + (cc < 0 ? -cc : cc) + 16;
    const uint32_t negc2 = (uint32_t)(-cc) * 0x00010001u;
    const int ma2 = a.sc.match + go + ge, mi2 = a.sc.mismatch + go + ge, wild2 = go + ge;
    const uint32_t lane0_mask = lane == 0 ? 0xffffffffu : 0u;
    const uint32_t one = a.one;
    const int u_edge = BIAS - 2 * go;
    const int u_free = BIAS - ge - cc;

    const SmxCkLayout ckl = smx_ck_layout(m, n);
    const uint32_t *const ck_lines = reinterpret_cast<const uint32_t *>(a.trace + pr.trace_off);
    const uint32_t *const ck_state = ck_lines + (ckl.ck_off >> 2);
    const int nbands = ckl.nbands, npairs = ckl.npairs;
    const bool split = a.split_last != 0 && pr.cls == 10 && (nbands & 1);  // one-warp class only, as in the forward kernel
    const int R0 = (nbands - 1) * 256;
    const bool split_hi = split && (m - R0 > 128);

    SmxWalk w;
    w.out = a.runs + pr.cigar_off; w.nruns = 0; w.cur_op = 0; w.cur_len = 0;
    w.clip = (pr.mode & SMX_MODE_CLIP) ? mode : 0;
    w.go = a.sc.go; w.ge = a.sc.ge; w.ma = a.sc.match; w.mi = a.sc.mismatch;
    w.acc = 0; w.q_used = 0; w.r_used = 0; w.best = 0; w.have = 0; w.cut_runs = 0; w.cut_q = 0; w.cut_r = 0;

    int i = r.end_q, j = r.end_r;
    if (lane == 0 && mode == SMX_MODE_SG_QE_DE) {
        if (i + 1 == m) w.emit(SMX_OP_D, n - 1 - j);
        else if (j + 1 == n) w.emit(SMX_OP_I, m - 1 - i);
    }
    int where = 0;  // 0 = DIAG (H), 1 = INS (E, emits 'D'), 2 = DEL (F, emits 'I')

    // Re-runs chunks [SMX_CK_CHUNKS * q, c_end] of band pair `pair` from its checkpoint q; direction words -> tile.
    auto recompute = [&](auto kc, auto splitc, const int pair, const int q, const int c_end) {
        constexpr int K = decltype(kc)::value;
        constexpr bool SPLIT = decltype(splitc)::value;
        const int band_lo = SPLIT ? nbands - 1 : 2 * pair, band_hi = SPLIT ? nbands - 1 : 2 * pair + 1;
        const bool has_hi = SPLIT ? split_hi : band_hi < nbands;
        const int i0_lo = SPLIT ? R0 + lane * K : band_lo * 32 * K + lane * K;
        const int i0_hi = SPLIT ? R0 + 128 + lane * K : band_hi * 32 * K + lane * K;
        const uint32_t *top_lo = ck_lines + (long long)(band_lo > 0 ? band_lo - 1 : 0) * ckl.nal;
        const uint32_t *top_hi = ck_lines + (long long)band_lo * ckl.nal;
        uint32_t plo[K], phi[K], Hc[K], El[K];
        const uint32_t *cp = ck_state + ((long long)pair * ckl.nck + q) * (SMX_CK_WORDS * 32) + lane;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int il = i0_lo + k, ih = i0_hi + k;
            plo[k] = il < m ? smx16h_profile(smx_code(s1[il]), ma2, mi2, wild2) : (uint32_t)wild2 * 0x01010101u;
            phi[k] = ih < m ? smx16h_profile(smx_code(s1[ih]), ma2, mi2, wild2) : (uint32_t)wild2 * 0x01010101u;
            Hc[k] = cp[k * 32];
            El[k] = cp[(8 + k) * 32];
        }
        uint32_t hd = cp[16 * 32], out_h = cp[17 * 32], out_f = cp[18 * 32], sel_prev = cp[19 * 32];
        int rch1 = (int)cp[20 * 32], rch2 = (int)cp[21 * 32];
        int row_best = 0, row_best_j = 0;
        const int sb0 = q * BLK;
        for (int chunk = q * SMX_CK_CHUNKS; chunk <= c_end; ++chunk) {
            const int s0 = chunk * 32;
            __syncwarp();
            const int jc = s0 + lane;
            const int jb = jc - LAG;
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
            // tile words of this chunk: [half][s - sb0][lane]; K = 4: 16-bit words
            uint32_t *tp_lo = SPLIT ? reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned short *>(tile) + (s0 - sb0) * 32 + lane) : tile + (s0 - sb0) * 32 + lane;
            uint32_t *tp_hi = SPLIT ? reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned short *>(tile) + (BLK + s0 - sb0) * 32 + lane) : tile + (BLK + s0 - sb0) * 32 + lane;
            if (s0 >= (has_hi ? 32 + LAG : 32) && s0 + 33 <= n) {
                if (has_hi) smx16h_fast_chunk<K, true, false, SPLIT>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, nullptr, nullptr, 0u, 0u, 0, 0, 0, 0, row_best, row_best_j, one);
                else smx16h_fast_chunk<K, false, false, SPLIT>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, nullptr, nullptr, 0u, 0u, 0, 0, 0, 0, row_best, row_best_j, one);
            } else {
                if (has_hi) smx16h_edge_chunk<K, true, SPLIT>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, nullptr, nullptr, 0u, 0u, 0, s0 - lane, n, 0, 0, row_best, row_best_j, one);
                else smx16h_edge_chunk<K, false, SPLIT>(Hc, El, plo, phi, hd, out_h, out_f, sel_prev, s_top[wib], lane0_mask, negc2, tp_lo, tp_hi, nullptr, nullptr, 0u, 0u, 0, s0 - lane, n, 0, 0, row_best, row_best_j, one);
            }
            rch2 = rch1;
            rch1 = rch0;
        }
        __syncwarp();
    };

    for (;;) {  // warp-uniform: (i, j, where) are lane 0's, broadcast after every segment
        if (i < 0 || j < 0) break;
        const bool in_split = split && i >= R0;
        const int K = in_split ? 4 : 8;
        const int pair = in_split ? npairs - 1 : i >> 9;
        int half, lane_t, k;
        if (in_split) { const int rin = i - R0; half = rin >> 7; lane_t = (rin & 127) >> 2; k = rin & 3; }
        else { const int rin = i & 511; half = rin >> 8; lane_t = (rin & 255) >> 3; k = rin & 7; }
        const int s = j + lane_t + (half ? LAG : 0);
        const int q = s / BLK, sb0 = q * BLK;
        if (in_split) recompute(std::integral_constant<int, 4>{}, std::true_type{}, pair, q, s >> 5);
        else recompute(std::integral_constant<int, 8>{}, std::false_type{}, pair, q, s >> 5);
        if (lane == 0) {
            const int wbytes = K >> 1;          // bytes per lane-step word
            const int line = 32 * wbytes;       // bytes per step
            const uint8_t *const tb = reinterpret_cast<const uint8_t *>(tile);
            const int row_lo = in_split ? R0 : pair << 9;  // first row of the pair (or of the split band)
            int srel = s - sb0;
            const uint8_t *wp = tb + (half * BLK + srel) * line + lane_t * wbytes;
            while (srel >= 0 && j >= 0) {
                const uint32_t nib = ((uint32_t)wp[k >> 1] >> ((k & 1) * 4)) & 15u;
                int di = 0, dj = 0;
                if (where == 0) {
                    if (!(nib & 1u)) {  // diagonal
                        w.emit(s1[i] == s2[j] ? SMX_OP_EQ : SMX_OP_X, 1);
                        di = 1; dj = 1;
                    } else where = (nib & 2u) ? 1 : 2;  // "E beat F" in bit 1 (smx_dp16.cuh)
                } else if (where == 1) {
                    w.emit(SMX_OP_D, 1);
                    dj = 1;
                    where = (nib & SMX_BIT_EOPEN) ? 0 : 1;
                } else {
                    w.emit(SMX_OP_I, 1);
                    di = 1;
                    where = (nib & SMX_BIT_FOPEN) ? 0 : 2;
                }
                if (dj) { --j; wp -= line; --srel; }
                if (di) {
                    --i;
                    if (k > 0) --k;
                    else if (lane_t > 0) { --lane_t; k = K - 1; wp -= line + wbytes; --srel; }
                    else if (half == 1) {  // the pair's upper band: last lane, last row, LAG steps earlier
                        half = 0; lane_t = 31; k = K - 1;
                        srel = j + 31 - sb0;
                        wp = tb + srel * line + 31 * wbytes;
                    } else break;  // left the pair (or the matrix) upwards
                    if (i < row_lo) break;
                }
            }
        }
        i = __shfl_sync(SMX_FULL_MASK, i, 0);
        j = __shfl_sync(SMX_FULL_MASK, j, 0);
        where = __shfl_sync(SMX_FULL_MASK, where, 0);
    }
    if (lane != 0) return;
    if (i < 0 && j >= 0) w.emit(SMX_OP_D, j + 1);
    else if (j < 0 && i >= 0) w.emit(SMX_OP_I, i + 1);
    w.flush();
    r.beg_q = 0;
    r.beg_r = 0;
    r.n_runs = w.nruns; r.run_start = 0; r.n_s = 0; r.n_h = 0;
    if (w.clip == 2) {
        if (!w.have || w.best < 0) { r.n_runs = 0; r.n_s = m; r.n_h = n; }
        else { r.n_runs = w.cut_runs; r.n_s = m - w.cut_q; r.n_h = n - w.cut_r; }
    } else if (w.clip == 1) {
        if (!w.have || w.acc - w.best < 0) { r.n_runs = 0; r.n_s = m; r.n_h = n; }
        else { r.run_start = w.cut_runs; r.n_runs = w.nruns - w.cut_runs; r.n_s = w.cut_q; r.n_h = w.cut_r; }
    }
    a.results[pid] = r;
}
