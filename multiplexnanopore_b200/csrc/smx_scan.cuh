// Diagonal match-count scan: counts[k] = #{ i : ref_ext[k + i] == query[i] }.
// Replaces af.k_mer_offset_analysis_2 (savemoney/modules/cython_functions/alignment_functions.pyx:46-64)
// together with the window construction around it (ref_query_alignment.py:351-366, 385):
//   circular: ref_ext = reference tiled, k in [0, Nr)
//   linear  : ref_ext = reference zero-padded by Nq-1 on both sides, k in [0, Nr+Nq-1)
//
// The sequence pool is packed once per upload into two BIT PLANES (smx_pack_kernel: for every 32 bases one word
// with the low bits and one with the high bits of the 2-bit codes, plus one flag byte for bytes other than
// A/C/G/T), so that one 32-bit word compares 32 bases:  mismatch = (r_lo ^ q_lo) | (r_hi ^ q_hi).
//
// One CTA = 1024 consecutive diagonals of one (reference, query) pair.  The packed query and the reference window
// the tile touches are staged in shared memory.  Diagonal k needs the window shifted by k mod 32 bits; the 32
// diagonals k = k0 + s + 32 y (y = 0..31) share ONE shifted stream S_s, each reading it one word further.  A group
// of 8 lanes owns the shift s: lane t holds S_s[4t .. 4t+3] (+ the step) in registers for its 4 diagonals; stepping
// to the next 32 query bases renames three registers and takes the fourth from lane t+1 with one warp shuffle per
// plane (the last lane of the group builds the fresh word with a funnel shift).  Per 128 cells a lane spends
// 8 LOP3 + 4 POPC + 4 adds + ~8 instructions of stream upkeep, 2.5 x fewer than a 16-bases-per-word formulation.
// Tiles that touch a flagged word fall back to an exact byte-compare loop in the same kernel (equality in the
// reference is on raw characters).
#pragma once
#include "smx_common.cuh"

#define SMX_SCAN_THREADS 256
#define SMX_SCAN_DPT 4                                   // diagonals per thread
#define SMX_SCAN_TILE (SMX_SCAN_THREADS * SMX_SCAN_DPT)  // diagonals per CTA

struct SmxScanPair {
    int64_t ref_off, qry_off, counts_off;
    int32_t nr, nq;
    int32_t max_k, tile0;  // tile0 = index of this pair's first tile in the launch
};

struct SmxScanArgs {
    const uint8_t *pool;
    const uint2 *pool2;      // bit planes of the pool: .x = low bits, .y = high bits of 32 bases
    const uint8_t *special;  // per 32-base word: holds a byte other than A/C/G/T
    const SmxScanPair *pairs;
    const int32_t *tile_pair;  // tile -> pair index
    int32_t *counts;
    int32_t topology;  // 0 circular, 1 linear
};

// one thread per 32 bases of the pool
__global__ void __launch_bounds__(256) smx_pack_kernel(const uint8_t *pool, int64_t n, uint2 *pool2, uint8_t *special, int64_t n_words)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t lo = 0, hi = 0;
    int sp = 0;
    const int64_t base = w * 32;
    if (base + 32 <= n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(pool + base);  // the pool allocation is 16-byte aligned
        const uint4 raw0 = src[0], raw1 = src[1];
        const uint32_t words[8] = {raw0.x, raw0.y, raw0.z, raw0.w, raw1.x, raw1.y, raw1.z, raw1.w};
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const int c = smx_code((words[b >> 2] >> (8 * (b & 3))) & 0xffu);
            sp |= c > 3;
            lo |= (uint32_t)(c & 1) << b;
            hi |= (uint32_t)((c >> 1) & 1) << b;
        }
    } else {
        for (int b = 0; b < 32 && base + b < n; ++b) {
            const int c = smx_code(pool[base + b]);
            sp |= c > 3;
            lo |= (uint32_t)(c & 1) << b;
            hi |= (uint32_t)((c >> 1) & 1) << b;
        }
    }
    pool2[w] = make_uint2(lo, hi);
    special[w] = (uint8_t)sp;
}

// bit planes of the 32 bases starting at pool position p (any alignment)
__device__ __forceinline__ uint2 smx_fetch32(const uint2 *pool2, const uint8_t *special, int64_t p, int &sp)
{
    const int64_t w = p >> 5;
    const int sh = (int)(p & 31);
    const uint2 a = pool2[w], b = pool2[w + 1];
    sp |= special[w] | special[w + 1];
    return make_uint2(__funnelshift_r(a.x, b.x, sh), __funnelshift_r(a.y, b.y, sh));
}

__device__ __forceinline__ uint32_t smx_lowmask(int nbits) { return nbits >= 32 ? 0xffffffffu : ((1u << nbits) - 1u); }

// one step of a lane: 32 query bases against its four stream words, then slide the stream by one word.
// U = sub-step inside a group of four (compile time): logical stream word d lives in register slot (d + U) & 3, so
// sliding renames registers instead of moving them.
template <bool LINEAR, int U, bool MASKED>
__device__ __forceinline__ void smx_scan_step(uint32_t (&alo)[4], uint32_t (&ahi)[4], uint32_t (&av)[4], int (&mism)[4], const uint2 q,
                                              const uint32_t qm, const uint2 rnext, const uint32_t vnext, uint2 &rprev, uint32_t &vprev,
                                              const int s, const bool last_lane)
{
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int slot = (d + U) & 3;
        uint32_t mis = (alo[slot] ^ q.x) | (ahi[slot] ^ q.y);
        if (LINEAR) mis |= ~av[slot];
        if (MASKED) mis &= qm;
        mism[d] += __popc(mis);
    }
    // lane t takes S_s[4t + 4 + w] from lane t+1 (its logical word 0); the last lane of the group builds it
    const int s0 = U & 3;
    uint32_t nlo = __shfl_down_sync(SMX_FULL_MASK, alo[s0], 1, 8);
    uint32_t nhi = __shfl_down_sync(SMX_FULL_MASK, ahi[s0], 1, 8);
    const uint32_t flo = __funnelshift_r(rprev.x, rnext.x, s);
    const uint32_t fhi = __funnelshift_r(rprev.y, rnext.y, s);
    alo[s0] = last_lane ? flo : nlo;
    ahi[s0] = last_lane ? fhi : nhi;
    rprev = rnext;
    if (LINEAR) {
        const uint32_t nv = __shfl_down_sync(SMX_FULL_MASK, av[s0], 1, 8);
        const uint32_t fv = __funnelshift_r(vprev, vnext, s);
        av[s0] = last_lane ? fv : nv;
        vprev = vnext;
    }
}

// smem layout: [query planes (uint2 per 32 bases, padded to a multiple of 4 words) | reference window planes |
//               reference window validity (linear)]
template <bool LINEAR>
__global__ void __launch_bounds__(SMX_SCAN_THREADS) smx_scan_kernel(const SmxScanArgs a)
{
    extern __shared__ uint2 smem2[];
    const int tile = blockIdx.x;
    const int pi = a.tile_pair[tile];
    const SmxScanPair pr = a.pairs[pi];
    const int nr = pr.nr, nq = pr.nq;
    const int k0 = (tile - pr.tile0) * SMX_SCAN_TILE;
    const int tid = threadIdx.x;
    const int shift0 = LINEAR ? (nq - 1) : 0;  // ref_ext[x] = reference[x - shift0] (linear) or reference[x mod nr]
    const int qwords = (nq + 31) >> 5;
    const int qwords4 = (qwords + 3) & ~3;
    const int rwords = qwords4 + 34;           // window words [0, 32 + qwords4] plus one for the funnel shift
    uint2 *qpk = smem2;
    uint2 *rpk = smem2 + qwords4;
    uint32_t *rvalid = reinterpret_cast<uint32_t *>(rpk + rwords);  // linear only

    // ---- stage the packed query and reference window ----
    int special = 0;
    for (int w = tid; w < qwords4; w += SMX_SCAN_THREADS)
        qpk[w] = w < qwords ? smx_fetch32(a.pool2, a.special, pr.qry_off + 32 * w, special) : make_uint2(0u, 0u);
    for (int w = tid; w < rwords; w += SMX_SCAN_THREADS) {
        const int x = k0 + w * 32;
        uint2 v = make_uint2(0u, 0u);
        if (!LINEAR) {
            if (nr >= 64) {
                // 32 bases from position x mod nr; a word that crosses the end continues at position 0
                const int p0 = x % nr;
                v = smx_fetch32(a.pool2, a.special, pr.ref_off + p0, special);
                const int left = nr - p0;  // bases before the wrap
                if (left < 32) {
                    const uint2 h = smx_fetch32(a.pool2, a.special, pr.ref_off, special);
                    const uint32_t m = smx_lowmask(left);
                    v.x = (v.x & m) | (h.x << left);
                    v.y = (v.y & m) | (h.y << left);
                }
            } else {
                for (int b = 0; b < 32; ++b) {
                    const int c = smx_code(a.pool[pr.ref_off + (x + b) % nr]);
                    special |= c > 3;
                    v.x |= (uint32_t)(c & 1) << b;
                    v.y |= (uint32_t)((c >> 1) & 1) << b;
                }
            }
        } else {
            // ref_ext positions outside [shift0, shift0 + nr) are padding that never matches (validity plane)
            const int pos = x - shift0;
            uint32_t valid = 0u;
            if (pos >= 0 && pos < nr) {
                v = smx_fetch32(a.pool2, a.special, pr.ref_off + pos, special);
                valid = smx_lowmask(nr - pos);
            } else if (pos < 0 && pos > -32) {
                const uint2 h = smx_fetch32(a.pool2, a.special, pr.ref_off, special);
                v.x = h.x << (-pos);
                v.y = h.y << (-pos);
                valid = smx_lowmask(nr) << (-pos);
            }
            rvalid[w] = valid;
        }
        rpk[w] = v;
    }
    special = __syncthreads_or(special);

    const int s = tid >> 3;   // shift of this group of 8 lanes
    const int t = tid & 7;    // lane in the group: stream words 4t .. 4t+3
    const int kb = k0 + s + 32 * SMX_SCAN_DPT * t;  // diagonals kb + 32 d, d = 0..3
    int mism[SMX_SCAN_DPT];
#pragma unroll
    for (int d = 0; d < SMX_SCAN_DPT; ++d) mism[d] = 0;

    if (!special) {
        // S_s[y] = window bits [32 y + s, 32 y + s + 32)
        uint32_t alo[4], ahi[4], av[4];
        {
            uint2 r0 = rpk[4 * t];
            uint32_t v0 = LINEAR ? rvalid[4 * t] : 0xffffffffu;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const uint2 r1 = rpk[4 * t + d + 1];
                const uint32_t v1 = LINEAR ? rvalid[4 * t + d + 1] : 0xffffffffu;
                alo[d] = __funnelshift_r(r0.x, r1.x, s);
                ahi[d] = __funnelshift_r(r0.y, r1.y, s);
                av[d] = __funnelshift_r(v0, v1, s);
                r0 = r1; v0 = v1;
            }
        }
        uint2 rprev = rpk[32];  // window word 32 + w of the step
        uint32_t vprev = LINEAR ? rvalid[32] : 0xffffffffu;
        const bool last_lane = t == 7;
        const uint32_t qmask_last = (nq & 31) ? smx_lowmask(nq & 31) : 0xffffffffu;
        // groups of four words that lie completely inside the query run unmasked
        const int full = (qwords - 1) & ~3;  // words [0, full) are complete and not the last word
        int w = 0;
        for (; w < full; w += 4) {
            smx_scan_step<LINEAR, 0, false>(alo, ahi, av, mism, qpk[w + 0], 0u, rpk[33 + w + 0], LINEAR ? rvalid[33 + w + 0] : 0u, rprev, vprev, s, last_lane);
            smx_scan_step<LINEAR, 1, false>(alo, ahi, av, mism, qpk[w + 1], 0u, rpk[33 + w + 1], LINEAR ? rvalid[33 + w + 1] : 0u, rprev, vprev, s, last_lane);
            smx_scan_step<LINEAR, 2, false>(alo, ahi, av, mism, qpk[w + 2], 0u, rpk[33 + w + 2], LINEAR ? rvalid[33 + w + 2] : 0u, rprev, vprev, s, last_lane);
            smx_scan_step<LINEAR, 3, false>(alo, ahi, av, mism, qpk[w + 3], 0u, rpk[33 + w + 3], LINEAR ? rvalid[33 + w + 3] : 0u, rprev, vprev, s, last_lane);
        }
        // last group: the final (possibly partial) word is masked, words beyond the query count nothing
        {
            const int last = qwords - 1;
            const uint32_t m0 = w + 0 < last ? 0xffffffffu : (w + 0 == last ? qmask_last : 0u);
            const uint32_t m1 = w + 1 < last ? 0xffffffffu : (w + 1 == last ? qmask_last : 0u);
            const uint32_t m2 = w + 2 < last ? 0xffffffffu : (w + 2 == last ? qmask_last : 0u);
            const uint32_t m3 = w + 3 < last ? 0xffffffffu : (w + 3 == last ? qmask_last : 0u);
            smx_scan_step<LINEAR, 0, true>(alo, ahi, av, mism, qpk[w + 0], m0, rpk[33 + w + 0], LINEAR ? rvalid[33 + w + 0] : 0u, rprev, vprev, s, last_lane);
            smx_scan_step<LINEAR, 1, true>(alo, ahi, av, mism, qpk[w + 1], m1, rpk[33 + w + 1], LINEAR ? rvalid[33 + w + 1] : 0u, rprev, vprev, s, last_lane);
            smx_scan_step<LINEAR, 2, true>(alo, ahi, av, mism, qpk[w + 2], m2, rpk[33 + w + 2], LINEAR ? rvalid[33 + w + 2] : 0u, rprev, vprev, s, last_lane);
            smx_scan_step<LINEAR, 3, true>(alo, ahi, av, mism, qpk[w + 3], m3, rpk[33 + w + 3], LINEAR ? rvalid[33 + w + 3] : 0u, rprev, vprev, s, last_lane);
        }
#pragma unroll
        for (int d = 0; d < SMX_SCAN_DPT; ++d) {
            const int k = kb + 32 * d;
            if (k < pr.max_k) a.counts[pr.counts_off + k] = nq - mism[d];
        }
    } else {
        // exact byte path (rare): raw character equality
        const uint8_t *ref = a.pool + pr.ref_off;
        const uint8_t *qry = a.pool + pr.qry_off;
#pragma unroll
        for (int d = 0; d < SMX_SCAN_DPT; ++d) {
            const int k = kb + 32 * d;
            if (k >= pr.max_k) continue;
            int c = 0;
            if (!LINEAR) {
                for (int i = 0; i < nq; ++i) c += (ref[(k + i) % nr] == qry[i]);
            } else {
                const int dd = k - shift0;
                const int i_lo = max(0, -dd), i_hi = min(nq, nr - dd);
                for (int i = i_lo; i < i_hi; ++i) c += (ref[dd + i] == qry[i]);
            }
            a.counts[pr.counts_off + k] = c;
        }
    }
}
