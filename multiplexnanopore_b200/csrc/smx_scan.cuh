// Diagonal match-count scan: counts[k] = #{ i : ref_ext[k + i] == query[i] }.
// Replaces af.k_mer_offset_analysis_2 (savemoney/modules/cython_functions/alignment_functions.pyx:46-64)
// together with the window construction around it (ref_query_alignment.py:351-366, 385):
//   circular: ref_ext = reference tiled, k in [0, Nr)
//   linear  : ref_ext = reference zero-padded by Nq-1 on both sides, k in [0, Nr+Nq-1)
//
// One CTA = 256 consecutive diagonals of one (reference, query) pair.  The query and the
// reference window the tile touches are staged in shared memory 2-bit packed (16 bases per
// 32-bit word); each thread walks its diagonal 16 bases at a time: funnel shift of two
// reference words, XOR, fold, POPC.  Tiles that touch a byte other than A/C/G/T fall back to an
// exact byte-compare loop in the same kernel (equality in the reference is on raw characters).
#pragma once
#include "smx_common.cuh"

#define SMX_SCAN_TILE 256

struct SmxScanPair {
    int64_t ref_off, qry_off, counts_off;
    int32_t nr, nq;
    int32_t max_k, tile0;  // tile0 = index of this pair's first tile in the launch
};

struct SmxScanArgs {
    const uint8_t *pool;
    const SmxScanPair *pairs;
    const int32_t *tile_pair;  // tile -> pair index
    int32_t *counts;
    int32_t topology;  // 0 circular, 1 linear
    int32_t smem_words_q;  // capacity (words) reserved for the packed query
};

// smem layout: [packed query words | packed reference window words] then (byte path) raw bytes reuse the same area
__global__ void __launch_bounds__(SMX_SCAN_TILE) smx_scan_kernel(const SmxScanArgs a)
{
    extern __shared__ uint32_t smem[];
    const int tile = blockIdx.x;
    const int pi = a.tile_pair[tile];
    const SmxScanPair pr = a.pairs[pi];
    const int nr = pr.nr, nq = pr.nq;
    const uint8_t *ref = a.pool + pr.ref_off;
    const uint8_t *qry = a.pool + pr.qry_off;
    const int k0 = (tile - pr.tile0) * SMX_SCAN_TILE;
    const int tid = threadIdx.x;
    const bool linear = a.topology == 1;
    // reference coordinate of ref_ext[x]: circular x mod nr; linear x - (nq-1) (out of range = padding)
    const int shift0 = linear ? (nq - 1) : 0;
    // window of ref_ext positions touched by the tile: [k0, k0 + 255 + nq - 1]
    const int wlen = SMX_SCAN_TILE + nq - 1;
    const int qwords = (nq + 15) >> 4;
    const int rwords = ((wlen + 15) >> 4) + 1;
    uint32_t *qpk = smem;
    uint32_t *rpk = smem + qwords;

    // ---- stage + pack (each thread packs whole words) ----
    int special = 0;
    for (int w = tid; w < qwords; w += SMX_SCAN_TILE) {
        uint32_t v = 0;
        for (int b = 0; b < 16; ++b) {
            const int i = w * 16 + b;
            if (i < nq) {
                const int c = smx_code(qry[i]);
                special |= (c > 3);
                v |= (uint32_t)(c & 3) << (2 * b);
            }
        }
        qpk[w] = v;
    }
    for (int w = tid; w < rwords; w += SMX_SCAN_TILE) {
        uint32_t v = 0;
        for (int b = 0; b < 16; ++b) {
            const int x = k0 + w * 16 + b;
            int pos = x - shift0;
            if (!linear) pos = x % nr;
            if (pos >= 0 && pos < nr && x < k0 + wlen) {
                const int c = smx_code(ref[pos]);
                special |= (c > 3);
                v |= (uint32_t)(c & 3) << (2 * b);
            }
        }
        rpk[w] = v;
    }
    special = __syncthreads_or(special);

    const int k = k0 + tid;
    if (k >= pr.max_k) return;
    // valid query range for this diagonal
    int i_lo = 0, i_hi = nq;
    if (linear) {
        const int d = k - (nq - 1);
        i_lo = max(0, -d);
        i_hi = min(nq, nr - d);
    }
    int count = 0;
    if (!special) {
        if (i_hi > i_lo) {
            const int sh = 2 * (tid & 15);
            const int xbase = tid >> 4;
            const int w_lo = i_lo >> 4, w_hi = (i_hi - 1) >> 4;
            uint32_t r0 = rpk[xbase + w_lo];
            for (int w = w_lo; w <= w_hi; ++w) {
                const uint32_t r1 = rpk[xbase + w + 1];
                const uint32_t s = __funnelshift_r(r0, r1, sh);
                r0 = r1;
                const uint32_t x = qpk[w] ^ s;
                uint32_t mis = (x | (x >> 1)) & 0x55555555u;
                // mask bases outside [i_lo, i_hi)
                const int base = w << 4;
                if (base < i_lo) mis |= (1u << (2 * (i_lo - base))) - 1u;
                if (base + 16 > i_hi) mis |= ~((1u << (2 * (i_hi - base))) - 1u);
                count += 16 - __popc(mis & 0x55555555u);
            }
        }
    } else {
        // exact byte path (rare): raw character equality
        for (int i = i_lo; i < i_hi; ++i) {
            const int x = k + i;
            const int pos = linear ? x - shift0 : x % nr;
            count += (ref[pos] == qry[i]);
        }
    }
    a.counts[pr.counts_off + k] = count;
}
