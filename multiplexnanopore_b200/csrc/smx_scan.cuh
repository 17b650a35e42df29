// Diagonal match-count scan: counts[k] = #{ i : ref_ext[k + i] == query[i] }.
// Replaces af.k_mer_offset_analysis_2 (savemoney/modules/cython_functions/alignment_functions.pyx:46-64)
// together with the window construction around it (ref_query_alignment.py:351-366, 385):
//   circular: ref_ext = reference tiled, k in [0, Nr)
//   linear  : ref_ext = reference zero-padded by Nq-1 on both sides, k in [0, Nr+Nq-1)
//
// The sequence pool is packed once per upload to 2 bit/base (smx_pack_kernel, 16 bases per 32-bit word, plus one
// flag byte per word for bytes other than A/C/G/T).  One CTA = 1024 consecutive diagonals of one (reference, query)
// pair: the packed query and the reference window the tile touches are staged in shared memory with two word loads
// and one funnel shift per word; every thread owns 4 diagonals that are 16 apart (same shift, neighbouring words),
// so one query word and one new reference word per step feed 64 cells: funnel shift, XOR, fold, POPC.
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
    const uint32_t *pool2;   // 2-bit packed pool
    const uint8_t *special;  // per packed word: holds a byte other than A/C/G/T
    const SmxScanPair *pairs;
    const int32_t *tile_pair;  // tile -> pair index
    int32_t *counts;
    int32_t topology;  // 0 circular, 1 linear
};

// one thread per 16 bases of the pool
__global__ void __launch_bounds__(256) smx_pack_kernel(const uint8_t *pool, int64_t n, uint32_t *pool2, uint8_t *special, int64_t n_words)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t v = 0;
    int sp = 0;
    const int64_t base = w * 16;
    if (base + 16 <= n) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(pool + base);
        const uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int c = smx_code((words[b >> 2] >> (8 * (b & 3))) & 0xffu);
            sp |= c > 3;
            v |= (uint32_t)(c & 3) << (2 * b);
        }
    } else {
        for (int b = 0; b < 16 && base + b < n; ++b) {
            const int c = smx_code(pool[base + b]);
            sp |= c > 3;
            v |= (uint32_t)(c & 3) << (2 * b);
        }
    }
    pool2[w] = v;
    special[w] = (uint8_t)sp;
}

// 16 bases starting at pool position p (any alignment)
__device__ __forceinline__ uint32_t smx_fetch16(const uint32_t *pool2, const uint8_t *special, int64_t p, int &sp)
{
    const int64_t w = p >> 4;
    const int sh = 2 * (int)(p & 15);
    sp |= special[w] | special[w + 1];
    return __funnelshift_r(pool2[w], pool2[w + 1], sh);
}

__device__ __forceinline__ int smx_scan_word(uint32_t q, uint32_t s, int base, int i_lo, int i_hi)
{
    const uint32_t x = q ^ s;
    uint32_t mis = (x | (x >> 1)) & 0x55555555u;
    if (base < i_lo) mis |= (1u << (2 * (i_lo - base))) - 1u;
    if (base + 16 > i_hi) mis |= ~((1u << (2 * (i_hi - base))) - 1u);
    return 16 - __popc(mis & 0x55555555u);
}

// smem layout: [packed query words | packed reference window words]
__global__ void __launch_bounds__(SMX_SCAN_THREADS) smx_scan_kernel(const SmxScanArgs a)
{
    extern __shared__ uint32_t smem[];
    const int tile = blockIdx.x;
    const int pi = a.tile_pair[tile];
    const SmxScanPair pr = a.pairs[pi];
    const int nr = pr.nr, nq = pr.nq;
    const int k0 = (tile - pr.tile0) * SMX_SCAN_TILE;
    const int tid = threadIdx.x;
    const bool linear = a.topology == 1;
    const int shift0 = linear ? (nq - 1) : 0;  // ref_ext[x] = reference[x - shift0] (linear) or reference[x mod nr]
    const int wlen = SMX_SCAN_TILE + nq - 1;   // window of ref_ext positions touched by the tile: [k0, k0 + wlen)
    const int qwords = (nq + 15) >> 4;
    const int rwords = SMX_SCAN_TILE / 16 + qwords + 2;
    (void)wlen;
    uint32_t *qpk = smem;
    uint32_t *rpk = smem + qwords;

    // ---- stage the packed query and reference window ----
    int special = 0;
    for (int w = tid; w < qwords; w += SMX_SCAN_THREADS) qpk[w] = smx_fetch16(a.pool2, a.special, pr.qry_off + 16 * w, special);
    for (int w = tid; w < rwords; w += SMX_SCAN_THREADS) {
        const int x = k0 + w * 16;
        uint32_t v = 0;
        if (!linear) {
            if (nr >= 16) {
                // 16 bases from position x mod nr; a word that crosses the end continues at position 0
                const int p0 = x % nr;
                v = smx_fetch16(a.pool2, a.special, pr.ref_off + p0, special);
                const int left = nr - p0;  // bases before the wrap
                if (left < 16) v = (v & ((1u << (2 * left)) - 1u)) | (smx_fetch16(a.pool2, a.special, pr.ref_off, special) << (2 * left));
            } else {
                for (int b = 0; b < 16; ++b) { const int c = smx_code(a.pool[pr.ref_off + (x + b) % nr]); special |= c > 3; v |= (uint32_t)(c & 3) << (2 * b); }
            }
        } else {
            const int pos = x - shift0;  // positions outside [0, nr) are padding: masked by the per-diagonal range below
            if (pos >= 0 && pos < nr) v = smx_fetch16(a.pool2, a.special, pr.ref_off + pos, special);
            else if (pos < 0 && pos > -16) v = smx_fetch16(a.pool2, a.special, pr.ref_off, special) << (2 * (-pos));
        }
        rpk[w] = v;
    }
    special = __syncthreads_or(special);

    const int sh = 2 * (tid & 15);
    const int xb = SMX_SCAN_DPT * (tid >> 4);
    const int kb = k0 + (tid & 15) + 16 * xb;  // diagonals kb + 16 d, d = 0..3
    int cnt[SMX_SCAN_DPT], i_lo[SMX_SCAN_DPT], i_hi[SMX_SCAN_DPT];
    int w_a = 0, w_b = nq >> 4;  // interior words [w_a, w_b) are complete for all four diagonals
#pragma unroll
    for (int d = 0; d < SMX_SCAN_DPT; ++d) {
        cnt[d] = 0;
        i_lo[d] = 0; i_hi[d] = nq;
        if (linear) {
            const int dd = kb + 16 * d - (nq - 1);
            i_lo[d] = max(0, -dd);
            i_hi[d] = max(i_lo[d], min(nq, nr - dd));
        }
        w_a = max(w_a, (i_lo[d] + 15) >> 4);
        w_b = min(w_b, i_hi[d] >> 4);
    }
    if (!special) {
        if (w_b > w_a) {
            uint32_t r[SMX_SCAN_DPT + 1];
#pragma unroll
            for (int d = 0; d <= SMX_SCAN_DPT; ++d) r[d] = rpk[xb + w_a + d];
            for (int w = w_a; w < w_b; ++w) {
                const uint32_t q = qpk[w];
#pragma unroll
                for (int d = 0; d < SMX_SCAN_DPT; ++d) {
                    const uint32_t x = q ^ __funnelshift_r(r[d], r[d + 1], sh);
                    cnt[d] += 16 - __popc((x | (x >> 1)) & 0x55555555u);
                }
#pragma unroll
                for (int d = 0; d < SMX_SCAN_DPT; ++d) r[d] = r[d + 1];
                r[SMX_SCAN_DPT] = rpk[xb + w + SMX_SCAN_DPT + 1];
            }
        } else {
            w_b = w_a;
        }
        // edge words (partial, or outside the common interior), per diagonal
#pragma unroll
        for (int d = 0; d < SMX_SCAN_DPT; ++d) {
            if (i_hi[d] <= i_lo[d]) continue;
            const int first = i_lo[d] >> 4, last = (i_hi[d] - 1) >> 4;
            for (int w = first; w <= last; ++w) {
                if (w >= w_a && w < w_b) { w = w_b - 1; continue; }
                const uint32_t s = __funnelshift_r(rpk[xb + d + w], rpk[xb + d + w + 1], sh);
                cnt[d] += smx_scan_word(qpk[w], s, w << 4, i_lo[d], i_hi[d]);
            }
        }
    } else {
        // exact byte path (rare): raw character equality
        const uint8_t *ref = a.pool + pr.ref_off;
        const uint8_t *qry = a.pool + pr.qry_off;
#pragma unroll
        for (int d = 0; d < SMX_SCAN_DPT; ++d) {
            const int k = kb + 16 * d;
            if (k >= pr.max_k) continue;
            int c = 0;
            for (int i = i_lo[d]; i < i_hi[d]; ++i) {
                const int x = k + i;
                const int pos = linear ? x - shift0 : x % nr;
                c += (ref[pos] == qry[i]);
            }
            cnt[d] = c;
        }
    }
#pragma unroll
    for (int d = 0; d < SMX_SCAN_DPT; ++d) {
        const int k = kb + 16 * d;
        if (k < pr.max_k) a.counts[pr.counts_off + k] = cnt[d];
    }
}
