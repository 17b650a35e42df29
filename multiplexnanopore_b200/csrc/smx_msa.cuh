// Row N3, first pass: MyMSA.generate_msa (savemoney/modules/msa.py:882-982) -- the reads assigned to a plasmid, each
// with its column CIGAR against the same reference, laid out as one multiple alignment.
//
// The reference walks all CIGARs in lock step, one output column per iteration: a column of the reference when no read
// is at an 'I' or 'S'; otherwise an extra column in which the reads standing at 'S' (first choice) or at 'I' consume a
// query base and every other read gets a filler ('O' + ' ' next to a hard-clipped stretch, else 'N' + '-').  Restated
// column-parallel:
//   * the extra columns between reference positions r-1 and r (the "gap" r, r = 0..R, R = the sentinel behind the last
//     base) depend only on the reads' letter blocks in that gap.  A block of the form S*I* consumes the first |S| of the
//     gap's S columns and the first |I| of its I columns, and the gap has max|S| + max|I| columns, S columns first.
//     A block in which an 'I' precedes an 'S' changes the order of the columns; such gaps (rare: no CIGAR of the
//     alignment stage holds one) get their exact column-type sequence from a serial simulation on the host;
//   * given the type sequence of its gap every read decides alone: it consumes a column iff its next block letter is
//     the column's type.
// Kernels: smx_msa_expand (one warp per read: letters, position of every reference-consuming letter, query index at
// every gap), smx_msa_gaps (one warp per gap: max|S|, max|I| over the reads), smx_msa_fill (one thread per read x gap).
#pragma once
#include "smx_common.cuh"

struct SmxMsaArgs {
    const uint8_t *ref;         // [R]
    int32_t R, n_reads;
    const uint8_t *qry;         // query bases, pooled
    const int8_t *qscore;       // same offsets
    const int64_t *qry_off;
    const int32_t *qry_len;
    const uint32_t *cigar_rle;  // len << 4 | op
    const int64_t *cigar_off;   // [n_reads + 1]
    uint8_t *let;               // expanded letters, pooled
    const int64_t *let_off;     // [n_reads + 1]
    int32_t *pos;               // [n_reads][R + 1] letter index of the r-th reference-consuming letter (R: the sentinel)
    int32_t *qcnt;              // [n_reads][R + 1] query bases consumed before the block of gap r
    int32_t *n_s, *n_i;         // [R + 1] S / I columns of gap r (pure gaps)
    int32_t *err;               // [n_reads] 1 = CIGAR does not consume exactly the reference and the query
    // fill
    const int64_t *col_off;     // [R + 2] first column of gap r; col_off[R + 1] = total columns
    const int32_t *type_off;    // [R + 2] complex gaps: range of their column types in `types` (empty range: pure gap)
    const uint8_t *types;       // 'S' / 'I' per extra column of the complex gaps
    int64_t n_cols;
    uint8_t *out_ref, *out_seq, *out_cig;  // [n_cols], [n_reads][n_cols] x 2
    int8_t *out_qs;                        // [n_reads][n_cols]
};

__device__ __forceinline__ uint8_t smx_msa_letter(uint32_t op)
{
    return op == 7u ? '=' : op == 8u ? 'X' : op == 1u ? 'I' : op == 2u ? 'D' : op == 4u ? 'S' : op == 5u ? 'H' : '?';
}

__global__ void __launch_bounds__(128) smx_msa_expand(const SmxMsaArgs a)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= a.n_reads) return;
    const uint32_t *runs = a.cigar_rle + a.cigar_off[w];
    const int n_runs = (int)(a.cigar_off[w + 1] - a.cigar_off[w]);
    uint8_t *let = a.let + a.let_off[w];
    int32_t *pos = a.pos + (size_t)w * (a.R + 1), *qc = a.qcnt + (size_t)w * (a.R + 1);
    int li = 0, refc = 0, q = 0;       // letters written, reference / query bases consumed so far
    int pend_q = 0;                    // query bases consumed right behind the previous reference-consuming letter
    bool bad = false;
    for (int r = 0; r < n_runs; ++r) {
        const uint32_t op = runs[r] & 15u;
        const int len = (int)(runs[r] >> 4);
        const uint8_t L = smx_msa_letter(op);
        const bool is_ref = op == 7u || op == 8u || op == 2u || op == 5u, is_q = op == 7u || op == 8u || op == 1u || op == 4u;
        if (L == '?' || (is_ref && refc + len > a.R)) { bad = true; break; }
        for (int x = lane; x < len; x += 32) {
            let[li + x] = L;
            if (is_ref) {
                pos[refc + x] = li + x;
                qc[refc + x] = x == 0 ? pend_q : q + (is_q ? x : 0);   // x > 0: the block of that gap is empty
            }
        }
        li += len;
        if (is_q) q += len;
        if (is_ref) { refc += len; pend_q = q; }
    }
    if (lane == 0) {
        pos[a.R] = li; qc[a.R] = pend_q;
        a.err[w] = (bad || refc != a.R || q != a.qry_len[w] || (int64_t)li != a.let_off[w + 1] - a.let_off[w]) ? 1 : 0;
    }
}

// block of read i in gap r: letters [bs, be)
__device__ __forceinline__ void smx_msa_block(const SmxMsaArgs &a, int i, int r, int &bs, int &be)
{
    const int32_t *pos = a.pos + (size_t)i * (a.R + 1);
    bs = r == 0 ? 0 : pos[r - 1] + 1;
    be = pos[r];
}

__global__ void __launch_bounds__(128) smx_msa_gaps(const SmxMsaArgs a)
{
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r > a.R) return;
    int ms = 0, mi = 0;
    for (int i = lane; i < a.n_reads; i += 32) {
        int bs, be;
        smx_msa_block(a, i, r, bs, be);
        if (be <= bs) continue;
        const uint8_t *let = a.let + a.let_off[i];
        int s = 0, n_i = 0;
        for (int x = bs; x < be; ++x) { if (let[x] == 'S') ++s; else ++n_i; }   // S*I* blocks: counts are the lengths
        ms = max(ms, s); mi = max(mi, n_i);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { ms = max(ms, __shfl_xor_sync(SMX_FULL_MASK, ms, off)); mi = max(mi, __shfl_xor_sync(SMX_FULL_MASK, mi, off)); }
    if (lane == 0) { a.n_s[r] = ms; a.n_i[r] = mi; }
}

__device__ __forceinline__ void smx_msa_fill_one(const SmxMsaArgs &a, const int i, const int r);

__global__ void __launch_bounds__(256) smx_msa_fill(const SmxMsaArgs a)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;        // gap
    if (r > a.R) return;
    for (int i = blockIdx.y; i <= a.n_reads; i += gridDim.y) smx_msa_fill_one(a, i, r);   // row n_reads = the reference itself
}

__device__ __forceinline__ void smx_msa_fill_one(const SmxMsaArgs &a, const int i, const int r)
{
    const int64_t c0 = a.col_off[r];
    const int n_extra = (int)(a.col_off[r + 1] - c0) - (r < a.R ? 1 : 0);
    if (i == a.n_reads) {
        for (int c = 0; c < n_extra; ++c) a.out_ref[c0 + c] = '-';
        if (r < a.R) a.out_ref[c0 + n_extra] = a.ref[r];
        return;
    }
    const uint8_t *let = a.let + a.let_off[i];
    const int L = (int)(a.let_off[i + 1] - a.let_off[i]);
    int bs, be;
    smx_msa_block(a, i, r, bs, be);
    const uint8_t *qry = a.qry + a.qry_off[i];
    const int8_t *qsc = a.qscore + a.qry_off[i];
    int q = a.qcnt[(size_t)i * (a.R + 1) + r];
    uint8_t *o_seq = a.out_seq + (size_t)i * a.n_cols + c0, *o_cig = a.out_cig + (size_t)i * a.n_cols + c0;
    int8_t *o_qs = a.out_qs + (size_t)i * a.n_cols + c0;
    if (n_extra > 0) {
        const int t0 = a.type_off[r], t1 = a.type_off[r + 1];
        const int n_s = a.n_s[r];
        int p = bs;   // the read's pointer: its next unconsumed letter (the reference letter / sentinel once the block is used up)
        for (int c = 0; c < n_extra; ++c) {
            const uint8_t T = t1 > t0 ? a.types[t0 + c] : (c < n_s ? 'S' : 'I');
            if (p < be && let[p] == T) {
                o_cig[c] = T; o_seq[c] = qry[q]; o_qs[c] = qsc[q];
                ++p; ++q;
            } else {
                // msa.py:932-945 / :957-969: a skipped column inside or next to a hard-clipped stretch reads 'O' + ' ', else 'N' + '-';
                // letter p of my_cigar + "Z" + my_cigar[-1], and letter p - 1 (index -1 wraps to the appended last letter)
                const uint8_t cur = p < L ? let[p] : 'Z';
                const uint8_t prev = p > 0 ? let[p - 1] : (L > 0 ? let[L - 1] : 'Z');
                const bool hard = cur == 'H' || prev == 'H' || (cur == 'Z' && L > 0 && let[0] == 'H');
                o_cig[c] = hard ? 'O' : 'N'; o_seq[c] = hard ? ' ' : '-'; o_qs[c] = -1;
            }
        }
        // (a block that the column types did not use up cannot happen: the types come from the blocks)
    } else {
        q += be - bs;
    }
    if (r < a.R) {
        q = a.qcnt[(size_t)i * (a.R + 1) + r] + (be - bs);
        const uint8_t Lr = let[be];
        o_cig[n_extra] = Lr;
        if (Lr == '=' || Lr == 'X') { o_seq[n_extra] = qry[q]; o_qs[n_extra] = qsc[q]; }
        else { o_seq[n_extra] = Lr == 'D' ? '-' : ' '; o_qs[n_extra] = -1; }
    }
}
