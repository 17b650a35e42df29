// Row N3, second half: the per-column Bayesian consensus of MyMSA.calculate_consensus_core (savemoney/modules/msa.py:
// 1749-1812) -- SequenceBasecallQscorePDF.calc_consensus_error_rate (cython_functions/alignment_functions.pyx:93-132).
//
// For every column, every candidate true base B and every base b the reference multiplies, read by read in row order,
//     val *= (P(b, called) * pdf(b, called, q)) / (P(B, called) * pdf(B, called, q))          (long double)
// starting from val = P_N[b] / P_N[B]; then p(B) = 1 - 1 / sum_b val.  The factor depends only on (B, b, called, q): the
// host tabulates it in native long double (5 x 5 x 5 x 94 values), and the device runs the 25 product chains of a column
// -- one thread per (column, B, b), the reads of the column in order, x87 multiplication restated on integers
// (smx_ld80.hpp) -- for the prior-weighted and the uniform start value at once.  Sums and 1 - 1 / sum are seven native
// operations per column and B on the host.  The matrices are those of smx_msa_fill: [read][column] bytes.
#pragma once
#include "smx_common.cuh"
#include "smx_consensus_host.hpp"

struct SmxConsArgs {
    const uint8_t *seq;      // [rows][cols] called characters
    const int8_t *qs;        // [rows][cols] q-scores (-1: none)
    const uint8_t *cig;      // [rows][cols] column letters; H and O rows do not vote (msa.py:1762)
    int32_t rows;
    int64_t cols;
    const uint64_t *tab_m;   // [25 pairs][5 called][94] significands of the factor table
    const int32_t *tab_e;    //                         exponents
    const int32_t *col_class;  // [cols] index of the column's reference letter in the start-value tables
    const uint64_t *init_m;  // [2][classes][25]: start values with prior, then without
    const int32_t *init_e;
    int32_t n_classes;
    uint64_t *out_m;         // [2][cols][25]
    int32_t *out_e;
    int32_t *n_events;       // [cols] voting reads
    int32_t *err;            // set when a voting read holds a character outside ATCG- or a q-score above 92 (the reference indexes out of bounds there)
};

// grid (ceil(cols / 128), 25): blockIdx.y = B * 5 + b; consecutive threads = consecutive columns (coalesced row reads)
__global__ void __launch_bounds__(128) smx_consensus_kernel(const SmxConsArgs a)
{
    __shared__ uint64_t s_m[SMX_CONS_BASES * SMX_CONS_QIDX];
    __shared__ int32_t s_e[SMX_CONS_BASES * SMX_CONS_QIDX];
    const int pair = blockIdx.y;
    for (int x = threadIdx.x; x < SMX_CONS_BASES * SMX_CONS_QIDX; x += blockDim.x) {
        s_m[x] = a.tab_m[(size_t)pair * SMX_CONS_BASES * SMX_CONS_QIDX + x];
        s_e[x] = a.tab_e[(size_t)pair * SMX_CONS_BASES * SMX_CONS_QIDX + x];
    }
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.cols) return;
    const int cls = a.col_class[c];
    SmxLd v_with, v_without;
    v_with.m = a.init_m[(size_t)cls * 25 + pair];
    v_with.e = a.init_e[(size_t)cls * 25 + pair];
    v_without.m = a.init_m[((size_t)a.n_classes + cls) * 25 + pair];
    v_without.e = a.init_e[((size_t)a.n_classes + cls) * 25 + pair];
    int events = 0;
    const bool bad = !smx_cons_column(a.seq, a.qs, a.cig, a.rows, a.cols, c, s_m, s_e, v_with, v_without, events);
    const size_t o = (size_t)c * 25 + pair;
    a.out_m[o] = v_with.m;
    a.out_e[o] = v_with.e;
    a.out_m[(size_t)a.cols * 25 + o] = v_without.m;
    a.out_e[(size_t)a.cols * 25 + o] = v_without.e;
    if (pair == 0) a.n_events[c] = events;
    if (bad) atomicExch(a.err, 1);
}
