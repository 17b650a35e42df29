// Row A1: the whole alignment stage behind one C-ABI call.
// Replaces execute_alignment_core / execute_alignment_core_loop
// (savemoney/post_analysis/post_analysis_core.py:56-87) and RecommendedGrouping.set_distance_matrix /
// calc_distance (savemoney/pre_survey/pre_survey_core.py:237-271):
//   for every (query, reference) pair and both strands:
//     A3 scan (GPU) -> A4 threshold + conserved runs (GPU) -> A5/A6 chain (host threads)
//     -> A7 plan -> A8 batched DP + traceback (GPU) -> A9-A11 clip / merge / rotate / score (host threads)
// Included by smx_api.cu (single translation unit); uses the handle and the phase functions there.
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <thread>

#include "smx_chain.hpp"
#include "smx_select.cuh"
#include "smx_stitch.hpp"

namespace smx {

enum PieceKind : uint8_t { PIECE_LITERAL = 0, PIECE_NW = 1, PIECE_SPECIAL = 2, PIECE_SG_HEAD = 3, PIECE_SG_TAIL = 4 };

struct Piece {
    uint8_t kind;
    char letter;       // PIECE_LITERAL
    int32_t len;       // PIECE_LITERAL: run length
    int32_t prob[2];   // DP problem ids (-1 = absent)
    int32_t n_ref, nq1, nq2;  // PIECE_SPECIAL: lengths of the reference slice and the two query parts
};

struct PairStrand {
    int32_t ref = 0, qry = 0, strand = 0;
    int32_t nr = 0, nq = 0;
    int64_t ref_pool = 0, qry_pool = 0;
    bool skip = false;       // empty result (no conserved region / omit_too_long)
    bool failed = false;     // the reference would raise here
    int32_t ref_off = 0, q_off = 0;
    std::vector<Region> regions;
    std::vector<Piece> pieces;
    // result
    std::vector<uint32_t> runs;
    int32_t score = 0, new_q_off = INT32_MIN;
};

struct PlannedProblem { int64_t s1_off, s2_off; int32_t m, n; uint8_t mode; };

inline double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename F>
inline void parallel_for(int n, int n_threads, F &&fn)
{
    if (n <= 0) return;
    n_threads = std::max(1, std::min(n_threads, n));
    if (n_threads == 1) { for (int i = 0; i < n; ++i) fn(i, 0); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&, t]() { for (;;) { const int i = next.fetch_add(1); if (i >= n) break; fn(i, t); } });
    for (auto &th : pool) th.join();
}

// numpy's pairwise float64 summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE)
inline double np_pairwise_sum(const double *a, int64_t n)
{
    if (n < 8) { double r = 0.; for (int64_t i = 0; i < n; ++i) r += a[i]; return r; }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// np.median(x) + sigma * np.std(x) over counts[lo:hi) exactly as numpy computes it for an int64 list
inline double np_threshold(const int32_t *c, int lo, int hi, double sigma)
{
    const int64_t n = hi - lo;
    std::vector<int32_t> v(c + lo, c + hi);
    std::nth_element(v.begin(), v.begin() + (n - 1) / 2, v.end());
    const double m_lo = v[(size_t)(n - 1) / 2];
    double med = m_lo;
    if (!(n & 1)) { std::nth_element(v.begin(), v.begin() + n / 2, v.end()); med = (m_lo + (double)v[(size_t)n / 2]) * 0.5; }
    std::vector<double> x((size_t)n);
    for (int64_t i = 0; i < n; ++i) x[(size_t)i] = (double)c[lo + i];
    const double mean = np_pairwise_sum(x.data(), n) / (double)n;
    for (int64_t i = 0; i < n; ++i) { const double d = (double)c[lo + i] - mean; x[(size_t)i] = d * d; }
    const double var = np_pairwise_sum(x.data(), n) / (double)n;
    return med + sigma * std::sqrt(var);
}

// host re-decision for a flagged pair-strand: selection + runs from its counts (exact semantics)
inline void host_regions(const uint8_t *ref, const uint8_t *qry, int nr, int nq, int topology, const int32_t *counts, int max_k,
                         int lo, int hi, double sigma, std::vector<SmxRegion> &out)
{
    out.clear();
    const double thr = np_threshold(counts, lo, hi, sigma);
    for (int k = 0; k < max_k; ++k) {
        if (!((double)counts[k] > thr)) continue;
        int start = -1;
        for (int i = 0; i <= nq; ++i) {
            bool eq = false;
            if (i < nq) {
                const int p = topology == 1 ? k - (nq - 1) + i : (k + i) % nr;
                if (p >= 0 && p < nr) eq = ref[p] == qry[i];
            }
            if (eq) { if (start < 0) start = i; }
            else if (start >= 0) { if (i - start > SMX_RUN_MIN) out.push_back(SmxRegion{k, start, i - 1}); start = -1; }
        }
    }
}

}  // namespace smx

struct SmxPipelineState {
    std::vector<smx::PairStrand> ps;
    std::vector<int64_t> out_off;
    int64_t total_runs = 0;
    double t_pool = 0, t_scan = 0, t_select = 0, t_chain = 0, t_dp = 0, t_stitch = 0, t_total = 0;
    double gpu_scan_ms = 0, gpu_select_ms = 0, gpu_dp_ms = 0, gpu_chain_ms = 0;
    int64_t scan_cells = 0, dp_cells = 0, n_dp = 0, n_host_redo = 0, n_failed = 0;
    int launches = 0;
    double dp_kernel_ms[16] = {0};
    int64_t dp_class_cells[14] = {0};
    int64_t dp_ck_cells = 0, dp_ck_problems = 0;  // cells / problems of the packed classes that ran checkpointed
    int dp_kernel_launches[16] = {0};
    bool valid = false;
};
