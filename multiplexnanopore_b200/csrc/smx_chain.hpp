// Row A6 on the host (C++): chain search over conserved regions.
// Restates SearchLongestPath.exec_circular / exec_linear / search_longest_path / get_1d_connection /
// get_connection_constraint and ConnectionMatrix (savemoney/modules/ref_query_alignment.py:649-709,
// 792-948; literal semantics in SURVEY.md Appendix C) with bitset adjacency instead of numpy
// matrices.  Results must be identical to the reference including its order-dependent choices:
// strict '>' relaxations in frontier order, first arg-max, and the linear-mode quirk (only the last
// listed region is used as origin).
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace smx {

struct Region { int32_t rs, re, qs, qe; };

class ChainSearch {
public:
    // returns false when the reference would raise (not exactly one (0,0) region)
    bool search(const std::vector<Region> &R, std::vector<int> &trace, int64_t &score)
    {
        const int n = (int)R.size();
        W_ = (n + 63) >> 6;
        int init = -1, n_init = 0;
        for (int i = 0; i < n; ++i) if (R[i].rs == 0 && R[i].qs == 0) { init = i; ++n_init; }
        if (n_init != 1) return false;
        adj_.assign((size_t)n * W_, 0);
        tmp_.assign((size_t)n * W_, 0);
        one_d(R, 0, adj_);
        one_d(R, 1, tmp_);
        for (size_t x = 0; x < adj_.size(); ++x) adj_[x] |= tmp_[x];
        // constraint + ignored nodes
        ignored_.assign(n, 0);
        for (int i = 0; i < n; ++i) ignored_[i] = (R[i].re - R[i].rs < 0) || (R[i].qe - R[i].qs < 0);
        for (int i = 0; i < n; ++i) {
            uint64_t *row = &adj_[(size_t)i * W_];
            if (ignored_[i]) { std::fill(row, row + W_, 0); continue; }
            for (int j = 0; j < n; ++j) {
                if (!((row[j >> 6] >> (j & 63)) & 1)) continue;
                // weight 0 edges do not exist in the reference ((W > 0) tests): w[j] = re - rs
                if (ignored_[j] || !(R[i].re < R[j].rs && R[i].qe < R[j].qs) || (R[j].re - R[j].rs) <= 0)
                    row[j >> 6] &= ~(1ull << (j & 63));
            }
        }
        // remove_unreachable_branches: nodes without input lose their outputs, to a fixed point
        indeg_.assign(n, 0);
        int prev_dead = -1;
        for (;;) {
            std::fill(indeg_.begin(), indeg_.end(), 0);
            for (int i = 0; i < n; ++i) {
                const uint64_t *row = &adj_[(size_t)i * W_];
                for (int w = 0; w < W_; ++w) {
                    uint64_t b = row[w];
                    while (b) { const int j = (w << 6) + __builtin_ctzll(b); b &= b - 1; ++indeg_[j]; }
                }
            }
            int dead = 0;
            for (int j = 0; j < n; ++j) if (indeg_[j] == 0 && j != init) ++dead;
            if (dead == prev_dead) break;
            for (int j = 0; j < n; ++j)
                if (indeg_[j] == 0 && j != init) { uint64_t *row = &adj_[(size_t)j * W_]; std::fill(row, row + W_, 0); }
            prev_dead = dead;
        }
        // label-correcting longest path in the reference's frontier order
        score_.assign(n, 0);
        parent_.assign(n, -1);
        score_[init] = R[init].re - R[init].rs;
        frontier_.clear();
        frontier_.push_back(init);
        // indeg_ holds the remaining unprocessed inputs per node (init excluded from the dead test only)
        while (!frontier_.empty()) {
            next_.clear();
            for (int i : frontier_) {
                const uint64_t *row = &adj_[(size_t)i * W_];
                for (int w = 0; w < W_; ++w) {
                    uint64_t b = row[w];
                    while (b) {
                        const int j = (w << 6) + __builtin_ctzll(b);
                        b &= b - 1;
                        const int64_t t = score_[i] + (R[j].re - R[j].rs);
                        if (t > score_[j]) { score_[j] = t; parent_[j] = i; }
                        if (--indeg_[j] == 0) next_.push_back(j);
                    }
                }
            }
            frontier_.swap(next_);
        }
        int best = 0;
        for (int j = 1; j < n; ++j) if (score_[j] > score_[best]) best = j;
        score = score_[best];
        trace.clear();
        if (best != init && parent_[best] < 0) {
            // an unreached node can only win with score 0 against a zero-length origin: the reference
            // would return trace None here and crash later
            return false;
        }
        for (int x = best; x >= 0; x = parent_[x]) trace.push_back(x);
        std::reverse(trace.begin(), trace.end());
        return true;
    }

    // exec_circular (:811-862): every region as origin, three-stage choice among the best
    bool circular(const std::vector<Region> &R, int n_ref, int n_q, std::vector<int> &best_trace)
    {
        const int n = (int)R.size();
        std::vector<std::vector<int>> traces((size_t)n);
        std::vector<int64_t> scores((size_t)n, 0);
        std::vector<Region> S;
        std::vector<int> keep, tr;
        for (int r = 0; r < n; ++r) {
            S.clear(); keep.clear();
            for (int i = 0; i < n; ++i) {
                Region x{R[i].rs - R[r].rs, R[i].re - R[r].rs, R[i].qs - R[r].qs, R[i].qe - R[r].qs};
                if (x.rs < 0) x.rs += n_ref;
                if (x.re < 0) x.re += n_ref;
                if (x.qs < 0) x.qs += n_q;
                if (x.qe < 0) x.qe += n_q;
                if (x.re - x.rs < 0) continue;  // spans the rotated reference origin: removed
                S.push_back(x); keep.push_back(i);
            }
            int64_t sc = 0;
            if (!search(S, tr, sc)) return false;
            traces[r].resize(tr.size());
            for (size_t x = 0; x < tr.size(); ++x) traces[r][x] = keep[tr[x]];
            scores[r] = sc;
        }
        const int64_t best = *std::max_element(scores.begin(), scores.end());
        std::vector<int> cand;
        for (int r = 0; r < n; ++r) if (scores[r] == best) cand.push_back(r);
        if (cand.size() == 1) { best_trace = traces[cand[0]]; return true; }
        std::vector<int64_t> g1(cand.size()), g2(cand.size());
        for (size_t c = 0; c < cand.size(); ++c) {
            const std::vector<int> &t = traces[cand[c]];
            int a = 0, b = 0;
            for (size_t x = 1; x < t.size(); ++x) {
                if (R[t[x]].qs < R[t[a]].qs) a = (int)x;   // argmin: first minimum
                if (R[t[x]].qe > R[t[b]].qe) b = (int)x;   // argmax: first maximum
            }
            int64_t d = ((int64_t)R[t[a]].rs - R[t[b]].re) % n_ref;
            if (d < 0) d += n_ref;  // Python's non-negative modulo
            g1[c] = d;
            int64_t tail = (int64_t)R[t.front()].qs - R[t.back()].qe;
            int64_t g = tail;
            for (size_t x = 0; x + 1 < t.size(); ++x) g = std::max<int64_t>(g, (int64_t)R[t[x + 1]].qs - R[t[x]].qe);
            g2[c] = g;
        }
        const int64_t top = *std::max_element(g1.begin(), g1.end());
        int pick = -1;
        for (size_t c = 0; c < cand.size(); ++c)
            if (g1[c] == top && (pick < 0 || g2[c] > g2[pick])) pick = (int)c;  // first arg-max of g2 among g1 == top
        best_trace = traces[cand[pick]];
        return true;
    }

    // exec_linear (:793-810): the search that is returned uses the LAST region as origin, no wrap
    bool linear(const std::vector<Region> &R, std::vector<int> &trace)
    {
        std::vector<Region> S(R.size());
        const Region &o = R.back();
        for (size_t i = 0; i < R.size(); ++i) S[i] = Region{R[i].rs - o.rs, R[i].re - o.rs, R[i].qs - o.qs, R[i].qe - o.qs};
        int64_t sc = 0;
        return search(S, trace, sc);
    }

private:
    // get_1d_connection (:903-942) on (start, end) of the reference (dim 0) or query (dim 1) intervals
    void one_d(const std::vector<Region> &R, int dim, std::vector<uint64_t> &A)
    {
        const int n = (int)R.size();
        ev_.resize((size_t)2 * n);
        for (int i = 0; i < n; ++i) {
            const int s = dim == 0 ? R[i].rs : R[i].qs, e = dim == 0 ? R[i].re : R[i].qe;
            // value * 10 - 1 for starts reproduces "start - 0.1" ordering with integers
            ev_[(size_t)2 * i] = Event{(int64_t)s * 10 - 1, 2 * i};
            ev_[(size_t)2 * i + 1] = Event{(int64_t)e * 10, 2 * i + 1};
        }
        std::sort(ev_.begin(), ev_.end(), [](const Event &x, const Event &y) { return x.key != y.key ? x.key < y.key : x.flat < y.flat; });
        state_.assign(n, 0);
        closed_.assign(W_, 0);   // bit set: state >= 2
        linked_.assign(W_, 0);   // bit set: state == 3
        for (const Event &e : ev_) {
            const int idx = e.flat >> 1;
            if (state_[idx] == 0) {
                for (int w = 0; w < W_; ++w) {
                    uint64_t b = closed_[w];
                    linked_[w] |= b;
                    while (b) { const int i = (w << 6) + __builtin_ctzll(b); b &= b - 1; A[(size_t)i * W_ + (idx >> 6)] |= 1ull << (idx & 63); }
                }
                state_[idx] = 1;
            } else {
                for (int w = 0; w < W_; ++w) { closed_[w] &= ~linked_[w]; linked_[w] = 0; }  // state 3 -> -1
                state_[idx] += 1;
                if (state_[idx] == 2) closed_[idx >> 6] |= 1ull << (idx & 63);
            }
        }
    }

    struct Event { int64_t key; int flat; };
    int W_ = 0;
    std::vector<uint64_t> adj_, tmp_, closed_, linked_;
    std::vector<uint8_t> ignored_;
    std::vector<int> indeg_, parent_, frontier_, next_, state_;
    std::vector<int64_t> score_;
    std::vector<Event> ev_;
};

}  // namespace smx
