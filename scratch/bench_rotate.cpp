// CPU micro-benchmark of smx::rotate_to_ref0_runs (host stitch stage, row A11) on run arrays shaped like C2 pair-strands:
//   g++ -O2 -std=c++17 -o /tmp/bench_rotate scratch/bench_rotate.cpp && /tmp/bench_rotate
#include <chrono>
#include <cstdio>
#include <random>
#include <string>
#include "../multiplexnanopore_b200/csrc/smx_stitch.hpp"

int main()
{
    std::mt19937_64 rng(5);
    const int N = 4000;
    std::vector<smx::Runs> cases(N);
    std::vector<long long> ro(N), qo(N);
    for (int c = 0; c < N; ++c) {
        smx::Runs &v = cases[c];
        long long nref = 0, nq = 0;
        const int target = 5000 + (int)(rng() % 5000);
        uint32_t last = 0;
        while (nref < target) {
            const unsigned u = rng() % 100;
            uint32_t op = u < 55 ? 7u : u < 70 ? 8u : u < 85 ? 1u : 2u;   // = X I D
            if (op == last) op = 7u == op ? 8u : 7u;
            const uint32_t len = op == 7u ? 1 + rng() % 25 : 1 + rng() % 2;
            v.push_back((len << 4) | op);
            if (op != 1u) nref += len;
            if (op != 2u) nq += len;
            last = op;
        }
        ro[c] = (long long)(rng() % (unsigned long long)nref);
        qo[c] = (long long)(rng() % (unsigned long long)nq);
    }
    double runs = 0;
    for (auto &v : cases) runs += (double)v.size();
    long long sink = 0;
    for (int rep = 0; rep < 3; ++rep) {
        std::vector<smx::Runs> work = cases;
        const auto t0 = std::chrono::steady_clock::now();
        for (int c = 0; c < N; ++c) sink += smx::rotate_to_ref0_runs(work[c], ro[c], qo[c]);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("rotate_to_ref0_runs: %.2f us per pair-strand (%.0f runs on average), %.2f ns per run\n", 1e6 * dt / N, runs / N, 1e9 * dt / runs);
    }
    return (int)(sink & 1);
}
