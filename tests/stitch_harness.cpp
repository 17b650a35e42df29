// CPU harness for tests/test_stitch_cpu.py: exposes the host-side merge of smx_stitch.hpp (row A10) so that the
// run-length implementation can be compared with the column-string one without a GPU.  Test infrastructure only.
#include "../multiplexnanopore_b200/csrc/smx_stitch.hpp"
#include <cstring>

extern "C" long long smx_test_merge(const uint32_t *k1, long long nk1, long long ns1, long long nh1, const uint32_t *k2, long long nk2, long long ns2,
                                    long long nh2, long long n_ref, long long nq1, long long nq2, int go, int ge, int ma, int mi, int column_strings,
                                    uint32_t *out, long long cap)
{
    smx::ClippedSg r1, r2;
    r1.kept = k1; r1.n_kept = nk1; r1.n_s = ns1; r1.n_h = nh1;
    r2.kept = k2; r2.n_kept = nk2; r2.n_s = ns2; r2.n_h = nh2;
    const smx::Scoring sc{go, ge, ma, mi};
    smx::Runs runs;
    if (!smx::merge_special_runs(r1, r2, n_ref, nq1, nq2, sc, runs, column_strings != 0)) return -1;
    if ((long long)runs.size() > cap) return -2;
    if (!runs.empty()) memcpy(out, runs.data(), sizeof(uint32_t) * runs.size());
    return (long long)runs.size();
}

// rotate_to_ref0 on runs (in place, returns the run count) and on the column string; both return new_query_seq_offset
extern "C" long long smx_test_rotate(uint32_t *runs, long long n, long long cap, long long ref_off, long long q_off, int column_strings, long long *new_off)
{
    smx::Runs v(runs, runs + n);
    if (column_strings) {
        std::string cig;
        smx::expand_runs(v.data(), (int64_t)v.size(), cig);
        *new_off = smx::rotate_to_ref0(cig, ref_off, q_off);
        smx::encode_runs(cig, v);
    } else {
        *new_off = smx::rotate_to_ref0_runs(v, ref_off, q_off);
    }
    if ((long long)v.size() > cap) return -2;
    if (!v.empty()) memcpy(runs, v.data(), sizeof(uint32_t) * v.size());
    return (long long)v.size();
}

// set_cigar_score on runs
extern "C" long long smx_test_score(const uint32_t *runs, long long n, int go, int ge, int ma, int mi)
{
    smx::Runs v(runs, runs + n);
    return (long long)smx::cigar_score_runs(v, smx::Scoring{go, ge, ma, mi});
}
