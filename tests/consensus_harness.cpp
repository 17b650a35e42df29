// CPU harness for tests/test_consensus_cpu.py: the integer restatement of the x87 multiplication that the consensus kernel
// runs on the device (multiplexnanopore_b200/csrc/smx_ld80.hpp) against the FPU's own long double multiplication.
#include "../multiplexnanopore_b200/csrc/smx_ld80.hpp"

namespace {
struct Rng {
    uint64_t s;
    uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
};
long double make(uint64_t m, int e)   // m * 2^(e - 63) through the raw format
{
    SmxLd v; v.m = m; v.e = e; return smx_ld_to_native(v);
}
bool same(const SmxLd a, const SmxLd b)
{
    const bool a_nan = a.e == SMX_LD_ESPECIAL && a.m != 0x8000000000000000ull, b_nan = b.e == SMX_LD_ESPECIAL && b.m != 0x8000000000000000ull;
    if (a_nan || b_nan) return a_nan && b_nan;
    return a.m == b.m && a.e == b.e;
}
}  // namespace

// regime 0: normal operands, exponents anywhere; 1: products around the denormal range; 2: products around the overflow
// threshold; 3: denormal operands; 4: zeros, infinities, NaNs mixed in; 5: significands with long runs of zeros / ones (ties)
extern "C" long long smx_test_ldmul(unsigned long long seed, long long n, int regime, long double *bad_a, long double *bad_b)
{
    Rng r{seed * 0x9E3779B97F4A7C15ull + 0x1234567ull};
    long long bad = 0;
    for (long long i = 0; i < n; ++i) {
        uint64_t ma = r.next() | 0x8000000000000000ull, mb = r.next() | 0x8000000000000000ull;
        int ea = (int)(r.next() % 32766) - 16382, eb = (int)(r.next() % 32766) - 16382;
        if (regime == 1) { ea = (int)(r.next() % 200) - 8300; eb = -16382 - ea + (int)(r.next() % 160) - 100; if (eb < -16382) eb = -16382; }
        if (regime == 2) { ea = (int)(r.next() % 200) + 8100; eb = 16383 - ea + (int)(r.next() % 8) - 4; if (eb > 16383) eb = 16383; }
        if (regime == 3) { ea = -16382; ma = r.next() >> (1 + r.next() % 63); if (r.next() & 1) { eb = -16382; mb = r.next() >> (1 + r.next() % 63); } else eb = (int)(r.next() % 16500); }
        if (regime == 5) {
            const int ka = 1 + (int)(r.next() % 62), kb = 1 + (int)(r.next() % 62);
            ma = 0x8000000000000000ull | ((r.next() & 1) ? ((1ull << ka) - 1) : (1ull << ka)) | ((r.next() & 1) ? 1ull : 0ull);
            mb = 0x8000000000000000ull | ((r.next() & 1) ? ~((1ull << kb) - 1) : (1ull << kb));
            if (r.next() & 1) { ea = (int)(r.next() % 100) - 8240; eb = -16382 - ea - (int)(r.next() % 70); if (eb < -16382) eb = -16382; }
        }
        volatile long double a = make(ma, ea), b = make(mb, eb);
        if (regime == 4) {
            const unsigned k = (unsigned)(r.next() % 6);
            if (k == 0) a = 0.0L;
            if (k == 1) b = 0.0L;
            if (k == 2) a = __builtin_infl();
            if (k == 3) { a = __builtin_infl(); b = 0.0L; }
            if (k == 4) a = __builtin_nanl("");
            if (k == 5) { a = __builtin_infl(); b = __builtin_infl(); }
        }
        volatile long double want = a * b;
        const SmxLd got = smx_ld_mul(smx_ld_from_native(a), smx_ld_from_native(b));
        if (!same(got, smx_ld_from_native(want))) {
            if (bad == 0 && bad_a && bad_b) { *bad_a = a; *bad_b = b; }
            ++bad;
        }
    }
    return bad;
}

// a chain of n factors drawn around `centre` (as the product over the reads of a column): restated vs native, every step
extern "C" long long smx_test_ldchain(unsigned long long seed, long long n, double centre)
{
    Rng r{seed * 0x9E3779B97F4A7C15ull + 0x777ull};
    volatile long double v = 1.0L;
    SmxLd w = smx_ld_from_native(1.0L);
    long long bad = 0;
    for (long long i = 0; i < n; ++i) {
        const double u = (double)(r.next() >> 11) / 9007199254740992.0;
        volatile long double num = (long double)(centre * (0.25 + 1.5 * u)) * (long double)(0.001 + u), den = (long double)(0.5 + u) * (long double)(0.002 + 0.5 * u);
        volatile long double f = num / den;
        v = v * f;
        w = smx_ld_mul(w, smx_ld_from_native(f));
        if (!same(w, smx_ld_from_native(v))) ++bad;
    }
    return bad;
}

// CPU replay of smx_consensus_run: the host code of the product (smx_consensus_host.hpp: tables, start values, final
// sums) around the kernel's per-thread body run in a loop.  Returns 0, or -7 like the entry point.
#include "../multiplexnanopore_b200/csrc/smx_consensus_host.hpp"
extern "C" int smx_test_consensus(const uint8_t *seq, const int8_t *qs, const uint8_t *cig, int n_rows, long long n_cols, const double *P25, const double *pdf,
                                  const int32_t *col_class, const double *pn_with, const double *pn_without, int n_classes, double *p_with, double *p_without,
                                  int32_t *n_events)
{
    std::vector<uint64_t> tm, im;
    std::vector<int32_t> te, ie;
    smx_cons_build_table(P25, pdf, tm, te);
    smx_cons_build_init(pn_with, pn_without, n_classes, im, ie);
    const size_t C = (size_t)n_cols;
    std::vector<uint64_t> om(50 * C);
    std::vector<int32_t> oe(50 * C);
    bool ok = true;
    for (size_t c = 0; c < C; ++c)
        for (int pair = 0; pair < 25; ++pair) {
            const int cls = col_class[c];
            SmxLd v_with, v_without;
            v_with.m = im[(size_t)cls * 25 + pair]; v_with.e = ie[(size_t)cls * 25 + pair];
            v_without.m = im[((size_t)n_classes + cls) * 25 + pair]; v_without.e = ie[((size_t)n_classes + cls) * 25 + pair];
            int events = 0;
            ok &= smx_cons_column(seq, qs, cig, n_rows, n_cols, (int64_t)c, tm.data() + (size_t)pair * SMX_CONS_BASES * SMX_CONS_QIDX,
                                  te.data() + (size_t)pair * SMX_CONS_BASES * SMX_CONS_QIDX, v_with, v_without, events);
            om[c * 25 + pair] = v_with.m; oe[c * 25 + pair] = v_with.e;
            om[C * 25 + c * 25 + pair] = v_without.m; oe[C * 25 + c * 25 + pair] = v_without.e;
            if (pair == 0) n_events[c] = events;
        }
    if (!ok) return -7;
    smx_cons_finish(om.data(), oe.data(), C, p_with, p_without);
    return 0;
}
