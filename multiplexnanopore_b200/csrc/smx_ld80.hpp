// x87 extended precision (the `long double` of the reference's Cython consensus code, alignment_functions.pyx:70-132,
// on x86-64 Linux: 64-bit significand with an explicit integer bit, 15-bit exponent, round to nearest even, gradual
// underflow) for NON-NEGATIVE values, restated on integers so that a CUDA thread multiplies exactly like the FPU does.
// Only the multiplication runs on the device (the product over the reads of a column is the whole cost of
// SequenceBasecallQscorePDF.calc_consensus_error_rate); tables, sums and the final 1 - 1 / sum are computed by the host
// in native long double.  Compiles as plain C++ (tests/consensus_harness.cpp checks it against the FPU) and as CUDA.
#pragma once
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define SMX_LD_HD __host__ __device__ __forceinline__
#else
#define SMX_LD_HD inline
#endif

// value = m * 2^(e - 63).  Finite: -16382 <= e <= 16383, m >= 2^63 unless e == -16382 (denormal or zero: m < 2^63).
// e == SMX_LD_ESPECIAL: infinity (m == 2^63) or NaN (anything else).
struct SmxLd {
    uint64_t m;
    int32_t e;
};
#define SMX_LD_EMIN (-16382)
#define SMX_LD_EMAX 16383
#define SMX_LD_ESPECIAL 16384

SMX_LD_HD void smx_ld_mul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo)
{
#ifdef __CUDA_ARCH__
    hi = __umul64hi(a, b);
    lo = a * b;
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    hi = (uint64_t)(p >> 64);
    lo = (uint64_t)p;
#endif
}

SMX_LD_HD int smx_ld_clz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

SMX_LD_HD SmxLd smx_ld_nan() { SmxLd r; r.m = 0xC000000000000000ull; r.e = SMX_LD_ESPECIAL; return r; }
SMX_LD_HD SmxLd smx_ld_inf() { SmxLd r; r.m = 0x8000000000000000ull; r.e = SMX_LD_ESPECIAL; return r; }
SMX_LD_HD SmxLd smx_ld_zero() { SmxLd r; r.m = 0; r.e = SMX_LD_EMIN; return r; }

// a * b, rounded once to nearest even in the destination format (normal or denormal), overflow to infinity
SMX_LD_HD SmxLd smx_ld_mul(const SmxLd a, const SmxLd b)
{
    if (a.e == SMX_LD_ESPECIAL || b.e == SMX_LD_ESPECIAL) {
        const bool a_nan = a.e == SMX_LD_ESPECIAL && a.m != 0x8000000000000000ull, b_nan = b.e == SMX_LD_ESPECIAL && b.m != 0x8000000000000000ull;
        if (a_nan || b_nan) return smx_ld_nan();
        if (a.m == 0 || b.m == 0) return smx_ld_nan();  // infinity * 0
        return smx_ld_inf();
    }
    if (a.m == 0 || b.m == 0) return smx_ld_zero();
    uint64_t hi, lo;
    smx_ld_mul64(a.m, b.m, hi, lo);
    const int L = hi ? 127 - smx_ld_clz64(hi) : 63 - smx_ld_clz64(lo);  // leading bit of the 128-bit product
    int e = a.e + b.e + L - 126;
    int s = L - 63;                                   // right shift that leaves 64 significant bits
    if (e < SMX_LD_EMIN) { s += SMX_LD_EMIN - e; e = SMX_LD_EMIN; }  // denormal: fewer bits survive
    uint64_t q;
    bool up;
    if (s <= 0) {            // only with L == 63 and no denormal shift: exact
        q = lo; up = false;
    } else if (s < 64) {
        q = (hi << (64 - s)) | (lo >> s);
        const uint64_t rem = lo & ((1ull << s) - 1), half = 1ull << (s - 1);
        up = rem > half || (rem == half && (q & 1));
    } else if (s == 64) {
        q = hi;
        up = lo > 0x8000000000000000ull || (lo == 0x8000000000000000ull && (q & 1));
    } else if (s < 128) {
        const int t = s - 64;
        q = hi >> t;
        const uint64_t rem_hi = hi & ((1ull << t) - 1), half_hi = 1ull << (t - 1);
        up = rem_hi > half_hi || (rem_hi == half_hi && (lo != 0 || (q & 1)));
    } else if (s == 128) {
        q = 0;
        up = hi > 0x8000000000000000ull || (hi == 0x8000000000000000ull && lo != 0);
    } else {
        q = 0; up = false;
    }
    SmxLd r;
    if (up) {
        ++q;
        if (q == 0) { q = 0x8000000000000000ull; ++e; }   // carried out of 64 bits
    }
    if (e > SMX_LD_EMAX) return smx_ld_inf();
    r.m = q;
    r.e = q == 0 ? SMX_LD_EMIN : e;
    return r;
}

// host only: conversion from / to the FPU's own format (x86-64: 10 significant bytes, little endian)
inline SmxLd smx_ld_from_native(long double v)
{
    static_assert(sizeof(long double) >= 10, "x87 extended precision expected");
    unsigned char raw[16] = {0};
    std::memcpy(raw, &v, 10);
    uint64_t m;
    uint16_t se;
    std::memcpy(&m, raw, 8);
    std::memcpy(&se, raw + 8, 2);
    const int f = se & 0x7fff;
    SmxLd r;
    r.m = m;
    r.e = f == 0 ? SMX_LD_EMIN : f == 0x7fff ? SMX_LD_ESPECIAL : f - 16383;
    return r;
}
inline long double smx_ld_to_native(const SmxLd a)
{
    unsigned char raw[16] = {0};
    uint16_t se;
    if (a.e == SMX_LD_ESPECIAL) se = 0x7fff;
    else if (a.m < 0x8000000000000000ull) se = 0;          // denormal or zero
    else se = (uint16_t)(a.e + 16383);
    std::memcpy(raw, &a.m, 8);
    std::memcpy(raw + 8, &se, 2);
    long double v = 0;
    std::memcpy(&v, raw, 10);
    return v;
}
