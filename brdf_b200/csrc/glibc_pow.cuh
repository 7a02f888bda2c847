// glibc_pow.cuh -- libm's double-precision pow(), reproduced operation by operation.
//
// The reference's model callback evaluates pow(cosine, n) with the host's libm once per sample
// (BRDFFunc, brdfdata.cpp:981, 986), and levmar's trajectory is sensitive to the last bit of every
// residual.  CUDA's pow() is a different algorithm (up to 2 ulp, glibc's stays below
// 0.52 ulp), so the two disagree in the last bit of a sizeable fraction of the values.  The batched
// fit's levmar-exact mode (batched_fit.cu, BRDFGPU_JAC_FD_EXACT) therefore evaluates the SAME
// function the reference's host does: glibc >= 2.28's pow (sysdeps/ieee754/dbl-64/e_pow.c: table-driven
// log in double-double, y*log(x) with an exact product split, table-driven exp), in the variant every
// x86-64 CPU with FMA + AVX2 runs (__pow_fma, selected by libm's ifunc).  The sequence of
// multiplications, additions and fused multiply-adds below follows that variant's machine code
// instruction by instruction (the compiler's contraction choices are part of the result); the two
// constant tables are glibc's own, read from libm.so.6 by profiles/tools/extract_pow_tables.py.
// tests/test_glibc_pow.py compares the host instantiation with libm's pow() bit for bit on tens of
// millions of arguments, tests/test_gpu_exact.py does the same for the device instantiation.
//
// Not reproduced: errno and the floating-point exception flags (nothing on the path reads them).
#pragma once

#include <cstdint>
#include <cstring>

#include "glibc_pow_tables.inc"

#ifdef __CUDACC__
#define BG_POW_HD __host__ __device__ __forceinline__
#else
#define BG_POW_HD inline
#endif

namespace brdfgpu {
namespace glibcpow {

struct LogEntry {
    double invc, logc, logctail;
};

#ifdef __CUDA_ARCH__
#define BG_POW_TABLE static __device__ const
#else
#define BG_POW_TABLE static const
#endif
// (two copies: one in device global memory, read through L1, one for the host instantiation)
static __device__ const LogEntry d_log_tab[128] = BG_POWLOG_TAB;
static __device__ const unsigned long long d_exp_tab[256] = BG_EXP_TAB;
static __device__ const double d_log_poly[7] = BG_POWLOG_POLY;
static __device__ const double d_exp_poly[4] = BG_EXP_POLY;
static const LogEntry h_log_tab[128] = BG_POWLOG_TAB;
static const unsigned long long h_exp_tab[256] = BG_EXP_TAB;
static const double h_log_poly[7] = BG_POWLOG_POLY;
static const double h_exp_poly[4] = BG_EXP_POLY;

#ifdef __CUDA_ARCH__
#define BG_LOG_TAB d_log_tab
#define BG_EXP_TABLE d_exp_tab
#define BG_LOG_POLY d_log_poly
#define BG_EXP_POLYN d_exp_poly
#else
#define BG_LOG_TAB h_log_tab
#define BG_EXP_TABLE h_exp_tab
#define BG_LOG_POLY h_log_poly
#define BG_EXP_POLYN h_exp_poly
#endif

// single-rounding primitives: nothing below may be contracted or re-associated by a compiler
BG_POW_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
BG_POW_HD double mul_(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;  // (volatile: the host compiler may not fuse it into a neighbouring addition)
    return r;
#endif
}
BG_POW_HD double add_(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
BG_POW_HD double sub_(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double r = a - b;
    return r;
#endif
}
BG_POW_HD uint64_t bits_(double v) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return u;
#endif
}
BG_POW_HD double dbl_(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}

constexpr uint64_t kOne = 0x3ff0000000000000ull, kInf = 0x7ff0000000000000ull;
constexpr uint64_t kOff = 0x3fe6955500000000ull;  // OFF of log_inline
constexpr uint32_t kSignBias = 0x800u << 7;       // SIGN_BIAS = 0x800 << EXP_TABLE_BITS

// 0: not an integer, 1: odd integer, 2: even integer (e_pow.c checkint)
BG_POW_HD int checkint(uint64_t iy) {
    const int e = (int)((iy >> 52) & 0x7ff);
    if (e < 0x3ff) return 0;
    if (e > 0x3ff + 52) return 2;
    if (iy & ((1ull << (0x3ff + 52 - e)) - 1)) return 0;
    if (iy & (1ull << (0x3ff + 52 - e))) return 1;
    return 2;
}
BG_POW_HD bool zeroinfnan(uint64_t i) { return 2 * i - 1 >= 2 * kInf - 1; }

// Is exp_inline's argument outside the range its main path covers (|x| < 2^-54 or |x| >= 512)?
BG_POW_HD bool exp_needs_care(double x) { return ((uint32_t)(bits_(x) >> 52) & 0x7ff) - 0x3c9u >= 0x3fu; }

// The main path of exp_inline: sign_bias == 0 and 2^-54 <= |x| < 512, no branches (callers check exp_needs_care).
BG_POW_HD double exp_ordinary(double x, double xtail) {
    double kd = fma_(x, BG_EXP_INVLN2N, BG_EXP_SHIFT);
    const uint64_t ki = bits_(kd);
    kd = sub_(kd, BG_EXP_SHIFT);
    double r = fma_(kd, BG_EXP_NEGLN2HIN, x);
    r = fma_(kd, BG_EXP_NEGLN2LON, r);
    r = add_(xtail, r);
    const uint64_t idx = 2 * (ki % 128);
    const double tail = dbl_(BG_EXP_TABLE[idx]);
    const double scale = dbl_(BG_EXP_TABLE[idx + 1] + (ki << (52 - 7)));
    const double r2 = mul_(r, r);
    const double a = fma_(r, BG_EXP_POLYN[1], BG_EXP_POLYN[0]);
    const double b = fma_(r, BG_EXP_POLYN[3], BG_EXP_POLYN[2]);
    double tmp = fma_(a, r2, add_(r, tail));
    tmp = fma_(b, mul_(r2, r2), tmp);
    return fma_(tmp, scale, scale);
}

// exp(x + xtail) * (-1)^(sign_bias != 0), e_pow.c exp_inline + specialcase
BG_POW_HD double exp_inline(double x, double xtail, uint32_t sign_bias) {
    uint32_t abstop = (uint32_t)(bits_(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x3fu) {
        if (abstop - 0x3c9u >= 0x80000000u) {  // |x| < 2^-54: 1 + x, rounded once
            const double one = add_(1.0, x);
            return sign_bias ? -one : one;
        }
        if (abstop >= 0x409u) {  // |x| >= 1024: certain underflow / overflow
            if (bits_(x) >> 63) return sign_bias ? -0.0 : 0.0;                      // __math_uflow
            return sign_bias ? -dbl_(kInf) : dbl_(kInf);                            // __math_oflow
        }
        abstop = 0;  // large |x|: the scale needs care below
    }
    // x = ln2/N * k + r
    double kd = fma_(x, BG_EXP_INVLN2N, BG_EXP_SHIFT);
    const uint64_t ki = bits_(kd);
    kd = sub_(kd, BG_EXP_SHIFT);
    double r = fma_(kd, BG_EXP_NEGLN2HIN, x);
    r = fma_(kd, BG_EXP_NEGLN2LON, r);
    r = add_(xtail, r);
    const uint64_t idx = 2 * (ki % 128);
    const uint64_t top = (ki + sign_bias) << (52 - 7);
    const double tail = dbl_(BG_EXP_TABLE[idx]);
    uint64_t sbits = BG_EXP_TABLE[idx + 1] + top;
    const double r2 = mul_(r, r);
    // tmp = tail + r + r2 * (C2 + r * C3) + r2 * r2 * (C4 + r * C5), contracted as the FMA build contracts it
    const double a = fma_(r, BG_EXP_POLYN[1], BG_EXP_POLYN[0]);
    const double b = fma_(r, BG_EXP_POLYN[3], BG_EXP_POLYN[2]);
    double tmp = fma_(a, r2, add_(r, tail));
    tmp = fma_(b, mul_(r2, r2), tmp);
    if (abstop == 0) {  // specialcase()
        if ((ki & 0x80000000ull) == 0) {  // k > 0: the exponent of the scale may have overflowed
            sbits -= 1009ull << 52;
            const double scale = dbl_(sbits);
            return mul_(0x1p1009, fma_(scale, tmp, scale));
        }
        sbits += 1022ull << 52;  // k < 0: result in or near the subnormal range
        const double scale = dbl_(sbits);
        const double st = mul_(scale, tmp);
        double y = add_(scale, st);
        const double ay = y < 0.0 ? -y : y;
        if (ay < 1.0) {  // round to the final precision before scaling down (no double rounding)
            const double one = y < 0.0 ? -1.0 : 1.0;
            double lo = add_(sub_(scale, y), st);
            const double hi = add_(one, y);
            lo = add_(add_(sub_(one, hi), y), lo);
            y = sub_(add_(hi, lo), one);
            if (y == 0.0) y = dbl_(sbits & 0x8000000000000000ull);
        }
        return mul_(0x1p-1022, y);
    }
    const double scale = dbl_(sbits);
    return fma_(tmp, scale, scale);
}

// log(x) for x = 2^k z, as hi + *tail (e_pow.c log_inline, FMA branch)
BG_POW_HD double log_inline(uint64_t ix, double* tail) {
    const uint64_t tmp = ix - kOff;
    const int i = (int)((tmp >> (52 - 7)) % 128);
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffull << 52));
    const double z = dbl_(iz);
    const double kd = (double)k;
    const double invc = BG_LOG_TAB[i].invc, logc = BG_LOG_TAB[i].logc, logctail = BG_LOG_TAB[i].logctail;
    const double r = fma_(z, invc, -1.0);
    // k ln2 + log(c) + r
    const double t1 = fma_(kd, BG_POWLOG_LN2HI, logc);
    const double t2 = add_(t1, r);
    const double lo1 = fma_(kd, BG_POWLOG_LN2LO, logctail);
    const double lo2 = add_(sub_(t1, t2), r);
    const double ar = mul_(BG_LOG_POLY[0], r);
    const double ar2 = mul_(r, ar);
    const double ar3 = mul_(r, ar2);
    const double hi = add_(t2, ar2);
    const double lo3 = fma_(ar, r, -ar2);
    const double lo4 = add_(sub_(t2, hi), ar2);
    // p = ar3 * (A1 + r A2 + ar2 * (A3 + r A4 + ar2 * (A5 + r A6)))
    const double p12 = fma_(r, BG_LOG_POLY[2], BG_LOG_POLY[1]);
    const double p34 = fma_(r, BG_LOG_POLY[4], BG_LOG_POLY[3]);
    const double p56 = fma_(r, BG_LOG_POLY[6], BG_LOG_POLY[5]);
    const double inner = fma_(p56, ar2, p34);
    const double poly = fma_(ar2, inner, p12);
    double lo = add_(lo1, lo2);
    lo = add_(lo, lo3);
    lo = add_(lo, lo4);
    lo = fma_(ar3, poly, lo);
    const double y = add_(hi, lo);
    *tail = add_(sub_(hi, y), lo);
    return y;
}

}  // namespace glibcpow

// ---- pow(x, y) in two halves, for callers that raise the same base to many exponents -------------------------
// For an ordinary base (positive, normal, finite) and an ordinary exponent (2^-65 <= |y| < 2^63) pow() is
// exp_inline(y * log(x)) with log(x) = hi + lo from log_inline, which depends on x alone: it can be computed once per
// base and reused, with the very same bits as a fresh call.  Everything else goes through glibc_pow().
BG_POW_HD bool pow_base_is_ordinary(double x) { return (uint32_t)(glibcpow::bits_(x) >> 52) - 0x001u < 0x7ffu - 0x001u; }
BG_POW_HD bool pow_exponent_is_ordinary(double y) { return ((uint32_t)(glibcpow::bits_(y) >> 52) & 0x7ff) - 0x3beu < 0x43eu - 0x3beu; }
BG_POW_HD void pow_log_of_base(double x, double* hi, double* lo) { *hi = glibcpow::log_inline(glibcpow::bits_(x), lo); }
// y * (hi + lo) as ehi + elo (e_pow.c, FMA branch)
BG_POW_HD void pow_scaled_log(double hi, double lo, double y, double* ehi, double* elo) {
    *ehi = glibcpow::mul_(y, hi);
    *elo = glibcpow::fma_(y, lo, glibcpow::fma_(hi, y, -*ehi));
}

// pow(x, y) with the bits glibc's __pow_fma returns
BG_POW_HD double glibc_pow(double x, double y) {
    using namespace glibcpow;
    uint32_t sign_bias = 0;
    uint64_t ix = bits_(x);
    const uint64_t iy = bits_(y);
    uint32_t topx = (uint32_t)(ix >> 52);
    const uint32_t topy = (uint32_t)(iy >> 52);
    if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
        // x < 2^-1022 (zero, subnormal, negative), inf or nan; or |y| < 2^-65, |y| >= 2^63, nan
        if (zeroinfnan(iy)) {
            if (2 * iy == 0) return 1.0;  // (signalling NaNs do not occur on this path)
            if (ix == kOne) return 1.0;
            if (2 * ix > 2 * kInf || 2 * iy > 2 * kInf) return add_(x, y);
            if (2 * ix == 2 * kOne) return 1.0;
            if ((2 * ix < 2 * kOne) == !(iy >> 63)) return 0.0;  // |x| < 1 && y == inf, or |x| > 1 && y == -inf
            return mul_(y, y);
        }
        if (zeroinfnan(ix)) {
            double x2 = mul_(x, x);
            if ((ix >> 63) && checkint(iy) == 1) x2 = -x2;
            return (iy >> 63) ? 1.0 / x2 : x2;
        }
        if (ix >> 63) {  // finite x < 0
            const int yint = checkint(iy);
            if (yint == 0) return dbl_(0xfff8000000000000ull);  // __math_invalid: (x - x) / (x - x), x86's default NaN
            if (yint == 1) sign_bias = kSignBias;
            ix &= 0x7fffffffffffffffull;
            topx &= 0x7ff;
        }
        if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
            if (ix == kOne) return 1.0;
            if ((topy & 0x7ff) < 0x3beu) return ix > kOne ? add_(1.0, y) : sub_(1.0, y);  // |y| < 2^-65
            return ((ix > kOne) == (topy < 0x800u)) ? dbl_(kInf) : 0.0;                    // |y| >= 2^63
        }
        if (topx == 0) {  // subnormal x: normalise
            ix = bits_(mul_(x, 0x1p52));
            ix &= 0x7fffffffffffffffull;
            ix -= 52ull << 52;
        }
    }
    double lo;
    const double hi = log_inline(ix, &lo);
    const double ehi = mul_(y, hi);
    const double elo = fma_(y, lo, fma_(hi, y, -ehi));
    return exp_inline(ehi, elo, sign_bias);
}

}  // namespace brdfgpu
