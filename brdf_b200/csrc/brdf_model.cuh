// brdf_model.cuh -- per-sample BRDF model arithmetic shared by every fit kernel.
//
// Reference callback (brdfdata.cpp:969-989):
//     Blinn-Phong (model 1): hx = p0*cosphi + p1*pow(costhetadash, p2)
//     Phong       (model 0): hx = p0*cosphi + ((p2+2)/2*pi)*p1*pow(costheta, p2)
// and levmar's difference Jacobian around it (misc_core.c:137-211).
//
// Device data layout (24 B per sample per pass): c = cosphi, L = log(t) with t the model's cosine,
// x = measurement.  t**n is evaluated as exp(n*L): the log is paid once when the sample set is
// made resident, and one pass needs one exp per evaluation point (two for a forward-difference
// Jacobian: n and n+d; the kd and ks columns reuse t**n) where levmar spends m+1 = 4 pow() calls.
//
// Two per-sample paths:
//   fast    |n*L| <= 700 for every exponent of the pass: branch-free exp (range reduction, degree-10
//           polynomial, exponent add), ~15 fp64 instructions.  Max relative error 1e-15.
//   careful everything else -- L is NaN (t negative or NaN), t == 0 or inf, exponent zero or not
//           finite, results that under/overflow: goes through the raw cosine and pow(), so libm's
//           special cases survive: pow(t<0, integer n) finite, pow(t<0, non-integer)
//           NaN, pow(0, n>0) = 0, pow(x, 0) = 1.
// The difference quotients are formed algebraically instead of subtracting two full model values:
//     column kd : ((kd+d0) - kd)/d0 * c                    (the model is linear in kd and ks, so this
//     column ks : ((ks+d1) - ks)/d1 * coef * t**n           is levmar's value up to its own rounding)
//     column n  : (coef(n+d2)*ks*t**(n+d2) - coef(n)*ks*t**n) / d2     (the truncation error of the
//                 finite step is part of the reference's fixed point, SURVEY.md Q11, and is kept)
// Residuals stay within 1e-12 relative of the reference callback; FMA contraction is allowed here.
#pragma once

#include "lm_engine.cuh"

namespace brdfgpu {

constexpr double kPi = 3.1415926535897932384626433832795;  // CV_PI, brdfdata.cpp:981
constexpr double kFastExpLimit = 700.0;

enum JacMode { kJacForward = 0, kJacCentral = 1, kJacAnalytic = 2 };

// Everything that depends on p only, computed once per pass (uniform across threads).
struct PassParams {
    double kd, ks, n, coef;  // coef = 1 (Blinn-Phong) or (n+2)/2*pi (Phong)
    double cks;              // coef*ks
    double n_hi, n_lo;       // n + d2, n - d2
    double g0, g1;           // ((kd+d0)-kd)/d0 [central: ((kd+d0)-(kd-d0))*0.5/d0], same for ks; analytic: 1
    double a_hi, a_lo;       // coef(n_hi)*ks/d2 and coef(n)*ks/d2 [central: coef(n_lo)*ks*0.5/d2]
    double dcoef;            // d coef / d n (analytic, Phong: pi/2)
    // literal difference data for the careful path: p_j + d_j, p_j - d_j, 1/d_j (0.5/d_j central)
    double kd_hi, ks_hi, kd_lo, ks_lo, coef_hi, coef_lo, inv[3];
    int model;
};

BG_HDI double model_coef(int model, double n) { return model == 1 ? 1.0 : ((n + 2.0) / 2.0 * kPi); }

BG_HDI PassParams make_pass_params(const double* p, int model, double delta, int jac_mode) {
    PassParams q;
    q.model = model;
    q.kd = p[0]; q.ks = p[1]; q.n = p[2];
    q.coef = model_coef(model, q.n);
    q.cks = q.coef * q.ks;
    q.dcoef = model == 1 ? 0.0 : (kPi / 2.0);
    double d[3];
    const double half = (jac_mode == kJacCentral) ? 0.5 : 1.0;
    for (int j = 0; j < 3; ++j) {  // d_j = max(|1e-4 p_j|, delta), misc_core.c:154-158, 193-196
        double dj = 1E-04 * p[j];
        dj = lm_abs(dj);
        if (dj < delta) dj = delta;
        d[j] = dj;
        q.inv[j] = (dj == 1.0) ? half : half / dj;  // (delta = 1, the reference's global preset: no fp64 division on the control path)
    }
    q.kd_hi = p[0] + d[0]; q.ks_hi = p[1] + d[1]; q.n_hi = p[2] + d[2];
    q.kd_lo = p[0] - d[0]; q.ks_lo = p[1] - d[1]; q.n_lo = p[2] - d[2];
    q.coef_hi = model_coef(model, q.n_hi);
    q.coef_lo = model_coef(model, q.n_lo);
    if (jac_mode == kJacForward) {
        q.g0 = (q.kd_hi - q.kd) * q.inv[0];
        q.g1 = (q.ks_hi - q.ks) * q.inv[1];
        q.a_hi = (q.coef_hi * q.ks) * q.inv[2];
        q.a_lo = q.cks * q.inv[2];
        q.n_lo = q.n;
    } else if (jac_mode == kJacCentral) {
        q.g0 = (q.kd_hi - q.kd_lo) * q.inv[0];
        q.g1 = (q.ks_hi - q.ks_lo) * q.inv[1];
        q.a_hi = (q.coef_hi * q.ks) * q.inv[2];
        q.a_lo = (q.coef_lo * q.ks) * q.inv[2];
    } else {
        q.g0 = 1.0; q.g1 = 1.0; q.a_hi = 0.0; q.a_lo = 0.0;
        q.n_hi = q.n_lo = q.n;
    }
    return q;
}

// The three numbers a trial-point (cost-only) evaluation needs.
struct CostPoint {
    double kd, cks, n;
};

BG_HDI CostPoint make_cost_point(const double* p, int model) {
    CostPoint q;
    q.kd = p[0];
    q.n = p[2];
    q.cks = model_coef(model, p[2]) * p[1];
    return q;
}

#ifdef __CUDACC__
// Coefficients of exp_core in constant memory: DFMA takes them as c[bank][offset] operands, so no
// instruction is spent on materialising 64-bit immediates inside the sample loops.
// near-minimax degree 10 on |r| <= ln2/2 (Chebyshev interpolation, max rel. error 9e-16)
static __constant__ double kExpPoly[10] = {
    0x1.26e46de8d8e82p-22, 0x1.7303e941557e0p-19, 0x1.a01b8fdc8d247p-16, 0x1.a01970086f448p-13, 0x1.6c16c0c5a5d65p-10,
    0x1.11111130a260cp-7,  0x1.555555558aa3dp-5,  0x1.555555554a290p-3,  0x1.ffffffffffe8fp-2,  0x1.0000000000024p+0};
static __constant__ double kExpRed[4] = {1.4426950408889634074, 6755399441055744.0 /* 1.5 * 2^52 */,
                                         -6.93147180369123816490e-01, -1.90821492927058770002e-10};

// exp(y) for |y| <= 700 (no NaN/Inf handling, no subnormal results by construction; any other
// input gives garbage but never traps -- callers overwrite those lanes on their careful path)
__device__ __forceinline__ double exp_core(double y) {
    const double magic = kExpRed[1];  // round-to-nearest integer lands in the low word
    const double t = __fma_rn(y, kExpRed[0], magic);
    const int k = __double2loint(t);
    const double kf = t - magic;
    double r = __fma_rn(kf, kExpRed[2], y);
    r = __fma_rn(kf, kExpRed[3], r);
    double v = kExpPoly[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) v = __fma_rn(v, r, kExpPoly[i]);
    v = __fma_rn(v, r, 1.0);
    return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

#ifdef BG_EXP_TABLE
// Table-driven variant (global_fit.cu): exp(y) = 2^k T[i] (1 + p(r)) with 64 k + i = round(y 64/ln2), T[i] = 2^(i/64)
// from shared memory and a degree-5 p on |r| <= ln2/128 -- 10 fp64 instructions per exponential instead of 14 and a
// dependency chain of 8 instead of 15, max relative error 2.1e-16 (1.9 ulp, against 4 ulp of the polynomial; exact
// emulation in profiles/tools/make_exp_table.py, which also generated the constants).  The table is replicated so that
// lane l reads copy l mod 16 whose entries all sit in bank pair l mod 16: an LDS.64 of a warp is conflict-free whatever
// the 32 indices are (8 KB; with 8 copies a Jacobian sweep of the persistent fit takes 11 % longer).  Every kernel of
// such a translation unit calls exp_table_load() first.
// Measured (profiles/r02_persist_control.md): the resident sweeps of the persistent fit take the same ~2.7 k cycles
// per trial point with either variant -- ptxas runs the chains of a trip nearly one after the other, so the sweeps
// wait on FP64 latency (13 cycles per dependent step, 4 warps per scheduler), not on the FP64 issue rate this saves.
#include "exp_table.inc"
#ifndef BG_EXP_TABLE_COPIES
#define BG_EXP_TABLE_COPIES 16
#endif
constexpr int kExpTabCopies = BG_EXP_TABLE_COPIES;
__shared__ unsigned long long s_exp_tab[64 * kExpTabCopies];

__device__ __forceinline__ void exp_table_load() {
    for (int i = threadIdx.x; i < 64 * kExpTabCopies; i += blockDim.x) s_exp_tab[i] = kExp2TabBiased[i / kExpTabCopies];
    __syncthreads();
}
#define BG_EXP_TABLE_LOAD() exp_table_load()

// N independent exponentials, step-major (see the polynomial variant below).  The integer side is spelled out: and +
// multiply-add for the address of the entry, ONE multiply-add for the scaling by 2^k (the table holds 2^(i/64) with
// i << 14 taken off the high word, so that adding j << 14 = (64 k + i) << 14 leaves k << 20 on top of the true entry),
// and the result falls out of the last FMA.
template <int N>
__device__ __forceinline__ void exp_core_n(const double* y, double* out) {
    const double magic = kExpTabRed[1];
    const unsigned lane_bytes = (unsigned)__cvta_generic_to_shared(s_exp_tab) + ((threadIdx.x & (kExpTabCopies - 1)) << 3);
    double t[N], r[N], T[N], q0[N], q1[N], r2[N];
#pragma unroll
    for (int k = 0; k < N; ++k) t[k] = __fma_rn(y[k], kExpTabRed[0], magic);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const unsigned j = (unsigned)__double2loint(t[k]);
        unsigned lo, hi, addr;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(j & 63u), "n"(8 * kExpTabCopies), "r"(lane_bytes));  // (kept as ONE IMAD)
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
        T[k] = __hiloint2double((int)(hi + j * 0x4000u), (int)lo);  // 2^k T[i]
    }
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = __fma_rn(t[k] - magic, kExpTabRed[2], y[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = __fma_rn(t[k] - magic, kExpTabRed[3], r[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) q1[k] = __fma_rn(kExpTabPoly[0], r[k], kExpTabPoly[1]);
#pragma unroll
    for (int k = 0; k < N; ++k) q0[k] = __fma_rn(kExpTabPoly[2], r[k], kExpTabPoly[3]);
#pragma unroll
    for (int k = 0; k < N; ++k) r2[k] = r[k] * r[k];
#pragma unroll
    for (int k = 0; k < N; ++k) q0[k] = __fma_rn(q1[k], r2[k], q0[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = __fma_rn(q0[k], r2[k], r[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = __fma_rn(T[k], r[k], T[k]);
}
#else
#define BG_EXP_TABLE_LOAD() ((void)0)
// N independent exponentials, written step-major (every step for all N before the next step) so the
// instruction stream interleaves N dependency chains: one warp then keeps the FP64 pipe busy by
// itself instead of relying on other warps to cover the DFMA latency.
template <int N>
__device__ __forceinline__ void exp_core_n(const double* y, double* out) {
    const double magic = kExpRed[1];
    double t[N], r[N], v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) t[k] = __fma_rn(y[k], kExpRed[0], magic);
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = __fma_rn(t[k] - magic, kExpRed[2], y[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = __fma_rn(t[k] - magic, kExpRed[3], r[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = __fma_rn(kExpPoly[0], r[k], kExpPoly[1]);
#pragma unroll
    for (int i = 2; i < 10; ++i) {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = __fma_rn(v[k], r[k], kExpPoly[i]);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = __fma_rn(v[k], r[k], 1.0);
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = __hiloint2double(__double2hiint(v[k]) + (__double2loint(t[k]) << 20), __double2loint(v[k]));
}

#endif  // BG_EXP_TABLE

// t**n / ln t with libm semantics through the raw cosine.  Deliberately NOT inlined and with
// by-value arguments only: the careful path is rare and pow() is ~300 instructions per site.
static __device__ __noinline__ double pow_careful(double traw, double n) { return pow(traw, n); }
static __device__ __noinline__ double log_careful(double traw) { return log(traw); }

// log once per sample when a set becomes resident; NaN marks "go through pow()".
__device__ __forceinline__ double log_or_flag(double t) {
    if (t >= 0.0) return log(t);  // t == 0 -> -inf, t == +inf -> +inf
    return __longlong_as_double(0x7ff8000000000000LL);
}

// Accumulator layout of one evaluation: JtJ upper triangle, Jte, ||e||^2.  Whether some residual
// is non-finite is only needed when ||e||^2 itself is not finite (lmbc_core.c:748,915) and is
// then counted by a separate pass.
enum { A00 = 0, A01, A02, A11, A12, A22, G0, G1, G2, ESQ, NACC };

__device__ __forceinline__ void accumulate_normal(double j0, double j1, double j2, double e, double* acc) {
    acc[A00] = __fma_rn(j0, j0, acc[A00]); acc[A01] = __fma_rn(j0, j1, acc[A01]); acc[A02] = __fma_rn(j0, j2, acc[A02]);
    acc[A11] = __fma_rn(j1, j1, acc[A11]); acc[A12] = __fma_rn(j1, j2, acc[A12]); acc[A22] = __fma_rn(j2, j2, acc[A22]);
    acc[G0] = __fma_rn(j0, e, acc[G0]); acc[G1] = __fma_rn(j1, e, acc[G1]); acc[G2] = __fma_rn(j2, e, acc[G2]);
    acc[ESQ] = __fma_rn(e, e, acc[ESQ]);
}

// Does an exponential with argument y = exponent * log(t) need libm semantics?  exp_core is valid for
// |y| < 700; everything else -- results that under/overflow, the NaN flag of t < 0, t == 0 or inf
// (y = +-inf, or NaN against a zero exponent), non-finite exponents -- goes through pow().
// A zero exponent with an ordinary t gives y = 0 and exp_core(0) == 1 == pow(t, 0) exactly.
// (decided on the high word -- |y| >= 700, Inf and NaN all have one at or above 700.0's -- so the test runs on the
// integer pipe and leaves the FP64 pipe to the arithmetic)
__device__ __forceinline__ bool needs_care(double y) { return (unsigned)(__double2hiint(y) & 0x7fffffff) >= 0x4085e000u; }

// model value on the careful path: products and sum rounded separately, as the reference callback
// compiled without contraction does.  One spelling for every pass, so the residual of a sample at a
// point has the same bits whether a cost sweep or a Jacobian sweep computes it (speculative Jacobians).
__device__ __forceinline__ double model_value_careful(double kd, double c, double cks, double pw) {
    return __dadd_rn(__dmul_rn(kd, c), __dmul_rn(cks, pw));
}

// careful path of one Jacobian sample: literal levmar differences of full model values
template <int JAC>
__device__ __forceinline__ void jac_terms_careful(const PassParams& q, double c, double traw, double x, double* out4) {
    const double pw = pow_careful(traw, q.n);
    const double hx = model_value_careful(q.kd, c, q.cks, pw);
    double j0, j1, j2;
    if (JAC == kJacForward) {  // jac[i][j] = (f(p + d_j e_j) - f(p)) * (1/d_j), misc_core.c:160-170
        const double pw_hi = pow_careful(traw, q.n_hi);
        j0 = ((q.kd_hi * c + q.cks * pw) - hx) * q.inv[0];
        j1 = ((q.kd * c + (q.coef * q.ks_hi) * pw) - hx) * q.inv[1];
        j2 = ((q.kd * c + (q.coef_hi * q.ks) * pw_hi) - hx) * q.inv[2];
    } else if (JAC == kJacCentral) {  // (f(p + d_j e_j) - f(p - d_j e_j)) * (0.5/d_j), misc_core.c:198-209
        const double pw_hi = pow_careful(traw, q.n_hi);
        const double pw_lo = pow_careful(traw, q.n_lo);
        j0 = ((q.kd_hi * c + q.cks * pw) - (q.kd_lo * c + q.cks * pw)) * q.inv[0];
        j1 = ((q.kd * c + (q.coef * q.ks_hi) * pw) - (q.kd * c + (q.coef * q.ks_lo) * pw)) * q.inv[1];
        j2 = ((q.kd * c + (q.coef_hi * q.ks) * pw_hi) - (q.kd * c + (q.coef_lo * q.ks) * pw_lo)) * q.inv[2];
    } else {  // exact partials: d/dkd = c, d/dks = coef t^n, d/dn = ks t^n (dcoef + coef ln t)
        j0 = c;
        j1 = q.coef * pw;
        j2 = q.ks * pw * (q.dcoef + q.coef * log_careful(traw));
    }
    out4[0] = __dsub_rn(x, hx); out4[1] = j0; out4[2] = j1; out4[3] = j2;
}

// N samples of a fused residual + Jacobian + normal-equation pass.  The fast path of all N samples
// is straight-line code (2N..3N independent exp chains the scheduler can interleave); samples that
// need libm semantics are redone afterwards under one rarely-taken branch.
// `q` feeds the fast path (callers pass a register copy: only the handful of fields used here stay
// live), `cold` the careful path (callers pass the original in shared / parameter memory).
template <int JAC, int N>
__device__ __forceinline__ void accumulate_jac_n(const PassParams& q, const PassParams& cold, const double* c,
                                                 const double* L, const double* x, const double* __restrict__ traw,
                                                 const long* idx, double* acc) {
    double e[N], j0[N], j1[N], j2[N];
    bool slow = false;
    const double g1c = (q.model == 1) ? q.g1 : q.g1 * q.coef;
    constexpr int NE = (JAC == kJacForward) ? 2 : (JAC == kJacCentral) ? 3 : 1;  // exponentials per sample
    double y[NE * N], pw[NE * N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        y[k] = q.n * L[k];
        if (NE >= 2) y[N + k] = q.n_hi * L[k];
        if (NE >= 3) y[2 * N + k] = q.n_lo * L[k];
    }
#pragma unroll
    for (int k = 0; k < NE * N; ++k) slow |= needs_care(y[k]);
    exp_core_n<NE * N>(y, pw);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        e[k] = x[k] - __fma_rn(q.kd, c[k], q.cks * pw[k]);
        j0[k] = q.g0 * c[k];
        j1[k] = g1c * pw[k];
        if (JAC == kJacForward) j2[k] = __fma_rn(q.a_hi, pw[N + k], -(q.a_lo * pw[k]));
        else if (JAC == kJacCentral) j2[k] = __fma_rn(q.a_hi, pw[N + k], -(q.a_lo * pw[2 * N + k]));
        else j2[k] = (q.ks * pw[k]) * __fma_rn(q.coef, L[k], q.dcoef);
    }
    if (slow) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            bool mine = needs_care(y[k]);
            if (NE >= 2) mine |= needs_care(y[N + k]);
            if (NE >= 3) mine |= needs_care(y[2 * N + k]);
            if (mine) {
                double o[4];
                jac_terms_careful<JAC>(cold, c[k], traw[idx[k]], x[k], o);
                // the residual takes the careful value exactly when a cost pass would (its test sees t**n only):
                // a sample whose t**(n+d) alone leaves the fast range keeps the fast residual, so ||e||^2 of a
                // Jacobian pass and of a cost pass at the same point agree bit for bit (speculative Jacobians)
                if (needs_care(y[k])) e[k] = o[0];
                j0[k] = o[1]; j1[k] = o[2]; j2[k] = o[3];
            }
        }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) accumulate_normal(j0[k], j1[k], j2[k], e[k], acc);
}

// One sample (tails, the batched kernels' per-lane loops)
template <int JAC>
__device__ __forceinline__ void accumulate_jac(const PassParams& q, double c, double L, double x,
                                               const double* __restrict__ traw, long i, double* acc) {
    accumulate_jac_n<JAC, 1>(q, q, &c, &L, &x, traw, &i, acc);
}

// residual e = x - f(p) (same arithmetic in every pass).  Q is PassParams or CostPoint.
template <class Q>
__device__ __forceinline__ double residual_careful(const Q& q, double c, double traw, double x) {
    return __dsub_rn(x, model_value_careful(q.kd, c, q.cks, pow_careful(traw, q.n)));
}

template <int N, class Q>
__device__ __forceinline__ void residuals_n(const Q& q, const double* c, const double* L, const double* x,
                                            const double* __restrict__ traw, const long* idx, double* e) {
    bool slow = false;
    double y[N], pw[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        y[k] = q.n * L[k];
        slow |= needs_care(y[k]);
    }
    exp_core_n<N>(y, pw);
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = x[k] - __fma_rn(q.kd, c[k], q.cks * pw[k]);
    if (slow) {
#pragma unroll
        for (int k = 0; k < N; ++k)
            if (needs_care(y[k])) e[k] = residual_careful(q, c[k], traw[idx[k]], x[k]);
    }
}

// the same N samples at TWO trial points: 2N interleaved exp chains, samples loaded once
template <int N, class Q>
__device__ __forceinline__ void residuals_n_x2(const Q& qa, const Q& qb, const double* c, const double* L, const double* x,
                                               const double* __restrict__ traw, const long* idx, double* ea, double* eb) {
    bool slow = false;
    double y[2 * N], pw[2 * N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        y[k] = qa.n * L[k];
        y[N + k] = qb.n * L[k];
    }
#pragma unroll
    for (int k = 0; k < 2 * N; ++k) slow |= needs_care(y[k]);
    exp_core_n<2 * N>(y, pw);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        ea[k] = x[k] - __fma_rn(qa.kd, c[k], qa.cks * pw[k]);
        eb[k] = x[k] - __fma_rn(qb.kd, c[k], qb.cks * pw[N + k]);
    }
    if (slow) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (needs_care(y[k])) ea[k] = residual_careful(qa, c[k], traw[idx[k]], x[k]);
            if (needs_care(y[N + k])) eb[k] = residual_careful(qb, c[k], traw[idx[k]], x[k]);
        }
    }
}

template <class Q>
__device__ __forceinline__ double residual_of(const Q& q, double c, double L, double x,
                                              const double* __restrict__ traw, long i) {
    double e;
    residuals_n<1>(q, &c, &L, &x, traw, &i, &e);
    return e;
}

// One sample of a trial-point pass: only ||x - f(p)||^2.
template <class Q>
__device__ __forceinline__ void accumulate_cost(const Q& q, double c, double L, double x,
                                                const double* __restrict__ traw, long i, double* esq) {
    const double e = residual_of(q, c, L, x, traw, i);
    *esq = __fma_rn(e, e, *esq);
}

// A PAIR of samples (the 16-byte unit every streaming loop loads)
template <class Q>
__device__ __forceinline__ void accumulate_cost_pair(const Q& q, double2 c, double2 L, double2 x,
                                                     const double* __restrict__ traw, long i0, double* esq) {
    const double cc[2] = {c.x, c.y}, ll[2] = {L.x, L.y}, xx[2] = {x.x, x.y};
    const long idx[2] = {i0, i0 + 1};
    double e[2];
    residuals_n<2>(q, cc, ll, xx, traw, idx, e);
    *esq = __fma_rn(e[0], e[0], *esq);
    *esq = __fma_rn(e[1], e[1], *esq);
}

// two pairs at once: four independent exp chains
template <class Q>
__device__ __forceinline__ void accumulate_cost_2pairs(const Q& q, double2 c, double2 L, double2 x, long i0, double2 d,
                                                       double2 M, double2 y, long j0, const double* __restrict__ traw,
                                                       double* esq_a, double* esq_b) {
    const double cc[4] = {c.x, c.y, d.x, d.y}, ll[4] = {L.x, L.y, M.x, M.y}, xx[4] = {x.x, x.y, y.x, y.y};
    const long idx[4] = {i0, i0 + 1, j0, j0 + 1};
    double e[4];
    residuals_n<4>(q, cc, ll, xx, traw, idx, e);
    *esq_a = __fma_rn(e[0], e[0], *esq_a);
    *esq_a = __fma_rn(e[1], e[1], *esq_a);
    *esq_b = __fma_rn(e[2], e[2], *esq_b);
    *esq_b = __fma_rn(e[3], e[3], *esq_b);
}

template <int JAC>
__device__ __forceinline__ void accumulate_jac_pair(const PassParams& q, const PassParams& cold, double2 c, double2 L,
                                                    double2 x, const double* __restrict__ traw, long i0, double* acc) {
    const double cc[2] = {c.x, c.y}, ll[2] = {L.x, L.y}, xx[2] = {x.x, x.y};
    const long idx[2] = {i0, i0 + 1};
    accumulate_jac_n<JAC, 2>(q, cold, cc, ll, xx, traw, idx, acc);
}
#endif  // __CUDACC__

}  // namespace brdfgpu
