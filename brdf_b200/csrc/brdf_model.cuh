// brdf_model.cuh -- per-sample BRDF model arithmetic shared by every fit kernel.
//
// Reference callback (brdfdata.cpp:969-989):
//     Blinn-Phong (model 1): hx = p0*cosphi + p1*pow(costhetadash, p2)
//     Phong       (model 0): hx = p0*cosphi + ((p2+2)/2*pi)*p1*pow(costheta, p2)
// and levmar's difference Jacobian around it (misc_core.c:137-211).
//
// Device data layout (24 B per sample per pass): c = cosphi, L = log(t) with t the model's cosine,
// x = measurement.  t**n is evaluated as exp(n*L): one exp per evaluation point instead of a pow,
// the log being paid once when the sample set is made resident.  L is NaN for samples whose t is
// negative or NaN; those (and every sample when n is not finite) take the slow path through the
// raw cosine and pow(), so libm's sign/integer special cases survive:
//     pow(t<0, integer n) finite, pow(t<0, non-integer) NaN, pow(0, n>0) = 0, pow(x, 0) = 1.
// Accuracy: |exp(n*L) - pow(t,n)| <= ~|n*L| * 2^-52 relative (< 2e-13 before exp overflows),
// far inside the 1e-6 residual gate.  FMA contraction is allowed in the fit kernels (same gate).
#pragma once

#include "lm_engine.cuh"

namespace brdfgpu {

constexpr double kPi = 3.1415926535897932384626433832795;  // CV_PI, brdfdata.cpp:981

enum JacMode { kJacForward = 0, kJacCentral = 1, kJacAnalytic = 2 };

// Everything that depends on p only, computed once per pass by every thread (uniform).
struct PassParams {
    double kd, ks, n, coef;  // coef = 1 (Blinn-Phong) or (n+2)/2*pi (Phong)
    // difference steps d_j = max(|1e-4 p_j|, delta) and 1/d_j (or 0.5/d_j central), misc_core.c:154-167,206
    double inv[3];
    double kd_hi, ks_hi, n_hi, coef_hi;  // p_j + d_j
    double kd_lo, ks_lo, n_lo, coef_lo;  // p_j - d_j (central only)
    double dcoef;                        // d coef / d n (analytic, Phong: pi/2)
    int slow_all;                        // exponent not finite somewhere: every sample through pow()
};

BG_HDI double model_coef(int model, double n) { return model == 1 ? 1.0 : ((n + 2.0) / 2.0 * kPi); }

BG_HDI PassParams make_pass_params(const double* p, int model, double delta, int jac_mode) {
    PassParams q;
    q.kd = p[0]; q.ks = p[1]; q.n = p[2];
    q.coef = model_coef(model, q.n);
    q.dcoef = model == 1 ? 0.0 : (kPi / 2.0);
    double d[3];
    for (int j = 0; j < 3; ++j) {
        double dj = 1E-04 * p[j];
        dj = lm_abs(dj);
        if (dj < delta) dj = delta;
        d[j] = dj;
        q.inv[j] = (jac_mode == kJacCentral ? 0.5 : 1.0) / dj;
    }
    q.kd_hi = p[0] + d[0]; q.ks_hi = p[1] + d[1]; q.n_hi = p[2] + d[2];
    q.kd_lo = p[0] - d[0]; q.ks_lo = p[1] - d[1]; q.n_lo = p[2] - d[2];
    q.coef_hi = model_coef(model, q.n_hi);
    q.coef_lo = model_coef(model, q.n_lo);
    q.slow_all = !(lm_finite(q.n) && lm_finite(q.n_hi) && lm_finite(q.n_lo));
    return q;
}

#ifdef __CUDACC__
// t**n for one sample.  L = log t (NaN => use traw).
__device__ __forceinline__ double pow_sample(double n, double L, const double* __restrict__ traw, long i,
                                             int slow_all) {
    if (slow_all || L != L) return pow(traw[i], n);
    if (n == 0.0) return 1.0;
    return exp(n * L);
}

// log once per sample when a set becomes resident; NaN marks "go through pow()".
__device__ __forceinline__ double log_or_flag(double t) {
    if (t >= 0.0) return log(t);  // t == 0 -> -inf, t == +inf -> +inf
    return __longlong_as_double(0x7ff8000000000000LL);
}

// Accumulator layout of one evaluation: JtJ upper triangle, Jte, ||e||^2, #non-finite residuals.
enum { A00 = 0, A01, A02, A11, A12, A22, G0, G1, G2, ESQ, NBAD, NACC };

// Prediction hx for one sample.
__device__ __forceinline__ double model_eval(const PassParams& q, double c, double pw) {
    return q.kd * c + (q.coef * q.ks) * pw;
}

// One sample of a fused residual + Jacobian + normal-equation pass.
template <int JAC>
__device__ __forceinline__ void accumulate_jac(const PassParams& q, double c, double L, double x,
                                               const double* __restrict__ traw, long i, double* acc) {
    const double pw = pow_sample(q.n, L, traw, i, q.slow_all);
    const double cks = q.coef * q.ks;
    const double hx = q.kd * c + cks * pw;
    const double e = x - hx;
    double j0, j1, j2;
    if (JAC == kJacForward) {
        // jac[i][j] = (f(p + d_j e_j) - f(p)) * (1/d_j), misc_core.c:160-170
        const double pw_hi = pow_sample(q.n_hi, L, traw, i, q.slow_all);
        j0 = ((q.kd_hi * c + cks * pw) - hx) * q.inv[0];
        j1 = ((q.kd * c + (q.coef * q.ks_hi) * pw) - hx) * q.inv[1];
        j2 = ((q.kd * c + (q.coef_hi * q.ks) * pw_hi) - hx) * q.inv[2];
    } else if (JAC == kJacCentral) {
        // jac[i][j] = (f(p + d_j e_j) - f(p - d_j e_j)) * (0.5/d_j), misc_core.c:198-209
        const double pw_hi = pow_sample(q.n_hi, L, traw, i, q.slow_all);
        const double pw_lo = pow_sample(q.n_lo, L, traw, i, q.slow_all);
        j0 = ((q.kd_hi * c + cks * pw) - (q.kd_lo * c + cks * pw)) * q.inv[0];
        j1 = ((q.kd * c + (q.coef * q.ks_hi) * pw) - (q.kd * c + (q.coef * q.ks_lo) * pw)) * q.inv[1];
        j2 = ((q.kd * c + (q.coef_hi * q.ks) * pw_hi) - (q.kd * c + (q.coef_lo * q.ks) * pw_lo)) * q.inv[2];
    } else {
        // exact partials: d/dkd = c, d/dks = coef t^n, d/dn = ks t^n (dcoef + coef ln t)
        j0 = c;
        j1 = q.coef * pw;
        double lt = L;
        if (q.slow_all || L != L) lt = log(traw[i]);
        j2 = q.ks * pw * (q.dcoef + q.coef * lt);
    }
    acc[A00] += j0 * j0; acc[A01] += j0 * j1; acc[A02] += j0 * j2;
    acc[A11] += j1 * j1; acc[A12] += j1 * j2; acc[A22] += j2 * j2;
    acc[G0] += j0 * e; acc[G1] += j1 * e; acc[G2] += j2 * e;
    acc[ESQ] += e * e;
    acc[NBAD] += lm_finite(e) ? 0.0 : 1.0;
}

// One sample of a trial-point pass: only ||x - f(p)||^2 and the non-finite count.
__device__ __forceinline__ void accumulate_cost(const PassParams& q, double c, double L, double x,
                                                const double* __restrict__ traw, long i, double* acc2) {
    const double pw = pow_sample(q.n, L, traw, i, q.slow_all);
    const double e = x - (q.kd * c + (q.coef * q.ks) * pw);
    acc2[0] += e * e;
    acc2[1] += lm_finite(e) ? 0.0 : 1.0;
}
#endif  // __CUDACC__

}  // namespace brdfgpu
