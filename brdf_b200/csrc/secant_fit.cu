// secant_fit.cu -- dlevmar_dif on resident samples: unconstrained LM with a SECANT Jacobian
// (reference: levmar/lm_core.c:438-842; the commented-out alternative at brdfdata.cpp:1059,1120).
//
// Unlike the box-constrained driver, dlevmar_dif keeps the n x m Jacobian as persistent state: it is
// rebuilt by finite differences only when the damping grew past nu = 16 or after K = max(m, 10)
// updates (lm_core.c:578); in between every trial step corrects it with Broyden's rank-one update
// (lm_core.c:759-769)
//     J_i += ((f(p + Dp)_i - f(p)_i - J_i . Dp) / ||Dp||^2) Dp .
// So here the Jacobian lives in HBM as three fp64 columns next to hx = f(p), and an LM iteration is
//     S1  k_secant_fd      (after a rebuild)  read c, L, x (24 B)        -> write J, hx (32 B), sums
//     K3  k_cost           trial point        read c, L, x (24 B)        -> ||x - f(p+Dp)||^2
//     S2  k_secant_update  Broyden + sums     read c, L, x, hx, J (56 B) -> write J [, hx] (24/32 B), sums
// each a streaming pass with the same fp64 accumulators and cross-CTA reduction as K2.  When the next
// iteration rebuilds the Jacobian anyway, the rank-one update is dead work and S2 is skipped.  The
// scalar control flow below follows lm_core.c line by line; with a communicator attached the sums of
// every pass are all-reduced, so several GPUs fit one problem.
#include <cstring>

#include "reduce.cuh"

namespace brdfgpu {

struct SecantState {
    double *j0, *j1, *j2, *hx;
};

// per-sample model value and difference-Jacobian row at p (fast path: exp; careful path: pow)
template <int JAC>
__device__ __forceinline__ void model_row(const PassParams& q, double c, double L, double x, const double* __restrict__ traw,
                                          long i, double& hx, double& r0, double& r1, double& r2) {
    constexpr int NE = (JAC == kJacCentral) ? 3 : 2;
    double y[NE], pw[NE];
    y[0] = q.n * L; y[1] = q.n_hi * L;
    if (NE == 3) y[2] = q.n_lo * L;
    bool slow = needs_care(y[0]) || needs_care(y[1]);
    if (NE == 3) slow = slow || needs_care(y[2]);
    if (slow) {
        double o[4];
        jac_terms_careful<JAC>(q, c, traw[i], x, o);
        hx = x - o[0]; r0 = o[1]; r1 = o[2]; r2 = o[3];
        return;
    }
    exp_core_n<NE>(y, pw);
    hx = __fma_rn(q.kd, c, q.cks * pw[0]);
    r0 = q.g0 * c;
    r1 = ((q.model == 1) ? q.g1 : q.g1 * q.coef) * pw[0];
    r2 = (JAC == kJacCentral) ? __fma_rn(q.a_hi, pw[1], -(q.a_lo * pw[2])) : __fma_rn(q.a_hi, pw[1], -(q.a_lo * pw[0]));
}

// S1: difference Jacobian at p, stored; hx = f(p) stored; J^T J, J^T e, ||e||^2 accumulated
template <int JAC>
__global__ void __launch_bounds__(kPassThreads, 2) k_secant_fd(SampleView v, PassParams q, SecantState st, double* partials,
                                                             unsigned* ticket, Publish pub) {
    __shared__ double red[(kPassThreads / 32) * NACC];
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    // sample PAIRS: 16-byte loads of c, L, x and 16-byte stores of the three Jacobian columns and hx
    const long nth = (long)gridDim.x * blockDim.x, npair = v.n >> 1;
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(v.c);
    const double2* __restrict__ l2 = reinterpret_cast<const double2*>(v.L);
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(v.x);
    double2 *j0 = reinterpret_cast<double2*>(st.j0), *j1 = reinterpret_cast<double2*>(st.j1),
            *j2 = reinterpret_cast<double2*>(st.j2), *h2 = reinterpret_cast<double2*>(st.hx);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += nth) {
        const double2 c = __ldg(c2 + i), L = __ldg(l2 + i), x = __ldg(x2 + i);
        double2 hx, r0, r1, r2;
        model_row<JAC>(q, c.x, L.x, x.x, v.traw, 2 * i, hx.x, r0.x, r1.x, r2.x);
        model_row<JAC>(q, c.y, L.y, x.y, v.traw, 2 * i + 1, hx.y, r0.y, r1.y, r2.y);
        j0[i] = r0; j1[i] = r1; j2[i] = r2; h2[i] = hx;
        accumulate_normal(r0.x, r1.x, r2.x, x.x - hx.x, acc);
        accumulate_normal(r0.y, r1.y, r2.y, x.y - hx.y, acc);
    }
    if ((v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long i = v.n - 1;
        const double c = v.c[i], L = v.L[i], x = v.x[i];
        double hx, r0, r1, r2;
        model_row<JAC>(q, c, L, x, v.traw, i, hx, r0, r1, r2);
        st.j0[i] = r0; st.j1[i] = r1; st.j2[i] = r2; st.hx[i] = hx;
        accumulate_normal(r0, r1, r2, x - hx, acc);
    }
    block_reduce_to<NACC>(acc, red, partials + (long)blockIdx.x * NACC);
    last_block_finish<NACC>(partials, ticket, red, pub);
}

// S2: Broyden rank-one update of the stored Jacobian with the trial point pnew = p + Dp
// (lm_core.c:759-769), then the normal equations of the next iteration from the updated J and the
// residual that will be current: x - f(pnew) if the step is accepted (then hx is replaced too,
// lm_core.c:793-797), else x - f(p).
__global__ void __launch_bounds__(kPassThreads, 2) k_secant_update(SampleView v, CostPoint qnew, double d0, double d1,
                                                                 double d2, double dp_l2, int accepted, SecantState st,
                                                                 double* partials, unsigned* ticket, Publish pub) {
    __shared__ double red[(kPassThreads / 32) * NACC];
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    // one sample of the update: tmp = (f(pnew) - hx - sum_l J[l] Dp[l]) / ||Dp||^2, summed from l = 0 as levmar does
    auto update = [&](double c, double L, double x, long i, double hx_old, double& r0, double& r1, double& r2, double& hx_new) {
        hx_new = x - residual_of(qnew, c, L, x, v.traw, i);
        const double dot = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(r0, d0)), __dmul_rn(r1, d1)), __dmul_rn(r2, d2));
        const double tmp = __ddiv_rn(__dsub_rn(__dsub_rn(hx_new, hx_old), dot), dp_l2);
        r0 = __dadd_rn(r0, __dmul_rn(tmp, d0));
        r1 = __dadd_rn(r1, __dmul_rn(tmp, d1));
        r2 = __dadd_rn(r2, __dmul_rn(tmp, d2));
        accumulate_normal(r0, r1, r2, x - (accepted ? hx_new : hx_old), acc);
    };
    const long nth = (long)gridDim.x * blockDim.x, npair = v.n >> 1;
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(v.c);
    const double2* __restrict__ l2 = reinterpret_cast<const double2*>(v.L);
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(v.x);
    double2 *j0 = reinterpret_cast<double2*>(st.j0), *j1 = reinterpret_cast<double2*>(st.j1),
            *j2 = reinterpret_cast<double2*>(st.j2), *h2 = reinterpret_cast<double2*>(st.hx);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += nth) {
        const double2 c = __ldg(c2 + i), L = __ldg(l2 + i), x = __ldg(x2 + i), hx_old = h2[i];
        double2 r0 = j0[i], r1 = j1[i], r2 = j2[i], hx_new;
        update(c.x, L.x, x.x, 2 * i, hx_old.x, r0.x, r1.x, r2.x, hx_new.x);
        update(c.y, L.y, x.y, 2 * i + 1, hx_old.y, r0.y, r1.y, r2.y, hx_new.y);
        j0[i] = r0; j1[i] = r1; j2[i] = r2;
        if (accepted) h2[i] = hx_new;
    }
    if ((v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long i = v.n - 1;
        double r0 = st.j0[i], r1 = st.j1[i], r2 = st.j2[i], hx_new;
        update(v.c[i], v.L[i], v.x[i], i, st.hx[i], r0, r1, r2, hx_new);
        st.j0[i] = r0; st.j1[i] = r1; st.j2[i] = r2;
        if (accepted) st.hx[i] = hx_new;
    }
    block_reduce_to<NACC>(acc, red, partials + (long)blockIdx.x * NACC);
    last_block_finish<NACC>(partials, ticket, red, pub);
}

// hx = f(p) only (accepted step right before a Jacobian rebuild needs nothing else: S1 rewrites hx)
static int secant_blocks(const brdfgpu_ctx* ctx, long n) {
    long want = ((n >> 1) + kPassThreads - 1) / kPassThreads;
    const long cap = (long)ctx->sm_count * 2;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

static void unpack_sums(const double* r, double* JtJ, double* Jte) {
    JtJ[0] = r[A00]; JtJ[1] = r[A01]; JtJ[2] = r[A02];
    JtJ[3] = r[A01]; JtJ[4] = r[A11]; JtJ[5] = r[A12];
    JtJ[6] = r[A02]; JtJ[7] = r[A12]; JtJ[8] = r[A22];
    Jte[0] = r[G0]; Jte[1] = r[G1]; Jte[2] = r[G2];
}

int global_fit_secant(brdfgpu_ctx* ctx, brdfgpu_samples* s, double* p, int m, int itmax, const double* opts, double* info,
                      double* covar) {
    if (m != 3) {
        set_error(ctx, "the BRDF models have exactly 3 parameters (kd, ks, n)");
        return BRDFGPU_LM_ERROR;
    }
    if (s->n < m) {  // lm_core.c:481-484
        fprintf(stderr, "brdfgpu dlevmar_dif: cannot solve a problem with fewer measurements [%ld] than unknowns [%d]\n", s->n, m);
        return BRDFGPU_LM_ERROR;
    }
    if (!s->jac || s->jac_capacity < s->n) {  // J (3 columns) + hx, allocated on first use and kept with the sample set
        cudaFree(s->jac);
        s->jac = nullptr;
        const size_t cap = (size_t)(s->n + (s->n & 1));
        BG_CUDA_OK(ctx, cudaMalloc(&s->jac, sizeof(double) * 4 * cap));
        s->jac_capacity = s->n;
    }
    const size_t stride = (size_t)(s->n + (s->n & 1));
    SecantState st{s->jac, s->jac + stride, s->jac + 2 * stride, s->jac + 3 * stride};
    const SampleView v = view_of(s);
    const bool pub_ok = ctx->nranks == 1;
    const int blocks = secant_blocks(ctx, s->n);

    const LmOptions o = lm_options(opts, itmax);
    double delta = opts ? opts[4] : kDiffDelta;  // lm_core.c:524-531
    int jkind = kJacForward;
    if (delta < 0.0) { delta = -delta; jkind = kJacCentral; }
    const int K = m >= 10 ? m : 10;  // lm_core.c:463

    double JtJ[9] = {0}, Jte[3] = {0}, Dp[3], diag[3] = {0}, pDp[3];
    double mu = 0.0, ginf = 0.0, tmp, e_cur, e_new, e_init, p_L2 = 0.0, Dp_L2 = DBL_MAX, dF, dL;
    int k, stop = 0, nu, nfev, njap = 0, nlss = 0;
    int updjac = 0, updp = 1, newjac = 0;
    unsigned long long n_updates = 0, n_updates_accepted = 0;  // Broyden passes launched (fit_stats)
    bool have_sums = false;  // JtJ / Jte of the current J and e are already on the host (from S1 / S2)

    auto publish = [&](void) { return pub_ok ? Publish{ctx->d_result, ctx->h_result_dev, ctx->h_seq_dev, ++ctx->seq}
                                             : Publish{ctx->d_result, nullptr, nullptr, 0}; };
    auto cost_at = [&](const double* q, double* out) -> int {
        if (launch_cost(ctx, s, q, pub_ok) != 0 || fetch_result(ctx, 1, pub_ok) != 0) return BRDFGPU_LM_ERROR;
        *out = ctx->h_result[0];
        return 0;
    };

    if (cost_at(p, &e_cur) != 0) return BRDFGPU_LM_ERROR;  // lm_core.c:545-551
    nfev = 1;
    e_init = e_cur;
    if (!lm_finite(e_cur)) stop = 7;
    nu = 20;  // forces a difference Jacobian on entry (lm_core.c:553)

    for (k = 0; k < itmax && !stop; ++k) {
        if (e_cur <= o.eps3) { stop = 6; break; }

        if ((updp && nu > 16) || updjac == K) {  // lm_core.c:578-603: rebuild J by differences
            const PassParams q = make_pass_params(p, s->model, delta, jkind);
            const Publish pb = publish();
            if (jkind == kJacForward) k_secant_fd<kJacForward><<<blocks, kPassThreads, 0, ctx->stream>>>(v, q, st, ctx->d_partials, ctx->d_sync, pb);
            else k_secant_fd<kJacCentral><<<blocks, kPassThreads, 0, ctx->stream>>>(v, q, st, ctx->d_partials, ctx->d_sync, pb);
            ++ctx->launches;
            BG_CUDA_OK(ctx, cudaGetLastError());
            if (fetch_result(ctx, NACC, pub_ok) != 0) return BRDFGPU_LM_ERROR;
            ++njap;
            nfev += (jkind == kJacForward) ? m : 2 * m;
            nu = 2; updjac = 0; updp = 0; newjac = 1;
            have_sums = true;
        }

        if (newjac) {  // lm_core.c:605-670: J^T J, J^T e from the current J and e
            newjac = 0;
            if (!have_sums) {
                set_error(ctx, "secant fit: internal error (normal equations requested without a pass)");
                return BRDFGPU_LM_ERROR;
            }
            unpack_sums(ctx->h_result, JtJ, Jte);
            have_sums = false;
            p_L2 = ginf = 0.0;
            for (int i = 0; i < m; ++i) {
                if (ginf < (tmp = lm_abs(Jte[i]))) ginf = tmp;
                diag[i] = JtJ[i * m + i];
                p_L2 += p[i] * p[i];
            }
        }

        if (ginf <= o.eps1) { Dp_L2 = 0.0; stop = 1; break; }  // lm_core.c:683

        if (k == 0) {  // lm_core.c:690-696
            tmp = -DBL_MAX;
            for (int i = 0; i < m; ++i)
                if (diag[i] > tmp) tmp = diag[i];
            mu = o.tau * tmp;
        }

        for (int i = 0; i < m; ++i) JtJ[i * m + i] += mu;  // lm_core.c:700-701
        ++nlss;
        if (solve_lu<3>(JtJ, Jte, Dp, m)) {
            Dp_L2 = 0.0;
            for (int i = 0; i < m; ++i) {
                pDp[i] = p[i] + (tmp = Dp[i]);
                Dp_L2 += tmp * tmp;
            }
            if (Dp_L2 <= o.eps2_sq * p_L2) { stop = 2; break; }
            if (Dp_L2 >= (p_L2 + o.eps2) / (kEpsilon * kEpsilon)) { stop = 4; break; }

            if (cost_at(pDp, &e_new) != 0) return BRDFGPU_LM_ERROR;  // lm_core.c:738-745
            ++nfev;
            if (!lm_finite(e_new)) { stop = 7; break; }

            dF = e_cur - e_new;
            dL = 0.0;
            for (int i = 0; i < m; ++i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);
            const bool accepted = dL > 0.0 && dF > 0.0;          // lm_core.c:774
            const bool broyden = updp || dF > 0.0;               // lm_core.c:759
            if (broyden) {
                ++updjac;
                newjac = 1;
                // does the next iteration rebuild J anyway?  Then the rank-one update is dead work;
                // hx (and e) at an accepted point are rewritten by that rebuild.
                const int updp_next = accepted ? 1 : updp, nu_next = accepted ? 2 : nu * 2;
                const bool rebuild_next = (updp_next && nu_next > 16) || updjac == K;
                const bool last_iteration = k + 1 >= itmax;
                if (!rebuild_next && !last_iteration) {
                    const CostPoint qn = make_cost_point(pDp, s->model);
                    k_secant_update<<<blocks, kPassThreads, 0, ctx->stream>>>(v, qn, Dp[0], Dp[1], Dp[2], Dp_L2, accepted ? 1 : 0, st,
                                                                             ctx->d_partials, ctx->d_sync, publish());
                    ++ctx->launches;
                    ++n_updates;
                    n_updates_accepted += accepted ? 1 : 0;
                    BG_CUDA_OK(ctx, cudaGetLastError());
                    if (fetch_result(ctx, NACC, pub_ok) != 0) return BRDFGPU_LM_ERROR;
                    have_sums = true;
                } else if (!rebuild_next) {
                    newjac = 0;  // nothing will consume it
                }
            }

            if (accepted) {  // lm_core.c:774-801
                tmp = (2.0 * dF / dL - 1.0);
                tmp = 1.0 - tmp * tmp * tmp;
                mu = mu * ((tmp >= kOneThird) ? tmp : kOneThird);
                nu = 2;
                for (int i = 0; i < m; ++i) p[i] = pDp[i];
                e_cur = e_new;
                updp = 1;
                continue;
            }
        }

        // step rejected or singular system (lm_core.c:806-817)
        mu *= nu;
        const int nu2 = (int)((unsigned)nu << 1);
        if (nu2 <= nu) { stop = 5; break; }
        nu = nu2;
        for (int i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
    }
    if (k >= itmax) stop = 3;
    for (int i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
    if (info) {
        const LmCounters c{nfev, njap, nlss};
        lm_fill_info<3>(info, JtJ, m, e_init, e_cur, ginf, Dp_L2, mu, k, stop, c);
    }
    if (covar) {
        long n_all = s->n;
        if (ctx->nranks > 1) {
            double cnt = (double)s->n;
            BG_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_result, &cnt, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            if (comm_allreduce_device(ctx, ctx->d_result, 1) != 0) return BRDFGPU_LM_ERROR;
            BG_CUDA_OK(ctx, cudaMemcpyAsync(&cnt, ctx->d_result, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            n_all = (long)cnt;
        }
        lm_covar<3>(JtJ, covar, e_cur, m, n_all);
    }
    ctx->fit_stats[0] = (unsigned long long)njap; ctx->fit_stats[1] = (unsigned long long)(nfev - njap * ((jkind == kJacForward) ? m : 2 * m));
    ctx->fit_stats[2] = ctx->fit_stats[1];
    for (int i = 3; i < 22; ++i) ctx->fit_stats[i] = 0;
    ctx->fit_stats[22] = n_updates; ctx->fit_stats[23] = n_updates_accepted;
    return (stop != 4 && stop != 7) ? k : BRDFGPU_LM_ERROR;
}

}  // namespace brdfgpu
