// lm_machine.cuh -- lm_bc_der (lm_engine.cuh) as a RESUMABLE state machine.
//
// Same algorithm, same arithmetic, same order of operations as lm_bc_der<MM> -- levmar's
// dlevmar_bc_der (lmbc_core.c:369-1022) with its line search (:179-337) and projected-gradient
// walk (:871-946) -- but instead of calling an evaluator from inside the control flow, the machine
// stops whenever it needs a function value or a Jacobian and says so:
//
//     m.start(...);                              // -> m.want == kWantCost, point in m.q
//     while (m.want != kWantNothing) {
//         if (m.want == kWantCost) m.feed_cost(||x - f(m.q)||^2, some_residual_nonfinite);
//         else                     m.feed_jac(JtJ, Jte);          // at m.q
//     }                                          // -> m.p, m.info(), m.ret()
//
// Why: in the batched kernel one THREAD owns one fit.  With the straight-line engine every call
// site of the evaluator is a different piece of code, so lanes of a warp that are at different
// places of the algorithm (one in the line search, one in the projected-gradient walk, ...) could
// never evaluate together.  With the machine the warp alternates between ONE evaluation section
// that all lanes execute in lock step and one short, divergent control section.
//
// Pinned bit-for-bit against the reference's levmar on its demo problems by
// tests/test_lm_engine_host.py (through brdfgpu_lm_bc_machine), like lm_bc_der itself.
// Not supported here (use lm_bc_der): diagonal scaling, covariance.
#pragma once

#include "lm_engine.cuh"

namespace brdfgpu {

enum { kWantNothing = 0, kWantCost = 1, kWantJac = 2 };

template <int MM>
struct BcMachine {
    // ---- problem ----
    int m;
    LmOptions o;
    const double* lb;  // may be nullptr
    const double* ub;  // may be nullptr
    // ---- request ----
    int want;
    double q[MM];  // point the pending evaluation is asked at
    // ---- state of lm_bc_der ----
    double p[MM], JtJ[MM * MM], Jte[MM], Dp[MM], diag[MM], pDp[MM];
    double mu, ginf, t, t0, e_cur, e_new, e_init, p_L2, Dp_L2;
    int k, stop, nu, gprevtaken;
    LmCounters cnt;
    // ---- state of the line search ----
    double ls_fc, ls_slp, ls_rmnlmb, ls_lambda, ls_tlmbda, ls_plmbda, ls_pfpls;
    double trial_pt[MM], trial_e;
    bool ls_firstback, trial_known;
    int ls_it;
    // ---- where to resume ----
    enum Pc { kStart, kTrial, kLineSearch, kProjGrad, kDone };
    int pc;

    BG_HDI Box box() const { return Box{lb, ub}; }

    BG_HDI void ask_cost(const double* pt, int resume_at) {
        LM_FOR(i) q[i] = pt[i];
        want = kWantCost;
        pc = resume_at;
    }

    // p0 is projected onto the box first (lmbc_core.c:516), like the callers of lm_bc_der do
    BG_HDI void start(int m_, const double* p0, const double* lb_, const double* ub_, const LmOptions& o_) {
        m = m_; o = o_; lb = lb_; ub = ub_;
        LM_FOR(i) p[i] = p0[i];
        box_project<MM>(p, box(), m);
        LM_FOR(i) {
            LM_FOR(j) JtJ[i * m + j] = 0.0;
            diag[i] = 0.0; Jte[i] = 0.0; Dp[i] = 0.0; pDp[i] = 0.0;
        }
        mu = ginf = t = t0 = e_new = p_L2 = 0.0;
        Dp_L2 = DBL_MAX;
        k = 0; stop = 0; nu = 2; gprevtaken = 0;
        cnt.nfev = cnt.njev = cnt.nlss = 0;
        trial_known = false;
        ask_cost(p, kStart);
    }

    BG_HDI int ret() const { return (stop != 4 && stop != 7) ? k : kLmError; }
    BG_HDI void fill_info(double* info) const { lm_fill_info<MM>(info, JtJ, m, e_init, e_cur, ginf, Dp_L2, mu, k, stop, cnt); }

    // ------------------------------------------------------------------------------------------
    BG_HDI void finish() {  // `done:` of lm_bc_der
        if (k >= o.itmax) stop = 3;
        LM_FOR(i) JtJ[i * m + i] = diag[i];
        want = kWantNothing;
        pc = kDone;
    }

    // top of the outer loop: either finishes or asks for the Jacobian at p
    BG_HDI void iteration_top() {
        if (!(k < o.itmax && !stop)) { finish(); return; }
        if (e_cur <= o.eps3) { stop = 6; finish(); return; }
        LM_FOR(i) q[i] = p[i];
        want = kWantJac;
    }
    BG_HDI void end_iteration() {  // the body ended through a `break` of the inner loop: ++k, next
        ++k;
        iteration_top();
    }

    // augment, solve, project (the top of the inner for(;;) of lm_bc_der) until a trial point exists
    BG_HDI void solve_and_ask_trial() {
        for (;;) {
            LM_FOR(i) JtJ[i * m + i] += mu;
            const int solved = solve_lu<MM>(JtJ, Jte, Dp, m);
            ++cnt.nlss;
            if (!solved) {  // :788-804
                mu *= nu;
                const int nu2 = (int)((unsigned)nu << 1);
                if (nu2 <= nu) { stop = 5; end_iteration(); return; }
                nu = nu2;
                LM_FOR(i) JtJ[i * m + i] = diag[i];
                continue;
            }
            double tmp;
            LM_FOR(i) pDp[i] = p[i] + Dp[i];
            box_project<MM>(pDp, box(), m);
            Dp_L2 = 0.0;
            LM_FOR(i) {
                Dp[i] = tmp = pDp[i] - p[i];
                Dp_L2 += tmp * tmp;
            }
            if (Dp_L2 <= o.eps2_sq * p_L2) { stop = 2; end_iteration(); return; }
            if (Dp_L2 >= (p_L2 + o.eps2) / (kEpsilon * kEpsilon)) { stop = 4; end_iteration(); return; }
            ask_cost(pDp, kTrial);
            return;
        }
    }

    BG_HDI void feed_jac(const double* JtJ_in, const double* Jte_in) {
        double tmp;
        LM_FOR(i) {
            LM_FOR(j) JtJ[i * m + j] = JtJ_in[i * m + j];
            Jte[i] = Jte_in[i];
        }
        ++cnt.njev;
        int j = 0, numactive = 0;  // :639-646
        p_L2 = ginf = 0.0;
        LM_FOR(i) {
            if (ub && p[i] == ub[i]) { ++numactive; if (Jte[i] > 0.0) ++j; }
            else if (lb && p[i] == lb[i]) { ++numactive; if (Jte[i] < 0.0) ++j; }
            else if (ginf < (tmp = lm_abs(Jte[i]))) ginf = tmp;
            diag[i] = JtJ[i * m + i];
            p_L2 += p[i] * p[i];
        }
        if (j == numactive && ginf <= o.eps1) { Dp_L2 = 0.0; stop = 1; finish(); return; }  // outer `break`: no ++k
        if (k == 0) {  // :666-674
            if (!lb && !ub) {
                tmp = -DBL_MAX;
                LM_FOR(i) if (diag[i] > tmp) tmp = diag[i];
                mu = o.tau * tmp;
            } else {
                mu = 0.5 * o.tau * e_cur;
            }
        }
        solve_and_ask_trial();
    }

    // take the line-search / projected-gradient point (:948-967)
    BG_HDI void take_point() {
        double tmp;
        Dp_L2 = 0.0;
        LM_FOR(i) {
            tmp = pDp[i] - p[i];
            Dp_L2 += tmp * tmp;
        }
        if (Dp_L2 <= o.eps2_sq * p_L2) { stop = 2; end_iteration(); return; }
        LM_FOR(i) p[i] = pDp[i];
        e_cur = e_new;
        end_iteration();
    }

    // ---- projected-gradient walk (:871-946), one candidate per evaluation ----
    BG_HDI void pg_begin() {
        double tmp = 0.0;
        LM_FOR(i) tmp += Jte[i] * Jte[i];
        tmp = sqrt(tmp);
        tmp = 100.0 / (1.0 + tmp);
        t0 = (tmp <= 1.0) ? tmp : 1.0;  // tini = 1
        t = gprevtaken ? t : t0;
        pg_step();
    }
    BG_HDI void pg_step() {
        if (!(t > 1e-18)) {  // tming: nothing found
            gprevtaken = 0;
            end_iteration();
            return;
        }
        double cand[MM];
        LM_FOR(i) cand[i] = p[i] - t * Jte[i];
        box_project<MM>(cand, box(), m);
        ask_cost(cand, kProjGrad);
    }
    BG_HDI void pg_fed(double e, bool bad) {
        const double alpha = 1e-4, beta = 0.9;
        double tmp;
        Dp_L2 = 0.0;
        LM_FOR(i) {
            pDp[i] = q[i];
            Dp[i] = tmp = pDp[i] - p[i];
            Dp_L2 += tmp * tmp;
        }
        e_new = e;
        ++cnt.nfev;
        if (!lm_finite(e_new) && bad) { stop = 7; finish(); return; }  // `goto done`: no ++k
        double gTd = 0.0;
        LM_FOR(i) gTd += Jte[i] * Dp[i];
        if (gprevtaken && e_new <= e_cur + 2.0 * 0.99999 * gTd) {  // starting t too small
            t = t0 * beta;
            gprevtaken = 0;
            pg_step();
            return;
        }
        if (e_new <= e_cur + 2.0 * alpha * gTd) {
            gprevtaken = 1;
            take_point();
            return;
        }
        t *= beta;
        pg_step();
    }

    // ---- line search (lm_line_search, :179-337); step = Dp (shortened in place), g = Jte ----
    BG_HDI void ls_begin() {
        const double steptl = 1e3 * sqrt(DBL_EPSILON);
        double tmp = sqrt(p_L2);
        const double stepmx = 1e3 * ((tmp >= 1.0) ? tmp : 1.0);
        LM_FOR(i) trial_pt[i] = pDp[i];
        trial_known = true;
        trial_e = e_new;
        ls_firstback = true;
        ls_tlmbda = ls_plmbda = ls_pfpls = 0.0;
        ls_fc = e_cur * 0.5;
        double tt = 0.0, sln, rln;
        LM_FOR_REV(i) tt += Dp[i] * Dp[i];
        sln = sqrt(tt);
        if (sln > stepmx) {
            const double scl = stepmx / sln;
            LM_FOR_REV(i) Dp[i] *= scl;
            sln = stepmx;
        }
        ls_slp = rln = 0.0;
        LM_FOR_REV(i) {
            ls_slp += Jte[i] * Dp[i];
            const double a = (lm_abs(p[i]) >= 1.0) ? lm_abs(p[i]) : 1.0;
            const double b = lm_abs(Dp[i]) / a;
            if (rln < b) rln = b;
        }
        ls_rmnlmb = steptl / rln;
        ls_lambda = 1.0;
        ls_it = kLsItMax;
        ls_step();
    }
    BG_HDI void ls_finished(int rc) {
        if (rc != 0 || !lm_finite(e_new)) {
            pg_begin();
        } else {
            gprevtaken = 0;
            take_point();
        }
    }
    BG_HDI void ls_step() {
        for (;;) {
            if (ls_it-- <= 0) { ls_finished(1); return; }
            LM_FOR_REV(i) pDp[i] = p[i] + ls_lambda * Dp[i];
            box_project<MM>(pDp, box(), m);
            bool reuse = trial_known;
            if (reuse) LM_FOR(i) reuse = reuse && pDp[i] == trial_pt[i];
            if (!reuse) {
                ask_cost(pDp, kLineSearch);
                return;
            }
            trial_known = false;
            if (ls_eval(trial_e)) return;
        }
    }
    // consumes one function value of the line search; true when the machine moved on (asked / finished)
    BG_HDI bool ls_eval(double tval) {
        const double alpha = 1e-4;
        ++cnt.nfev;
        const double fpls = 0.5 * tval;
        e_new = tval;
        if (fpls <= ls_fc + ls_slp * alpha * ls_lambda) { ls_finished(0); return true; }
        if (ls_lambda < ls_rmnlmb) { ls_finished(1); return true; }
        if (!lm_finite(fpls)) {
            ls_lambda *= 0.1;
            ls_firstback = true;
        } else {
            if (ls_firstback) {
                ls_tlmbda = -ls_lambda * ls_slp / ((fpls - ls_fc - ls_slp) * 2.0);
                ls_firstback = false;
            } else {
                const double t1 = fpls - ls_fc - ls_lambda * ls_slp;
                const double t2 = ls_pfpls - ls_fc - ls_plmbda * ls_slp;
                const double t3 = 1.0 / (ls_lambda - ls_plmbda);
                const double a3 = 3.0 * t3 * (t1 / (ls_lambda * ls_lambda) - t2 / (ls_plmbda * ls_plmbda));
                const double b = t3 * (t2 * ls_lambda / (ls_plmbda * ls_plmbda) - t1 * ls_plmbda / (ls_lambda * ls_lambda));
                const double disc = b * b - a3 * ls_slp;
                if (disc > b * b)
                    ls_tlmbda = (-b + ((a3 < 0) ? -sqrt(disc) : sqrt(disc))) / a3;
                else
                    ls_tlmbda = (-b + ((a3 < 0) ? sqrt(disc) : -sqrt(disc))) / a3;
                if (ls_tlmbda > ls_lambda * 0.5) ls_tlmbda = ls_lambda * 0.5;
            }
            ls_plmbda = ls_lambda;
            ls_pfpls = fpls;
            if (ls_tlmbda < ls_lambda * 0.1) ls_lambda *= 0.1;
            else ls_lambda = ls_tlmbda;
        }
        return false;
    }

    // ------------------------------------------------------------------------------------------
    BG_HDI void feed_cost(double e, bool bad) {
        const double gamma = 0.99995, rho = 1e-8;
        double tmp;
        switch (pc) {
            case kStart:  // :522-534
                e_cur = e;
                cnt.nfev = 1;
                e_init = e_cur;
                if (!lm_finite(e_cur)) stop = 7;
                iteration_top();
                return;
            case kTrial: {
                e_new = e;
                ++cnt.nfev;
                if (!lm_finite(e_new) && bad) { stop = 7; end_iteration(); return; }  // :748
                if (e_new <= gamma * e_cur) {  // LM step accepted, :753-785
                    double dL = 0.0;
                    LM_FOR(i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);
                    if (dL > 0.0) {
                        const double dF = e_cur - e_new;
                        tmp = (2.0 * dF / dL - 1.0);
                        tmp = 1.0 - tmp * tmp * tmp;
                        mu = mu * ((tmp >= kOneThird) ? tmp : kOneThird);
                    } else {
                        tmp = 0.1 * e_new;
                        mu = (mu >= tmp) ? tmp : mu;
                    }
                    nu = 2;
                    LM_FOR(i) p[i] = pDp[i];
                    e_cur = e_new;
                    gprevtaken = 0;
                    end_iteration();
                    return;
                }
                double gTd = 0.0;  // rejected: descent direction? (:810-816)
                LM_FOR(i) {
                    Jte[i] = -Jte[i];
                    gTd += Jte[i] * Dp[i];
                }
                if (gTd <= -rho * pow(Dp_L2, kPow / 2.0)) ls_begin();
                else pg_begin();
                return;
            }
            case kLineSearch:
                if (!ls_eval(e)) ls_step();
                return;
            case kProjGrad:
                pg_fed(e, bad);
                return;
            default:
                return;
        }
    }
};

}  // namespace brdfgpu
