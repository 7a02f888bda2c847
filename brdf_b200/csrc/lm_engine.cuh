// lm_engine.cuh -- the Levenberg-Marquardt control loop on REDUCED quantities.
//
// levmar (reference: levmar/lmbc_core.c:369-1022, lm_core.c:64-432) keeps n-sized arrays e, hx and
// an n x m Jacobian in host memory and walks them serially.  Here nothing n-sized exists: the
// control code only ever consumes
//     J^T J (m x m), J^T e (m)            from a Jacobian evaluation at p,
//     ||x - f(p')||^2 (+ "some residual is non-finite")   from a trial evaluation,
// which an Evaluator produces by streaming the samples on the GPU.  The same source is compiled
// three times:
//     host      -- HostEval launches one kernel per evaluation (global mode, NCCL all-reduce inside)
//     device    -- GridEval inside one persistent cooperative kernel (global mode, single launch)
//     device    -- GroupEval, one lane group per fit (batched mode)
// so all modes make bit-identical decisions from identical sums.
//
// Evaluator concept:
//     void   jac (const double* p, double* JtJ /*m*m row-major, full*/, double* Jte /*m*/);
//     double cost(const double* p, bool& elems_nonfinite);
//     static constexpr int kCostBatch;   // > 1: also batch_points() / cost_many() / batch_cost() / batch_bad() (PgBatch)
//     static constexpr bool kLanePgWalk; // true: also pg_walk(), the projected-gradient walk one candidate per lane
//     static constexpr bool kSpecJac;    // optional, true: also cost_site() / trial_outcome() / ls_outcome() (below)
//
// Speculative Jacobians (evaluators with kSpecJac).  A point that is evaluated because it may become
// the next iterate -- the LM trial point, a line-search probe, the first projected-gradient candidate
// -- needs ||x - f||^2 now and, if it is taken, J^T J / J^T e at the very same point one step later.
// One fused sweep yields both, so such an evaluator may answer a cost request with a Jacobian sweep
// and keep the sums; jac() at a bit-identical point then costs nothing.  The engine only names the
// SITE of each evaluation and reports what happened to it; whether to speculate is the evaluator's
// business (the persistent kernel predicts "same as last time" per site).  Values, decisions and
// the nfev / njev counters are the same with and without speculation.
#pragma once

#include <cfloat>
#include <cmath>
#include <type_traits>

#ifdef __CUDACC__
#define BG_HD __host__ __device__
#define BG_HDI __host__ __device__ __forceinline__
#else
#define BG_HD
#define BG_HDI inline
#endif

namespace brdfgpu {

// levmar/lmbc.c:35-38, lm.c:35-36, levmar.h:95-101
constexpr double kEpsilon = 1E-12;
constexpr double kOneThird = 0.3333333334;
constexpr int kLsItMax = 150;
constexpr double kPow = 2.1;
constexpr double kInitMu = 1E-03;
constexpr double kStopThresh = 1E-17;
constexpr double kDiffDelta = 1E-06;
constexpr int kLmError = -1;

// misc.h:68 (not fabs: identical -0.0 / NaN behaviour)
BG_HDI double lm_abs(double v) { return (v >= 0.0) ? v : -v; }
BG_HDI bool lm_finite(double v) { return (v - v) == 0.0; }  // false for NaN and +-Inf
// y + a*x with ONE spelling per target (device: fused; host: levmar's two roundings), for points that are formed
// at two places and compared bit for bit (the line-search probe and its announcement)
// BG_LM_LEVMAR_ARITH (the levmar-exact batched kernel, compiled with --fmad=false): levmar's two roundings on the
// device as well, and libm's pow() through its reproduction in glibc_pow.cuh.
BG_HDI double lm_axpy(double a, double x, double y) {
#if defined(__CUDA_ARCH__) && !defined(BG_LM_LEVMAR_ARITH)
    return __fma_rn(a, x, y);
#else
    return y + a * x;
#endif
}
#ifdef BG_LM_LEVMAR_ARITH
#define BG_LM_POW(x, y) ::brdfgpu::glibc_pow((x), (y))
#else
#define BG_LM_POW(x, y) pow((x), (y))
#endif

struct LmOptions {
    double tau, eps1, eps2, eps2_sq, eps3;
    int itmax;
};

BG_HDI LmOptions lm_options(const double* opts, int itmax) {
    LmOptions o;
    if (opts) {  // lmbc_core.c:470-476
        o.tau = opts[0]; o.eps1 = opts[1]; o.eps2 = opts[2]; o.eps2_sq = opts[2] * opts[2]; o.eps3 = opts[3];
    } else {     // :477-483
        o.tau = kInitMu; o.eps1 = kStopThresh; o.eps2 = kStopThresh;
        o.eps2_sq = kStopThresh * kStopThresh; o.eps3 = kStopThresh;
    }
    o.itmax = itmax;
    return o;
}

// Loops over the m parameters are written against the compile-time bound MM with an `i < m` guard
// and unrolled, and nothing below indexes a local array with a run-time value: on the device every
// local array then lives in registers.  (Local memory costs ~350 cycles per dependent access on
// B200 -- profiles/micro/lat.cu -- and this code sits on the critical path of every evaluation.)
#define LM_FOR(i) _Pragma("unroll") for (int i = 0; i < MM; ++i) if (i < m)
#define LM_FOR_REV(i) _Pragma("unroll") for (int i = MM; i-- > 0;) if (i < m)
#define LM_FOR2(i, lo, hi) _Pragma("unroll") for (int i = 0; i < MM; ++i) if (i >= (lo) && i < (hi))

// ---------------------------------------------------------------------------------------------
// m x m solve: Crout LU, implicit row scaling, partial pivoting (Axb_core.c:1196-1270).
// Returns 0 when a row of A is all zero; a zero pivot is replaced by DBL_EPSILON.
// ---------------------------------------------------------------------------------------------
template <int MM>
BG_HDI int lu_factor(double (&a)[MM * MM], int (&perm)[MM], int m) {
    double rowscale[MM];
    int piv = -1;
    bool zero_row = false;
    LM_FOR(i) {
        double big = 0.0;
        LM_FOR(j) {
            const double t = lm_abs(a[i * MM + j]);
            if (t > big) big = t;
        }
        if (big == 0.0) zero_row = true;
        rowscale[i] = 1.0 / big;
    }
    if (zero_row) return 0;
    LM_FOR(j) {
        LM_FOR2(i, 0, j) {
            double s = a[i * MM + j];
            LM_FOR2(k, 0, i) s -= a[i * MM + k] * a[k * MM + j];
            a[i * MM + j] = s;
        }
        double big = 0.0;
        LM_FOR2(i, j, m) {
            double s = a[i * MM + j];
            LM_FOR2(k, 0, j) s -= a[i * MM + k] * a[k * MM + j];
            a[i * MM + j] = s;
            const double t = rowscale[i] * lm_abs(s);
            if (t >= big) { big = t; piv = i; }
        }
        // row interchange j <-> piv, written per candidate row so every index is static.  (piv < j
        // only happens when NaNs froze the pivot search at an earlier column: kept as levmar does)
        LM_FOR(r) {
            if (r != j && piv == r) {
                LM_FOR(k) {
                    const double t = a[r * MM + k];
                    a[r * MM + k] = a[j * MM + k];
                    a[j * MM + k] = t;
                }
                rowscale[r] = rowscale[j];
            }
        }
        perm[j] = piv;
        if (a[j * MM + j] == 0.0) a[j * MM + j] = DBL_EPSILON;
        if (j != m - 1) {
            const double t = 1.0 / a[j * MM + j];
            LM_FOR2(i, j + 1, m) a[i * MM + j] *= t;
        }
    }
    return 1;
}

template <int MM>
BG_HDI void lu_substitute(const double (&a)[MM * MM], const int (&perm)[MM], double (&x)[MM], int m) {
    int first = 0;
    LM_FOR(i) {
        // s = x[perm[i]]; x[perm[i]] = x[i]
        double s = x[i];
        LM_FOR(r) {
            if (r != i && perm[i] == r) {
                s = x[r];
                x[r] = x[i];
            }
        }
        if (first != 0) {
            LM_FOR2(k, 0, i) if (k >= first - 1) s -= a[i * MM + k] * x[k];
        } else if (s != 0.0) {
            first = i + 1;
        }
        x[i] = s;
    }
    LM_FOR_REV(i) {
        double s = x[i];
        LM_FOR2(j, i + 1, m) s -= a[i * MM + j] * x[j];
        x[i] = s / a[i * MM + i];
    }
}

// A, B, x: m x m row-major / m (caller layout, stride m)
template <int MM>
BG_HDI int solve_lu(const double* A, const double* B, double* x, int m) {
    double a[MM * MM], xx[MM];
    int perm[MM];
    LM_FOR(i) {
        LM_FOR(j) a[i * MM + j] = A[i * m + j];
        xx[i] = B[i];
    }
    if (!lu_factor<MM>(a, perm, m)) {
        LM_FOR(i) x[i] = xx[i];  // levmar leaves B's copy in x on failure (never read by the callers)
        return 0;
    }
    lu_substitute<MM>(a, perm, xx, m);
    LM_FOR(i) x[i] = xx[i];
    return 1;
}

// covariance C = sumsq/(n-m) * (JtJ)^-1 through the same LU (misc_core.c:426-591). 0 on failure.
template <int MM>
BG_HDI int lm_covar(const double* JtJ, double* C, double sumsq, int m, long n) {
    double a[MM * MM], x[MM];
    int perm[MM];
    LM_FOR(i) LM_FOR(j) a[i * MM + j] = JtJ[i * m + j];
    if (!lu_factor<MM>(a, perm, m)) return 0;
    for (int l = 0; l < m; ++l) {
        LM_FOR(i) x[i] = (i == l) ? 1.0 : 0.0;
        lu_substitute<MM>(a, perm, x, m);
        LM_FOR(i) C[i * m + l] = x[i];
    }
    const double fact = sumsq / (double)(n - m);
    for (int i = 0; i < m * m; ++i) C[i] *= fact;
    return m;
}

// ---------------------------------------------------------------------------------------------
// box helpers (lmbc_core.c:59-88, 94-142)
// ---------------------------------------------------------------------------------------------
// (levmar's __MEDIAN3 with the same five comparisons, written as selects: in the lane-parallel projected-gradient walks
// every lane projects a different candidate, and branches would diverge)
BG_HDI double lm_median3(double lo, double v, double hi) {
    const double below = (hi >= lo) ? lo : ((hi <= v) ? v : hi);   // lo >= v
    const double above = (hi >= v) ? v : ((hi <= lo) ? lo : hi);   // lo <  v (or unordered)
    return (lo >= v) ? below : above;
}

struct Box {
    const double* lb;  // may be nullptr
    const double* ub;  // may be nullptr
};

template <int MM = 8>
BG_HDI void box_project(double* p, const Box& b, int m) {
    if (!b.lb && !b.ub) return;
    LM_FOR_REV(i) {
        if (b.lb && b.ub) p[i] = lm_median3(b.lb[i], p[i], b.ub[i]);
        else if (b.ub) { if (p[i] > b.ub[i]) p[i] = b.ub[i]; }
        else { if (p[i] < b.lb[i]) p[i] = b.lb[i]; }
    }
}

struct LmCounters {
    int nfev, njev, nlss;
};

// evaluation sites named to evaluators with kSpecJac: the LM trial point, or line-search probe k >= 2
constexpr int kSiteTrial = 0;
constexpr int kSitePgFirst = -1;  // first candidate of a projected-gradient walk (sequential form of the walk)
template <class E, class = void>
struct SpecJac : std::false_type {};
template <class E>
struct SpecJac<E, std::void_t<decltype(E::kSpecJac)>> : std::integral_constant<bool, E::kSpecJac> {};

// Evaluators with `static constexpr bool kSharedState = true` lend the engine a workspace (double* workspace(), at least
// MM * MM + 4 * MM doubles) for J^T J, J^T e, Dp, diag and pDp.  In the batched kernels every lane of a group runs the
// engine on identical values, so that state is the same in all lanes: held once in shared memory instead of in every
// lane's registers it frees ~42 registers per thread (occupancy is what those latency-bound kernels run on).
template <class E, class = void>
struct SharedState : std::false_type {};
template <class E>
struct SharedState<E, std::void_t<decltype(E::kSharedState)>> : std::integral_constant<bool, E::kSharedState> {};

template <class Eval>
BG_HDI double eval_cost_site(Eval& ev, int site, const double* p, bool& bad) {
    if constexpr (SpecJac<Eval>::value) return ev.cost_site(site, p, bad);
    else return ev.cost(p, bad);
}
template <class Eval>
BG_HDI void note_trial_outcome(Eval& ev, bool accepted) {
    if constexpr (SpecJac<Eval>::value) ev.trial_outcome(accepted);
}
template <class Eval>
BG_HDI void note_pg_outcome(Eval& ev, bool took_first_candidate) {
    if constexpr (SpecJac<Eval>::value) ev.pg_outcome(took_first_candidate);
}
template <class Eval>
BG_HDI void note_ls_outcome(Eval& ev, int accepted_probe /* 0: the search failed */) {
    if constexpr (SpecJac<Eval>::value) ev.ls_outcome(accepted_probe);
}
// Before the LM trial point is evaluated: what the iteration evaluates next if the trial is rejected -- the line
// search's probe at lambda = 0.1 (its first backtrack when the step was far too long), announced every time,
// and, only when the evaluator asks for it, the first candidate of the projected-gradient walk (from p, J^T e and
// the step length t, which costs a square root and a division to know this early).
template <class Eval>
BG_HDI void note_probe_hint(Eval& ev, const double* probe) {
    if constexpr (SpecJac<Eval>::value) ev.probe_hint(probe);
}
template <class Eval>
BG_HDI bool wants_candidate_hint(Eval& ev) {
    if constexpr (SpecJac<Eval>::value) return ev.wants_candidate_hint();
    else return false;
}
template <class Eval>
BG_HDI void note_candidate_hint(Eval& ev, const double* p, const double* Jte, double t, const double* lb, const double* ub) {
    if constexpr (SpecJac<Eval>::value) ev.candidate_hint(p, Jte, t, lb, ub);
}
// The line search is about to evaluate its LAST probe (lambda is already below the minimum step): if
// that probe is not accepted the search fails and the projected-gradient walk starts from p with
// gradient g at step length t -- its first candidate is known now, so the evaluator may evaluate it
// (and its Jacobian) in the same sweep as the probe.
template <class Eval>
BG_HDI void note_ls_fallback(Eval& ev, const double* p, const double* g, double t, const double* lb, const double* ub) {
    if constexpr (SpecJac<Eval>::value) ev.ls_fallback(p, g, t, lb, ub);
}

// Calls the evaluator in the caller's (unscaled) coordinates; with diagonal scaling D the control
// loop works on q = D^-1 p and J_q = J_p D (lmbc_core.c:360-366, 555-570), i.e.
// JtJ_ij *= d_i d_j and Jte_i *= d_i -- the scaled sums are formed from the unscaled ones here
// instead of scaling n rows of J.
template <int MM, class Eval>
BG_HDI void eval_jac_scaled(Eval& ev, const double* q, const double* dscl, int m, double* JtJ, double* Jte) {
    if (!dscl) {
        ev.jac(q, JtJ, Jte);
        return;
    }
    double ps[MM];
    LM_FOR_REV(i) ps[i] = q[i] * dscl[i];
    ev.jac(ps, JtJ, Jte);
    LM_FOR(i) {
        Jte[i] *= dscl[i];
        LM_FOR(j) JtJ[i * m + j] *= dscl[i] * dscl[j];
    }
}

template <int MM, class Eval>
BG_HDI double eval_cost_scaled(Eval& ev, const double* q, const double* dscl, int m, bool& bad) {
    if (!dscl) return ev.cost(q, bad);
    double ps[MM];
    LM_FOR_REV(i) ps[i] = q[i] * dscl[i];
    return ev.cost(ps, bad);
}

// the LM trial point: the same products q_i * d_i eval_jac_scaled forms, so a kept Jacobian is found again
template <int MM, class Eval>
BG_HDI double eval_trial_scaled(Eval& ev, const double* q, const double* dscl, int m, bool& bad) {
    if (!dscl) return eval_cost_site(ev, kSiteTrial, q, bad);
    double ps[MM];
    LM_FOR_REV(i) ps[i] = q[i] * dscl[i];
    return eval_cost_site(ev, kSiteTrial, ps, bad);
}

// Candidate points of the projected-gradient walk.  Evaluators with kCostBatch > 1 own the storage
// (shared memory in the persistent kernel) and evaluate up to kCostBatch points per call:
//     double* batch_points();                       // kCostBatch x m, filled by the engine (unscaled)
//     void    cost_many(int cnt, const double* dscl, int m);
//     double  batch_cost(int c);  bool batch_bad(int c);
// Everybody else evaluates one point per call through cost().
template <int MM, class Eval, bool MANY = (Eval::kCostBatch > 1)>
struct PgBatch {
    double pts[MM], e;
    bool b;
    BG_HDI double* points(Eval&) { return pts; }
    BG_HDI void run(Eval& ev, int, const double* dscl, int m, bool first = false) {
        if constexpr (SpecJac<Eval>::value) {  // (evaluators without speculation keep ONE inlined copy of cost())
            if (first && !dscl) {
                e = eval_cost_site(ev, kSitePgFirst, pts, b);
                return;
            }
        }
        e = eval_cost_scaled<MM>(ev, pts, dscl, m, b);
    }
    BG_HDI double cost(Eval&, int) const { return e; }
    BG_HDI bool bad(Eval&, int) const { return b; }
};
template <int MM, class Eval>
struct PgBatch<MM, Eval, true> {
    BG_HDI double* points(Eval& ev) { return ev.batch_points(); }
    BG_HDI void run(Eval& ev, int cnt, const double* dscl, int m, bool = false) { ev.cost_many(cnt, dscl, m); }
    BG_HDI double cost(Eval& ev, int c) const { return ev.batch_cost(c); }
    BG_HDI bool bad(Eval& ev, int c) const { return ev.batch_bad(c); }
};

// info[] of lmbc_core.c:978-991 / lm_core.c:405-418
template <int MM>
BG_HDI void lm_fill_info(double* info, const double* JtJ, int m, double e0, double e, double ginf,
                         double dp2, double mu, int k, int stop, const LmCounters& c) {
    if (!info) return;
    double big = -DBL_MAX;
    LM_FOR(i) if (big < JtJ[i * m + i]) big = JtJ[i * m + i];
    info[0] = e0; info[1] = e; info[2] = ginf; info[3] = dp2; info[4] = mu / big;
    info[5] = (double)k; info[6] = (double)stop; info[7] = (double)c.nfev;
    info[8] = (double)c.njev; info[9] = (double)c.nlss;
}

// ---------------------------------------------------------------------------------------------
// Schnabel/Koontz/Weiss backtracking line search with box projection (lmbc_core.c:179-337).
// `step` may be shortened in place.  Returns iretcd (0 = acceptable point in xnew / fnew).
// ---------------------------------------------------------------------------------------------
template <int MM, class Eval>
BG_HDI int lm_line_search(Eval& ev, int m, const double* xc, double fc, const double* g, double* step,
                          double alpha, double* xnew, double& fnew_sumsq, const Box& box,
                          const double* dscl, double stepmx, double steptl, LmCounters& cnt,
                          const double* known_pt, double known_f, double pg_t) {
    bool firstback = true, bad;
    double sln, slp, rln, rmnlmb, lambda, tlmbda = 0.0, plmbda = 0.0, pfpls = 0.0, fpls, t;

    fc *= 0.5;
    t = 0.0;
    LM_FOR_REV(i) t += step[i] * step[i];
    sln = sqrt(t);
    if (sln > stepmx) {
        const double scl = stepmx / sln;
        LM_FOR_REV(i) step[i] *= scl;
        sln = stepmx;
    }
    slp = rln = 0.0;
    LM_FOR_REV(i) {
        slp += g[i] * step[i];
        const double a = (lm_abs(xc[i]) >= 1.0) ? lm_abs(xc[i]) : 1.0;
        const double b = (a == 1.0) ? lm_abs(step[i]) : lm_abs(step[i]) / a;  // x / 1.0 == x: no division needed
        if (rln < b) rln = b;
    }
    rmnlmb = steptl / rln;
    lambda = 1.0;

    for (int it = kLsItMax; it-- > 0;) {
        LM_FOR_REV(i) xnew[i] = lm_axpy(lambda, step[i], xc[i]);
        box_project<MM>(xnew, box, m);

        // levmar's first probe (lambda = 1) is usually the very point whose rejection started this
        // search -- p + (pDp - p) == pDp exactly for ordinary steps -- and it evaluates it again
        // (:260).  Evaluations are deterministic here, so the known value is reused: one sweep over
        // the samples less, same numbers, and the evaluation is still counted in nfev.
        bool reuse = known_pt != nullptr && !dscl;
        if (reuse) LM_FOR(i) reuse = reuse && xnew[i] == known_pt[i];
        if (reuse) {
            t = known_f;
            known_pt = nullptr;
        } else if (!dscl) {
            if (lambda < rmnlmb) note_ls_fallback(ev, xc, g, pg_t, box.lb, box.ub);
            t = eval_cost_site(ev, kLsItMax - it, xnew, bad);  // probe number 1, 2, ... as the site
        } else {  // :262-266 scales the point in place and back (not an exact round trip)
            LM_FOR_REV(i) xnew[i] *= dscl[i];
            t = ev.cost(xnew, bad);
            LM_FOR_REV(i) xnew[i] /= dscl[i];
        }
        ++cnt.nfev;
        fpls = 0.5 * t;
        fnew_sumsq = t;

        if (fpls <= fc + slp * alpha * lambda) {
            note_ls_outcome(ev, kLsItMax - it);
            return 0;
        }
        if (lambda < rmnlmb) {
            note_ls_outcome(ev, 0);
            return 1;
        }

        if (!lm_finite(fpls)) {
            lambda *= 0.1;
            firstback = true;
        } else {
            if (firstback) {
                tlmbda = -lambda * slp / ((fpls - fc - slp) * 2.0);
                firstback = false;
            } else {
                const double t1 = fpls - fc - lambda * slp;
                const double t2 = pfpls - fc - plmbda * slp;
                const double t3 = 1.0 / (lambda - plmbda);
                const double a3 = 3.0 * t3 * (t1 / (lambda * lambda) - t2 / (plmbda * plmbda));
                const double b = t3 * (t2 * lambda / (plmbda * plmbda) - t1 * plmbda / (lambda * lambda));
                const double disc = b * b - a3 * slp;
                if (disc > b * b)
                    tlmbda = (-b + ((a3 < 0) ? -sqrt(disc) : sqrt(disc))) / a3;
                else
                    tlmbda = (-b + ((a3 < 0) ? sqrt(disc) : -sqrt(disc))) / a3;
                if (tlmbda > lambda * 0.5) tlmbda = lambda * 0.5;
            }
            plmbda = lambda;
            pfpls = fpls;
            if (tlmbda < lambda * 0.1) lambda *= 0.1;
            else lambda = tlmbda;
        }
    }
    note_ls_outcome(ev, 0);
    return 1;
}

// First step length of a projected-gradient walk, t0 = min(100 / (1 + ||g||), tini) (lmbc_core.c:876-879).
// 100 / (1 + ||g||) exceeds tini = 1 whenever ||g||^2 < 9000 (then 1 + ||g|| < 96), so the square root and the
// division -- ~600 cycles on the control warp -- are only paid for large gradients; same value in every case.
template <int MM>
struct MMTag {};
template <int MM>
BG_HDI double lm_pg_first_step(const double* g, int m, double tini, MMTag<MM>) {
    double tmp = 0.0;
    LM_FOR(i) tmp += g[i] * g[i];
    if (tmp < 9000.0 && tini == 1.0) return tini;  // (NaN falls through to the literal formula)
    tmp = sqrt(tmp);
    tmp = 100.0 / (1.0 + tmp);
    return (tmp <= tini) ? tmp : tini;
}

// =============================================================================================
// Box-constrained LM (projected LM step / line search / projected gradient), lmbc_core.c:369-1022.
// p: in/out (m).  lb/ub/dscl may be nullptr.  JtJ_out (m*m, may be nullptr) receives the
// unaugmented normal matrix at exit (for the covariance).  Returns #iterations or kLmError.
// The start must already have been range-checked by the caller (n >= m, lb <= ub, dscl > 0).
// `lbs/ubs`: when dscl is given the caller passes bounds already divided by dscl (:536-540).
// =============================================================================================
template <int MM, class Eval>
BG_HDI int lm_bc_der(Eval& ev, int m, double* p, const double* lb, const double* ub, const double* dscl,
                     const LmOptions& o, double* info, double* JtJ_out) {
    const double alpha = 1e-4, beta = 0.9, gamma = 0.99995, rho = 1e-8;
    const double tini = 1.0, tming = 1e-18;
    constexpr bool kShared = SharedState<Eval>::value;
    double local_state[kShared ? 1 : MM * MM + 4 * MM];
    double* const ws = [&]() -> double* {
        if constexpr (kShared) return ev.workspace();
        else return local_state;
    }();
    double *const JtJ = ws, *const Jte = ws + MM * MM, *const Dp = Jte + MM, *const diag = Dp + MM, *const pDp = diag + MM;
    double mu = 0.0, ginf = 0.0, t = 0.0, t0, tmp;
    double e_cur, e_new = 0.0, e_init, p_L2 = 0.0, Dp_L2 = DBL_MAX, dF, dL, gTd;
    int k, stop = 0, nu = 2, gprevtaken = 0, numactive, j;
    bool bad = false;
    LmCounters cnt = {0, 0, 0};
    const Box box = {lb, ub};

    LM_FOR(i) {
        LM_FOR(jj) JtJ[i * m + jj] = 0.0;
        diag[i] = 0.0;
    }

    // e = x - f(p) at the (projected) start, :522-534.  p is still in caller coordinates here.
    e_cur = ev.cost(p, bad);
    cnt.nfev = 1;
    e_init = e_cur;
    if (!lm_finite(e_cur)) stop = 7;

    if (dscl) LM_FOR_REV(i) p[i] /= dscl[i];

    for (k = 0; k < o.itmax && !stop; ++k) {
        if (e_cur <= o.eps3) { stop = 6; break; }

        eval_jac_scaled<MM>(ev, p, dscl, m, JtJ, Jte);
        ++cnt.njev;

        // ||J^T e||_inf over free variables, ||p||^2 (:639-646)
        j = numactive = 0;
        p_L2 = ginf = 0.0;
        LM_FOR(i) {
            if (ub && p[i] == ub[i]) { ++numactive; if (Jte[i] > 0.0) ++j; }
            else if (lb && p[i] == lb[i]) { ++numactive; if (Jte[i] < 0.0) ++j; }
            else if (ginf < (tmp = lm_abs(Jte[i]))) ginf = tmp;
            diag[i] = JtJ[i * m + i];
            p_L2 += p[i] * p[i];
        }
        if (j == numactive && ginf <= o.eps1) { Dp_L2 = 0.0; stop = 1; break; }

        if (k == 0) {  // :666-674
            if (!lb && !ub) {
                tmp = -DBL_MAX;
                LM_FOR(i) if (diag[i] > tmp) tmp = diag[i];
                mu = o.tau * tmp;
            } else {
                mu = 0.5 * o.tau * e_cur;  // Kanzow's starting mu
            }
        }

        for (;;) {
            bool use_pg = false;

            LM_FOR(i) JtJ[i * m + i] += mu;
            const int solved = solve_lu<MM>(JtJ, Jte, Dp, m);
            ++cnt.nlss;

            if (!solved) {  // :788-804
                mu *= nu;
                const int nu2 = (int)((unsigned)nu << 1);
                if (nu2 <= nu) { stop = 5; break; }
                nu = nu2;
                LM_FOR(i) JtJ[i * m + i] = diag[i];
                continue;
            }

            LM_FOR(i) pDp[i] = p[i] + Dp[i];
            box_project<MM>(pDp, box, m);
            Dp_L2 = 0.0;
            LM_FOR(i) {
                Dp[i] = tmp = pDp[i] - p[i];
                Dp_L2 += tmp * tmp;
            }
            if (Dp_L2 <= o.eps2_sq * p_L2) { stop = 2; break; }
            // almost singular (:726-730): (p_L2 + eps2) / 1e-24; the division is only made when the cheap lower bound of
            // that threshold is reached (same decision: x * 9.9e23 < x / 1e-24 for every x >= 0)
            if (Dp_L2 >= (p_L2 + o.eps2) * 9.9e23 && Dp_L2 >= (p_L2 + o.eps2) / (kEpsilon * kEpsilon)) { stop = 4; break; }

            bool t0_known = false;
            if constexpr (SpecJac<Eval>::value) {
                if (!dscl) {
                    const double lambda01 = 1.0 * 0.1;  // the line search's lambda after one clipped backtrack
                    double probe[MM];
                    LM_FOR_REV(i) probe[i] = lm_axpy(lambda01, Dp[i], p[i]);
                    box_project<MM>(probe, box, m);
                    note_probe_hint(ev, probe);
                    if (wants_candidate_hint(ev)) {
                        // first step length of a projected-gradient walk from here (:876-879), needed early
                        t0 = lm_pg_first_step(Jte, m, tini, MMTag<MM>());
                        t0_known = true;
                        note_candidate_hint(ev, p, Jte, gprevtaken ? t : t0, lb, ub);
                    }
                }
            }
            e_new = eval_trial_scaled<MM>(ev, pDp, dscl, m, bad);
            ++cnt.nfev;
            // :748 -- overflow of the sum alone is tolerated, non-finite residuals are not
            if (!lm_finite(e_new) && bad) { stop = 7; break; }
            note_trial_outcome(ev, e_new <= gamma * e_cur);

            if (e_new <= gamma * e_cur) {  // LM step accepted, :753-785
                dL = 0.0;
                LM_FOR(i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);
                if (dL > 0.0) {
                    dF = e_cur - e_new;
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
                break;
            }

            // rejected: descent direction? (:810-816)
            gTd = 0.0;
            LM_FOR(i) {
                Jte[i] = -Jte[i];
                gTd += Jte[i] * Dp[i];
            }
            if (!t0_known) {  // first step length of a projected-gradient walk from here (:876-879)
                t0 = lm_pg_first_step(Jte, m, tini, MMTag<MM>());
            }
            // levmar: gTd <= -rho * pow(Dp_L2, kPow/2) (:816).  D**1.05 lies between D and D*D, so the power (a
            // ~2000-cycle libm call on the control warp) is only needed when gTd falls between the two bounds --
            // it practically never does (rho = 1e-8); the decision is levmar's in every case, NaN and Inf included
            bool descent;
            {
                const double dd = Dp_L2 * Dp_L2;
                const double hi = (Dp_L2 >= dd) ? Dp_L2 : dd, lo = (Dp_L2 >= dd) ? dd : Dp_L2;
                if (gTd <= -rho * hi * 1.000001) descent = true;
                else if (gTd > -rho * lo * 0.999999) descent = false;
                else descent = gTd <= -rho * BG_LM_POW(Dp_L2, kPow / 2.0);
            }
            if (descent) {
                const double steptl = 1e3 * sqrt(DBL_EPSILON);
                tmp = sqrt(p_L2);
                const double stepmx = 1e3 * ((tmp >= 1.0) ? tmp : 1.0);
                // the rejected trial point and its (finite or overflowed-sum) cost, for the probe at lambda = 1
                double trial_pt[MM];
                LM_FOR(i) trial_pt[i] = pDp[i];
                const int rc = lm_line_search<MM>(ev, m, p, e_cur, Jte, Dp, alpha, pDp, e_new, box, dscl,
                                                  stepmx, steptl, cnt, trial_pt, e_new, gprevtaken ? t : t0);
                if (rc != 0 || !lm_finite(e_new)) use_pg = true;
                else gprevtaken = 0;
            } else {
                use_pg = true;
            }

            if (use_pg) {  // projected gradient search, :871-946
                bool found = false, fatal = false;  // t0: computed above, before the line search

                // levmar walks t, t*beta, t*beta^2, ... one function evaluation at a time (:885-934).
                // The candidate points of that walk depend only on p and J^T e, so an evaluator may
                // take up to Eval::kCostBatch of them per call (one sweep over the samples, one
                // reduction); they are consumed strictly in levmar's order with levmar's tests, and
                // candidates past the stopping one are discarded and not counted in nfev.
                constexpr int KB = Eval::kCostBatch;
                bool pg_done = false;
                int width = 1;  // most walks stop at their first candidates: speculate 1, 2, 4, ... KB points
                t = gprevtaken ? t : t0;
                if constexpr (Eval::kLanePgWalk) {
                    // the evaluator runs the same walk with one candidate per LANE of its control warp
                    // (generation, projection and the acceptance tests of a batch in parallel; the first
                    // lane with an event decides, exactly as the sequential order would)
                    if (!dscl) {
                        const int outcome = ev.pg_walk(p, Jte, e_cur, lb, ub, t, t0, gprevtaken, pDp, Dp, Dp_L2, e_new, cnt.nfev);
                        if (outcome == 2) { stop = 7; fatal = true; }
                        found = outcome == 1;
                        pg_done = true;
                    }
                }
                if (!pg_done) {
                PgBatch<MM, Eval> batch;
                double* pts = batch.points(ev);
                constexpr bool kSites = SpecJac<Eval>::value;
                [[maybe_unused]] bool first_round = true, took_first = false;
                while (t > tming && !pg_done) {
                    int nc = 0;
                    double tt = t;
                    while (nc < width && tt > tming) {
                        double cand[MM];
                        LM_FOR(i) cand[i] = p[i] - tt * Jte[i];
                        box_project<MM>(cand, box, m);
                        LM_FOR(i) pts[nc * m + i] = cand[i];
                        ++nc;
                        tt *= beta;
                    }
                    batch.run(ev, nc, dscl, m, kSites && first_round);
                    bool restarted = false;
                    double tc = t;  // the same recurrence reproduces every candidate's t
                    for (int c = 0; c < nc; ++c, tc *= beta) {
                        t = tc;
                        Dp_L2 = 0.0;
                        LM_FOR(i) {
                            pDp[i] = pts[c * m + i];
                            Dp[i] = tmp = pDp[i] - p[i];
                            Dp_L2 += tmp * tmp;
                        }
                        e_new = batch.cost(ev, c);
                        ++cnt.nfev;
                        if (!lm_finite(e_new) && batch.bad(ev, c)) { stop = 7; fatal = true; pg_done = true; break; }

                        gTd = 0.0;
                        LM_FOR(i) gTd += Jte[i] * Dp[i];

                        if (gprevtaken && e_new <= e_cur + 2.0 * 0.99999 * gTd) {  // starting t too small
                            t = t0 * beta;  // t = t0, then the loop increment of :885 still applies (:926-930)
                            gprevtaken = 0;
                            restarted = true;
                            break;
                        }
                        if (e_new <= e_cur + 2.0 * alpha * gTd) {
                            found = true;
                            pg_done = true;
                            if constexpr (kSites) took_first = first_round && c == 0;
                            break;
                        }
                    }
                    if constexpr (kSites) first_round = false;
                    if (!pg_done && !restarted) t = tt;
                    width = (2 * width < KB) ? 2 * width : KB;
                }
                if constexpr (kSites) note_pg_outcome(ev, took_first);
                }
                if (fatal) goto done;
                if (!found) { gprevtaken = 0; break; }
                gprevtaken = 1;
            }

            // take the line-search / projected-gradient point (:948-967)
            Dp_L2 = 0.0;
            LM_FOR(i) {
                tmp = pDp[i] - p[i];
                Dp_L2 += tmp * tmp;
            }
            if (Dp_L2 <= o.eps2_sq * p_L2) { stop = 2; break; }
            LM_FOR(i) p[i] = pDp[i];
            e_cur = e_new;
            break;
        }
    }

done:
    if (k >= o.itmax) stop = 3;
    LM_FOR(i) JtJ[i * m + i] = diag[i];
    lm_fill_info<MM>(info, JtJ, m, e_init, e_cur, ginf, Dp_L2, mu, k, stop, cnt);
    if (JtJ_out) LM_FOR(i) LM_FOR(jj) JtJ_out[i * m + jj] = JtJ[i * m + jj];
    if (dscl) LM_FOR(i) p[i] *= dscl[i];
    return (stop != 4 && stop != 7) ? k : kLmError;
}

// =============================================================================================
// Unconstrained LM with a full Jacobian evaluation per outer iteration, lm_core.c:64-432.
// =============================================================================================
template <int MM, class Eval>
BG_HDI int lm_der(Eval& ev, int m, double* p, const LmOptions& o, double* info, double* JtJ_out) {
    double JtJ[MM * MM], Jte[MM], Dp[MM], diag[MM], pDp[MM];
    double mu = 0.0, ginf = 0.0, tmp, e_cur, e_new, e_init, p_L2 = 0.0, Dp_L2 = DBL_MAX, dF, dL;
    int k, stop = 0, nu = 2;
    bool bad = false;
    LmCounters cnt = {0, 0, 0};

    LM_FOR(i) {
        LM_FOR(jj) JtJ[i * m + jj] = 0.0;
        diag[i] = 0.0;
    }

    e_cur = ev.cost(p, bad);
    cnt.nfev = 1;
    e_init = e_cur;
    if (!lm_finite(e_cur)) stop = 7;

    for (k = 0; k < o.itmax && !stop; ++k) {
        if (e_cur <= o.eps3) { stop = 6; break; }

        ev.jac(p, JtJ, Jte);
        ++cnt.njev;

        p_L2 = ginf = 0.0;
        LM_FOR(i) {
            if (ginf < (tmp = lm_abs(Jte[i]))) ginf = tmp;
            diag[i] = JtJ[i * m + i];
            p_L2 += p[i] * p[i];
        }
        if (ginf <= o.eps1) { Dp_L2 = 0.0; stop = 1; break; }

        if (k == 0) {
            tmp = -DBL_MAX;
            LM_FOR(i) if (diag[i] > tmp) tmp = diag[i];
            mu = o.tau * tmp;
        }

        for (;;) {
            LM_FOR(i) JtJ[i * m + i] += mu;
            ++cnt.nlss;
            if (solve_lu<MM>(JtJ, Jte, Dp, m)) {
                Dp_L2 = 0.0;
                LM_FOR(i) {
                    pDp[i] = p[i] + (tmp = Dp[i]);
                    Dp_L2 += tmp * tmp;
                }
                if (Dp_L2 <= o.eps2_sq * p_L2) { stop = 2; break; }
                // almost singular (:726-730): (p_L2 + eps2) / 1e-24; the division is only made when the cheap lower bound of
            // that threshold is reached (same decision: x * 9.9e23 < x / 1e-24 for every x >= 0)
            if (Dp_L2 >= (p_L2 + o.eps2) * 9.9e23 && Dp_L2 >= (p_L2 + o.eps2) / (kEpsilon * kEpsilon)) { stop = 4; break; }

                e_new = eval_cost_site(ev, kSiteTrial, pDp, bad);
                ++cnt.nfev;
                if (!lm_finite(e_new)) { stop = 7; break; }

                dL = 0.0;
                LM_FOR(i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);
                dF = e_cur - e_new;
                note_trial_outcome(ev, dL > 0.0 && dF > 0.0);
                if (dL > 0.0 && dF > 0.0) {
                    tmp = (2.0 * dF / dL - 1.0);
                    tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= kOneThird) ? tmp : kOneThird);
                    nu = 2;
                    LM_FOR(i) p[i] = pDp[i];
                    e_cur = e_new;
                    break;
                }
            }
            mu *= nu;
            const int nu2 = (int)((unsigned)nu << 1);
            if (nu2 <= nu) { stop = 5; break; }
            nu = nu2;
            LM_FOR(i) JtJ[i * m + i] = diag[i];
        }
    }
    if (k >= o.itmax) stop = 3;
    LM_FOR(i) JtJ[i * m + i] = diag[i];
    lm_fill_info<MM>(info, JtJ, m, e_init, e_cur, ginf, Dp_L2, mu, k, stop, cnt);
    if (JtJ_out) LM_FOR(i) LM_FOR(jj) JtJ_out[i * m + jj] = JtJ[i * m + jj];
    return (stop != 4 && stop != 7) ? k : kLmError;
}

}  // namespace brdfgpu
