/*
 * lm_oracle.c -- CPU oracle, solver half.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * A from-scratch plain-C restatement of the levmar 2.6 routines the BRDF fit runs through
 * (reference: levmar/lmbc_core.c, lm_core.c, misc_core.c, Axb_core.c, built without LAPACK as
 * shipped, levmar/levmar.h:31).  The arithmetic -- operation order, summation order, comparison
 * direction -- follows the cited lines so that results are bit-identical to the reference build
 * (checked against oracle/_ref/liblevmar_ref.so in tests/test_oracle_kat.py and tests/test_oracle_golden.py); the code
 * organisation (one shared iteration context, helper routines, no macros/templating over the
 * real type) is this repository's own.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* constants: levmar/lmbc.c:35-38, lm.c:35-36, levmar.h:95-101, misc.h:54-61 */
#define K_EPSILON    1E-12
#define K_ONE_THIRD  0.3333333334
#define K_LSITMAX    150
#define K_POW        2.1
#define K_INIT_MU    1E-03
#define K_STOP_THR   1E-17
#define K_DIFF_DELTA 1E-06
#define K_BLOCK      32
#define K_BLOCK_SQ   (K_BLOCK * K_BLOCK)

/* misc.h:68 -- not fabs(): keeps -0.0 and NaN handling identical */
static inline double absval(double v) { return (v >= 0.0) ? v : -v; }
static inline int is_finite(double v) { return isfinite(v); } /* compiler.h:36 `finite` */

/* ------------------------------------------------------------------------------------------------
 * e = x - y (x may be NULL = 0), returns sum of squares.  misc_core.c:721-807: four running sums,
 * blocks of eight walked from the top block downwards, then the 1..7 leftover elements; element
 * (i) of a block feeds sum[(7 - (i mod 8)) mod 4], leftover k of r feeds sum[(7 - r + k) mod 4].
 * ------------------------------------------------------------------------------------------------ */
double oracle_L2nrmxmy(double *e, const double *x, const double *y, int n)
{
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int nblk = (n >> 3) << 3;
    int top, k;

    for (top = nblk - 1; top > 0; top -= 8) {
        for (k = 0; k < 8; ++k) {
            const int idx = top - k;
            const double d = x ? x[idx] - y[idx] : -y[idx];
            e[idx] = d;
            acc[k & 3] += d * d;
        }
    }
    if (nblk < n) {
        const int r = n - nblk;
        for (k = 0; k < r; ++k) {
            const int idx = nblk + k;
            const double d = x ? x[idx] - y[idx] : -y[idx];
            e[idx] = d;
            acc[(7 - r + k) & 3] += d * d;
        }
    }
    return acc[0] + acc[1] + acc[2] + acc[3];
}

/* b = a^T a for row-major n x m a; 32-wide blocking over rows (misc_core.c:82-134). */
void oracle_trans_mat_mat_mult(const double *a, double *b, int n, int m)
{
    int jj, kk, i, j, k;

    for (jj = 0; jj < m; jj += K_BLOCK) {
        const int jhi = (jj + K_BLOCK <= m) ? jj + K_BLOCK : m;
        for (i = 0; i < m; ++i)
            for (j = (jj >= i ? jj : i); j < jhi; ++j)
                b[i * m + j] = 0.0;

        for (kk = 0; kk < n; kk += K_BLOCK) {
            const int khi = (kk + K_BLOCK <= n) ? kk + K_BLOCK : n;
            for (i = 0; i < m; ++i)
                for (j = (jj >= i ? jj : i); j < jhi; ++j) {
                    double s = 0.0;
                    for (k = kk; k < khi; ++k)
                        s += a[k * m + i] * a[k * m + j];
                    b[i * m + j] += s;
                }
        }
    }
    for (i = 0; i < m; ++i)
        for (j = 0; j < i; ++j)
            b[i * m + j] = b[j * m + i];
}

/* step rule shared by both difference schemes: d = max(|1e-4 * p_j|, delta)  (misc_core.c:154-158) */
static double fd_step(double pj, double delta)
{
    double d = 1E-04 * pj;
    d = absval(d);
    if (d < delta) d = delta;
    return d;
}

/* forward differences, misc_core.c:137-172 */
void oracle_fdif_forw_jac(oracle_func_t func, double *p, double *hx, double *hxx, double delta,
                          double *jac, int m, int n, void *adata)
{
    int i, j;
    for (j = 0; j < m; ++j) {
        const double saved = p[j];
        double d = fd_step(p[j], delta);
        p[j] += d;
        func(p, hxx, m, n, adata);
        p[j] = saved;
        d = 1.0 / d;
        for (i = 0; i < n; ++i)
            jac[i * m + j] = (hxx[i] - hx[i]) * d;
    }
}

/* central differences, misc_core.c:175-211 */
void oracle_fdif_cent_jac(oracle_func_t func, double *p, double *hxm, double *hxp, double delta,
                          double *jac, int m, int n, void *adata)
{
    int i, j;
    for (j = 0; j < m; ++j) {
        const double saved = p[j];
        double d = fd_step(p[j], delta);
        p[j] -= d;
        func(p, hxm, m, n, adata);
        p[j] = saved + d;
        func(p, hxp, m, n, adata);
        p[j] = saved;
        d = 0.5 / d;
        for (i = 0; i < n; ++i)
            jac[i * m + j] = (hxp[i] - hxm[i]) * d;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Crout LU with implicit row scaling and partial pivoting on a private copy, then the permuted
 * forward and the back substitution.  Factorisation: Axb_core.c:1196-1247 (== misc_core.c:457-506);
 * returns 0 if a row is entirely zero (:1202-1208); a zero pivot becomes DBL_EPSILON (:1240-1241).
 * ------------------------------------------------------------------------------------------------ */
static int lu_factor(double *a, int *perm, double *rowscale, int m)
{
    int i, j, k, piv = -1;

    for (i = 0; i < m; ++i) {
        double big = 0.0, t;
        for (j = 0; j < m; ++j)
            if ((t = absval(a[i * m + j])) > big) big = t;
        if (big == 0.0) return 0;
        rowscale[i] = 1.0 / big;
    }
    for (j = 0; j < m; ++j) {
        double big = 0.0, s, t;
        for (i = 0; i < j; ++i) {
            s = a[i * m + j];
            for (k = 0; k < i; ++k) s -= a[i * m + k] * a[k * m + j];
            a[i * m + j] = s;
        }
        for (i = j; i < m; ++i) {
            s = a[i * m + j];
            for (k = 0; k < j; ++k) s -= a[i * m + k] * a[k * m + j];
            a[i * m + j] = s;
            if ((t = rowscale[i] * absval(s)) >= big) { big = t; piv = i; }
        }
        if (j != piv) {
            for (k = 0; k < m; ++k) {
                t = a[piv * m + k];
                a[piv * m + k] = a[j * m + k];
                a[j * m + k] = t;
            }
            rowscale[piv] = rowscale[j];
        }
        perm[j] = piv;
        if (a[j * m + j] == 0.0) a[j * m + j] = DBL_EPSILON;
        if (j != m - 1) {
            t = 1.0 / a[j * m + j];
            for (i = j + 1; i < m; ++i) a[i * m + j] *= t;
        }
    }
    return 1;
}

/* substitution phase, Axb_core.c:1252-1270 (== misc_core.c:515-533) */
static void lu_substitute(const double *a, const int *perm, double *x, int m)
{
    int i, j, first = 0;

    for (i = 0; i < m; ++i) {
        double s;
        j = perm[i];
        s = x[j];
        x[j] = x[i];
        if (first != 0) {
            for (j = first - 1; j < i; ++j) s -= a[i * m + j] * x[j];
        } else if (s != 0.0) {
            first = i + 1;
        }
        x[i] = s;
    }
    for (i = m - 1; i >= 0; --i) {
        double s = x[i];
        for (j = i + 1; j < m; ++j) s -= a[i * m + j] * x[j];
        x[i] = s / a[i * m + i];
    }
}

int oracle_Ax_eq_b_LU(const double *A, const double *B, double *x, int m)
{
    double *a = (double *)malloc(((size_t)m * m + m) * sizeof(double) + (size_t)m * sizeof(int));
    double *rowscale;
    int *perm, ok;

    if (!a) { fprintf(stderr, "oracle_Ax_eq_b_LU: out of memory\n"); exit(1); }
    rowscale = a + (size_t)m * m;
    perm = (int *)(rowscale + m);
    memcpy(a, A, (size_t)m * m * sizeof(double));
    memcpy(x, B, (size_t)m * sizeof(double));
    ok = lu_factor(a, perm, rowscale, m);
    if (ok) lu_substitute(a, perm, x, m);
    free(a);
    return ok;
}

/* C = sumsq/(n-m) * JtJ^-1 via LU, column by column (misc_core.c:426-542, 564-591). 0 on failure. */
int oracle_covar(const double *JtJ, double *C, double sumsq, int m, int n)
{
    double *a = (double *)malloc(((size_t)m * m + 2 * (size_t)m) * sizeof(double) + (size_t)m * sizeof(int));
    double *x, *rowscale, fact;
    int *perm, i, l;

    if (!a) return 0;
    x = a + (size_t)m * m;
    rowscale = x + m;
    perm = (int *)(rowscale + m);
    memcpy(a, JtJ, (size_t)m * m * sizeof(double));
    if (!lu_factor(a, perm, rowscale, m)) {
        fprintf(stderr, "oracle_covar: singular matrix\n");
        free(a);
        return 0;
    }
    for (l = 0; l < m; ++l) {
        for (i = 0; i < m; ++i) x[i] = 0.0;
        x[l] = 1.0;
        lu_substitute(a, perm, x, m);
        for (i = 0; i < m; ++i) C[i * m + l] = x[i];
    }
    free(a);
    fact = sumsq / (double)(n - m);
    for (i = 0; i < m * m; ++i) C[i] *= fact;
    return m;
}

/* ------------------------------------------------------------------------------------------------
 * Normal equations  JtJ = J^T J,  Jte = J^T e.
 * small (n*m below/at the 32*32 threshold): rows walked downwards, lower triangle accumulated then
 * mirrored (lmbc_core.c:592-616); large: blocked product + row-major J^T e (lmbc_core.c:617-632).
 * The bc/der drivers switch at nm < 1024 (lmbc_core.c:573, lm_core.c:197), dif at nm <= 1024
 * (lm_core.c:594) -- `small` is decided by the caller.
 * ------------------------------------------------------------------------------------------------ */
static void normal_equations(const double *jac, const double *e, double *JtJ, double *Jte,
                             int m, int n, int small)
{
    int i, j, l;

    if (small) {
        for (i = m * m; i-- > 0;) JtJ[i] = 0.0;
        for (i = m; i-- > 0;) Jte[i] = 0.0;
        for (l = n; l-- > 0;) {
            const double *row = jac + (size_t)l * m;
            for (i = m; i-- > 0;) {
                const double a = row[i];
                for (j = i + 1; j-- > 0;) JtJ[i * m + j] += row[j] * a;
                Jte[i] += a * e[l];
            }
        }
        for (i = m; i-- > 0;)
            for (j = i + 1; j < m; ++j) JtJ[i * m + j] = JtJ[j * m + i];
    } else {
        oracle_trans_mat_mat_mult(jac, JtJ, n, m);
        for (i = 0; i < m; ++i) Jte[i] = 0.0;
        for (i = 0; i < n; ++i) {
            const double *row = jac + (size_t)i * m;
            const double ei = e[i];
            for (l = 0; l < m; ++l) Jte[l] += row[l] * ei;
        }
    }
}

/* ---------------- box helpers ---------------- */

/* median of (lo, v, hi) with the reference's exact comparison tree (lmbc_core.c:59-61): the
 * outcome for NaN / crossed inputs depends on it. */
static double median3(double lo, double v, double hi)
{
    if (lo >= v) {
        if (hi >= lo) return lo;
        return (hi <= v) ? v : hi;
    }
    if (hi >= v) return v;
    return (hi <= lo) ? lo : hi;
}

/* lmbc_core.c:68-88 */
static void box_project(double *p, const double *lb, const double *ub, int m)
{
    int i;
    if (!lb && !ub) return;
    for (i = m; i-- > 0;) {
        if (lb && ub) p[i] = median3(lb[i], p[i], ub[i]);
        else if (ub) { if (p[i] > ub[i]) p[i] = ub[i]; }
        else { if (p[i] < lb[i]) p[i] = lb[i]; }
    }
}

/* lmbc_core.c:94-142: infinite (+-DBL_MAX) bounds are left alone */
static void box_scale(double *lb, double *ub, const double *scl, int m, int divide)
{
    int i;
    for (i = m; i-- > 0;) {
        if (ub && ub[i] != DBL_MAX) ub[i] = divide ? ub[i] / scl[i] : ub[i] * scl[i];
        if (lb && lb[i] != -DBL_MAX) lb[i] = divide ? lb[i] / scl[i] : lb[i] * scl[i];
    }
}

/* misc_core.c:661-671 */
static int box_consistent(const double *lb, const double *ub, int m)
{
    int i;
    if (!lb || !ub) return 1;
    for (i = 0; i < m; ++i)
        if (lb[i] > ub[i]) return 0;
    return 1;
}

/* overflow-avoiding 2-norm, lmbc_core.c:155-168 (Blue's method, the no-LAPACK branch) */
static double scaled_norm(const double *v, int n)
{
    double big = 0.0, s = 0.0;
    int i;
    for (i = n; i-- > 0;) {
        if (v[i] > big) big = v[i];
        else if (v[i] < -big) big = -v[i];
    }
    for (i = n; i-- > 0;) {
        const double t = v[i] / big;
        s += t * t;
    }
    return big * sqrt(s);
}

/* ---------------- shared evaluation context for the bc driver ---------------- */
struct bc_ctx {
    oracle_func_t func;
    void *adata;
    double *x, *hx;          /* measurements, model output / residual scratch (n) */
    double *lb, *ub, *dscl;  /* may be NULL */
    double *scaled;          /* m scratch for dscl*p */
    int m, n;
    int nfev;
};

/* evaluate func at q (in scaled coordinates when dscl is given), hx <- x - func, return ||.||^2
 * (lmbc_core.c:728-738, 894-905) */
static double bc_cost_at(struct bc_ctx *c, double *q)
{
    int i;
    if (!c->dscl) {
        c->func(q, c->hx, c->m, c->n, c->adata);
    } else {
        for (i = c->m; i-- > 0;) c->scaled[i] = q[i] * c->dscl[i];
        c->func(c->scaled, c->hx, c->m, c->n, c->adata);
    }
    ++c->nfev;
    return oracle_L2nrmxmy(c->hx, c->x, c->hx, c->n);
}

/* ------------------------------------------------------------------------------------------------
 * Backtracking line search (Schnabel/Koontz/Weiss uncmin lnsrch with box projection),
 * lmbc_core.c:179-337.  `step` may be shortened in place (:234-240).  Returns iretcd (0 = found).
 * ------------------------------------------------------------------------------------------------ */
static int line_search(struct bc_ctx *c, const double *xc, double fc, const double *g, double *step,
                       double alpha, double *xnew, double *fnew_sumsq, double stepmx, double steptl)
{
    const int m = c->m;
    int i, it, firstback = 1;
    double sln, slp, rln, rmnlmb, lambda, tlmbda = 0.0, plmbda = 0.0, pfpls = 0.0, fpls, t;

    fc *= 0.5;
    for (i = m, t = 0.0; i-- > 0;) t += step[i] * step[i];
    sln = sqrt(t);
    if (sln > stepmx) {
        const double scl = stepmx / sln;
        for (i = m; i-- > 0;) step[i] *= scl;
        sln = stepmx;
    }
    for (i = m, slp = rln = 0.0; i-- > 0;) {
        double a, b;
        slp += g[i] * step[i];
        a = (absval(xc[i]) >= 1.0) ? absval(xc[i]) : 1.0;
        b = absval(step[i]) / a;
        if (rln < b) rln = b;
    }
    rmnlmb = steptl / rln;
    lambda = 1.0;

    for (it = K_LSITMAX; it-- > 0;) {
        for (i = m; i-- > 0;) xnew[i] = xc[i] + lambda * step[i];
        box_project(xnew, c->lb, c->ub, m);

        if (!c->dscl) {
            c->func(xnew, c->hx, m, c->n, c->adata);
            ++c->nfev;
        } else { /* :262-266 scales xnew in place and back */
            for (i = m; i-- > 0;) xnew[i] *= c->dscl[i];
            c->func(xnew, c->hx, m, c->n, c->adata);
            ++c->nfev;
            for (i = m; i-- > 0;) xnew[i] /= c->dscl[i];
        }
        t = oracle_L2nrmxmy(c->hx, c->x, c->hx, c->n);
        fpls = 0.5 * t;
        *fnew_sumsq = t;

        if (fpls <= fc + slp * alpha * lambda) return 0;
        if (lambda < rmnlmb) return 1;

        if (!is_finite(fpls)) {
            lambda *= 0.1;
            firstback = 1;
        } else {
            if (firstback) {
                tlmbda = -lambda * slp / ((fpls - fc - slp) * 2.0);
                firstback = 0;
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
    return 1;
}

/* info[] layout, lmbc_core.c:978-991 / lm_core.c:405-418 */
static void fill_info(double *info, const double *JtJ, int m, double e0, double e, double ginf,
                      double dp2, double mu, int k, int stop, int nfev, int njev, int nlss)
{
    double big = -DBL_MAX;
    int i;
    if (!info) return;
    for (i = 0; i < m; ++i)
        if (big < JtJ[i * m + i]) big = JtJ[i * m + i];
    info[0] = e0; info[1] = e; info[2] = ginf; info[3] = dp2; info[4] = mu / big;
    info[5] = (double)k; info[6] = (double)stop; info[7] = (double)nfev;
    info[8] = (double)njev; info[9] = (double)nlss;
}

static void read_opts(const double *opts, double *tau, double *eps1, double *eps2, double *eps2_sq,
                      double *eps3)
{
    if (opts) {
        *tau = opts[0]; *eps1 = opts[1]; *eps2 = opts[2]; *eps2_sq = opts[2] * opts[2]; *eps3 = opts[3];
    } else {
        *tau = K_INIT_MU; *eps1 = K_STOP_THR; *eps2 = K_STOP_THR;
        *eps2_sq = K_STOP_THR * K_STOP_THR; *eps3 = K_STOP_THR;
    }
}

/* ================================================================================================
 * Box-constrained LM with analytic Jacobian: projected LM step, else line search along the LM
 * step when it is a descent direction, else projected-gradient search.  lmbc_core.c:369-1022.
 * ================================================================================================ */
int oracle_dlevmar_bc_der(oracle_func_t func, oracle_jacf_t jacf, double *p, double *x, int m, int n,
                          double *lb, double *ub, double *dscl, int itmax, double *opts, double *info,
                          double *work, double *covar, void *adata)
{
    const double alpha = 1e-4, beta = 0.9, gamma = 0.99995, rho = 1e-8;
    const double tini = 1.0, tming = 1e-18;
    const int nm = n * m;
    double tau, eps1, eps2, eps2_sq, eps3;
    double *e, *hx, *Jte, *jac, *JtJ, *Dp, *diag, *pDp;
    double mu = 0.0, ginf = 0.0, t = 0.0, t0, tmp;
    double e_cur, e_new = 0.0, e_init, p_L2 = 0.0, Dp_L2 = DBL_MAX, dF, dL, gTd;
    int i, j, k, ownwork = 0, stop = 0, nu = 2, njev = 0, nlss = 0, gprevtaken = 0, numactive;
    struct bc_ctx c;

    if (n < m) {
        fprintf(stderr, "oracle_dlevmar_bc_der(): cannot solve a problem with fewer measurements [%d] than unknowns [%d]\n", n, m);
        return ORACLE_LM_ERROR;
    }
    if (!jacf) {
        fprintf(stderr, "oracle_dlevmar_bc_der(): no Jacobian function\n");
        return ORACLE_LM_ERROR;
    }
    if (!box_consistent(lb, ub, m)) {
        fprintf(stderr, "oracle_dlevmar_bc_der(): at least one lower bound exceeds the upper one\n");
        return ORACLE_LM_ERROR;
    }
    c.scaled = NULL;
    if (dscl) {
        for (i = m; i-- > 0;)
            if (dscl[i] <= 0.0) {
                fprintf(stderr, "oracle_dlevmar_bc_der(): scaling constants should be positive (scale %d: %g <= 0)\n", i, dscl[i]);
                return ORACLE_LM_ERROR;
            }
        c.scaled = (double *)malloc((size_t)m * sizeof(double));
        if (!c.scaled) return ORACLE_LM_ERROR;
    }
    read_opts(opts, &tau, &eps1, &eps2, &eps2_sq, &eps3);

    if (!work) {
        work = (double *)malloc(((size_t)2 * n + 4 * (size_t)m + (size_t)n * m + (size_t)m * m) * sizeof(double));
        if (!work) { free(c.scaled); return ORACLE_LM_ERROR; }
        ownwork = 1;
    }
    e = work; hx = e + n; Jte = hx + n; jac = Jte + m; JtJ = jac + nm; Dp = JtJ + m * m;
    diag = Dp + m; pDp = diag + m;

    c.func = func; c.adata = adata; c.x = x; c.hx = hx; c.lb = lb; c.ub = ub; c.dscl = dscl;
    c.m = m; c.n = n; c.nfev = 0;

    /* feasibility of the start (:513-520) */
    for (i = 0; i < m; ++i) pDp[i] = p[i];
    box_project(p, lb, ub, m);
    for (i = 0; i < m; ++i)
        if (pDp[i] != p[i])
            fprintf(stderr, "Warning: component %d of starting point not feasible in oracle_dlevmar_bc_der()! [%g projected to %g]\n", i, pDp[i], p[i]);

    /* e = x - f(p) (:522-534) */
    func(p, hx, m, n, adata);
    c.nfev = 1;
    e_cur = oracle_L2nrmxmy(e, x, hx, n);
    e_init = e_cur;
    if (!is_finite(e_cur)) stop = 7;

    if (dscl) {
        for (i = m; i-- > 0;) p[i] /= dscl[i];
        box_scale(lb, ub, dscl, m, 1);
    }

    for (k = 0; k < itmax && !stop; ++k) {
        if (e_cur <= eps3) { stop = 6; break; }

        /* Jacobian at p (:555-570) */
        if (!dscl) {
            jacf(p, jac, m, n, adata);
            ++njev;
        } else {
            for (i = m; i-- > 0;) c.scaled[i] = p[i] * dscl[i];
            jacf(c.scaled, jac, m, n, adata);
            ++njev;
            for (i = n; i-- > 0;)
                for (j = m; j-- > 0;) jac[(size_t)i * m + j] *= dscl[j];
        }
        normal_equations(jac, e, JtJ, Jte, m, n, nm < K_BLOCK_SQ);

        /* ||J^T e||_inf over free variables, ||p||^2 (:639-646) */
        for (i = j = numactive = 0, p_L2 = ginf = 0.0; i < m; ++i) {
            if (ub && p[i] == ub[i]) { ++numactive; if (Jte[i] > 0.0) ++j; }
            else if (lb && p[i] == lb[i]) { ++numactive; if (Jte[i] < 0.0) ++j; }
            else if (ginf < (tmp = absval(Jte[i]))) ginf = tmp;
            diag[i] = JtJ[i * m + i];
            p_L2 += p[i] * p[i];
        }
        if (j == numactive && ginf <= eps1) { Dp_L2 = 0.0; stop = 1; break; }

        if (k == 0) { /* :666-674 */
            if (!lb && !ub) {
                for (i = 0, tmp = -DBL_MAX; i < m; ++i)
                    if (diag[i] > tmp) tmp = diag[i];
                mu = tau * tmp;
            } else {
                mu = 0.5 * tau * e_cur;
            }
        }

        for (;;) {
            int solved, use_pg = 0;

            for (i = 0; i < m; ++i) JtJ[i * m + i] += mu;
            solved = oracle_Ax_eq_b_LU(JtJ, Jte, Dp, m);
            ++nlss;

            if (!solved) { /* :788-804 */
                int nu2;
                mu *= nu;
                nu2 = nu << 1;
                if (nu2 <= nu) { stop = 5; break; }
                nu = nu2;
                for (i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
                continue;
            }

            for (i = 0; i < m; ++i) pDp[i] = p[i] + Dp[i];
            box_project(pDp, lb, ub, m);
            for (i = 0, Dp_L2 = 0.0; i < m; ++i) {
                Dp[i] = tmp = pDp[i] - p[i];
                Dp_L2 += tmp * tmp;
            }
            if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
            if (Dp_L2 >= (p_L2 + eps2) / (K_EPSILON * K_EPSILON)) { stop = 4; break; }

            e_new = bc_cost_at(&c, pDp);
            if (!is_finite(e_new) && !is_finite(scaled_norm(hx, n))) { stop = 7; break; }

            if (e_new <= gamma * e_cur) { /* LM step accepted, :753-785 */
                for (i = 0, dL = 0.0; i < m; ++i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);
                if (dL > 0.0) {
                    dF = e_cur - e_new;
                    tmp = (2.0 * dF / dL - 1.0);
                    tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= K_ONE_THIRD) ? tmp : K_ONE_THIRD);
                } else {
                    tmp = 0.1 * e_new;
                    mu = (mu >= tmp) ? tmp : mu;
                }
                nu = 2;
                for (i = 0; i < m; ++i) p[i] = pDp[i];
                for (i = 0; i < n; ++i) e[i] = hx[i];
                e_cur = e_new;
                gprevtaken = 0;
                break;
            }

            /* LM step rejected: is it at least a descent direction? (:810-816) */
            for (i = 0, gTd = 0.0; i < m; ++i) {
                Jte[i] = -Jte[i];
                gTd += Jte[i] * Dp[i];
            }
            if (gTd <= -rho * pow(Dp_L2, K_POW / 2.0)) {
                double stepmx, steptl = 1e3 * sqrt(DBL_EPSILON);
                int rc;
                tmp = sqrt(p_L2);
                stepmx = 1e3 * ((tmp >= 1.0) ? tmp : 1.0);
                rc = line_search(&c, p, e_cur, Jte, Dp, alpha, pDp, &e_new, stepmx, steptl);
                if (rc != 0 || !is_finite(e_new)) use_pg = 1;
                else gprevtaken = 0;
            } else {
                use_pg = 1;
            }

            if (use_pg) { /* projected gradient search, :871-946 */
                int found = 0;
                for (i = 0, tmp = 0.0; i < m; ++i) tmp += Jte[i] * Jte[i];
                tmp = sqrt(tmp);
                tmp = 100.0 / (1.0 + tmp);
                t0 = (tmp <= tini) ? tmp : tini;

                for (t = gprevtaken ? t : t0; t > tming; t *= beta) {
                    for (i = 0; i < m; ++i) pDp[i] = p[i] - t * Jte[i];
                    box_project(pDp, lb, ub, m);
                    for (i = 0, Dp_L2 = 0.0; i < m; ++i) {
                        Dp[i] = tmp = pDp[i] - p[i];
                        Dp_L2 += tmp * tmp;
                    }
                    e_new = bc_cost_at(&c, pDp);
                    if (!is_finite(e_new) && !is_finite(scaled_norm(hx, n))) { stop = 7; goto done; }

                    for (i = 0, gTd = 0.0; i < m; ++i) gTd += Jte[i] * Dp[i];

                    if (gprevtaken && e_new <= e_cur + 2.0 * 0.99999 * gTd) {
                        t = t0;
                        gprevtaken = 0;
                        continue; /* note: the loop increment then applies t *= beta (:926-930) */
                    }
                    if (e_new <= e_cur + 2.0 * alpha * gTd) { found = 1; break; }
                }
                if (!found) { gprevtaken = 0; break; } /* search failed, next outer iteration */
                gprevtaken = 1;
            }

            /* take the line-search / projected-gradient point (:948-967) */
            for (i = 0, Dp_L2 = 0.0; i < m; ++i) {
                tmp = pDp[i] - p[i];
                Dp_L2 += tmp * tmp;
            }
            if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
            for (i = 0; i < m; ++i) p[i] = pDp[i];
            for (i = 0; i < n; ++i) e[i] = hx[i];
            e_cur = e_new;
            break;
        }
    }

done:
    if (k >= itmax) stop = 3;
    for (i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
    fill_info(info, JtJ, m, e_init, e_cur, ginf, Dp_L2, mu, k, stop, c.nfev, njev, nlss);

    if (covar) {
        oracle_covar(JtJ, covar, e_cur, m, n);
        if (dscl)
            for (i = m; i-- > 0;)
                for (j = m; j-- > 0;) covar[i * m + j] *= (dscl[i] * dscl[j]);
    }
    if (ownwork) free(work);
    if (dscl) {
        for (i = 0; i < m; ++i) p[i] *= dscl[i];
        box_scale(lb, ub, dscl, m, 0);
        free(c.scaled);
    }
    return (stop != 4 && stop != 7) ? k : ORACLE_LM_ERROR;
}

/* ---- finite-difference front end of the bc driver, lmbc_core.c:1027-1129 ---- */
struct fd_wrap {
    oracle_func_t func;
    void *adata;
    double *hx, *hxx;
    double delta;
    int forward;
};

static void fd_wrap_func(double *p, double *hx, int m, int n, void *data)
{
    struct fd_wrap *w = (struct fd_wrap *)data;
    w->func(p, hx, m, n, w->adata);
}

static void fd_wrap_jacf(double *p, double *jac, int m, int n, void *data)
{
    struct fd_wrap *w = (struct fd_wrap *)data;
    if (w->forward) {
        w->func(p, w->hx, m, n, w->adata);
        oracle_fdif_forw_jac(w->func, p, w->hx, w->hxx, w->delta, jac, m, n, w->adata);
    } else {
        oracle_fdif_cent_jac(w->func, p, w->hx, w->hxx, w->delta, jac, m, n, w->adata);
    }
}

int oracle_dlevmar_bc_dif(oracle_func_t func, double *p, double *x, int m, int n,
                          double *lb, double *ub, double *dscl, int itmax, double *opts, double *info,
                          double *work, double *covar, void *adata)
{
    struct fd_wrap w;
    int ret;

    w.forward = !opts || opts[4] >= 0.0;
    w.func = func;
    w.adata = adata;
    w.hx = (double *)malloc((size_t)2 * n * sizeof(double));
    if (!w.hx) return ORACLE_LM_ERROR;
    w.hxx = w.hx + n;
    w.delta = opts ? absval(opts[4]) : K_DIFF_DELTA;

    ret = oracle_dlevmar_bc_der(fd_wrap_func, fd_wrap_jacf, p, x, m, n, lb, ub, dscl, itmax, opts,
                                info, work, covar, &w);
    if (info) /* each Jacobian costs m+1 (forward) or 2m (central) evaluations, :1119-1124 */
        info[7] += info[8] * (w.forward ? (m + 1) : (2 * m));
    free(w.hx);
    return ret;
}

/* ================================================================================================
 * Unconstrained LM, analytic Jacobian.  lm_core.c:64-432.
 * ================================================================================================ */
int oracle_dlevmar_der(oracle_func_t func, oracle_jacf_t jacf, double *p, double *x, int m, int n,
                       int itmax, double *opts, double *info, double *work, double *covar, void *adata)
{
    const int nm = n * m;
    double tau, eps1, eps2, eps2_sq, eps3;
    double *e, *hx, *Jte, *jac, *JtJ, *Dp, *diag, *pDp;
    double mu = 0.0, ginf = 0.0, tmp, e_cur, e_new, e_init, p_L2 = 0.0, Dp_L2 = DBL_MAX, dF, dL;
    int i, k, ownwork = 0, stop = 0, nu = 2, nfev, njev = 0, nlss = 0;

    if (n < m) {
        fprintf(stderr, "oracle_dlevmar_der(): cannot solve a problem with fewer measurements [%d] than unknowns [%d]\n", n, m);
        return ORACLE_LM_ERROR;
    }
    if (!jacf) {
        fprintf(stderr, "oracle_dlevmar_der(): no Jacobian function\n");
        return ORACLE_LM_ERROR;
    }
    read_opts(opts, &tau, &eps1, &eps2, &eps2_sq, &eps3);
    if (!work) {
        work = (double *)malloc(((size_t)2 * n + 4 * (size_t)m + (size_t)n * m + (size_t)m * m) * sizeof(double));
        if (!work) return ORACLE_LM_ERROR;
        ownwork = 1;
    }
    e = work; hx = e + n; Jte = hx + n; jac = Jte + m; JtJ = jac + nm; Dp = JtJ + m * m;
    diag = Dp + m; pDp = diag + m;

    func(p, hx, m, n, adata);
    nfev = 1;
    e_cur = oracle_L2nrmxmy(e, x, hx, n);
    e_init = e_cur;
    if (!is_finite(e_cur)) stop = 7;

    for (k = 0; k < itmax && !stop; ++k) {
        if (e_cur <= eps3) { stop = 6; break; }

        jacf(p, jac, m, n, adata);
        ++njev;
        normal_equations(jac, e, JtJ, Jte, m, n, nm < K_BLOCK_SQ);

        for (i = 0, p_L2 = ginf = 0.0; i < m; ++i) {
            if (ginf < (tmp = absval(Jte[i]))) ginf = tmp;
            diag[i] = JtJ[i * m + i];
            p_L2 += p[i] * p[i];
        }
        if (ginf <= eps1) { Dp_L2 = 0.0; stop = 1; break; }

        if (k == 0) {
            for (i = 0, tmp = -DBL_MAX; i < m; ++i)
                if (diag[i] > tmp) tmp = diag[i];
            mu = tau * tmp;
        }

        for (;;) {
            int nu2;
            for (i = 0; i < m; ++i) JtJ[i * m + i] += mu;
            ++nlss;
            if (oracle_Ax_eq_b_LU(JtJ, Jte, Dp, m)) {
                for (i = 0, Dp_L2 = 0.0; i < m; ++i) {
                    pDp[i] = p[i] + (tmp = Dp[i]);
                    Dp_L2 += tmp * tmp;
                }
                if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
                if (Dp_L2 >= (p_L2 + eps2) / (K_EPSILON * K_EPSILON)) { stop = 4; break; }

                func(pDp, hx, m, n, adata);
                ++nfev;
                e_new = oracle_L2nrmxmy(hx, x, hx, n);
                if (!is_finite(e_new)) { stop = 7; break; }

                for (i = 0, dL = 0.0; i < m; ++i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);
                dF = e_cur - e_new;
                if (dL > 0.0 && dF > 0.0) {
                    tmp = (2.0 * dF / dL - 1.0);
                    tmp = 1.0 - tmp * tmp * tmp;
                    mu = mu * ((tmp >= K_ONE_THIRD) ? tmp : K_ONE_THIRD);
                    nu = 2;
                    for (i = 0; i < m; ++i) p[i] = pDp[i];
                    for (i = 0; i < n; ++i) e[i] = hx[i];
                    e_cur = e_new;
                    break;
                }
            }
            mu *= nu;
            nu2 = nu << 1;
            if (nu2 <= nu) { stop = 5; break; }
            nu = nu2;
            for (i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
        }
    }
    if (k >= itmax) stop = 3;
    for (i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
    fill_info(info, JtJ, m, e_init, e_cur, ginf, Dp_L2, mu, k, stop, nfev, njev, nlss);
    if (covar) oracle_covar(JtJ, covar, e_cur, m, n);
    if (ownwork) free(work);
    return (stop != 4 && stop != 7) ? k : ORACLE_LM_ERROR;
}

/* ================================================================================================
 * Unconstrained secant LM: difference Jacobian refreshed only when nu > 16 after a parameter
 * update or after K = max(m, 10) rank-one updates, Broyden update otherwise.  lm_core.c:438-842.
 * One damping attempt per outer iteration (no inner loop), unlike the _der driver.
 * ================================================================================================ */
int oracle_dlevmar_dif(oracle_func_t func, double *p, double *x, int m, int n,
                       int itmax, double *opts, double *info, double *work, double *covar, void *adata)
{
    const int nm = n * m, K = (m >= 10) ? m : 10;
    double tau, eps1, eps2, eps2_sq, eps3, delta;
    double *e, *hx, *Jte, *jac, *JtJ, *Dp, *diag, *pDp, *wrk, *wrk2;
    double mu = 0.0, ginf = 0.0, tmp, e_cur, e_new, e_init, p_L2 = 0.0, Dp_L2 = DBL_MAX, dF, dL;
    int i, j, k, l, ownwork = 0, stop = 0, nu, nu2, nfev, njap = 0, nlss = 0;
    int forward = 1, updjac = 0, updp = 1, newjac = 0;

    if (n < m) {
        fprintf(stderr, "oracle_dlevmar_dif(): cannot solve a problem with fewer measurements [%d] than unknowns [%d]\n", n, m);
        return ORACLE_LM_ERROR;
    }
    read_opts(opts, &tau, &eps1, &eps2, &eps2_sq, &eps3);
    if (opts) {
        delta = opts[4];
        if (delta < 0.0) { delta = -delta; forward = 0; }
    } else {
        delta = K_DIFF_DELTA;
    }
    if (!work) {
        work = (double *)malloc(((size_t)4 * n + 4 * (size_t)m + (size_t)n * m + (size_t)m * m) * sizeof(double));
        if (!work) return ORACLE_LM_ERROR;
        ownwork = 1;
    }
    e = work; hx = e + n; Jte = hx + n; jac = Jte + m; JtJ = jac + nm; Dp = JtJ + m * m;
    diag = Dp + m; pDp = diag + m; wrk = pDp + m; wrk2 = wrk + n;

    func(p, hx, m, n, adata);
    nfev = 1;
    e_cur = oracle_L2nrmxmy(e, x, hx, n);
    e_init = e_cur;
    if (!is_finite(e_cur)) stop = 7;

    nu = 20; /* forces a difference Jacobian on entry (:564) */

    for (k = 0; k < itmax && !stop; ++k) {
        if (e_cur <= eps3) { stop = 6; break; }

        if ((updp && nu > 16) || updjac == K) {
            if (forward) {
                oracle_fdif_forw_jac(func, p, hx, wrk, delta, jac, m, n, adata);
                ++njap; nfev += m;
            } else {
                oracle_fdif_cent_jac(func, p, wrk, wrk2, delta, jac, m, n, adata);
                ++njap; nfev += 2 * m;
            }
            nu = 2; updjac = 0; updp = 0; newjac = 1;
        }

        if (newjac) {
            newjac = 0;
            normal_equations(jac, e, JtJ, Jte, m, n, nm <= K_BLOCK_SQ);
            for (i = 0, p_L2 = ginf = 0.0; i < m; ++i) {
                if (ginf < (tmp = absval(Jte[i]))) ginf = tmp;
                diag[i] = JtJ[i * m + i];
                p_L2 += p[i] * p[i];
            }
        }

        if (ginf <= eps1) { Dp_L2 = 0.0; stop = 1; break; }

        if (k == 0) {
            for (i = 0, tmp = -DBL_MAX; i < m; ++i)
                if (diag[i] > tmp) tmp = diag[i];
            mu = tau * tmp;
        }

        for (i = 0; i < m; ++i) JtJ[i * m + i] += mu;
        ++nlss;
        if (oracle_Ax_eq_b_LU(JtJ, Jte, Dp, m)) {
            for (i = 0, Dp_L2 = 0.0; i < m; ++i) {
                pDp[i] = p[i] + (tmp = Dp[i]);
                Dp_L2 += tmp * tmp;
            }
            if (Dp_L2 <= eps2_sq * p_L2) { stop = 2; break; }
            if (Dp_L2 >= (p_L2 + eps2) / (K_EPSILON * K_EPSILON)) { stop = 4; break; }

            func(pDp, wrk, m, n, adata);
            ++nfev;
            e_new = oracle_L2nrmxmy(wrk2, x, wrk, n);
            if (!is_finite(e_new)) { stop = 7; break; }

            dF = e_cur - e_new;
            if (updp || dF > 0) { /* Broyden rank-one update of J (:759-769) */
                for (i = 0; i < n; ++i) {
                    double *row = jac + (size_t)i * m;
                    for (l = 0, tmp = 0.0; l < m; ++l) tmp += row[l] * Dp[l];
                    tmp = (wrk[i] - hx[i] - tmp) / Dp_L2;
                    for (j = 0; j < m; ++j) row[j] += tmp * Dp[j];
                }
                ++updjac;
                newjac = 1;
            }

            for (i = 0, dL = 0.0; i < m; ++i) dL += Dp[i] * (mu * Dp[i] + Jte[i]);

            if (dL > 0.0 && dF > 0.0) {
                tmp = (2.0 * dF / dL - 1.0);
                tmp = 1.0 - tmp * tmp * tmp;
                mu = mu * ((tmp >= K_ONE_THIRD) ? tmp : K_ONE_THIRD);
                nu = 2;
                for (i = 0; i < m; ++i) p[i] = pDp[i];
                for (i = 0; i < n; ++i) { e[i] = wrk2[i]; hx[i] = wrk[i]; }
                e_cur = e_new;
                updp = 1;
                continue;
            }
        }

        mu *= nu;
        nu2 = nu << 1;
        if (nu2 <= nu) { stop = 5; break; }
        nu = nu2;
        for (i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
    }
    if (k >= itmax) stop = 3;
    for (i = 0; i < m; ++i) JtJ[i * m + i] = diag[i];
    fill_info(info, JtJ, m, e_init, e_cur, ginf, Dp_L2, mu, k, stop, nfev, njap, nlss);
    if (covar) oracle_covar(JtJ, covar, e_cur, m, n);
    if (ownwork) free(work);
    return (stop != 4 && stop != 7) ? k : ORACLE_LM_ERROR;
}
