/*
 * brdf_oracle.c -- CPU oracle, model callback and fit drivers.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates the levmar callback the reference hands to dlevmar_bc_dif (brdfdata.cpp:962-989) and the
 * two driver presets around it (brdfdata.cpp:991-1075 global, :1077-1136 per face).
 */
#include <math.h>
#include <stdlib.h>

#include "oracle.h"

/* OpenCV's CV_PI (core/cvdef.h), the constant brdfdata.cpp:981 multiplies by */
#define ORACLE_PI 3.1415926535897932384626433832795

/* brdfdata.cpp:969-989.  `x` is the model output hx despite its name.  Any other modelInfo leaves
 * x untouched, as in the reference. */
void oracle_BRDFFunc(double *p, double *x, int m, int n, void *data)
{
    const struct oracle_extraData *d = (const struct oracle_extraData *)data;
    const double *angles = d->angles;
    int i;
    (void)m;

    for (i = 0; i < n; ++i) {
        const double cosPhi = angles[i];
        if (d->modelInfo == 0) {          /* Phong: note (n+2)/2*pi, not /(2 pi)  (:981) */
            const double cosTheta = angles[i + n * 2];
            x[i] = p[0] * cosPhi + ((p[2] + 2.0) / 2.0 * ORACLE_PI) * p[1] * (pow(cosTheta, p[2]));
        } else if (d->modelInfo == 1) {   /* Blinn-Phong (:986) */
            const double cosThetaDash = angles[i + n];
            x[i] = p[0] * cosPhi + p[1] * (pow(cosThetaDash, p[2]));
        }
    }
}

/* Exact partial derivatives of the callback above; row-major n x m, columns >= 3 are zero. */
void oracle_BRDFJac(double *p, double *jac, int m, int n, void *data)
{
    const struct oracle_extraData *d = (const struct oracle_extraData *)data;
    const double *angles = d->angles;
    int i, j;

    for (i = 0; i < n; ++i) {
        double *row = jac + (long)i * m;
        const double c = angles[i];
        for (j = 0; j < m; ++j) row[j] = 0.0;
        row[0] = c;
        if (d->modelInfo == 0) {
            const double t = angles[i + n * 2];
            const double tn = pow(t, p[2]);
            const double coef = (p[2] + 2.0) / 2.0 * ORACLE_PI;
            row[1] = coef * tn;
            row[2] = p[1] * tn * (ORACLE_PI / 2.0 + coef * log(t));
        } else {
            const double t = angles[i + n];
            const double tn = pow(t, p[2]);
            row[1] = tn;
            row[2] = p[1] * tn * log(t);
        }
    }
}

static int run_preset(const double *phi, const double *thetaDash, const double *theta,
                      const double *I, long n, int model, const double p0[3], int itmax,
                      const double opts_in[5], double *p_out, double *info_out)
{
    struct oracle_extraData data;
    double p[3], opts[5], info[ORACLE_LM_INFO_SZ], lower[3] = {0, 0, 0}, upper[3] = {100, 100, 100};
    double *x = (double *)malloc((size_t)n * sizeof(double));
    int ret;
    long i;

    data.angles = (double *)malloc((size_t)3 * n * sizeof(double));
    if (!x || !data.angles) { free(x); free(data.angles); return ORACLE_LM_ERROR; }
    for (i = 0; i < n; ++i) {
        x[i] = I[i];
        data.angles[i] = phi[i];
        data.angles[n + i] = thetaDash[i];
        data.angles[2 * n + i] = theta ? theta[i] : 0.0;
    }
    data.modelInfo = model;
    for (i = 0; i < 3; ++i) p[i] = p0[i];
    for (i = 0; i < 5; ++i) opts[i] = opts_in[i];

    ret = oracle_dlevmar_bc_dif(oracle_BRDFFunc, p, x, 3, (int)n, lower, upper, NULL, itmax, opts,
                                info, NULL, NULL, &data);
    for (i = 0; i < 3; ++i) p_out[i] = p[i];
    if (info_out)
        for (i = 0; i < ORACLE_LM_INFO_SZ; ++i) info_out[i] = info[i];
    free(data.angles);
    free(x);
    return ret;
}

/* Per-face preset: p0 = (0.5, 1, 1), itmax 100, opts {1e-3, 1e-15, 1e-15, 1e-20, 1e-6}
 * (brdfdata.cpp:1085, 1107-1119).  Unlike :1102 the Phong block is written where the callback
 * reads it (the reference's offset bug leaves it uninitialised). */
int oracle_solve_equation(const double *phi, const double *thetaDash, const double *theta,
                          const double *I, int nimg, int model, double *p_out, double *info_out)
{
    static const double p0[3] = {0.5, 1.0, 1.0};
    static const double opts[5] = {1E-03, 1E-15, 1E-15, 1E-20, 1E-06};
    return run_preset(phi, thetaDash, theta, I, nimg, model, p0, 100, opts, p_out, info_out);
}

/* Global preset: p0 = (0,0,0), itmax 2000, opts {1e-3, 1e-15, 1e-10, 1e-50, delta = 1}
 * (brdfdata.cpp:1002, 1046-1058).  Samples are taken in the order given ("aligned" order,
 * SURVEY.md Q6). */
int oracle_solve_equation_single(const double *phi, const double *thetaDash, const double *theta,
                                 const double *I, long nsamples, int model, double *p_out,
                                 double *info_out)
{
    static const double p0[3] = {0.0, 0.0, 0.0};
    static const double opts[5] = {1E-03, 1E-15, 1E-10, 1E-50, 1.0};
    return run_preset(phi, thetaDash, theta, I, nsamples, model, p0, 2000, opts, p_out, info_out);
}
