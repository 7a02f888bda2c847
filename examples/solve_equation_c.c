/*
 * solve_equation_c.c -- the reference's two levmar call sites, in plain C, against libbrdfgpu.
 *
 * This is CBRDFdata::SolveEquation_SingleBRDF (brdfdata.cpp:991-1075) and CBRDFdata::SolveEquation
 * (brdfdata.cpp:1077-1136) with the Eigen/OpenCV containers replaced by malloc'ed arrays: same
 * angles layout [cosphi(n) ; costhetadash(n) ; costheta(n)], same extraData, same options, same
 * dlevmar_bc_dif argument list -- only the two names carry the brdfgpu_ prefix.  It shows that the
 * drop-in boundary needs nothing but a C compiler on the host side:
 *
 *     gcc -O2 -Iinclude examples/solve_equation_c.c -Lbrdf_b200 -lbrdfgpu -Wl,-rpath,$PWD/brdf_b200 -lm -o solve_equation_c
 *     ./solve_equation_c [nfaces]
 *
 * Output: one line per fit with the parameters and levmar's info[] summary (exit code 0 when all fits
 * returned >= 0).  tests/test_gpu_c_example.py builds and runs it and compares with the CPU oracle.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "brdfgpu.h"

/* deterministic pseudo-random numbers in [0,1) (splitmix64) */
static unsigned long long state = 88172645463325252ULL;
static double uniform(void) {
    unsigned long long z = (state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv) {
    const int nfaces = argc > 1 ? atoi(argv[1]) : 2000; /* faces x 16 LEDs */
    const int n = nfaces * 16, m = 3;
    const double truth[3] = {0.6, 0.35, 12.0};
    double *x = malloc(sizeof(double) * n), *angles = malloc(sizeof(double) * 3 * n);
    if (!x || !angles) return 2;
    for (int i = 0; i < n; ++i) {
        const double cphi = uniform(), ctd = uniform(), cth = uniform();
        angles[i] = cphi; angles[n + i] = ctd; angles[2 * n + i] = cth; /* brdfdata.cpp:1029-1042 */
        double v = truth[0] * cphi + truth[1] * pow(ctd, truth[2]) + (uniform() - 0.5) * 0.01;
        v = floor(255.0 * v);
        x[i] = (v < 0 ? 0 : v > 255 ? 255 : v) / 255.0; /* 8-bit photographs */
    }
    brdfgpu_extraData data = {angles, BRDFGPU_MODEL_BLINN_PHONG}; /* struct extraData, brdfdata.cpp:962-966 */
    int failures = 0;

    { /* ---- SolveEquation_SingleBRDF: one fit over all samples (brdfdata.cpp:1002,1046-1058) ---- */
        double p[3] = {0.0, 0.0, 0.0}, opts[BRDFGPU_LM_OPTS_SZ], info[BRDFGPU_LM_INFO_SZ];
        double lower[] = {0, 0, 0}, upper[] = {100, 100, 100};
        opts[0] = BRDFGPU_LM_INIT_MU; opts[1] = 1E-15; opts[2] = 1E-10; opts[3] = 1E-50; opts[4] = 1;
        int error = brdfgpu_dlevmar_bc_dif(brdfgpu_BRDFFunc, p, x, m, n, lower, upper, NULL, 2000, opts, info, NULL, NULL, &data);
        printf("global  ret=%d p=%.12g %.12g %.12g sumsq=%.12g iters=%g reason=%g nfev=%g\n", error, p[0], p[1], p[2], info[1],
               info[5], info[6], info[7]);
        failures += error < 0;
    }
    /* ---- SolveEquation: the first 4 faces, 16 samples each (brdfdata.cpp:1085,1107-1119) ---- */
    for (int f = 0; f < 4; ++f) {
        double a16[48], x16[16];
        for (int i = 0; i < 16; ++i) {
            a16[i] = angles[16 * f + i]; a16[16 + i] = angles[n + 16 * f + i]; a16[32 + i] = angles[2 * n + 16 * f + i];
            x16[i] = x[16 * f + i];
        }
        brdfgpu_extraData d16 = {a16, BRDFGPU_MODEL_BLINN_PHONG};
        double p[3] = {0.5, 1.0, 1.0}, opts[BRDFGPU_LM_OPTS_SZ], info[BRDFGPU_LM_INFO_SZ];
        double lower[] = {0, 0, 0}, upper[] = {100, 100, 100};
        opts[0] = BRDFGPU_LM_INIT_MU; opts[1] = 1E-15; opts[2] = 1E-15; opts[3] = 1E-20; opts[4] = BRDFGPU_LM_DIFF_DELTA;
        int error = brdfgpu_dlevmar_bc_dif(brdfgpu_BRDFFunc, p, x16, m, 16, lower, upper, NULL, 100, opts, info, NULL, NULL, &d16);
        printf("face %d  ret=%d p=%.12g %.12g %.12g sumsq=%.12g iters=%g reason=%g nfev=%g\n", f, error, p[0], p[1], p[2], info[1],
               info[5], info[6], info[7]);
        failures += error < 0;
    }
    /* model prediction through the callback itself (BRDFFunc semantics) */
    {
        double p[3] = {0.6, 0.35, 12.0}, hx[4];
        brdfgpu_extraData d4 = {angles, BRDFGPU_MODEL_BLINN_PHONG};
        double a4[12];
        for (int i = 0; i < 4; ++i) { a4[i] = angles[i]; a4[4 + i] = angles[n + i]; a4[8 + i] = angles[2 * n + i]; }
        d4.angles = a4;
        brdfgpu_BRDFFunc(p, hx, 3, 4, &d4);
        printf("BRDFFunc hx=%.15g %.15g %.15g %.15g\n", hx[0], hx[1], hx[2], hx[3]);
    }
    free(x); free(angles);
    return failures ? 1 : 0;
}
