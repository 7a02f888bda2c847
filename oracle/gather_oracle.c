/*
 * gather_oracle.c -- CPU oracle, sample-gather half.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates the reference's gather (brdfdata.cpp:629-681 pixel<->face map, :945-960 radiance fetch,
 * :799-943 per-sample cosines, :314-330 face normals, :130-147 ambient subtraction, :683-756 LED
 * table).  The projection is the Tsai .cal camera (the literal GL path of the
 * reference cannot land a single centroid inside the photographs).
 *
 * PARITY UNPINNED at this boundary: the reference C++ needs OpenCV/Eigen/libigl/GL and cannot be
 * built here, and it ships no test for these functions, so no OUTPUT of the reference pins this file.
 * What narrows the arithmetic down instead is the reference's own build record: its qmake Makefile
 * (reference Makefile:17) compiles with `-O2 -std=gnu++11`, no -march/-mfma/-ffast-math -> x86-64
 * baseline = SSE2 packets of two doubles, no FMA contraction possible, no reassociation; and its
 * dependency list (reference Makefile, the ../libigl/external/eigen/Eigen/src/Core/... entries)
 * names CoreEvaluators.h, AssignEvaluator.h, arch/CUDA/Half.h, arch/AVX512, arch/ZVector,
 * arch/Default/ConjHelper.h and no IndexedView.h / Reshaped.h / arch/GPU: that file set is Eigen
 * 3.3.x (>= 3.3.5; libigl's external/eigen pin is 3.3.7), not 3.2 (no evaluators) and not 3.4.
 * Eigen 3.3's Redux.h / Dot.h then fix the evaluation order of the three Eigen expressions used:
 *
 *   v.normalize() on a RowVector3d (brdfdata.cpp:327, 838, 843, 890, 934)
 *       Dot.h: z = squaredNorm(); if (z > 0) v /= sqrt(z)  -- a true division per component.
 *       squaredNorm() = cwiseAbs2().sum() on a plain fixed-size vector: the evaluator has
 *       PacketAccessBit (3.3 allows unaligned packets) and LinearAccessBit, find_best_packet<double,3>
 *       is Packet2d, so redux_impl<LinearVectorizedTraversal, CompleteUnrolling> runs:
 *       predux({x0*x0, x1*x1}) then func(res, x2*x2)         =>  (x0*x0 + x1*x1) + x2*x2
 *   lightDir.cwiseProduct(face_normals.row(i)).sum(), H.cwiseProduct(face_normals.row(i)).sum()
 *   (brdfdata.cpp:893, 937)
 *       face_normals is a column-major MatrixXd, so .row(i) is a Block with a run-time inner stride:
 *       its evaluator has no PacketAccessBit, hence neither has the CwiseBinaryOp.  The expression
 *       takes its compile-time size (3) from its LEFT operand (CwiseBinaryOp traits use Lhs), so
 *       redux_impl<DefaultTraversal, CompleteUnrolling> -> redux_novec_unroller<0, 3>, which splits
 *       by halves, HalfLength = 3/2 = 1:  func(unroller<0,1>, unroller<1,2>)
 *                                                             =>  a0*b0 + (a1*b1 + a2*b2)
 *   face_normals.row(i).cwiseProduct(lightDir).sum() (brdfdata.cpp:841; glutcallbacks.cpp:383, 393)
 *       same operands the other way round: the LEFT operand is the Block, size Dynamic, so
 *       redux_impl<DefaultTraversal, NoUnrolling>: res = coeff(0); for i = 1.. res = func(res, coeff(i))
 *                                                             =>  (a0*b0 + a1*b1) + a2*b2
 *   R.cwiseProduct(P).sum(), viewDir.cwiseProduct(R).sum() on two plain RowVector3d (brdfdata.cpp:849;
 *   glutcallbacks.cpp:420): vectorised like squaredNorm       =>  (a0*b0 + a1*b1) + a2*b2
 *
 * (Eigen's headers are not installed in this image, so the Redux.h reasoning above is from the 3.3
 * sources as published, not re-checked against a local copy.)  Round 1 summed every dot left to
 * right; that order is kept as oracle_set_dot_order(ORACLE_DOT_SEQUENTIAL) / BRDFGPU_GATHER_SEQ_DOT and
 * differs from the canonical one only in GetCosLN / GetCosNH.  The Tsai projection is not Eigen code
 * (SURVEY.md 8c defines it): its dots stay (d0*n0 + d1*n1) + d2*n2.
 *
 * The CUDA gather must reproduce this file bit-for-bit:
 *   - every operation is a single IEEE-754 double +,-,*,/ or sqrt, never fused (-ffp-contract=off);
 *   - centroid = (((0 + v0) + v1) + v2) / 3.0  per component (brdfdata.cpp:653-660);
 *   - double -> int conversion truncates (brdfdata.cpp:677).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#include "oracle.h"

enum { CAM_CX = 0, CAM_CY, CAM_F, CAM_SX, CAM_N = 4, CAM_O = 7, CAM_A = 10, CAM_P = 13 };

static double dot3(const double *a, const double *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* fixed-size vector .cwiseProduct(row of a column-major MatrixXd).sum(): Eigen 3.3's unrolled scalar
 * reduction splits by halves (see the header); ORACLE_DOT_SEQUENTIAL restores round 1's guess */
static int g_dot_order = ORACLE_DOT_EIGEN33;
void oracle_set_dot_order(int order) { g_dot_order = order; }
static double dot3_row(const double *a, const double *b)
{
    if (g_dot_order == ORACLE_DOT_SEQUENTIAL) return dot3(a, b);
    return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
}

static void normalize3(double *v)
{
    const double z = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
    if (z > 0.0) {
        const double r = sqrt(z);
        v[0] /= r; v[1] /= r; v[2] /= r;
    }
}

static void centroid(const double *V, const int *F, int face, double *c)
{
    int j, k;
    for (k = 0; k < 3; ++k) {
        double s = 0.0;
        for (j = 0; j < 3; ++j) s += V[(size_t)F[(size_t)face * 3 + j] * 3 + k];
        c[k] = s / 3.0;
    }
}

/* brdfdata.cpp:695-755: 4x4 serpentine grid, x fixed */
void oracle_led_table(double *led)
{
    const double x = 303.5, min_y = -157.1, max_y = -2.3, min_z = 555.3, max_z = 645.8;
    const double y_step = (max_y - min_y) / 3, z_step = (max_z - min_z) / 3;
    const double ys[4] = {max_y, max_y - y_step, min_y + y_step, min_y};
    const double zs[4] = {min_z, min_z + z_step, max_z - z_step, max_z};
    int i;
    for (i = 0; i < 16; ++i) {
        const int row = i / 4, col = i % 4;
        led[i * 3 + 0] = x;
        led[i * 3 + 1] = (row % 2 == 0) ? ys[col] : ys[3 - col];
        led[i * 3 + 2] = zs[row];
    }
}

/* brdfdata.cpp:314-330 */
void oracle_face_normals(const double *V, const int *F, int nF, double *FN)
{
    int i, k;
    for (i = 0; i < nF; ++i) {
        const double *v0 = V + (size_t)F[(size_t)i * 3 + 0] * 3;
        const double *v1 = V + (size_t)F[(size_t)i * 3 + 1] * 3;
        const double *v2 = V + (size_t)F[(size_t)i * 3 + 2] * 3;
        double e1[3], e2[3], nrm[3];
        for (k = 0; k < 3; ++k) { e1[k] = v1[k] - v0[k]; e2[k] = v2[k] - v0[k]; }
        nrm[0] = e1[1] * e2[2] - e1[2] * e2[1];
        nrm[1] = e1[2] * e2[0] - e1[0] * e2[2];
        nrm[2] = e1[0] * e2[1] - e1[1] * e2[0];
        normalize3(nrm);
        for (k = 0; k < 3; ++k) FN[(size_t)i * 3 + k] = nrm[k];
    }
}

/* brdfdata.cpp:140-146: img = sat_u8(sat_u8(img - dark) - dark), the dark frame goes twice */
void oracle_subtract_ambient(unsigned char *img, const unsigned char *dark, long nbytes)
{
    long i;
    for (i = 0; i < nbytes; ++i) {
        int v = (int)img[i] - (int)dark[i];
        if (v < 0) v = 0;
        v -= (int)dark[i];
        if (v < 0) v = 0;
        img[i] = (unsigned char)v;
    }
}

/* Tsai pin-hole projection of a world point (kappa1 ignored: the reference never parses it,
 * brdfdata.cpp:195-247).  Returns 1 and the pixel when the point is in front of the camera and
 * inside the W x H image; image rows run top-down. */
static int project_tsai(const double *c, const double *cam, int W, int H, int *col, int *row)
{
    double d[3], xc, yc, zc, u, v;
    d[0] = c[0] - cam[CAM_P + 0];
    d[1] = c[1] - cam[CAM_P + 1];
    d[2] = c[2] - cam[CAM_P + 2];
    xc = dot3(d, cam + CAM_N);
    yc = dot3(d, cam + CAM_O);
    zc = dot3(d, cam + CAM_A);
    if (!(zc > 0.0)) return 0;
    u = cam[CAM_CX] + ((cam[CAM_SX] * cam[CAM_F]) * xc) / zc;
    v = cam[CAM_CY] + (cam[CAM_F] * yc) / zc;
    if (!(u >= 0.0 && v >= 0.0 && u < (double)W && v < (double)H)) return 0;
    *col = (int)u;
    *row = (int)v;
    return 1;
}

/* ---- options beyond the reference; flags == 0 is the reference's behaviour ----
 *   ORACLE_GATHER_DEPTH_TEST      a pixel goes to the face whose centroid is nearest to the camera (smallest
 *                                 zc = (C - p).a; equal depths: the later face), not simply to the last face
 *   ORACLE_GATHER_CULL_BACKFACES  faces seen from behind, N.(p - C) <= 0, are not mapped
 *   ORACLE_GATHER_KAPPA1          Tsai's radial lens distortion: the undistorted sensor point (Xu, Yu) =
 *                                 (f xc/zc, f yc/zc) is pulled to (Xd, Yd) with Xu = Xd (1 + kappa1 r^2),
 *                                 r^2 = Xd^2 + Yd^2, by FIVE fixed-point steps from (Xu, Yu); pixel =
 *                                 (cx + sx Xd, cy + Yd)
 * Returns 1, the pixel and the depth when the face is mapped. */
static int project_opts(const double *c, const double *N, const double *cam, double kappa1, int flags,
                        int W, int H, int *col, int *row, double *depth)
{
    double d[3], xc, yc, zc, u, v;
    d[0] = c[0] - cam[CAM_P + 0];
    d[1] = c[1] - cam[CAM_P + 1];
    d[2] = c[2] - cam[CAM_P + 2];
    xc = dot3(d, cam + CAM_N);
    yc = dot3(d, cam + CAM_O);
    zc = dot3(d, cam + CAM_A);
    if (!(zc > 0.0)) return 0;
    if (flags & ORACLE_GATHER_CULL_BACKFACES) {
        double toward[3];
        toward[0] = -d[0]; toward[1] = -d[1]; toward[2] = -d[2];
        if (!(dot3(N, toward) > 0.0)) return 0;
    }
    if (flags & ORACLE_GATHER_KAPPA1) {
        const double xu = (cam[CAM_F] * xc) / zc, yu = (cam[CAM_F] * yc) / zc;
        double xd = xu, yd = yu;
        int it;
        for (it = 0; it < 5; ++it) {
            const double s = 1.0 + kappa1 * (xd * xd + yd * yd);
            xd = xu / s;
            yd = yu / s;
        }
        u = cam[CAM_CX] + cam[CAM_SX] * xd;
        v = cam[CAM_CY] + yd;
    } else {
        u = cam[CAM_CX] + ((cam[CAM_SX] * cam[CAM_F]) * xc) / zc;
        v = cam[CAM_CY] + (cam[CAM_F] * yc) / zc;
    }
    if (!(u >= 0.0 && v >= 0.0 && u < (double)W && v < (double)H)) return 0;
    *col = (int)u;
    *row = (int)v;
    *depth = zc;
    return 1;
}

int oracle_calc_pixel2surface_opts(const double *V, const int *F, const double *FN, int nF, const double *cam,
                                   double kappa1, int flags, int W, int H, int *map)
{
    int i, hits = 0;
    double *zbuf = NULL;
    for (i = 0; i < W * H; ++i) map[i] = -1;
    if (flags & ORACLE_GATHER_DEPTH_TEST) {
        zbuf = (double *)malloc((size_t)W * H * sizeof(double));
        if (!zbuf) return -1;
    }
    for (i = 0; i < nF; ++i) {
        double c[3], z;
        int col, row;
        centroid(V, F, i, c);
        if (!project_opts(c, FN + (size_t)i * 3, cam, kappa1, flags, W, H, &col, &row, &z)) continue;
        ++hits;
        if (zbuf) {
            const int px = row * W + col;
            if (map[px] >= 0 && zbuf[px] < z) continue; /* something nearer is already there */
            zbuf[px] = z;
        }
        map[row * W + col] = i;
    }
    free(zbuf);
    return hits;
}

/* brdfdata.cpp:629-681: faces in index order, last writer wins, map starts at -1.
 * Returns the number of faces that landed inside the image. */
int oracle_calc_pixel2surface(const double *V, const int *F, int nF, const double *cam,
                              int W, int H, int *map)
{
    int i, hits = 0;
    for (i = 0; i < W * H; ++i) map[i] = -1;
    for (i = 0; i < nF; ++i) {
        double c[3];
        int col, row;
        centroid(V, F, i, c);
        if (project_tsai(c, cam, W, H, &col, &row)) {
            map[row * W + col] = i;
            ++hits;
        }
    }
    return hits;
}

/* brdfdata.cpp:857-899: cos(phi) = N . normalize(L_k - C) */
void oracle_cos_ln(const double *V, const int *F, const double *FN, const double *led, int nled,
                   int face, double *phi)
{
    int k;
    for (k = 0; k < nled; ++k) {
        double c[3], l[3];
        centroid(V, F, face, c);
        l[0] = led[k * 3 + 0] - c[0];
        l[1] = led[k * 3 + 1] - c[1];
        l[2] = led[k * 3 + 2] - c[2];
        normalize3(l);
        phi[k] = dot3_row(l, FN + (size_t)face * 3);
    }
}

/* brdfdata.cpp:902-943: cos(theta') = N . normalize(L_k - 2C + P_cam) */
void oracle_cos_nh(const double *V, const int *F, const double *FN, const double *led, int nled,
                   const double *cam, int face, double *thetaDash)
{
    int k;
    for (k = 0; k < nled; ++k) {
        double c[3], h[3];
        centroid(V, F, face, c);
        h[0] = led[k * 3 + 0] - 2 * c[0] + cam[CAM_P + 0];
        h[1] = led[k * 3 + 1] - 2 * c[1] + cam[CAM_P + 1];
        h[2] = led[k * 3 + 2] - 2 * c[2] + cam[CAM_P + 2];
        normalize3(h);
        thetaDash[k] = dot3_row(h, FN + (size_t)face * 3);
    }
}

/* brdfdata.cpp:799-855, reproduced literally including its two slips: the light
 * direction uses the centroid's x for all three components (:835) and the result is R.P, not R.V
 * (:849).  Only the Phong model (modelInfo 0) consumes it. */
void oracle_cos_rv(const double *V, const int *F, const double *FN, const double *led, int nled,
                   const double *cam, int face, double *theta)
{
    int k, a;
    (void)cam;
    for (k = 0; k < nled; ++k) {
        const double *N = FN + (size_t)face * 3;
        double c[3], l[3], P[3], R[3], s;
        centroid(V, F, face, c);
        l[0] = c[0] - led[k * 3 + 0];
        l[1] = c[0] - led[k * 3 + 1];
        l[2] = c[0] - led[k * 3 + 2];
        normalize3(l);
        s = dot3(N, l);
        for (a = 0; a < 3; ++a) P[a] = s * N[a];
        for (a = 0; a < 3; ++a) R[a] = l[a] - 2 * P[a];
        theta[k] = dot3(R, P);
    }
}

/* brdfdata.cpp:945-960 with the image row given directly (top-down; the GL flip H-1-y of :955
 * belongs to the literal GL projection, not to the Tsai one). BGR interleaved u8. */
void oracle_intensities_from_pixel(const unsigned char *const *images, int nimg, int W, int H,
                                   int x, int row, int channel, double *I)
{
    int k;
    (void)H;
    for (k = 0; k < nimg; ++k)
        I[k] = images[k][((size_t)row * W + x) * 3 + channel] / 255.0;
}

/* One camera, whole gather: map, then for every face that still owns its pixel (ascending face id)
 * the 3 x nimg cosines and the 3 x nimg intensities.  Sample s = fit*nimg + k.  `I` holds the three
 * BGR channels back to back with a channel stride of nF*nimg doubles. */
int oracle_gather(const double *V, const int *F, int nF, const double *cam, const double *led,
                  const unsigned char *const *images, int nimg, int W, int H,
                  int *map, int *fit_face, int *fit_pixel,
                  double *phi, double *thetaDash, double *theta, double *I)
{
    double *FN;
    int i, ch, nfit = 0;

    FN = (double *)malloc((size_t)nF * 3 * sizeof(double));
    if (!FN) return -1;
    oracle_face_normals(V, F, nF, FN);
    oracle_calc_pixel2surface(V, F, nF, cam, W, H, map);

    for (i = 0; i < nF; ++i) {
        double c[3];
        int col, row;
        centroid(V, F, i, c);
        if (!project_tsai(c, cam, W, H, &col, &row)) continue;
        if (map[row * W + col] != i) continue; /* a later face overwrote this pixel */
        fit_face[nfit] = i;
        fit_pixel[nfit] = row * W + col;
        oracle_cos_ln(V, F, FN, led, nimg, i, phi + (size_t)nfit * nimg);
        oracle_cos_nh(V, F, FN, led, nimg, cam, i, thetaDash + (size_t)nfit * nimg);
        oracle_cos_rv(V, F, FN, led, nimg, cam, i, theta + (size_t)nfit * nimg);
        for (ch = 0; ch < 3; ++ch)
            oracle_intensities_from_pixel(images, nimg, W, H, col, row, ch,
                                          I + (size_t)ch * nF * nimg + (size_t)nfit * nimg);
        ++nfit;
    }
    free(FN);
    return nfit;
}

/* glutcallbacks.cpp:346-445: the "show shaded BRDF" loop, one colour per face; the light source is the
 * eye (:352-353).  Operation order as written there: centroid by three additions and a division
 * (:356-366), lightDir / viewDir / h through Eigen's normalize() (:369-380), cwiseProduct().sum() dots,
 * `float cosRV` (:420), colour = kd*cosLN + ks*[coef*]pow(t, n) with coef = (n+2)/(2*CV_PI) for Phong
 * (:426-435 -- NOT the ((n+2)/2*pi) of BRDFFunc).  The literal cosLN of :385 passes the dot product as the
 * column argument of face_normals(i, .): it truncates to 0 for every |dot| < 1. */
void oracle_shade_faces(const double *V, const int *F, const double *FN, int nF, const double *eye,
                        const double *center, int model, int single, const double *brdf, int literal,
                        double *bgr)
{
    const double pi = 3.1415926535897932384626433832795;
    int i, k, ch;
    for (i = 0; i < nF; ++i) {
        const double *N = FN + (size_t)i * 3;
        double c[3], l[3], v[3], h[3], nl, cosLN, t;
        centroid(V, F, i, c);
        for (k = 0; k < 3; ++k) { l[k] = eye[k] - c[k]; v[k] = eye[k] - center[k]; }
        normalize3(l);
        normalize3(v);
        nl = dot3(N, l);
        cosLN = nl;
        if (literal) {
            int col = (int)nl;
            if (col < 0) col = 0;
            if (col > 2) col = 2;
            cosLN = N[col];
        }
        if (model == 1) {
            for (k = 0; k < 3; ++k) h[k] = l[k] + v[k];
            normalize3(h);
            t = dot3(N, h);
        } else {
            const double sf = -nl;
            double R[3];
            for (k = 0; k < 3; ++k) R[k] = l[k] - 2.0 * (sf * N[k]);
            t = (double)(float)dot3(v, R);
        }
        for (ch = 0; ch < 3; ++ch) {
            const double *q = brdf + (single ? (size_t)3 * ch : (size_t)9 * i + 3 * ch);
            const double pw = pow(t, q[2]);
            const double spec = (model == 1) ? q[1] * pw : (q[1] * ((q[2] + 2.0) / (2.0 * pi))) * pw;
            bgr[(size_t)i * 3 + ch] = q[0] * cosLN + spec;
        }
    }
}

/* oracle_gather with the options of oracle_calc_pixel2surface_opts */
int oracle_gather_opts(const double *V, const int *F, int nF, const double *cam, double kappa1, int flags,
                       const double *led, const unsigned char *const *images, int nimg, int W, int H,
                       int *map, int *fit_face, int *fit_pixel,
                       double *phi, double *thetaDash, double *theta, double *I)
{
    double *FN;
    int i, ch, nfit = 0;

    FN = (double *)malloc((size_t)nF * 3 * sizeof(double));
    if (!FN) return -1;
    oracle_face_normals(V, F, nF, FN);
    if (oracle_calc_pixel2surface_opts(V, F, FN, nF, cam, kappa1, flags, W, H, map) < 0) { free(FN); return -1; }

    for (i = 0; i < nF; ++i) {
        double c[3], z;
        int col, row;
        centroid(V, F, i, c);
        if (!project_opts(c, FN + (size_t)i * 3, cam, kappa1, flags, W, H, &col, &row, &z)) continue;
        if (map[row * W + col] != i) continue;
        fit_face[nfit] = i;
        fit_pixel[nfit] = row * W + col;
        oracle_cos_ln(V, F, FN, led, nimg, i, phi + (size_t)nfit * nimg);
        oracle_cos_nh(V, F, FN, led, nimg, cam, i, thetaDash + (size_t)nfit * nimg);
        oracle_cos_rv(V, F, FN, led, nimg, cam, i, theta + (size_t)nfit * nimg);
        for (ch = 0; ch < 3; ++ch)
            oracle_intensities_from_pixel(images, nimg, W, H, col, row, ch,
                                          I + (size_t)ch * nF * nimg + (size_t)nfit * nimg);
        ++nfit;
    }
    free(FN);
    return nfit;
}

/* ------------------------------------------------------------------------------------------------
 * The LITERAL projection of the reference: CalcPixel2SurfaceMapping reads the GL MODELVIEW / PROJECTION
 * matrices and the viewport back (brdfdata.cpp:662-669) and calls gluProject (:671).  gluProject lives in
 * libGLU (un-vendored, un-versioned system library; SGI's implementation as shipped by Mesa GLU 9.0,
 * src/libutil/project.c), restated here from its published source:
 *     out = M_modelview * (x, y, z, 1);  in = M_projection * out   (column-major matrices, each component
 *     in[0]*m[0*4+i] + in[1]*m[1*4+i] + in[2]*m[2*4+i] + in[3]*m[3*4+i], left to right)
 *     fail when in[3] == 0;  in[0..2] /= in[3];  in[k] = in[k] * 0.5 + 0.5;
 *     winx = in[0] * viewport[2] + viewport[0];  winy = in[1] * viewport[3] + viewport[1]
 * The reference then writes map.at<int>(winY, winX) = i whenever winY >= 0 && winX >= 0 (:676-677) -- WITHOUT an
 * upper bound (with the shipped matrices every such write lands outside the 800 x 600 map).  Here the
 * write is made only inside the map; rows are GL rows (bottom-up), and the radiance fetch flips them,
 * images[k](H-1-y, x) (brdfdata.cpp:955).
 * ------------------------------------------------------------------------------------------------ */
static void glu_mult_matrix_vec(const double *m, const double *in, double *out)
{
    int i;
    for (i = 0; i < 4; ++i)
        out[i] = in[0] * m[0 * 4 + i] + in[1] * m[1 * 4 + i] + in[2] * m[2 * 4 + i] + in[3] * m[3 * 4 + i];
}

static int project_gl(const double *c, const double *mv, const double *proj, const int *viewport, int W, int H,
                      int *col, int *row)
{
    double in[4], out[4], winx, winy;
    in[0] = c[0]; in[1] = c[1]; in[2] = c[2]; in[3] = 1.0;
    glu_mult_matrix_vec(mv, in, out);
    glu_mult_matrix_vec(proj, out, in);
    if (in[3] == 0.0) return 0;
    in[0] /= in[3]; in[1] /= in[3];
    in[0] = in[0] * 0.5 + 0.5;
    in[1] = in[1] * 0.5 + 0.5;
    winx = in[0] * viewport[2] + viewport[0];
    winy = in[1] * viewport[3] + viewport[1];
    if (!(winy >= 0 && winx >= 0)) return 0;                   /* brdfdata.cpp:676 */
    if (!(winx < (double)W && winy < (double)H)) return 0;     /* (the reference writes out of bounds here) */
    *col = (int)winx;
    *row = (int)winy;
    return 1;
}

int oracle_calc_pixel2surface_gl(const double *V, const int *F, int nF, const double *mv, const double *proj,
                                 const int *viewport, int W, int H, int *map)
{
    int i, hits = 0;
    for (i = 0; i < W * H; ++i) map[i] = -1;
    for (i = 0; i < nF; ++i) {
        double c[3];
        int col, row;
        centroid(V, F, i, c);
        if (project_gl(c, mv, proj, viewport, W, H, &col, &row)) {
            map[row * W + col] = i;
            ++hits;
        }
    }
    return hits;
}

/* the whole gather through the literal projection; cam supplies the camera position of GetCosNH (m_p) */
int oracle_gather_gl(const double *V, const int *F, int nF, const double *cam, const double *mv, const double *proj,
                     const int *viewport, const double *led, const unsigned char *const *images, int nimg, int W, int H,
                     int *map, int *fit_face, int *fit_pixel, double *phi, double *thetaDash, double *theta, double *I)
{
    double *FN;
    int i, ch, nfit = 0;
    FN = (double *)malloc((size_t)nF * 3 * sizeof(double));
    if (!FN) return -1;
    oracle_face_normals(V, F, nF, FN);
    oracle_calc_pixel2surface_gl(V, F, nF, mv, proj, viewport, W, H, map);
    for (i = 0; i < nF; ++i) {
        double c[3];
        int col, row;
        centroid(V, F, i, c);
        if (!project_gl(c, mv, proj, viewport, W, H, &col, &row)) continue;
        if (map[row * W + col] != i) continue;
        fit_face[nfit] = i;
        fit_pixel[nfit] = row * W + col;
        oracle_cos_ln(V, F, FN, led, nimg, i, phi + (size_t)nfit * nimg);
        oracle_cos_nh(V, F, FN, led, nimg, cam, i, thetaDash + (size_t)nfit * nimg);
        oracle_cos_rv(V, F, FN, led, nimg, cam, i, theta + (size_t)nfit * nimg);
        for (ch = 0; ch < 3; ++ch)   /* GetIntensities_FromPixel(x, y, c): row H-1-y, brdfdata.cpp:955 */
            oracle_intensities_from_pixel(images, nimg, W, H, col, H - 1 - row, ch,
                                          I + (size_t)ch * nF * nimg + (size_t)nfit * nimg);
        ++nfit;
    }
    free(FN);
    return nfit;
}

/* The matrices the reference's Display_ sets up before the mapping (glutcallbacks.cpp:626-642, 672-689): an
 * asymmetric frustum from the constant fields of view 78 / 49 degrees with the principal point of the .cal file,
 * offsets scaled by the WINDOW size, and gluLookAt(0,0,50, 0,0,0, 0,1,0).  GL keeps matrices in float32 and
 * glGetDoublev widens them: entries are computed in double here and rounded to float32 once; the
 * driver's own float arithmetic is not part of any contract. */
void oracle_reference_gl_matrices(double cx, double cy, int win_w, int win_h, double *mv, double *proj)
{
    const double DEG2RAD = 3.14159265 / 180, fov = 78, fovV = 49, front = 1.0, back = 1000.0;
    const double aspect = fov / fovV;
    const double tangent = tan(fovV / 2 * DEG2RAD), height = front * tangent, width = height * aspect;
    const double offset_y = 2.0 * (win_h / 2.0 - cy) / win_h, offset_x = 2.0 * (win_w / 2.0 - cx) / win_w;
    const double l = -width + offset_x, r = width + offset_x, b = -height - offset_y, t = height - offset_y;
    int i;
    for (i = 0; i < 16; ++i) { mv[i] = 0.0; proj[i] = 0.0; }
    proj[0] = 2 * front / (r - l); proj[5] = 2 * front / (t - b);
    proj[8] = (r + l) / (r - l); proj[9] = (t + b) / (t - b); proj[10] = -(back + front) / (back - front); proj[11] = -1.0;
    proj[14] = -2 * back * front / (back - front);
    mv[0] = mv[5] = mv[10] = mv[15] = 1.0;   /* looking down -z from (0,0,50) with +y up: a pure translation */
    mv[14] = -50.0;
    for (i = 0; i < 16; ++i) { mv[i] = (double)(float)mv[i]; proj[i] = (double)(float)proj[i]; }
}
