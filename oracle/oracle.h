/*
 * oracle.h -- CPU oracle for the BRDF-fitting hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's algorithm (ccalantzis/BRDF: brdfdata.cpp + the
 * vendored levmar 2.6) used exclusively as the checker by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under brdf_b200/ may include, link or
 * call it: the product path is CUDA only and fails loudly without its extension.
 *
 * Pinning: the solver half is checked bit-for-bit against the reference's own levmar compiled
 * unmodified into oracle/_ref/liblevmar_ref.so (tests/test_oracle_kat.py, tests/test_oracle_golden.py), against the known
 * answers of levmar/lmdemo.c (tests/test_oracle_kat.py) and against committed golden vectors that
 * oracle/_ref produced (tests/golden/, tests/golden/make_golden.py).  The gather half restates
 * brdfdata.cpp with the Tsai .cal projection; no reference test pins results
 * at that boundary and the C++ cannot be built here (OpenCV/Eigen/libigl/GL absent), so for the
 * gather this file is the definition: "parity unpinned" beyond the arithmetic cited per function.
 *
 * All reference citations are file:line relative to the reference repository root.
 */
#ifndef BRDF_ORACLE_H
#define BRDF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- levmar-style callbacks (levmar/levmar.h:106-127) ---- */
typedef void (*oracle_func_t)(double *p, double *hx, int m, int n, void *adata);
typedef void (*oracle_jacf_t)(double *p, double *jac, int m, int n, void *adata);

#define ORACLE_LM_ERROR   (-1)
#define ORACLE_LM_INFO_SZ 10
#define ORACLE_LM_OPTS_SZ 5

/* ---- solver (lm_oracle.c) ---- */
double oracle_L2nrmxmy(double *e, const double *x, const double *y, int n);          /* misc_core.c:721-807 */
void   oracle_trans_mat_mat_mult(const double *a, double *b, int n, int m);         /* misc_core.c:82-134  */
void   oracle_fdif_forw_jac(oracle_func_t func, double *p, double *hx, double *hxx,
                            double delta, double *jac, int m, int n, void *adata);  /* misc_core.c:137-172 */
void   oracle_fdif_cent_jac(oracle_func_t func, double *p, double *hxm, double *hxp,
                            double delta, double *jac, int m, int n, void *adata);  /* misc_core.c:175-211 */
int    oracle_Ax_eq_b_LU(const double *A, const double *B, double *x, int m);        /* Axb_core.c:1140-1277 */
int    oracle_covar(const double *JtJ, double *C, double sumsq, int m, int n);       /* misc_core.c:426-591 */

int oracle_dlevmar_bc_der(oracle_func_t func, oracle_jacf_t jacf, double *p, double *x, int m, int n,
                          double *lb, double *ub, double *dscl, int itmax, double *opts, double *info,
                          double *work, double *covar, void *adata);                /* lmbc_core.c:369-1022 */
int oracle_dlevmar_bc_dif(oracle_func_t func, double *p, double *x, int m, int n,
                          double *lb, double *ub, double *dscl, int itmax, double *opts, double *info,
                          double *work, double *covar, void *adata);                /* lmbc_core.c:1062-1129 */
int oracle_dlevmar_der(oracle_func_t func, oracle_jacf_t jacf, double *p, double *x, int m, int n,
                       int itmax, double *opts, double *info, double *work, double *covar,
                       void *adata);                                                /* lm_core.c:64-432 */
int oracle_dlevmar_dif(oracle_func_t func, double *p, double *x, int m, int n,
                       int itmax, double *opts, double *info, double *work, double *covar,
                       void *adata);                                                /* lm_core.c:438-842 */

/* ---- model callback (brdf_oracle.c; brdfdata.cpp:962-989) ---- */
struct oracle_extraData {
    double *angles;   /* SoA: [cosphi(0..n-1) ; costhetadash(n..2n-1) ; costheta(2n..3n-1)] */
    int modelInfo;    /* 0 = Phong, 1 = Blinn-Phong */
};
void oracle_BRDFFunc(double *p, double *x, int m, int n, void *data);
/* analytic Jacobian of the same model (not in the reference; used to check the analytic kernel mode) */
void oracle_BRDFJac(double *p, double *jac, int m, int n, void *data);

/* Fit drivers with the reference's two option presets (brdfdata.cpp:991-1075, 1077-1136). */
int oracle_solve_equation(const double *phi, const double *thetaDash, const double *theta,
                          const double *I, int nimg, int model, double *p_out, double *info_out);
int oracle_solve_equation_single(const double *phi, const double *thetaDash, const double *theta,
                                 const double *I, long nsamples, int model, double *p_out,
                                 double *info_out);

/* ---- gather (gather_oracle.c) ---- */
/* camera = {cx, cy, f, sx, nx,ny,nz, ox,oy,oz, ax,ay,az, px,py,pz}  (brdfdata.cpp:195-247) */
#define ORACLE_CAM_SZ 16
void oracle_led_table(double *led /*16x3 row-major*/);                               /* brdfdata.cpp:683-756 */
void oracle_face_normals(const double *V, const int *F, int nF, double *FN);         /* brdfdata.cpp:314-330 */
void oracle_subtract_ambient(unsigned char *img, const unsigned char *dark, long nbytes); /* :130-147 */
int  oracle_calc_pixel2surface(const double *V, const int *F, int nF, const double *cam,
                               int W, int H, int *map);                              /* brdfdata.cpp:629-681 */
void oracle_cos_ln(const double *V, const int *F, const double *FN, const double *led, int nled,
                   int face, double *phi);                                           /* brdfdata.cpp:857-899 */
void oracle_cos_nh(const double *V, const int *F, const double *FN, const double *led, int nled,
                   const double *cam, int face, double *thetaDash);                  /* brdfdata.cpp:902-943 */
void oracle_cos_rv(const double *V, const int *F, const double *FN, const double *led, int nled,
                   const double *cam, int face, double *theta);                      /* brdfdata.cpp:799-855 */
void oracle_intensities_from_pixel(const unsigned char *const *images, int nimg, int W, int H,
                                   int x, int row, int channel, double *I);          /* brdfdata.cpp:945-960 */
/* Full gather for one camera: returns the number of mapped faces (fits); arrays sized for nF. */
int  oracle_gather(const double *V, const int *F, int nF, const double *cam, const double *led,
                   const unsigned char *const *images, int nimg, int W, int H,
                   int *map, int *fit_face, int *fit_pixel,
                   double *phi, double *thetaDash, double *theta, double *I /* [3][nfit*nimg] */);

/* Evaluation order of the Eigen reductions in GetCosLN / GetCosNH (gather_oracle.c header): the canonical
 * order is Eigen 3.3's, a0*b0 + (a1*b1 + a2*b2); ORACLE_DOT_SEQUENTIAL = (a0*b0 + a1*b1) + a2*b2 */
#define ORACLE_DOT_EIGEN33 0
#define ORACLE_DOT_SEQUENTIAL 1
void oracle_set_dot_order(int order);

/* Options beyond the reference (depth test, back-face culling, Tsai kappa1); flags 0 == the functions above */
#define ORACLE_GATHER_DEPTH_TEST 1
#define ORACLE_GATHER_CULL_BACKFACES 2
#define ORACLE_GATHER_KAPPA1 4
int  oracle_calc_pixel2surface_opts(const double *V, const int *F, const double *FN, int nF, const double *cam,
                                    double kappa1, int flags, int W, int H, int *map);
int  oracle_gather_opts(const double *V, const int *F, int nF, const double *cam, double kappa1, int flags,
                        const double *led, const unsigned char *const *images, int nimg, int W, int H,
                        int *map, int *fit_face, int *fit_pixel,
                        double *phi, double *thetaDash, double *theta, double *I);

/* The reference's LITERAL projection (gluProject through the GL matrices, brdfdata.cpp:662-677) */
int  oracle_calc_pixel2surface_gl(const double *V, const int *F, int nF, const double *mv, const double *proj,
                                  const int *viewport, int W, int H, int *map);
int  oracle_gather_gl(const double *V, const int *F, int nF, const double *cam, const double *mv, const double *proj,
                      const int *viewport, const double *led, const unsigned char *const *images, int nimg, int W, int H,
                      int *map, int *fit_face, int *fit_pixel, double *phi, double *thetaDash, double *theta, double *I);
void oracle_reference_gl_matrices(double cx, double cy, int win_w, int win_h, double *mv, double *proj);

/* BRDF-shaded preview colours per face, (B, G, R) x nF (glutcallbacks.cpp:346-445).  literal != 0 keeps the
 * reference's cosLN = face_normals(i, (int)(N.lightDir)) (column index clamped to 0..2). */
void oracle_shade_faces(const double *V, const int *F, const double *FN, int nF, const double *eye,
                        const double *center, int model, int single, const double *brdf, int literal,
                        double *bgr);

#ifdef __cplusplus
}
#endif
#endif
