/*
 * brdfgpu.h -- C-ABI of the B200-native BRDF-fitting hot path (libbrdfgpu.so).
 *
 * Drop-in boundary for ONE path of ccalantzis/BRDF: sample gather -> BRDF model residual/Jacobian
 * -> Levenberg-Marquardt normal equations + small solve.  Plain pointers and sizes only; every
 * entry point names the reference interface it replaces (file:line relative to the reference
 * repository).  All computation runs in CUDA kernels compiled for sm_100a; there is no CPU
 * fallback: without a usable GPU the compute entry points return BRDFGPU_LM_ERROR and set
 * brdfgpu_last_error().
 *
 * Threading: like the reference's levmar (static LU scratch, levmar/levmar.h:36-42) a context is
 * not re-entrant; use one context per host thread / per GPU.
 */
#ifndef BRDFGPU_H
#define BRDFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* levmar/levmar.h:95-101 */
#define BRDFGPU_LM_OPTS_SZ 5
#define BRDFGPU_LM_INFO_SZ 10
#define BRDFGPU_LM_ERROR (-1)
#define BRDFGPU_LM_INIT_MU 1E-03
#define BRDFGPU_LM_STOP_THRESH 1E-17
#define BRDFGPU_LM_DIFF_DELTA 1E-06
/* The two BRDF models have exactly m == 3 parameters (kd, ks, n: BRDFFunc reads p[0..2] only,
 * brdfdata.cpp:980-986).  Every fit entry point on BRDF samples (sections 1-3) requires m == 3 and returns
 * BRDFGPU_LM_ERROR for any other m.  Only the reduced-evaluator control loop of section 6, which never sees a
 * model, is generic: 1 <= m <= BRDFGPU_MAX_PARAMS ("single-digit parameter count", BASELINE.json north_star). */
#define BRDFGPU_NUM_PARAMS 3
#define BRDFGPU_MAX_PARAMS 8

/* BRDF models, CBRDFdata::m_model (brdfdata.cpp:978,983; main.cpp:43 selects 1) */
#define BRDFGPU_MODEL_PHONG 0
#define BRDFGPU_MODEL_BLINN_PHONG 1

/* struct extraData, brdfdata.cpp:962-966: SoA angles = [cosphi(n) ; costhetadash(n) ; costheta(n)] */
typedef struct brdfgpu_extraData {
    double *angles;
    int modelInfo;
} brdfgpu_extraData;

typedef void (*brdfgpu_func_t)(double *p, double *hx, int m, int n, void *adata);
typedef void (*brdfgpu_jacf_t)(double *p, double *jac, int m, int n, void *adata);

/* ------------------------------------------------------------------------------------------------
 * 1. levmar-signature entry points (the two call sites brdfdata.cpp:1058 and :1119)
 * ------------------------------------------------------------------------------------------------ */

/* Replaces BRDFFunc (brdfdata.cpp:969-989).  Same signature and semantics: writes the model
 * prediction hx[0..n-1] for parameters p.  Evaluated by a CUDA kernel (angles are copied to the
 * GPU); its address is also the model selector the fit entry points below require as `func`. */
void brdfgpu_BRDFFunc(double *p, double *hx, int m, int n, void *adata);

/* Analytic Jacobian of BRDFFunc, row-major n x m (no reference counterpart; levmar jacf
 * signature, levmar/levmar.h:107).  Selector for brdfgpu_dlevmar_bc_der / _der. */
void brdfgpu_BRDFJac(double *p, double *jac, int m, int n, void *adata);

/* Replaces dlevmar_bc_dif (levmar/levmar.h:124-127, lmbc_core.c:1062-1129): identical argument
 * list, return value (#iterations or LM_ERROR) and info[0..9] meaning.  `func` must be
 * brdfgpu_BRDFFunc and `adata` a brdfgpu_extraData*; any other callback returns LM_ERROR (no CPU
 * path by design).  m must be BRDFGPU_NUM_PARAMS (3): levmar's signature is generic in m
 * (levmar/levmar.h:124-127) but the only callback this library accepts has three parameters; m != 3 returns
 * LM_ERROR with brdfgpu_last_error(NULL) = "... the BRDF models have exactly 3 parameters (kd, ks, n) ...",
 * checked before any GPU work.  n >= m.  `work` is accepted and ignored. */
int brdfgpu_dlevmar_bc_dif(brdfgpu_func_t func, double *p, double *x, int m, int n, double *lb,
                           double *ub, double *dscl, int itmax, double *opts, double *info,
                           double *work, double *covar, void *adata);

/* Replaces dlevmar_bc_der (levmar/levmar.h:118-122, lmbc_core.c:369-1022); jacf must be
 * brdfgpu_BRDFJac. */
int brdfgpu_dlevmar_bc_der(brdfgpu_func_t func, brdfgpu_jacf_t jacf, double *p, double *x, int m,
                           int n, double *lb, double *ub, double *dscl, int itmax, double *opts,
                           double *info, double *work, double *covar, void *adata);

/* Replaces dlevmar_dif (levmar/levmar.h:112-115, lm_core.c:438-842; the commented-out alternative
 * at brdfdata.cpp:1059,1120) and dlevmar_der (levmar.h:106-110, lm_core.c:64-432). */
int brdfgpu_dlevmar_dif(brdfgpu_func_t func, double *p, double *x, int m, int n, int itmax,
                        double *opts, double *info, double *work, double *covar, void *adata);
int brdfgpu_dlevmar_der(brdfgpu_func_t func, brdfgpu_jacf_t jacf, double *p, double *x, int m,
                        int n, int itmax, double *opts, double *info, double *work, double *covar,
                        void *adata);

/* ------------------------------------------------------------------------------------------------
 * 2. Contexts and device-resident sample sets (so benchmarks can exclude host<->device copies)
 * ------------------------------------------------------------------------------------------------ */
typedef struct brdfgpu_ctx brdfgpu_ctx;
typedef struct brdfgpu_samples brdfgpu_samples;
typedef struct brdfgpu_batch brdfgpu_batch;

/* device < 0: current CUDA device.  Returns 0 on success. */
int brdfgpu_create(int device, brdfgpu_ctx **out);
void brdfgpu_destroy(brdfgpu_ctx *ctx);
/* last error text of a context (ctx may be NULL for the process-default context / create errors) */
const char *brdfgpu_last_error(brdfgpu_ctx *ctx);
/* number of kernel launches this context has issued so far (bench.py's gpu_launches) */
unsigned long long brdfgpu_launch_count(brdfgpu_ctx *ctx);
/* What the last global fit of this context did (bench.py's roofline accounting):
 * out[0] sweeps over the samples that built a Jacobian, out[1] cost-only sweeps, out[2] trial
 * points evaluated by those (>= levmar's count: the projected-gradient walk is evaluated eight
 * candidates per sweep and the unused ones are discarded), out[3] samples held in shared memory for
 * the whole fit, out[4] CTAs of the persistent kernel (0 = host-driven fit), out[5..7] SM clock cycles CTA 0 spent
 * sweeping samples / in the grid-wide exchange / in the kernel altogether, out[8..11] the exchange by
 * phase, out[12..18] control-code cycles by the kind of sweep they led to, out[19] cost evaluations answered
 * by a Jacobian sweep (speculative Jacobians), out[20] Jacobians that were then obtained without a
 * sweep, out[21] iterations done in a single fused sweep, out[22..23] (secant fits, brdfgpu_fit_global_unc / _dlevmar_dif)
 * Broyden update passes launched / of those after an accepted step (up to 24 values). */
int brdfgpu_fit_stats(brdfgpu_ctx *ctx, unsigned long long *out, int count);
/* CUDA stream (cudaStream_t) the context launches on, for timing with CUDA events */
void *brdfgpu_stream(brdfgpu_ctx *ctx);
/* wait for everything queued on that stream */
int brdfgpu_synchronize(brdfgpu_ctx *ctx);

/* Sample set for a global fit: n samples (cosphi_i, t_i, x_i) with t = costhetadash (Blinn-Phong)
 * or costheta (Phong), i.e. the two angle blocks BRDFFunc reads for `model`
 * (brdfdata.cpp:977-986), plus the measurements.  Host pointers; x may be NULL (= zeros,
 * lmbc_core.c:373).  Device layout: three fp64 arrays, 24 B per sample per pass. */
int brdfgpu_samples_upload(brdfgpu_ctx *ctx, long n, const double *cosphi, const double *t,
                           const double *x, int model, brdfgpu_samples **out);
/* New values for an existing set of the same size (device buffers are reused: no allocation). */
int brdfgpu_samples_reload(brdfgpu_ctx *ctx, brdfgpu_samples *s, const double *cosphi, const double *t,
                           const double *x);
/* Same from DEVICE pointers (e.g. another library's tensors); data is copied device-to-device. */
int brdfgpu_samples_from_device(brdfgpu_ctx *ctx, long n, const double *d_cosphi, const double *d_t,
                                const double *d_x, int model, brdfgpu_samples **out);
/* Synthetic samples generated on the device (tests/synth.py recipe: counter-based splitmix64,
 * truth model + +-0.005 noise, 8-bit quantised), sample indices start .. start+n-1. */
int brdfgpu_samples_synth(brdfgpu_ctx *ctx, long n, unsigned long long seed, long start,
                          const double truth[3], int model, brdfgpu_samples **out);
long brdfgpu_samples_count(const brdfgpu_samples *s);
/* Copy the resident arrays back (any pointer may be NULL): raw cosines and measurements. */
int brdfgpu_samples_download(brdfgpu_ctx *ctx, const brdfgpu_samples *s, double *cosphi, double *t,
                             double *x);
void brdfgpu_samples_free(brdfgpu_ctx *ctx, brdfgpu_samples *s);

/* How the global fit is driven */
#define BRDFGPU_DRIVE_HOST 0       /* host control loop, one kernel per evaluation */
#define BRDFGPU_DRIVE_PERSISTENT 1 /* whole LM fit in one cooperative kernel, control loop on device */
/* Jacobian definition */
#define BRDFGPU_JAC_FD 0        /* finite differences exactly as levmar (forward, or central if opts[4]<0) */
#define BRDFGPU_JAC_ANALYTIC 1  /* exact partials (bc_der / der semantics) */
/* brdfgpu_batch_fit only: dlevmar_bc_dif + BRDFFunc reproduced operation by operation -- glibc's pow() bit for bit,
 * levmar's literal forward differences, its small-problem summation orders (lmbc_core.c:592-616, misc_core.c:721-807)
 * and its control arithmetic without multiply-add contraction -- so p, info[0..9] and the return value of every fit
 * EQUAL the reference's (not merely within tolerance), including the fits levmar abandons at itmax.  Forward
 * differences, at most 128 samples per fit (n m < 1024: levmar's small-problem branch).  1.5-1.9x the time of
 * BRDFGPU_JAC_FD, whose converged fits agree to the parity bars (1e-4 parameters, 1e-6 cost).  The reference's own
 * drivers below (brdfgpu_solve_equation, _solve_equation_batch, brdfgpu_calc_brdf_equation) use this mode. */
#define BRDFGPU_JAC_FD_EXACT 2

/* Global box-constrained fit on a resident sample set == dlevmar_bc_dif / _bc_der on the same data
 * (lmbc_core.c:369-1129).  With a communicator attached (brdfgpu_comm_init) the samples of all
 * ranks form ONE problem: every evaluation all-reduces the m(m+1)/2+m+2 partial sums. */
int brdfgpu_fit_global(brdfgpu_ctx *ctx, const brdfgpu_samples *s, double *p, int m,
                       const double *lb, const double *ub, const double *dscl, int itmax,
                       const double *opts, double *info, double *covar, int drive, int jac_mode);
/* Unconstrained variants == dlevmar_dif (secant, lm_core.c:438-842) / dlevmar_der (lm_core.c:64-432) */
int brdfgpu_fit_global_unc(brdfgpu_ctx *ctx, const brdfgpu_samples *s, double *p, int m, int itmax,
                           const double *opts, double *info, double *covar, int jac_mode);

/* Single evaluations (parity tests, roofline measurement).
 * residuals: e_i = x_i - f(p)_i for every sample (the vector levmar keeps in `e`, lmbc_core.c:526).
 * normal_eq: one fused residual + Jacobian + J^T J / J^T e pass at p; out[0..5] = JtJ upper triangle
 *            (00,01,02,11,12,22), out[6..8] = Jte, out[9] = ||e||^2, out[10] = #non-finite residuals.
 *            delta as opts[4] (negative = central); jac_mode as above.
 * cost:      out[0] = ||x - f(p)||^2, out[1] = #non-finite residuals. */
int brdfgpu_eval_residuals(brdfgpu_ctx *ctx, const brdfgpu_samples *s, const double *p, double *e);
int brdfgpu_eval_normal_eq(brdfgpu_ctx *ctx, const brdfgpu_samples *s, const double *p, double delta,
                           int jac_mode, double *out11);
int brdfgpu_eval_cost(brdfgpu_ctx *ctx, const brdfgpu_samples *s, const double *p, double *out2);
/* Launch `reps` back-to-back passes without host synchronisation in between (timing only):
 * kind 0 = normal_eq (FD forward), 1 = cost. */
int brdfgpu_eval_repeat(brdfgpu_ctx *ctx, const brdfgpu_samples *s, const double *p, double delta,
                        int kind, int reps);

/* ------------------------------------------------------------------------------------------------
 * 3. Reference fit drivers with raw pointers instead of Eigen / cv::Mat
 * ------------------------------------------------------------------------------------------------ */

/* CBRDFdata::SolveEquation (brdfdata.cpp:1077-1136): one per-face fit of nimg samples with the
 * reference preset p0=(0.5,1,1), bounds [0,100]^3, itmax 100, opts {1e-3,1e-15,1e-15,1e-20,1e-6}.
 * theta may be NULL for Blinn-Phong.  info may be NULL. */
int brdfgpu_solve_equation(const double *phi, const double *thetaDash, const double *theta,
                           const double *I, int nimg, int model, double *p, double *info);

/* CBRDFdata::SolveEquation_SingleBRDF (brdfdata.cpp:991-1075): one global fit over nsamples with
 * p0=(0,0,0), bounds [0,100]^3, itmax 2000, opts {1e-3,1e-15,1e-10,1e-50,delta=1}.  Samples are
 * used in the order given (SURVEY.md Q6 "aligned" order). */
int brdfgpu_solve_equation_single(const double *phi, const double *thetaDash, const double *theta,
                                  const double *I, long nsamples, int model, double *p, double *info);

/* The same with the reference's LITERAL flattening (brdfdata.cpp:1008-1042): phi / thetaDash / theta /
 * I are rows x nimg row-major (one row per face, as the reference's matrices); the measurements are taken row-major,
 * x[i*nimg + j] = I(i, j), the angle blocks through Eigen's linear (column-major) index, angles[k] = phi(k % rows, k / rows)
 * -- so sample k pairs the intensity of face k / nimg with the cosines of face k % rows.  Every row must hold values
 * (the reference leaves the rows of unmapped faces uninitialised).  For comparing against the reference as written;
 * brdfgpu_solve_equation_single is the aligned order. */
int brdfgpu_solve_equation_single_colmajor(const double *phi, const double *thetaDash, const double *theta,
                                           const double *I, long rows, int nimg, int model, double *p, double *info);

/* The per-pixel loop of CBRDFdata::CalcBRDFEquation (brdfdata.cpp:1195-1221) as one launch:
 * nfit independent dlevmar_bc_dif solves of nper samples each, fit f using rows
 * [f*nper, (f+1)*nper) of phi/thetaDash/theta/I.  p_out = nfit x 3, info_out = nfit x 10 (may be
 * NULL), ret_out = nfit levmar return values (may be NULL).  Options as brdfgpu_solve_equation
 * unless p0/opts/itmax are overridden through brdfgpu_batch_fit below.  For nper <= 128 the fits run in the
 * levmar-exact mode (BRDFGPU_JAC_FD_EXACT): p, info and ret equal dlevmar_bc_dif's bit for bit. */
int brdfgpu_solve_equation_batch(brdfgpu_ctx *ctx, long nfit, int nper, const double *phi,
                                 const double *thetaDash, const double *theta, const double *I,
                                 int model, double *p_out, double *info_out, int *ret_out);

/* Device-resident batched problem set */
int brdfgpu_batch_upload(brdfgpu_ctx *ctx, long nfit, int nper, const double *cosphi, const double *t,
                         const double *x, int model, brdfgpu_batch **out);
int brdfgpu_batch_synth(brdfgpu_ctx *ctx, long nfit, int nper, unsigned long long seed, long first_fit,
                        int model, brdfgpu_batch **out);
/* p0 (3), lb/ub (3, may be NULL), opts (5, may be NULL = levmar defaults).  Results stay on the
 * device until brdfgpu_batch_results. */
int brdfgpu_batch_fit(brdfgpu_ctx *ctx, brdfgpu_batch *b, const double *p0, const double *lb,
                      const double *ub, int itmax, const double *opts, int jac_mode);
int brdfgpu_batch_results(brdfgpu_ctx *ctx, const brdfgpu_batch *b, double *p_out, double *info_out,
                          int *ret_out);
long brdfgpu_batch_count(const brdfgpu_batch *b);
void brdfgpu_batch_free(brdfgpu_ctx *ctx, brdfgpu_batch *b);

/* ------------------------------------------------------------------------------------------------
 * 4. Sample gather (brdfdata.cpp:629-681, 799-960)
 * ------------------------------------------------------------------------------------------------ */
typedef struct brdfgpu_scene brdfgpu_scene;

/* camera = {cx, cy, f, sx, nx,ny,nz, ox,oy,oz, ax,ay,az, px,py,pz}: the .cal fields
 * CBRDFdata::WriteValue keeps (brdfdata.cpp:195-247) */
#define BRDFGPU_CAM_SZ 16

/* CBRDFdata::InitLEDs (brdfdata.cpp:683-756): the 16 hard-coded LED positions, 16 x 3 row-major */
void brdfgpu_led_table(double *led16x3);

/* Mesh + photographs resident on the device.  V: nV x 3 fp64 row-major, F: nF x 3 int32 vertex
 * ids (m_vertices/m_faces of brdfdata.h), images: nimg pointers to H x W x 3 u8 BGR (cv::imread
 * layout), dark: optional dark frame subtracted twice with saturation (brdfdata.cpp:130-147) on
 * the device, led: nimg x 3 LED positions (NULL = the reference table, brdfdata.cpp:683-756). */
int brdfgpu_scene_create(brdfgpu_ctx *ctx, const double *V, int nV, const int *F, int nF,
                         const unsigned char *const *images, int nimg, int W, int H,
                         const unsigned char *dark, const double *led, brdfgpu_scene **out);
void brdfgpu_scene_free(brdfgpu_ctx *ctx, brdfgpu_scene *sc);
/* dims5 = {nV, nF, nimg, W, H} of a scene */
int brdfgpu_scene_dims(const brdfgpu_scene *sc, int *dims5);
/* face normals as CalcFaceNormals (brdfdata.cpp:314-330), nF x 3 */
int brdfgpu_scene_face_normals(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, double *FN);
/* ambient-subtracted image k back to the host (H*W*3 bytes) */
int brdfgpu_scene_image(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, int k, unsigned char *out);

/* CBRDFdata::CalcPixel2SurfaceMapping (brdfdata.cpp:629-681) through the Tsai camera: map is
 * H x W int32, -1 = no face, else the LAST (= highest) face id whose centroid lands there. */
int brdfgpu_calc_pixel2surface(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, const double *cam, int *map);

/* Full gather for ncam cameras (GetCosLN/GetCosNH/GetCosRV + GetIntensities_FromPixel for every
 * face that owns its pixel, brdfdata.cpp:1195-1208 / 1147-1178).  Outputs (host, any may be NULL):
 *   maps      ncam x H x W int32
 *   nfit_cam  ncam counts; fits of camera v follow those of camera v-1, ascending face id inside
 *   fit_face / fit_pixel   total fits
 *   phi, thetaDash, theta  total fits x nimg
 *   I                      3 (B,G,R) x total fits x nimg
 * Returns the total number of fits (>= 0) or BRDFGPU_LM_ERROR.  `capacity` = rows the output
 * arrays can hold (<= ncam*nF always suffices); the channel stride of I is capacity*nimg. */
long brdfgpu_gather(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, const double *cams, int ncam,
                    long capacity, int *maps, long *nfit_cam, int *fit_face, int *fit_pixel,
                    double *phi, double *thetaDash, double *theta, double *I);

/* Options BEYOND the reference for every later gather / pixel-map call on this scene.
 * The reference maps a pixel to the LAST face whose centroid lands there, whether or not that face is
 * visible (no depth test, brdfdata.cpp:676-677), and never reads kappa1 (brdfdata.cpp:195-247); flags = 0
 * (the default) keeps exactly that.
 *   DEPTH_TEST      the pixel goes to the face whose centroid is nearest to the camera (equal depth: last face)
 *   CULL_BACKFACES  faces seen from behind (N.(p_cam - centroid) <= 0) are not mapped
 *   KAPPA1          Tsai's radial distortion Xu = Xd (1 + kappa1 r^2), one kappa1 per camera of the gather call
 *                   (the <kappa1> tag of the .cal files: brdfgpu_read_cal_kappa1), five fixed-point steps */
#define BRDFGPU_GATHER_DEPTH_TEST 1
#define BRDFGPU_GATHER_CULL_BACKFACES 2
#define BRDFGPU_GATHER_KAPPA1 4
/* Evaluation order of the final dot products of GetCosLN / GetCosNH (brdfdata.cpp:893, 937).  The default is what
 * Eigen 3.3 -- the version the reference's Makefile dependency list shows -- does for a fixed-size vector times a
 * row of a column-major MatrixXd: a0*b0 + (a1*b1 + a2*b2).  SEQ_DOT sums left to right, (a0*b0 + a1*b1) + a2*b2
 * (round 1's definition; differs in the last bit of some cosines).  oracle/gather_oracle.c has the derivation. */
#define BRDFGPU_GATHER_SEQ_DOT 8
int brdfgpu_scene_set_gather_options(brdfgpu_ctx *ctx, brdfgpu_scene *sc, int flags, const double *kappa1, int ncam);

/* The reference's LITERAL projection instead of the Tsai camera, for every later gather / pixel-map call on this
 * scene: CalcPixel2SurfaceMapping reads GL_MODELVIEW_MATRIX, GL_PROJECTION_MATRIX and GL_VIEWPORT back and calls
 * gluProject (brdfdata.cpp:662-671); pass those three arrays as the caller read them (column-major, 16 + 16 doubles,
 * 4 ints).  gluProject's arithmetic is libGLU's (SGI / Mesa GLU 9.0 project.c), reproduced operation by operation.
 * Map rows are GL rows (bottom-up) and the radiance fetch reads image row H-1-y (brdfdata.cpp:955).  The reference
 * writes map(winY, winX) for any winX, winY >= 0 with no upper bound (an out-of-bounds write on the shipped data,
 * SURVEY.md 2.4-Q1); here pixels outside the W x H map are dropped.  One view per call (one set of GL matrices); the
 * camera passed to the gather still supplies the position of GetCosNH.  Any pointer NULL: back to the Tsai camera. */
int brdfgpu_scene_set_gl_projection(brdfgpu_ctx *ctx, brdfgpu_scene *sc, const double *model_view16,
                                    const double *projection16, const int *viewport4);
/* The matrices the reference's Display_ leaves in GL before the mapping (glutcallbacks.cpp:626-642, 672-689): the
 * asymmetric frustum of the constant 78 / 49 degree fields of view around the .cal principal point (cx, cy), offsets
 * scaled by the WINDOW size (1920 x 1080 by default, main.cpp:22-23), near 1, far 1000, and
 * gluLookAt(0,0,50, 0,0,0, 0,1,0); entries rounded to float32 as GL stores them.  Host code. */
void brdfgpu_reference_gl_matrices(double cx, double cy, int window_width, int window_height, double *model_view16,
                                   double *projection16);

/* Gather that stays on the device and hands the samples straight to the fit stages. */
int brdfgpu_gather_resident(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, const double *cams, int ncam,
                            int model, int channel, brdfgpu_samples **global_out,
                            brdfgpu_batch **batch_out, long *nfit_out);

/* CBRDFdata::CalcBRDFEquation (brdfdata.cpp:1188-1227): gather + one per-face fit per mapped
 * face and colour channel.  brdf_surfaces: nF x 3 channels x 3 {kd, ks, n} (brdfdata.cpp:368-377;
 * untouched for unmapped faces).  Returns the number of fitted faces. */
long brdfgpu_calc_brdf_equation(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, const double *cam,
                                int model, double *brdf_surfaces);
/* CBRDFdata::CalcBRDFEquation_SingleBRDF (brdfdata.cpp:1138-1186): gather + one global fit per
 * channel; single_brdf = 3 channels x 3 raw p[0..2] (labels ignored), info = 3 x 10,
 * ret = 3 levmar return values. */
long brdfgpu_calc_brdf_equation_single(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, const double *cam,
                                       int model, double *single_brdf, double *info, int *ret);

/* The consumer of the fitted parameters: per-face colour of the BRDF-shaded preview
 * (glutcallbacks.cpp:346-445, the loop its own comment wants moved "auf grafikkarte").  The light sits at
 * the eye: lightDir = normalize(eye - centroid), viewDir = normalize(eye - center), h their normalised
 * sum; Blinn-Phong kd*cosLN + ks*pow(N.h, n), Phong kd*cosLN + ks*((n+2)/(2 pi))*pow((float)(viewDir.R), n).
 * brdf: single != 0 -> 3 channels x {kd, ks, n} (single_brdf), else nF x 3 x 3 (brdf_surfaces as
 * brdfgpu_calc_brdf_equation writes them).  literal_cosln != 0 keeps the reference's cosLN =
 * face_normals(i, (int)(N.lightDir)) (glutcallbacks.cpp:385: the dot product is used as a COLUMN INDEX, so
 * cosLN is the normal's x component whenever |N.lightDir| < 1; indices are clamped to 0..2); 0 uses the dot
 * product itself.  bgr_out: nF x 3 (B, G, R) fp64, the values the reference hands to glColor4d. */
int brdfgpu_shade_faces(brdfgpu_ctx *ctx, const brdfgpu_scene *sc, const double *eye, const double *center,
                        int model, int single, const double *brdf, int literal_cosln, double *bgr_out);

/* ------------------------------------------------------------------------------------------------
 * 4b. The reference's input files (host code, usable without a GPU)
 * ------------------------------------------------------------------------------------------------ */

/* CBRDFdata::LoadCameraParameters + WriteValue (brdfdata.cpp:149-247): the '<name>value</name>' scanner of
 * the reference, atof() on the value, the 16 fields of BRDFGPU_CAM_SZ kept (kappa1, camera_model and any
 * other tag ignored, as the reference ignores them).  Returns a bit mask of the fields that were present
 * (0xffff = all 16; absent fields are 0.0) or BRDFGPU_LM_ERROR when the file cannot be read. */
int brdfgpu_read_cal(const char *path, double *cam16);

/* The <kappa1> value of a .cal file, which the reference never reads (for BRDFGPU_GATHER_KAPPA1).
 * Returns 1 when the tag is present (else 0 and *kappa1 = 0) or BRDFGPU_LM_ERROR. */
int brdfgpu_read_cal_kappa1(const char *path, double *kappa1);

/* CBRDFdata::LoadModel -> igl::readOBJ (brdfdata.cpp:289-312): `v x y z` rows and the vertex index of the
 * first three corners of every `f` (v, v/vt, v//vn, v/vt/vn; 1-based or negative = relative), 0-based in F.
 * Two calls: V == NULL returns the counts in *nV / *nF; then V (nV x 3 fp64) and F (nF x 3 int32) sized
 * accordingly, with *nV / *nF holding those counts.  Returns 0 or BRDFGPU_LM_ERROR. */
int brdfgpu_read_obj(const char *path, double *V, int *F, int *nV, int *nF);

/* cv::imread(path, IMREAD_COLOR) for 8-bit non-interlaced PNG files (brdfdata.cpp:40, 122): H x W x 3 bytes,
 * B G R order, alpha dropped, grey replicated.  Two calls: bgr == NULL returns the size in *W / *H. */
int brdfgpu_read_png(const char *path, unsigned char *bgr, int *W, int *H);

/* The loading sequence of main.cpp:41-59 into a device-resident scene: LoadModel(obj_path),
 * LoadImages(image_folder) = image_folder + "1.png" .. "<nimg>.png" (the string is a prefix, as in the
 * reference: end it with '/'), SubtractAmbientLight with image_folder + "dark.png" when that file exists,
 * LoadCameraParameters(cal_path) into cam16 (cal_path may be NULL), InitLEDs.  PNG only: the reference's
 * current sources name ".jpeg" files that its repository does not contain. */
int brdfgpu_scene_load(brdfgpu_ctx *ctx, const char *image_folder, const char *obj_path, const char *cal_path,
                       int nimg, brdfgpu_scene **out, double *cam16);

/* ------------------------------------------------------------------------------------------------
 * 5. Multi-GPU (one process per GPU).  Global mode only needs it: batched fits and gather views
 *    shard with no communication.
 * ------------------------------------------------------------------------------------------------ */
#define BRDFGPU_UNIQUE_ID_BYTES 128
/* rank 0 creates the id, the launcher broadcasts it (torch.distributed / MPI / files) */
int brdfgpu_comm_unique_id(char *id128);
int brdfgpu_comm_init(brdfgpu_ctx *ctx, const char *id128, int rank, int nranks);
void brdfgpu_comm_destroy(brdfgpu_ctx *ctx);
/* all-reduce (sum, fp64) of a small host vector through the context's communicator: the exchange
 * step of the global fit, exposed for tests */
int brdfgpu_comm_allreduce(brdfgpu_ctx *ctx, double *buf, int count);

/* Fused in-kernel exchange: with peer buffers attached the persistent fit kernel all-reduces its
 * sums itself, by peer stores over NVLink / NVSwitch inside the same launch (no NCCL call, no host
 * round trip per evaluation).  Every rank exports one 72-byte record (a 64-byte CUDA-IPC handle + the
 * exchange tag its buffer has reached), the launcher gathers them (rank order) and every rank attaches
 * all of them.  One process per GPU.
 * Re-attaching (after brdfgpu_peer_detach, or after a fit was abandoned because a peer never delivered, which
 * leaves the ranks' tags out of step): every rank exports AGAIN and attaches the fresh records.  The new session
 * starts above the highest tag any rank ever wrote, so cells left in the buffers by the old session can never be
 * taken for fresh ones; no memset and no barrier between attach and the first fit are needed. */
#define BRDFGPU_IPC_HANDLE_BYTES 72
int brdfgpu_peer_export(brdfgpu_ctx *ctx, char *handle72);
int brdfgpu_peer_attach(brdfgpu_ctx *ctx, const char *handles /* nranks x 72 */, int rank, int nranks);
void brdfgpu_peer_detach(brdfgpu_ctx *ctx);

/* ------------------------------------------------------------------------------------------------
 * 6. The LM control loop on caller-supplied REDUCED evaluators (host logic of the global mode;
 *    lets the exact product control code run wherever the sums come from, e.g. CPU tests with
 *    world_size-2 gloo all-reduces).
 *      jac_cb : given p, fill JtJ (m x m row-major, full) and Jte (m) at p
 *      cost_cb: given p, return ||x - f(p)||^2 and set *nonfinite = #non-finite residuals
 * ------------------------------------------------------------------------------------------------ */
typedef void (*brdfgpu_reduced_jac_t)(const double *p, int m, double *JtJ, double *Jte, void *user);
typedef double (*brdfgpu_reduced_cost_t)(const double *p, int m, double *nonfinite, void *user);
int brdfgpu_lm_bc_reduced(brdfgpu_reduced_jac_t jac_cb, brdfgpu_reduced_cost_t cost_cb, void *user,
                          double *p, int m, long n, const double *lb, const double *ub,
                          const double *dscl, int itmax, const double *opts, double *info,
                          double *covar);
int brdfgpu_lm_unc_reduced(brdfgpu_reduced_jac_t jac_cb, brdfgpu_reduced_cost_t cost_cb, void *user,
                           double *p, int m, long n, int itmax, const double *opts, double *info,
                           double *covar);
/* Test hook: on != 0 makes brdfgpu_lm_bc_reduced hand the projected-gradient candidates to the
 * evaluator eight at a time, the way the persistent fit kernel receives them (results must not
 * change).  Returns the previous setting; after a batched run, the largest batch that occurred. */
int brdfgpu_lm_reduced_batching(int on);
/* the on-device 3x3..8x8 solve of dAx_eq_b_LU_noLapack (Axb_core.c:1140-1277), host instantiation */
int brdfgpu_Ax_eq_b_LU(const double *A, const double *B, double *x, int m);

/* library / build identification */
const char *brdfgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BRDFGPU_H */
