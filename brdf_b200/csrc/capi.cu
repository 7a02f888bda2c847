// capi.cu -- the C-ABI of include/brdfgpu.h: contexts, the levmar-signature entry points, resident
// sample sets, the reference's fit drivers with raw pointers, and the reduced-evaluator LM loop.
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace brdfgpu;

namespace brdfgpu {

static std::string g_create_error;

void set_error(brdfgpu_ctx* ctx, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    fprintf(stderr, "brdfgpu: %s\n", msg.c_str());
}

brdfgpu_ctx* default_ctx() {
    static brdfgpu_ctx* ctx = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!ctx && brdfgpu_create(-1, &ctx) != 0) ctx = nullptr;
    return ctx;
}

}  // namespace brdfgpu

static brdfgpu_ctx* ctx_or_default(brdfgpu_ctx* ctx) { return ctx ? ctx : default_ctx(); }

extern "C" const char* brdfgpu_version(void) { return "brdfgpu 0.1 (sm_100a)"; }

extern "C" int brdfgpu_create(int device, brdfgpu_ctx** out) {
    if (!out) return BRDFGPU_LM_ERROR;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1) {
        // no CPU fallback by design: the library is CUDA only
        set_error(nullptr, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
        return BRDFGPU_LM_ERROR;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0;
    if (device >= count) {
        set_error(nullptr, "device index out of range");
        return BRDFGPU_LM_ERROR;
    }
    brdfgpu_ctx* ctx = new brdfgpu_ctx;
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        // The fit kernels keep ~1 KB of per-thread stack; without this flag the driver shrinks the
        // local-memory pool again after every such launch and re-grows it for the next one (a
        // synchronising reallocation of tens of milliseconds -- host calls were 10-50x slower).
        unsigned flags = 0;
        if (cudaGetDeviceFlags(&flags) != cudaSuccess) { cudaGetLastError(); flags = 0; }
        if (!(flags & cudaDeviceLmemResizeToMax) && cudaSetDeviceFlags(flags | cudaDeviceLmemResizeToMax) != cudaSuccess)
            cudaGetLastError();  // an application that fixed its flags earlier keeps them
    }
    if (e == cudaSuccess) {
        // the gather allocates its temporaries with cudaMallocAsync: let the device's default pool keep them
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) cudaGetLastError();
        } else {
            cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->coop, cudaDevAttrCooperativeLaunch, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_partials, sizeof(double) * 2 * kMaxPassBlocks * kResultDoubles);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_sync, sizeof(unsigned) * 16);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_sync, 0, sizeof(unsigned) * 16);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_result, sizeof(double) * kResultDoubles);
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_result, sizeof(double) * kResultDoubles, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&ctx->h_result_dev, ctx->h_result, 0);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_seq, sizeof(unsigned long long), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *ctx->h_seq = 0;
        e = cudaHostGetDevicePointer(&ctx->h_seq_dev, (void*)ctx->h_seq, 0);
    }
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_cells, sizeof(uint4) * 2 * 160 * 32);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_fitio, sizeof(GlobalFitOut));
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_fitio, sizeof(GlobalFitOut), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_counts, sizeof(int) * kCountStagingInts, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_error(nullptr, std::string("context creation failed: ") + cudaGetErrorString(e));
        brdfgpu_destroy(ctx);
        return BRDFGPU_LM_ERROR;
    }
    *out = ctx;
    return 0;
}

extern "C" void brdfgpu_destroy(brdfgpu_ctx* ctx) {
    if (!ctx) return;
    brdfgpu_comm_destroy(ctx);
    cudaSetDevice(ctx->device);
    if (ctx->pooled) {
        ctx->pooled->pooled = false;
        brdfgpu_samples_free(ctx, ctx->pooled);
        ctx->pooled = nullptr;
    }
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    brdfgpu_peer_detach(ctx);
    cudaFree(ctx->peer_local);
    ctx->peer_local = nullptr;
    cudaFree(ctx->d_partials); cudaFree(ctx->d_sync); cudaFree(ctx->d_result); cudaFree(ctx->d_fitio); cudaFree(ctx->d_cells);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->h_seq) cudaFreeHost((void*)ctx->h_seq);
    if (ctx->h_fitio) cudaFreeHost(ctx->h_fitio);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* brdfgpu_last_error(brdfgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" unsigned long long brdfgpu_launch_count(brdfgpu_ctx* ctx) {
    ctx = ctx_or_default(ctx);
    return ctx ? ctx->launches : 0ull;
}
extern "C" int brdfgpu_fit_stats(brdfgpu_ctx* ctx, unsigned long long* out, int count) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !out) return BRDFGPU_LM_ERROR;
    for (int i = 0; i < count && i < 24; ++i) out[i] = ctx->fit_stats[i];
    return 0;
}
extern "C" void* brdfgpu_stream(brdfgpu_ctx* ctx) {
    ctx = ctx_or_default(ctx);
    return ctx ? (void*)ctx->stream : nullptr;
}
extern "C" int brdfgpu_synchronize(brdfgpu_ctx* ctx) {
    ctx = ctx_or_default(ctx);
    if (!ctx) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// resident sample sets
// ------------------------------------------------------------------------------------------------
static int samples_fill(brdfgpu_ctx* ctx, brdfgpu_samples* s, const double* c, const double* t, const double* x,
                        cudaMemcpyKind kind, bool wait_for_copies = true) {
    const size_t nb = sizeof(double) * (size_t)s->n;
    if (s->n == 0) return 0;
    BG_CUDA_OK(ctx, cudaMemcpyAsync(s->c, c, nb, kind, ctx->stream));
    BG_CUDA_OK(ctx, cudaMemcpyAsync(s->traw, t, nb, kind, ctx->stream));
    if (x) BG_CUDA_OK(ctx, cudaMemcpyAsync(s->x, x, nb, kind, ctx->stream));
    else BG_CUDA_OK(ctx, cudaMemsetAsync(s->x, 0, nb, ctx->stream));  // x == NULL: zeros, lmbc_core.c:373
    if (samples_prepare(ctx, s) != 0) return BRDFGPU_LM_ERROR;
    // host buffers may be reused by the caller as soon as we return: wait for the copies (the log
    // pass behind them is a few microseconds) -- unless the caller is itself a synchronous entry point
    // that queues the fit right behind them and waits for everything before IT returns
    if (kind == cudaMemcpyHostToDevice && wait_for_copies) BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Sample set backed by the context's reusable buffers (one user at a time; falls back to a fresh
// allocation when busy).
static int samples_upload_pooled(brdfgpu_ctx* ctx, long n, const double* cosphi, const double* t, const double* x,
                                 int model, brdfgpu_samples** out, bool wait_for_copies = true) {
    if (ctx->pooled_busy || n <= 0) return brdfgpu_samples_upload(ctx, n, cosphi, t, x, model, out);
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (!ctx->pooled || ctx->pooled_capacity < n) {
        if (ctx->pooled) {
            ctx->pooled->pooled = false;
            brdfgpu_samples_free(ctx, ctx->pooled);
            ctx->pooled = nullptr;
            ctx->pooled_capacity = 0;
        }
        const long cap = n + n / 4 + 1024;
        if (samples_alloc(ctx, cap, model, &ctx->pooled) != 0) return BRDFGPU_LM_ERROR;
        ctx->pooled_capacity = cap;
        ctx->pooled->pooled = true;
    }
    brdfgpu_samples* s = ctx->pooled;
    s->n = n;
    s->model = model;
    ctx->pooled_busy = true;
    if (samples_fill(ctx, s, cosphi, t, x, cudaMemcpyHostToDevice, wait_for_copies) != 0) {
        ctx->pooled_busy = false;
        return BRDFGPU_LM_ERROR;
    }
    *out = s;
    return 0;
}

extern "C" int brdfgpu_samples_upload(brdfgpu_ctx* ctx, long n, const double* cosphi, const double* t, const double* x,
                                      int model, brdfgpu_samples** out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !out || (n > 0 && (!cosphi || !t))) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_samples* s = nullptr;
    if (samples_alloc(ctx, n, model, &s) != 0) return BRDFGPU_LM_ERROR;
    if (samples_fill(ctx, s, cosphi, t, x, cudaMemcpyHostToDevice) != 0) {
        brdfgpu_samples_free(ctx, s);
        return BRDFGPU_LM_ERROR;
    }
    *out = s;
    return 0;
}

extern "C" int brdfgpu_samples_reload(brdfgpu_ctx* ctx, brdfgpu_samples* s, const double* cosphi, const double* t,
                                      const double* x) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s || (s->n > 0 && (!cosphi || !t))) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return samples_fill(ctx, s, cosphi, t, x, cudaMemcpyHostToDevice);
}

extern "C" int brdfgpu_samples_from_device(brdfgpu_ctx* ctx, long n, const double* d_cosphi, const double* d_t,
                                           const double* d_x, int model, brdfgpu_samples** out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !out || (n > 0 && (!d_cosphi || !d_t))) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_samples* s = nullptr;
    if (samples_alloc(ctx, n, model, &s) != 0) return BRDFGPU_LM_ERROR;
    if (samples_fill(ctx, s, d_cosphi, d_t, d_x, cudaMemcpyDeviceToDevice) != 0) {
        brdfgpu_samples_free(ctx, s);
        return BRDFGPU_LM_ERROR;
    }
    *out = s;
    return 0;
}

extern "C" int brdfgpu_samples_synth(brdfgpu_ctx* ctx, long n, unsigned long long seed, long start, const double truth[3],
                                     int model, brdfgpu_samples** out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !out || !truth) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_samples* s = nullptr;
    if (samples_alloc(ctx, n, model, &s) != 0) return BRDFGPU_LM_ERROR;
    if (synth_samples(ctx, s, seed, start, truth) != 0 || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        brdfgpu_samples_free(ctx, s);
        return BRDFGPU_LM_ERROR;
    }
    *out = s;
    return 0;
}

extern "C" long brdfgpu_samples_count(const brdfgpu_samples* s) { return s ? s->n : 0; }

extern "C" int brdfgpu_samples_download(brdfgpu_ctx* ctx, const brdfgpu_samples* s, double* cosphi, double* t, double* x) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s) return BRDFGPU_LM_ERROR;
    const size_t nb = sizeof(double) * (size_t)s->n;
    if (cosphi) BG_CUDA_OK(ctx, cudaMemcpyAsync(cosphi, s->c, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (t) BG_CUDA_OK(ctx, cudaMemcpyAsync(t, s->traw, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (x) BG_CUDA_OK(ctx, cudaMemcpyAsync(x, s->x, nb, cudaMemcpyDeviceToHost, ctx->stream));
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" void brdfgpu_samples_free(brdfgpu_ctx* ctx, brdfgpu_samples* s) {
    if (!s) return;
    if (s->pooled) {  // back to its context
        ctx = ctx_or_default(ctx);
        if (ctx && ctx->pooled == s) ctx->pooled_busy = false;
        return;
    }
    free_block(ctx, s->block, s->stream);
    cudaFree(s->jac);
    delete s;
}

// ------------------------------------------------------------------------------------------------
// global fits and single evaluations on resident samples
// ------------------------------------------------------------------------------------------------
extern "C" int brdfgpu_fit_global(brdfgpu_ctx* ctx, const brdfgpu_samples* s, double* p, int m, const double* lb,
                                  const double* ub, const double* dscl, int itmax, const double* opts, double* info,
                                  double* covar, int drive, int jac_mode) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s || !p) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return global_fit(ctx, s, p, m, lb, ub, dscl, itmax, opts, info, covar, drive, jac_mode, 0);
}

extern "C" int brdfgpu_fit_global_unc(brdfgpu_ctx* ctx, const brdfgpu_samples* s, double* p, int m, int itmax,
                                      const double* opts, double* info, double* covar, int jac_mode) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s || !p) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (jac_mode == BRDFGPU_JAC_ANALYTIC)
        return global_fit(ctx, s, p, m, nullptr, nullptr, nullptr, itmax, opts, info, covar, BRDFGPU_DRIVE_PERSISTENT,
                          jac_mode, 1);
    return global_fit_secant(ctx, const_cast<brdfgpu_samples*>(s), p, m, itmax, opts, info, covar);
}

extern "C" int brdfgpu_eval_residuals(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double* e) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s || !p || !e) return BRDFGPU_LM_ERROR;
    return global_residuals(ctx, s, p, e);
}
extern "C" int brdfgpu_eval_normal_eq(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta,
                                      int jac_mode, double* out11) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s || !p || !out11) return BRDFGPU_LM_ERROR;
    return global_normal_eq(ctx, s, p, delta, jac_mode, out11);
}
extern "C" int brdfgpu_eval_cost(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double* out2) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s || !p || !out2) return BRDFGPU_LM_ERROR;
    return global_cost(ctx, s, p, out2);
}
extern "C" int brdfgpu_eval_repeat(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta, int kind,
                                   int reps) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !s || !p) return BRDFGPU_LM_ERROR;
    return global_repeat(ctx, s, p, delta, kind, reps);
}

// ------------------------------------------------------------------------------------------------
// levmar-signature entry points
// ------------------------------------------------------------------------------------------------
extern "C" void brdfgpu_BRDFFunc(double* p, double* hx, int m, int n, void* adata) {
    (void)m;
    brdfgpu_ctx* ctx = default_ctx();
    const brdfgpu_extraData* d = static_cast<const brdfgpu_extraData*>(adata);
    if (!ctx || !d || !d->angles) {
        fprintf(stderr, "brdfgpu_BRDFFunc: no CUDA device or no data -- hx left untouched (there is no CPU path)\n");
        return;
    }
    model_predict(ctx, p, d->angles, d->modelInfo, n, hx);
}

extern "C" void brdfgpu_BRDFJac(double* p, double* jac, int m, int n, void* adata) {
    brdfgpu_ctx* ctx = default_ctx();
    const brdfgpu_extraData* d = static_cast<const brdfgpu_extraData*>(adata);
    if (!ctx || !d || !d->angles || m < 3) {
        fprintf(stderr, "brdfgpu_BRDFJac: no CUDA device or bad arguments -- jac left untouched\n");
        return;
    }
    model_jacobian(ctx, p, d->angles, d->modelInfo, n, m, jac);
}

// shared body of the four levmar-signature fits
static int levmar_entry(const char* name, brdfgpu_func_t func, brdfgpu_jacf_t jacf, bool need_jacf, double* p, double* x,
                        int m, int n, double* lb, double* ub, double* dscl, int itmax, double* opts, double* info,
                        double* covar, void* adata, bool constrained) {
    if (func != brdfgpu_BRDFFunc || (need_jacf && jacf != brdfgpu_BRDFJac)) {
        fprintf(stderr, "%s: only the brdfgpu_BRDFFunc / brdfgpu_BRDFJac callbacks are supported (GPU path, no CPU fallback)\n",
                name);
        return BRDFGPU_LM_ERROR;
    }
    if (m != BRDFGPU_NUM_PARAMS) {  // levmar is generic in m (levmar.h:124-127); BRDFFunc is not (brdfdata.cpp:980-986)
        set_error(nullptr, std::string(name) + ": the BRDF models have exactly 3 parameters (kd, ks, n); m = " + std::to_string(m) +
                               " is not supported");
        return BRDFGPU_LM_ERROR;
    }
    const brdfgpu_extraData* d = static_cast<const brdfgpu_extraData*>(adata);
    if (!d || !d->angles || !p) {
        fprintf(stderr, "%s: adata must point to a brdfgpu_extraData with angles\n", name);
        return BRDFGPU_LM_ERROR;
    }
    if (n < m) {  // lmbc_core.c:440-443 / lm_core.c:113-116
        fprintf(stderr, "%s: cannot solve a problem with fewer measurements [%d] than unknowns [%d]\n", name, n, m);
        return BRDFGPU_LM_ERROR;
    }
    if (d->modelInfo != 0 && d->modelInfo != 1) {
        fprintf(stderr, "%s: unknown model id %d\n", name, d->modelInfo);
        return BRDFGPU_LM_ERROR;
    }
    brdfgpu_ctx* ctx = default_ctx();
    if (!ctx) return BRDFGPU_LM_ERROR;
    brdfgpu_samples* s = nullptr;
    const double* t = d->angles + (d->modelInfo == 1 ? (size_t)n : 2 * (size_t)n);
    // the copies, the log pass and the fit are queued back to back; the caller's buffers are released by the wait below
    if (samples_upload_pooled(ctx, n, d->angles, t, x, d->modelInfo, &s, /*wait_for_copies=*/false) != 0) return BRDFGPU_LM_ERROR;
    int ret;
    const int jm = need_jacf ? BRDFGPU_JAC_ANALYTIC : BRDFGPU_JAC_FD;
    if (constrained) ret = brdfgpu_fit_global(ctx, s, p, m, lb, ub, dscl, itmax, opts, info, covar, BRDFGPU_DRIVE_PERSISTENT, jm);
    else ret = brdfgpu_fit_global_unc(ctx, s, p, m, itmax, opts, info, covar, jm);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) ret = BRDFGPU_LM_ERROR;  // (a fit that was refused never waited)
    brdfgpu_samples_free(ctx, s);
    return ret;
}

extern "C" int brdfgpu_dlevmar_bc_dif(brdfgpu_func_t func, double* p, double* x, int m, int n, double* lb, double* ub,
                                      double* dscl, int itmax, double* opts, double* info, double* work, double* covar,
                                      void* adata) {
    (void)work;
    return levmar_entry("brdfgpu_dlevmar_bc_dif", func, nullptr, false, p, x, m, n, lb, ub, dscl, itmax, opts, info, covar,
                        adata, true);
}
extern "C" int brdfgpu_dlevmar_bc_der(brdfgpu_func_t func, brdfgpu_jacf_t jacf, double* p, double* x, int m, int n,
                                      double* lb, double* ub, double* dscl, int itmax, double* opts, double* info,
                                      double* work, double* covar, void* adata) {
    (void)work;
    return levmar_entry("brdfgpu_dlevmar_bc_der", func, jacf, true, p, x, m, n, lb, ub, dscl, itmax, opts, info, covar,
                        adata, true);
}
extern "C" int brdfgpu_dlevmar_dif(brdfgpu_func_t func, double* p, double* x, int m, int n, int itmax, double* opts,
                                   double* info, double* work, double* covar, void* adata) {
    (void)work;
    return levmar_entry("brdfgpu_dlevmar_dif", func, nullptr, false, p, x, m, n, nullptr, nullptr, nullptr, itmax, opts,
                        info, covar, adata, false);
}
extern "C" int brdfgpu_dlevmar_der(brdfgpu_func_t func, brdfgpu_jacf_t jacf, double* p, double* x, int m, int n, int itmax,
                                   double* opts, double* info, double* work, double* covar, void* adata) {
    (void)work;
    return levmar_entry("brdfgpu_dlevmar_der", func, jacf, true, p, x, m, n, nullptr, nullptr, nullptr, itmax, opts, info,
                        covar, adata, false);
}

// ------------------------------------------------------------------------------------------------
// the reference's fit drivers (brdfdata.cpp:991-1136) with raw pointers
// ------------------------------------------------------------------------------------------------
extern "C" int brdfgpu_solve_equation(const double* phi, const double* thetaDash, const double* theta, const double* I,
                                      int nimg, int model, double* p, double* info) {
    int ret = BRDFGPU_LM_ERROR;
    if (brdfgpu_solve_equation_batch(nullptr, 1, nimg, phi, thetaDash, theta, I, model, p, info, &ret) != 0)
        return BRDFGPU_LM_ERROR;
    return ret;
}

extern "C" int brdfgpu_solve_equation_single(const double* phi, const double* thetaDash, const double* theta,
                                             const double* I, long nsamples, int model, double* p, double* info) {
    brdfgpu_ctx* ctx = default_ctx();
    if (!ctx || !phi || !I || !p) return BRDFGPU_LM_ERROR;
    const double* t = model == 1 ? thetaDash : theta;
    if (!t) return BRDFGPU_LM_ERROR;
    static const double lb[3] = {0, 0, 0}, ub[3] = {100, 100, 100};
    static const double opts[5] = {1E-03, 1E-15, 1E-10, 1E-50, 1.0};  // brdfdata.cpp:1055-1056
    brdfgpu_samples* s = nullptr;
    if (samples_upload_pooled(ctx, nsamples, phi, t, I, model, &s) != 0) return BRDFGPU_LM_ERROR;
    p[0] = p[1] = p[2] = 0.0;  // brdfdata.cpp:1002
    const int ret = brdfgpu_fit_global(ctx, s, p, 3, lb, ub, nullptr, 2000, opts, info, nullptr, BRDFGPU_DRIVE_PERSISTENT,
                                       BRDFGPU_JAC_FD);
    brdfgpu_samples_free(ctx, s);
    return ret;
}

// SolveEquation_SingleBRDF exactly as the reference flattens its matrices (brdfdata.cpp:1008-1042):
// the measurements row-major, x[i * nimg + j] = I(i, j), but the three angle blocks through Eigen's LINEAR index of a
// column-major MatrixXd, angles[k] = phi(k) = phi(k % rows, k / rows).  Sample k therefore pairs I(k / nimg, k % nimg)
// with the cosines of face k % rows and LED k / rows.
extern "C" int brdfgpu_solve_equation_single_colmajor(const double* phi, const double* thetaDash, const double* theta,
                                                      const double* I, long rows, int nimg, int model, double* p,
                                                      double* info) {
    if (!phi || !I || !p || rows < 1 || nimg < 1) return BRDFGPU_LM_ERROR;
    const double* t = model == 1 ? thetaDash : theta;
    if (!t) return BRDFGPU_LM_ERROR;
    const long n = rows * nimg;
    std::vector<double> c((size_t)n), tt((size_t)n);
    for (long k = 0; k < n; ++k) {
        const long src = (k % rows) * nimg + (k / rows);  // row-major storage of element (k % rows, k / rows)
        c[(size_t)k] = phi[src];
        tt[(size_t)k] = t[src];
    }
    return brdfgpu_solve_equation_single(c.data(), model == 1 ? tt.data() : nullptr, model == 1 ? nullptr : tt.data(), I, n, model, p,
                                         info);
}

extern "C" int brdfgpu_solve_equation_batch(brdfgpu_ctx* ctx, long nfit, int nper, const double* phi,
                                            const double* thetaDash, const double* theta, const double* I, int model,
                                            double* p_out, double* info_out, int* ret_out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !phi || !I || !p_out) return BRDFGPU_LM_ERROR;
    const double* t = model == 1 ? thetaDash : theta;
    if (!t) return BRDFGPU_LM_ERROR;
    static const double p0[3] = {0.5, 1.0, 1.0}, lb[3] = {0, 0, 0}, ub[3] = {100, 100, 100};  // brdfdata.cpp:1085,1112-1113
    static const double opts[5] = {1E-03, 1E-15, 1E-15, 1E-20, 1E-06};                        // brdfdata.cpp:1116-1117
    brdfgpu_batch* b = nullptr;
    if (brdfgpu_batch_upload(ctx, nfit, nper, phi, t, I, model, &b) != 0) return BRDFGPU_LM_ERROR;
    // the reference's drivers promise the reference's results: levmar-exact wherever levmar's small-problem branch applies
    int rc = brdfgpu_batch_fit(ctx, b, p0, lb, ub, 100, opts, nper <= 128 ? BRDFGPU_JAC_FD_EXACT : BRDFGPU_JAC_FD);
    if (rc == 0) rc = brdfgpu_batch_results(ctx, b, p_out, info_out, ret_out);
    brdfgpu_batch_free(ctx, b);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// device-resident batched problem sets
// ------------------------------------------------------------------------------------------------
extern "C" int brdfgpu_batch_upload(brdfgpu_ctx* ctx, long nfit, int nper, const double* cosphi, const double* t,
                                    const double* x, int model, brdfgpu_batch** out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !out || (nfit > 0 && (!cosphi || !t))) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_batch* b = nullptr;
    if (batch_alloc(ctx, nfit, nper, model, &b) != 0) return BRDFGPU_LM_ERROR;
    const size_t nb = sizeof(double) * (size_t)nfit * nper;
    cudaError_t e = cudaSuccess;
    if (nb) {
        e = cudaMemcpyAsync(b->c, cosphi, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b->traw, t, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = x ? cudaMemcpyAsync(b->x, x, nb, cudaMemcpyHostToDevice, ctx->stream) : cudaMemsetAsync(b->x, 0, nb, ctx->stream);
    }
    if (e != cudaSuccess || batch_prepare(ctx, b) != 0 || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        if (e != cudaSuccess) set_error(ctx, std::string("batch_upload: ") + cudaGetErrorString(e));
        brdfgpu_batch_free(ctx, b);
        return BRDFGPU_LM_ERROR;
    }
    *out = b;
    return 0;
}

extern "C" int brdfgpu_batch_synth(brdfgpu_ctx* ctx, long nfit, int nper, unsigned long long seed, long first_fit,
                                   int model, brdfgpu_batch** out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !out) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_batch* b = nullptr;
    if (batch_alloc(ctx, nfit, nper, model, &b) != 0) return BRDFGPU_LM_ERROR;
    if (synth_batch(ctx, b, seed, first_fit) != 0 || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        brdfgpu_batch_free(ctx, b);
        return BRDFGPU_LM_ERROR;
    }
    *out = b;
    return 0;
}

extern "C" int brdfgpu_batch_fit(brdfgpu_ctx* ctx, brdfgpu_batch* b, const double* p0, const double* lb, const double* ub,
                                 int itmax, const double* opts, int jac_mode) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !b || !p0) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (jac_mode == BRDFGPU_JAC_FD_EXACT) return batch_fit_exact(ctx, b, p0, lb, ub, itmax, opts);
    return batch_fit(ctx, b, p0, lb, ub, itmax, opts, jac_mode);
}

extern "C" int brdfgpu_batch_results(brdfgpu_ctx* ctx, const brdfgpu_batch* b, double* p_out, double* info_out,
                                     int* ret_out) {
    ctx = ctx_or_default(ctx);
    if (!ctx || !b) return BRDFGPU_LM_ERROR;
    const size_t nf = (size_t)b->nfit;
    if (p_out && nf) BG_CUDA_OK(ctx, cudaMemcpyAsync(p_out, b->p, sizeof(double) * 3 * nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (info_out && nf)
        BG_CUDA_OK(ctx, cudaMemcpyAsync(info_out, b->info, sizeof(double) * 10 * nf, cudaMemcpyDeviceToHost, ctx->stream));
    if (ret_out && nf) BG_CUDA_OK(ctx, cudaMemcpyAsync(ret_out, b->ret, sizeof(int) * nf, cudaMemcpyDeviceToHost, ctx->stream));
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" long brdfgpu_batch_count(const brdfgpu_batch* b) { return b ? b->nfit : 0; }

extern "C" void brdfgpu_batch_free(brdfgpu_ctx* ctx, brdfgpu_batch* b) {
    if (!b) return;
    free_block(ctx, b->block, b->stream);
    delete b;
}

// ------------------------------------------------------------------------------------------------
// the LM control loop on caller-supplied reduced evaluators (host instantiation of lm_engine.cuh)
// ------------------------------------------------------------------------------------------------
namespace {
struct CallbackEval {
    static constexpr int kCostBatch = 1;
    static constexpr bool kLanePgWalk = false;
    brdfgpu_reduced_jac_t jac_cb;
    brdfgpu_reduced_cost_t cost_cb;
    void* user;
    int m;
    void jac(const double* p, double* JtJ, double* Jte) { jac_cb(p, m, JtJ, Jte, user); }
    double cost(const double* p, bool& bad) {
        double nonfinite = 0.0;
        const double e = cost_cb(p, m, &nonfinite, user);
        bad = nonfinite != 0.0;
        return e;
    }
};
// the same callbacks, but the projected-gradient walk hands over its candidates eight at a time,
// exactly as the persistent fit kernel receives them
struct CallbackEvalMany : CallbackEval {
    static constexpr int kCostBatch = 8;
    int max_batch = 0;
    double pts[kCostBatch * kMaxM], es[kCostBatch];
    bool bads[kCostBatch];
    double* batch_points() { return pts; }
    void cost_many(int cnt, const double* dscl, int mm) {
        if (cnt > max_batch) max_batch = cnt;
        for (int c = 0; c < cnt; ++c) {
            double q[kMaxM];
            for (int i = mm; i-- > 0;) q[i] = dscl ? pts[c * mm + i] * dscl[i] : pts[c * mm + i];
            es[c] = cost(q, bads[c]);
        }
    }
    double batch_cost(int c) const { return es[c]; }
    bool batch_bad(int c) const { return bads[c]; }
};
}  // namespace

static int g_reduced_batched = 0;
extern "C" int brdfgpu_lm_reduced_batching(int on) {
    const int prev = g_reduced_batched;
    g_reduced_batched = on ? 1 : 0;
    return prev;
}

extern "C" int brdfgpu_lm_bc_reduced(brdfgpu_reduced_jac_t jac_cb, brdfgpu_reduced_cost_t cost_cb, void* user, double* p,
                                     int m, long n, const double* lb, const double* ub, const double* dscl, int itmax,
                                     const double* opts, double* info, double* covar) {
    if (!jac_cb || !cost_cb || !p || m < 1 || m > kMaxM || n < m) return BRDFGPU_LM_ERROR;
    if (lb && ub)
        for (int i = 0; i < m; ++i)
            if (lb[i] > ub[i]) return BRDFGPU_LM_ERROR;
    if (dscl)
        for (int i = 0; i < m; ++i)
            if (dscl[i] <= 0.0) return BRDFGPU_LM_ERROR;
    double lbs[kMaxM], ubs[kMaxM], JtJ[kMaxM * kMaxM], fit_info[10];
    for (int i = 0; i < m; ++i) {
        if (lb) lbs[i] = dscl ? lb[i] / dscl[i] : lb[i];
        if (ub) ubs[i] = dscl ? ub[i] / dscl[i] : ub[i];
    }
    const Box box{lb, ub};
    box_project(p, box, m);
    int ret;
    if (g_reduced_batched) {
        CallbackEvalMany ev;
        ev.jac_cb = jac_cb; ev.cost_cb = cost_cb; ev.user = user; ev.m = m;
        ret = lm_bc_der<kMaxM>(ev, m, p, lb ? lbs : nullptr, ub ? ubs : nullptr, dscl, lm_options(opts, itmax), fit_info, JtJ);
        g_reduced_batched = ev.max_batch > 0 ? ev.max_batch : 1;
    } else {
        CallbackEval ev{jac_cb, cost_cb, user, m};
        ret = lm_bc_der<kMaxM>(ev, m, p, lb ? lbs : nullptr, ub ? ubs : nullptr, dscl, lm_options(opts, itmax), fit_info, JtJ);
    }
    if (info)
        for (int i = 0; i < 10; ++i) info[i] = fit_info[i];
    if (covar) {
        lm_covar<kMaxM>(JtJ, covar, fit_info[1], m, n);
        if (dscl)
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < m; ++j) covar[i * m + j] *= dscl[i] * dscl[j];
    }
    return ret;
}

extern "C" int brdfgpu_lm_unc_reduced(brdfgpu_reduced_jac_t jac_cb, brdfgpu_reduced_cost_t cost_cb, void* user, double* p,
                                      int m, long n, int itmax, const double* opts, double* info, double* covar) {
    if (!jac_cb || !cost_cb || !p || m < 1 || m > kMaxM || n < m) return BRDFGPU_LM_ERROR;
    double JtJ[kMaxM * kMaxM], fit_info[10];
    CallbackEval ev{jac_cb, cost_cb, user, m};
    const int ret = lm_der<kMaxM>(ev, m, p, lm_options(opts, itmax), fit_info, JtJ);
    if (info)
        for (int i = 0; i < 10; ++i) info[i] = fit_info[i];
    if (covar) lm_covar<kMaxM>(JtJ, covar, fit_info[1], m, n);
    return ret;
}

extern "C" int brdfgpu_Ax_eq_b_LU(const double* A, const double* B, double* x, int m) {
    if (!A || !B || !x || m < 1 || m > kMaxM) return 0;
    return solve_lu<kMaxM>(A, B, x, m);
}
