// global_fit.cu -- global mode: ONE BRDF fit over all resident samples.
//
// Replaces, for the two BRDF models, the passes levmar makes over its n-sized host arrays
// (reference: BRDFFunc brdfdata.cpp:969-989 called m+2 times per iteration, the forward/central
// difference Jacobian misc_core.c:137-211, J^T J / J^T e lmbc_core.c:573-632 and the residual
// norm misc_core.c:721-807) by two streaming kernels:
//
//   K2 k_normal_eq : fused residual + difference/analytic Jacobian + J^T J, J^T e, ||e||^2
//   K3 k_cost      : ||x - f(p)||^2 at a trial point
//
// Both read 24 B per sample (cosphi, log t, x as fp64 SoA, 16-byte vector loads), keep the sums in
// fp64 registers, reduce with warp shuffles, write one partial per CTA and let the last CTA to
// finish add the partials in a fixed order, so results are deterministic for a given grid.
// The levmar control loop (lm_engine.cuh) consumes only those 11 numbers.  Two drivers:
//
//   host       : one kernel per evaluation, the last CTA publishes the sums straight into mapped
//                pinned memory and the host spins on a ticket (no stream synchronise per pass);
//                with a communicator the sums are all-reduced across ranks first.
//   persistent : the whole fit in ONE cooperative kernel -- every thread runs the same control
//                code on the same grid-reduced sums, a grid barrier per evaluation.
#include <cooperative_groups.h>

#include <cstring>
#include <vector>

#include "brdf_model.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace brdfgpu {

struct SampleView {
    const double *c, *L, *x, *traw;
    long n;
};

static SampleView view_of(const brdfgpu_samples* s) { return SampleView{s->c, s->L, s->x, s->traw, s->n}; }

// ------------------------------------------------------------------------------------------------
// streaming bodies: grid-stride over sample PAIRS (16-byte loads), unrolled for loads in flight
// ------------------------------------------------------------------------------------------------
// Loads run one grid-stride step ahead of the arithmetic (register double buffer), so every warp
// always has 3 x 16 B in flight while it works through ~100 fp64 instructions of the current pair.
struct Pair {
    double2 c, l, x;
};
__device__ __forceinline__ Pair load_pair(const double2* __restrict__ c2, const double2* __restrict__ l2,
                                          const double2* __restrict__ x2, long i) {
    Pair p;
    p.c = __ldg(c2 + i);
    p.l = __ldg(l2 + i);
    p.x = __ldg(x2 + i);
    return p;
}

template <int JAC>
__device__ __forceinline__ void stream_jac(const SampleView& v, const PassParams& q, long tid, long nthreads,
                                           double* acc) {
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(v.c);
    const double2* __restrict__ l2 = reinterpret_cast<const double2*>(v.L);
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(v.x);
    const long npair = v.n >> 1;
    long i = tid;
    Pair cur;
    if (i < npair) cur = load_pair(c2, l2, x2, i);
    while (i < npair) {
        const long nxt = i + nthreads;
        Pair ahead = cur;
        if (nxt < npair) ahead = load_pair(c2, l2, x2, nxt);
        accumulate_jac<JAC>(q, cur.c.x, cur.l.x, cur.x.x, v.traw, 2 * i, acc);
        accumulate_jac<JAC>(q, cur.c.y, cur.l.y, cur.x.y, v.traw, 2 * i + 1, acc);
        cur = ahead;
        i = nxt;
    }
    if ((v.n & 1) && tid == 0) {
        const long j = v.n - 1;
        accumulate_jac<JAC>(q, v.c[j], v.L[j], v.x[j], v.traw, j, acc);
    }
}

__device__ __forceinline__ void stream_cost(const SampleView& v, const PassParams& q, long tid, long nthreads,
                                            double* acc2 /* [0] = sum e^2 */) {
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(v.c);
    const double2* __restrict__ l2 = reinterpret_cast<const double2*>(v.L);
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(v.x);
    const long npair = v.n >> 1;
    long i = tid;
    Pair cur;
    if (i < npair) cur = load_pair(c2, l2, x2, i);
    while (i < npair) {
        const long nxt = i + nthreads;
        Pair ahead = cur;
        if (nxt < npair) ahead = load_pair(c2, l2, x2, nxt);
        accumulate_cost(q, cur.c.x, cur.l.x, cur.x.x, v.traw, 2 * i, acc2);
        accumulate_cost(q, cur.c.y, cur.l.y, cur.x.y, v.traw, 2 * i + 1, acc2);
        cur = ahead;
        i = nxt;
    }
    if ((v.n & 1) && tid == 0) {
        const long j = v.n - 1;
        accumulate_cost(q, v.c[j], v.L[j], v.x[j], v.traw, j, acc2);
    }
}

// number of non-finite residuals: only run when ||e||^2 came out non-finite, to tell an overflow of
// the sum from invalid model values (lmbc_core.c:748, 915: VECNORM is non-finite iff some element is)
__device__ __forceinline__ void stream_count_bad(const SampleView& v, const PassParams& q, long tid, long nthreads,
                                                 double* cnt) {
    for (long i = tid; i < v.n; i += nthreads) {
        const double e = residual_of(q, v.c[i], v.L[i], v.x[i], v.traw, i);
        *cnt += lm_finite(e) ? 0.0 : 1.0;
    }
}

// ------------------------------------------------------------------------------------------------
// reductions: registers -> warp shuffles -> shared -> one partial per CTA -> fixed-order final sum
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_to(const double* acc, double* red /*[nwarps*NV] shared*/,
                                                double* out /*[NV] global*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double v = acc[k];
#pragma unroll
        for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp * NV + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += red[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// every warp takes quantities k = warp, warp + nwarps, ...; lanes stride over the CTA partials
template <int NV>
__device__ __forceinline__ void final_reduce(const double* partials, int nblocks, double* out /*[NV]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int k = warp; k < NV; k += nwarps) {
        double s = 0.0;
        for (int b = lane; b < nblocks; b += 32) s += __ldcg(partials + (long)b * NV + k);  // L2: other CTAs wrote it
#pragma unroll
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) out[k] = s;
    }
}

struct Publish {
    double* result;             // device, NV doubles
    double* h_result;           // mapped pinned (device alias) or nullptr
    unsigned long long* h_seq;  // mapped pinned ticket or nullptr
    unsigned long long seq;
};

template <int NV>
__device__ __forceinline__ void last_block_finish(double* partials, unsigned* ticket, double* red, const Publish& pub) {
    __shared__ bool is_last;
    __threadfence();  // partial of this CTA visible before the ticket
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);  // wraps to 0: reusable
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    final_reduce<NV>(partials, gridDim.x, red);
    __syncthreads();
    if (threadIdx.x < NV) {
        pub.result[threadIdx.x] = red[threadIdx.x];
        if (pub.h_result) pub.h_result[threadIdx.x] = red[threadIdx.x];
    }
    if (pub.h_seq) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *pub.h_seq = pub.seq;
    }
}

template <int JAC>
__global__ void __launch_bounds__(kPassThreads, 3) k_normal_eq(SampleView v, PassParams q, double* partials,
                                                             unsigned* ticket, Publish pub) {
    __shared__ double red[(kPassThreads / 32) * NACC];
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    stream_jac<JAC>(v, q, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x, acc);
    block_reduce_to<NACC>(acc, red, partials + (long)blockIdx.x * NACC);
    last_block_finish<NACC>(partials, ticket, red, pub);
}

template <bool COUNT_BAD>
__global__ void __launch_bounds__(kPassThreads, 4) k_cost(SampleView v, PassParams q, double* partials, unsigned* ticket,
                                                        Publish pub) {
    __shared__ double red[(kPassThreads / 32)];
    double acc[1] = {0.0};
    if (COUNT_BAD) stream_count_bad(v, q, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x, acc);
    else stream_cost(v, q, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x, acc);
    block_reduce_to<1>(acc, red, partials + (long)blockIdx.x);
    last_block_finish<1>(partials, ticket, red, pub);
}

// e_i = x_i - f(p)_i, the vector levmar keeps in `e` (lmbc_core.c:526)
__global__ void k_residuals(SampleView v, PassParams q, double* e) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < v.n; i += (long)gridDim.x * blockDim.x) {
        e[i] = residual_of(q, v.c[i], v.L[i], v.x[i], v.traw, i);
    }
}

__global__ void k_prepare(const double* __restrict__ traw, double* __restrict__ L, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        L[i] = log_or_flag(traw[i]);
}

// BRDFFunc itself (brdfdata.cpp:969-989), straight pow(): the spot-check entry point
__global__ void k_predict(const double* __restrict__ c, const double* __restrict__ t, int n, double kd, double ks,
                          double nn, int model, double* __restrict__ hx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (model == 1) hx[i] = kd * c[i] + ks * pow(t[i], nn);
    else hx[i] = kd * c[i] + ((nn + 2.0) / 2.0 * kPi) * ks * pow(t[i], nn);
}

__global__ void k_model_jac(const double* __restrict__ c, const double* __restrict__ t, int n, int m, double kd,
                            double ks, double nn, int model, double* __restrict__ jac) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* row = jac + (long)i * m;
    for (int j = 0; j < m; ++j) row[j] = 0.0;
    const double tn = pow(t[i], nn);
    const double coef = model_coef(model, nn);
    row[0] = c[i];
    row[1] = coef * tn;
    row[2] = ks * tn * ((model == 1 ? 0.0 : kPi / 2.0) + coef * log(t[i]));
}

// ------------------------------------------------------------------------------------------------
// host-side plumbing
// ------------------------------------------------------------------------------------------------
static int wait_ticket(brdfgpu_ctx* ctx, unsigned long long seq) {
    // the last CTA stores the sums and then the ticket into mapped pinned memory
    for (unsigned spin = 0;; ++spin) {
        if (*ctx->h_seq == seq) return 0;
        if ((spin & 0xfffff) == 0xfffff) {
            cudaError_t e = cudaStreamQuery(ctx->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                set_error(ctx, std::string("kernel failed: ") + cudaGetErrorString(e));
                return BRDFGPU_LM_ERROR;
            }
            if (e == cudaSuccess && *ctx->h_seq != seq) {
                set_error(ctx, "kernel finished without publishing its result");
                return BRDFGPU_LM_ERROR;
            }
        }
    }
}

static int fetch_result(brdfgpu_ctx* ctx, int count, bool published) {
    if (ctx->nranks > 1) {
        if (comm_allreduce_device(ctx, ctx->d_result, count) != 0) return BRDFGPU_LM_ERROR;
        BG_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, count * sizeof(double), cudaMemcpyDeviceToHost,
                                         ctx->stream));
        BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (published) return wait_ticket(ctx, ctx->seq);
    BG_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_result, ctx->d_result, count * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int to_jac_kind(int jac_mode, double delta_signed) {
    if (jac_mode == BRDFGPU_JAC_ANALYTIC) return kJacAnalytic;
    return delta_signed < 0.0 ? kJacCentral : kJacForward;
}

static int launch_normal_eq(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta, int jkind,
                            bool publish) {
    const PassParams q = make_pass_params(p, s->model, delta, jkind);
    const int blocks = pass_blocks(ctx, s->n, 3);
    Publish pub{ctx->d_result, nullptr, nullptr, 0};
    if (publish) pub = Publish{ctx->d_result, ctx->h_result_dev, ctx->h_seq_dev, ++ctx->seq};
    const SampleView v = view_of(s);
    switch (jkind) {
        case kJacForward:
            k_normal_eq<kJacForward><<<blocks, kPassThreads, 0, ctx->stream>>>(v, q, ctx->d_partials, ctx->d_sync, pub);
            break;
        case kJacCentral:
            k_normal_eq<kJacCentral><<<blocks, kPassThreads, 0, ctx->stream>>>(v, q, ctx->d_partials, ctx->d_sync, pub);
            break;
        default:
            k_normal_eq<kJacAnalytic><<<blocks, kPassThreads, 0, ctx->stream>>>(v, q, ctx->d_partials, ctx->d_sync, pub);
            break;
    }
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

static int launch_cost(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, bool publish, bool count_bad = false) {
    const PassParams q = make_pass_params(p, s->model, 1.0, kJacAnalytic);
    const int blocks = pass_blocks(ctx, s->n, 4);
    Publish pub{ctx->d_result, nullptr, nullptr, 0};
    if (publish) pub = Publish{ctx->d_result, ctx->h_result_dev, ctx->h_seq_dev, ++ctx->seq};
    if (count_bad) k_cost<true><<<blocks, kPassThreads, 0, ctx->stream>>>(view_of(s), q, ctx->d_partials, ctx->d_sync, pub);
    else k_cost<false><<<blocks, kPassThreads, 0, ctx->stream>>>(view_of(s), q, ctx->d_partials, ctx->d_sync, pub);
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// number of non-finite residuals at p over all ranks
static int count_bad(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double* count) {
    const bool pub = ctx->nranks == 1;
    if (launch_cost(ctx, s, p, pub, true) != 0 || fetch_result(ctx, 1, pub) != 0) return BRDFGPU_LM_ERROR;
    *count = ctx->h_result[0];
    return 0;
}

int global_normal_eq(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta, int jac_mode,
                     double* out11) {
    const bool pub = ctx->nranks == 1;
    if (launch_normal_eq(ctx, s, p, lm_abs(delta), to_jac_kind(jac_mode, delta), pub) != 0) return BRDFGPU_LM_ERROR;
    if (fetch_result(ctx, NACC, pub) != 0) return BRDFGPU_LM_ERROR;
    for (int k = 0; k < NACC; ++k) out11[k] = ctx->h_result[k];
    return count_bad(ctx, s, p, out11 + NACC);
}

int global_cost(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double* out2) {
    const bool pub = ctx->nranks == 1;
    if (launch_cost(ctx, s, p, pub) != 0) return BRDFGPU_LM_ERROR;
    if (fetch_result(ctx, 1, pub) != 0) return BRDFGPU_LM_ERROR;
    out2[0] = ctx->h_result[0];
    return count_bad(ctx, s, p, out2 + 1);
}

int global_repeat(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta, int kind, int reps) {
    for (int r = 0; r < reps; ++r) {
        const int rc = (kind == 0) ? launch_normal_eq(ctx, s, p, lm_abs(delta), to_jac_kind(BRDFGPU_JAC_FD, delta), false)
                                   : launch_cost(ctx, s, p, false);
        if (rc != 0) return rc;
    }
    return 0;
}

int global_residuals(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double* e_host) {
    double* d_e = nullptr;
    BG_CUDA_OK(ctx, cudaMalloc(&d_e, sizeof(double) * (size_t)s->n));
    const PassParams q = make_pass_params(p, s->model, 1.0, kJacAnalytic);
    k_residuals<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(view_of(s), q, d_e);
    ++ctx->launches;
    cudaError_t e = cudaMemcpyAsync(e_host, d_e, sizeof(double) * (size_t)s->n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_e);
    BG_CUDA_OK(ctx, e);
    return 0;
}

int samples_alloc(brdfgpu_ctx* ctx, long n, int model, brdfgpu_samples** out) {
    if (n < 0 || (model != 0 && model != 1)) {
        set_error(ctx, "samples: bad size or model");
        return BRDFGPU_LM_ERROR;
    }
    brdfgpu_samples* s = new brdfgpu_samples;
    s->n = n;
    s->model = model;
    const size_t bytes = sizeof(double) * (size_t)(n > 0 ? n + (n & 1) : 2);
    cudaError_t e = cudaMalloc(&s->c, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->L, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->x, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&s->traw, bytes);
    if (e != cudaSuccess) {
        cudaFree(s->c); cudaFree(s->L); cudaFree(s->x); cudaFree(s->traw);
        delete s;
        set_error(ctx, std::string("samples: cudaMalloc: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    *out = s;
    return 0;
}

int samples_prepare(brdfgpu_ctx* ctx, brdfgpu_samples* s) {
    if (s->n == 0) return 0;
    const int blocks = (int)((s->n + 255) / 256 < (long)ctx->sm_count * 16 ? (s->n + 255) / 256 : (long)ctx->sm_count * 16);
    k_prepare<<<blocks, 256, 0, ctx->stream>>>(s->traw, s->L, s->n);
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

static int model_io(brdfgpu_ctx* ctx, const double* p, const double* angles_host, int model, int n, int m,
                    double* out_host, bool jac) {
    if (n <= 0) return 0;
    if (model != 0 && model != 1) return 0;  // BRDFFunc leaves hx untouched for other ids (brdfdata.cpp:978,983)
    double *d_c = nullptr, *d_t = nullptr, *d_o = nullptr;
    const size_t nb = sizeof(double) * (size_t)n, ob = jac ? nb * m : nb;
    cudaError_t e = cudaMalloc(&d_c, nb);
    if (e == cudaSuccess) e = cudaMalloc(&d_t, nb);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, ob);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_c, angles_host, nb, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_t, angles_host + (model == 1 ? (size_t)n : 2 * (size_t)n), nb, cudaMemcpyHostToDevice,
                            ctx->stream);
    if (e == cudaSuccess) {
        if (jac) k_model_jac<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_c, d_t, n, m, p[0], p[1], p[2], model, d_o);
        else k_predict<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_c, d_t, n, p[0], p[1], p[2], model, d_o);
        ++ctx->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_o, ob, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_c); cudaFree(d_t); cudaFree(d_o);
    BG_CUDA_OK(ctx, e);
    return 0;
}

int model_predict(brdfgpu_ctx* ctx, const double* p, const double* angles_host, int model, int n, double* hx_host) {
    return model_io(ctx, p, angles_host, model, n, 3, hx_host, false);
}
int model_jacobian(brdfgpu_ctx* ctx, const double* p, const double* angles_host, int model, int n, int m,
                   double* jac_host) {
    return model_io(ctx, p, angles_host, model, n, m, jac_host, true);
}

// ------------------------------------------------------------------------------------------------
// host driver: lm_engine on sums produced kernel by kernel
// ------------------------------------------------------------------------------------------------
struct HostEval {
    brdfgpu_ctx* ctx;
    const brdfgpu_samples* s;
    double delta;
    int jkind;
    bool failed;

    void jac(const double* p, double* JtJ, double* Jte) {
        const bool pub = ctx->nranks == 1;
        if (failed || launch_normal_eq(ctx, s, p, delta, jkind, pub) != 0 || fetch_result(ctx, NACC, pub) != 0) {
            failed = true;
            for (int i = 0; i < 9; ++i) JtJ[i] = NAN;
            for (int i = 0; i < 3; ++i) Jte[i] = NAN;
            return;
        }
        const double* r = ctx->h_result;
        JtJ[0] = r[A00]; JtJ[1] = r[A01]; JtJ[2] = r[A02];
        JtJ[3] = r[A01]; JtJ[4] = r[A11]; JtJ[5] = r[A12];
        JtJ[6] = r[A02]; JtJ[7] = r[A12]; JtJ[8] = r[A22];
        Jte[0] = r[G0]; Jte[1] = r[G1]; Jte[2] = r[G2];
    }
    double cost(const double* p, bool& bad) {
        const bool pub = ctx->nranks == 1;
        if (failed || launch_cost(ctx, s, p, pub) != 0 || fetch_result(ctx, 1, pub) != 0) {
            failed = true;
            bad = true;
            return NAN;
        }
        const double esq = ctx->h_result[0];
        bad = false;
        if (!lm_finite(esq)) {  // rare: tell an overflowed sum from invalid residuals
            double cnt = 0.0;
            if (count_bad(ctx, s, p, &cnt) != 0) failed = true;
            bad = failed || cnt != 0.0;
        }
        return esq;
    }
};

// ------------------------------------------------------------------------------------------------
// persistent driver: the whole fit inside one cooperative kernel
// ------------------------------------------------------------------------------------------------
constexpr int kPersistThreads = 512;  // one CTA per SM, 16 warps, <= 128 registers per thread

// Grid-wide barrier for the cooperative kernel: one atomic per CTA on a monotonically increasing
// counter, everybody spins on an acquire load.  (Cheaper than cooperative_groups' grid.sync(): the
// partial sums ride on the same release/acquire, so one barrier per evaluation is all there is.)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target, const int* abort_flag) {
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned seen, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            // a CTA that gave up on a peer rank (peer_exchange) leaves the control loop early: do not
            // wait for it forever
            if (abort_flag && (++spins & 0x3ffu) == 0 && *(volatile const int*)abort_flag) break;
        } while (seen < target);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// fused cross-GPU all-reduce (global mode on several GPUs): the exchange step of an evaluation
// happens INSIDE the fit kernel over NVLink peer memory, so a multi-GPU fit is still one launch per
// rank.  After the local grid reduction CTA 0 of every rank stores its rank's sums as flagged
// cells into slot [parity][rank] of EVERY rank's exchange buffer (peer stores through NVSwitch);
// every CTA then polls its own rank's buffer until all slots carry the current tag and adds them in
// rank order -- identical bits on all ranks, so all ranks take identical LM decisions with no
// broadcast.  Two parities: a rank can run at most one exchange ahead of the slowest CTA of any peer.
// ------------------------------------------------------------------------------------------------
constexpr long long kPeerSpinCycles = 6000000000LL;  // ~3 s at 2 GHz, then the fit is abandoned

__device__ __forceinline__ void peer_store_cell(uint4* dst, double v, unsigned tag) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(lo), "r"(tag), "r"(hi), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 peer_load_cell(const uint4* src) {
    uint4 c;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(src)
                 : "memory");
    return c;
}

// in: res[0..NV) = this rank's sums (shared, complete); out: res[0..NV) = sums over all ranks
template <int NV>
__device__ __forceinline__ void peer_exchange(PeerView& pv, double* res, double* stage /*shared [kMaxRanks*NV]*/,
                                              int* timeout_flag) {
    if (pv.nranks <= 1) return;
    pv.epoch = pv.epoch + 1u ? pv.epoch + 1u : 1u;  // never 0: the buffers start zeroed
    const unsigned tag = pv.epoch;
    const int par = (int)(tag & 1u);
    const int r = threadIdx.x / NV, k = threadIdx.x % NV;
    if (threadIdx.x < NV * pv.nranks) {
        if (blockIdx.x == 0)  // (also after a timeout: the peers must not wait for this rank in turn)
            peer_store_cell(pv.remote[r] + ((long)(par * kMaxRanks + pv.rank) * kPeerCellsPerRank + k), res[k], tag);
        const uint4* src = pv.local + ((long)(par * kMaxRanks + r) * kPeerCellsPerRank + k);
        const long long t0 = clock64();
        unsigned spins = 0;
        uint4 c = peer_load_cell(src);
        while (c.y != tag || c.w != tag) {
            if ((++spins & 0xffu) == 0 && (clock64() - t0 > kPeerSpinCycles || *(volatile int*)timeout_flag)) {
                *timeout_flag = 1;
                c.x = 0u; c.z = 0x7ff80000u;  // NaN: the control loop stops with reason 7
                break;
            }
            c = peer_load_cell(src);
        }
        stage[r * NV + k] = __hiloint2double((int)c.z, (int)c.x);
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double sum = 0.0;
        for (int q = 0; q < pv.nranks; ++q) sum += stage[q * NV + threadIdx.x];
        res[threadIdx.x] = sum;
    }
    __syncthreads();
}

struct GridEval {
    SampleView v;
    int model, jkind;
    double delta;
    double* partials;  // 2 x kMaxPassBlocks x NACC, double-buffered across evaluations
    double* red;       // shared [(kPersistThreads/32) * NACC]
    double* res;       // shared [NACC]
    int parity;
    unsigned* bar;     // grid barrier counter (zeroed before the launch)
    unsigned bar_target;
    PeerView peer;     // nranks == 1: no exchange
    double* stage;     // shared [kMaxRanks * NACC]
    int* timeout_flag; // global

    // one noinline instance per Jacobian kind: each gets its own register allocation
    template <int JAC>
    __device__ __noinline__ void jac_pass(const PassParams& q) {
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        stream_jac<JAC>(v, q, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x, acc);
        double* buf = partials + (long)parity * kMaxPassBlocks * NACC;
        block_reduce_to<NACC>(acc, red, buf + (long)blockIdx.x * NACC);
    }

    __device__ __forceinline__ void jac(const double* p, double* JtJ, double* Jte) {
        const PassParams q = make_pass_params(p, model, delta, jkind);
        if (jkind == kJacForward) jac_pass<kJacForward>(q);
        else if (jkind == kJacCentral) jac_pass<kJacCentral>(q);
        else jac_pass<kJacAnalytic>(q);
        double* buf = partials + (long)parity * kMaxPassBlocks * NACC;
        parity ^= 1;
        grid_barrier(bar, bar_target, peer.nranks > 1 ? timeout_flag : nullptr);
        final_reduce<NACC>(buf, gridDim.x, res);
        __syncthreads();
        peer_exchange<NACC>(peer, res, stage, timeout_flag);
        JtJ[0] = res[A00]; JtJ[1] = res[A01]; JtJ[2] = res[A02];
        JtJ[3] = res[A01]; JtJ[4] = res[A11]; JtJ[5] = res[A12];
        JtJ[6] = res[A02]; JtJ[7] = res[A12]; JtJ[8] = res[A22];
        Jte[0] = res[G0]; Jte[1] = res[G1]; Jte[2] = res[G2];
        __syncthreads();  // res is rewritten by the next pass's exchange before its barrier
    }

    __device__ __forceinline__ double scalar_pass(const PassParams& q, bool count_bad) {
        double acc[1] = {0.0};
        const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long)gridDim.x * blockDim.x;
        if (count_bad) stream_count_bad(v, q, tid, nth, acc);
        else stream_cost(v, q, tid, nth, acc);
        double* buf = partials + (long)parity * kMaxPassBlocks * NACC;
        parity ^= 1;
        block_reduce_to<1>(acc, red, buf + (long)blockIdx.x);
        grid_barrier(bar, bar_target, peer.nranks > 1 ? timeout_flag : nullptr);
        final_reduce<1>(buf, gridDim.x, res);
        __syncthreads();
        peer_exchange<1>(peer, res, stage, timeout_flag);
        const double out = res[0];
        __syncthreads();  // res is rewritten by the next pass's exchange before its barrier
        return out;
    }

    __device__ __noinline__ double cost(const double* p, bool& bad) {
        const PassParams q = make_pass_params(p, model, 1.0, kJacAnalytic);
        const double esq = scalar_pass(q, false);
        bad = false;
        if (!lm_finite(esq)) bad = scalar_pass(q, true) != 0.0;  // uniform across the grid: same sums everywhere
        return esq;
    }
};

__global__ void __launch_bounds__(kPersistThreads, 1) k_persistent_fit(SampleView v, int model, GlobalFitSpec spec,
                                                                        double* partials, unsigned* barrier,
                                                                        PeerView peer, GlobalFitOut* out) {
    __shared__ double red[(kPersistThreads / 32) * NACC];
    __shared__ double res[NACC];
    __shared__ double stage[kMaxRanks * NACC];
    if (blockIdx.x == 0 && threadIdx.x == 0) out->peer_timeout = 0;
    GridEval ev{v, model, spec.jac_mode, spec.delta, partials, red, res, 0, barrier, 0u, peer, stage, &out->peer_timeout};
    double p[3], info[10], JtJ[9];
    for (int i = 0; i < 3; ++i) p[i] = spec.p[i];
    int ret;
    if (spec.unconstrained)
        ret = lm_der<3>(ev, 3, p, spec.opt, info, JtJ);
    else
        ret = lm_bc_der<3>(ev, 3, p, spec.has_lb ? spec.lb : nullptr, spec.has_ub ? spec.ub : nullptr,
                           spec.has_dscl ? spec.dscl : nullptr, spec.opt, info, JtJ);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out->ret = ret;
        out->peer_epoch = ev.peer.epoch;
        for (int i = 0; i < 3; ++i) out->p[i] = p[i];
        for (int i = 0; i < 10; ++i) out->info[i] = info[i];
        for (int i = 0; i < 9; ++i) out->JtJ[i] = JtJ[i];
    }
}

static int persistent_grid(brdfgpu_ctx* ctx, long n) {
    if (ctx->persistent_blocks_per_sm == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent_fit, kPersistThreads, 0) != cudaSuccess ||
            per_sm < 1)
            per_sm = -1;
        ctx->persistent_blocks_per_sm = per_sm;
    }
    if (ctx->persistent_blocks_per_sm < 1) return 0;
    long want = (n / 2 + kPersistThreads - 1) / kPersistThreads;
    const long cap = (long)ctx->sm_count * ctx->persistent_blocks_per_sm;
    if (want < 1) want = 1;
    if (want > cap) want = cap;
    if (want > kMaxPassBlocks) want = kMaxPassBlocks;
    return (int)want;
}

static void finish_info(double* info, const double* fit_info, int m, int jkind, bool dif_accounting) {
    if (!info) return;
    for (int i = 0; i < 10; ++i) info[i] = fit_info[i];
    // dlevmar_bc_dif charges every Jacobian m+1 (forward) or 2m (central) function calls,
    // lmbc_core.c:1119-1124
    if (dif_accounting) info[7] += info[8] * (jkind == kJacCentral ? 2 * m : m + 1);
}

int global_fit(brdfgpu_ctx* ctx, const brdfgpu_samples* s, double* p, int m, const double* lb, const double* ub,
               const double* dscl, int itmax, const double* opts, double* info, double* covar, int drive, int jac_mode,
               int unconstrained) {
    if (m != 3) {
        set_error(ctx, "the BRDF models have exactly 3 parameters (kd, ks, n)");
        return BRDFGPU_LM_ERROR;
    }
    if (s->n < m) {  // lmbc_core.c:440-443
        fprintf(stderr, "brdfgpu fit: cannot solve a problem with fewer measurements [%ld] than unknowns [%d]\n", s->n, m);
        return BRDFGPU_LM_ERROR;
    }
    if (unconstrained) lb = ub = dscl = nullptr;
    if (lb && ub) {  // dlevmar_box_check, misc_core.c:661-671
        for (int i = 0; i < m; ++i)
            if (lb[i] > ub[i]) {
                fprintf(stderr, "brdfgpu fit: at least one lower bound exceeds the upper one\n");
                return BRDFGPU_LM_ERROR;
            }
    }
    if (dscl) {  // lmbc_core.c:456-463
        for (int i = 0; i < m; ++i)
            if (dscl[i] <= 0.0) {
                fprintf(stderr, "brdfgpu fit: at least one non-positive scaling constant\n");
                return BRDFGPU_LM_ERROR;
            }
    }
    const double delta_signed = opts ? opts[4] : kDiffDelta;  // lmbc_core.c:1105,1115
    const int jkind = to_jac_kind(jac_mode, delta_signed);
    const double delta = lm_abs(delta_signed);
    const LmOptions o = lm_options(opts, itmax);

    // scaled bounds (lmbc_core.c:536-540) live in local copies: the caller's arrays stay untouched
    double lbs[3], ubs[3], p_in[3];
    for (int i = 0; i < 3; ++i) {
        if (lb) lbs[i] = dscl ? lb[i] / dscl[i] : lb[i];
        if (ub) ubs[i] = dscl ? ub[i] / dscl[i] : ub[i];
    }
    // the start is projected onto the box first (lmbc_core.c:514-520), in caller coordinates
    {
        const Box box{lb, ub};
        for (int i = 0; i < 3; ++i) p_in[i] = p[i];
        box_project(p, box, m);
        for (int i = 0; i < 3; ++i)
            if (p_in[i] != p[i])
                fprintf(stderr, "Warning: component %d of starting point not feasible in brdfgpu fit! [%g projected to %g]\n",
                        i, p_in[i], p[i]);
    }

    double fit_info[10], JtJ[9];
    int ret;
    const bool can_persist = ctx->coop && (ctx->nranks == 1 || ctx->peer_attached);
    const int grid = (drive == BRDFGPU_DRIVE_PERSISTENT && can_persist) ? persistent_grid(ctx, s->n) : 0;
    if (grid > 0) {
        GlobalFitSpec spec;
        memset(&spec, 0, sizeof(spec));
        spec.m = m; spec.itmax = itmax; spec.jac_mode = jkind; spec.delta = delta; spec.opt = o;
        spec.has_lb = lb != nullptr; spec.has_ub = ub != nullptr; spec.has_dscl = dscl != nullptr;
        spec.unconstrained = unconstrained;
        for (int i = 0; i < 3; ++i) {
            spec.p[i] = p[i];
            if (lb) spec.lb[i] = lbs[i];
            if (ub) spec.ub[i] = ubs[i];
            if (dscl) spec.dscl[i] = dscl[i];
        }
        SampleView v = view_of(s);
        int model = s->model;
        double* partials = ctx->d_partials;
        GlobalFitOut* d_out = static_cast<GlobalFitOut*>(ctx->d_fitio);
        unsigned* barrier = ctx->d_sync + 4;
        BG_CUDA_OK(ctx, cudaMemsetAsync(barrier, 0, sizeof(unsigned), ctx->stream));
        PeerView peer;
        memset(&peer, 0, sizeof(peer));
        peer.nranks = 1;
        if (ctx->nranks > 1) {
            peer.local = ctx->peer_local;
            for (int r = 0; r < ctx->nranks; ++r) peer.remote[r] = ctx->peer_remote[r];
            peer.rank = ctx->rank;
            peer.nranks = ctx->nranks;
            peer.epoch = ctx->peer_epoch;
        }
        void* args[] = {&v, &model, &spec, &partials, &barrier, &peer, &d_out};
        BG_CUDA_OK(ctx, cudaLaunchCooperativeKernel((const void*)k_persistent_fit, dim3(grid), dim3(kPersistThreads), args,
                                                     0, ctx->stream));
        ++ctx->launches;
        BG_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_fitio, d_out, sizeof(GlobalFitOut), cudaMemcpyDeviceToHost, ctx->stream));
        BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        const GlobalFitOut* h = static_cast<const GlobalFitOut*>(ctx->h_fitio);
        if (ctx->nranks > 1) {
            ctx->peer_epoch = h->peer_epoch;
            if (h->peer_timeout) {
                set_error(ctx, "multi-GPU fit abandoned: a peer rank never delivered its sums (peer exchange timeout)");
                return BRDFGPU_LM_ERROR;
            }
        }
        ret = h->ret;
        for (int i = 0; i < 3; ++i) p[i] = h->p[i];
        for (int i = 0; i < 10; ++i) fit_info[i] = h->info[i];
        for (int i = 0; i < 9; ++i) JtJ[i] = h->JtJ[i];
    } else {
        HostEval ev{ctx, s, delta, jkind, false};
        if (unconstrained) ret = lm_der<3>(ev, 3, p, o, fit_info, JtJ);
        else ret = lm_bc_der<3>(ev, 3, p, lb ? lbs : nullptr, ub ? ubs : nullptr, dscl, o, fit_info, JtJ);
        if (ev.failed) return BRDFGPU_LM_ERROR;
    }
    finish_info(info, fit_info, m, jkind, jac_mode == BRDFGPU_JAC_FD);
    if (covar) {  // lmbc_core.c:994-1002
        // ||e||^2 over ALL samples of all ranks went into fit_info[1]; n is the global count
        long n_all = s->n;
        if (ctx->nranks > 1) {
            double cnt = (double)s->n;
            // sample counts differ per rank by at most one: sum them through the same exchange
            double* tmp = ctx->d_result;
            cudaMemcpyAsync(tmp, &cnt, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
            if (comm_allreduce_device(ctx, tmp, 1) != 0) return BRDFGPU_LM_ERROR;
            cudaMemcpyAsync(&cnt, tmp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            n_all = (long)cnt;
        }
        lm_covar<3>(JtJ, covar, fit_info[1], m, n_all);
        if (dscl)
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < m; ++j) covar[i * m + j] *= dscl[i] * dscl[j];
    }
    return ret;
}

}  // namespace brdfgpu
