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
//   persistent : the whole fit in ONE cooperative kernel -- the control warp of every CTA runs the same
//                control code on the same grid-reduced sums; one flagged-cell exchange per sweep (the data
//                is the barrier), samples resident in shared memory, speculative Jacobians and fused
//                sweeps (GridEval) so that an iteration needs one or two sweeps.
#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef BG_EXP_POLY  // (make EXTRA=-DBG_EXP_POLY: the polynomial exponential, for comparisons)
#define BG_EXP_TABLE 1  // every kernel below that evaluates the model calls exp_table_load() first
#endif
#include "brdf_model.cuh"
#include "common.cuh"
#include "reduce.cuh"

namespace cg = cooperative_groups;

namespace brdfgpu {

template <int JAC>
__device__ __forceinline__ void stream_jac(const SampleView& v, const PassParams& q, const PassParams& cold,
                                           long first_pair, long tid, long nthreads, double* acc) {
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(v.c);
    const double2* __restrict__ l2 = reinterpret_cast<const double2*>(v.L);
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(v.x);
    const long npair = v.n >> 1;
    long i = first_pair + tid;
    Pair cur;
    if (i < npair) cur = load_pair(c2, l2, x2, i);
    while (i < npair) {
        const long nxt = i + nthreads;
        Pair ahead = cur;
        if (nxt < npair) ahead = load_pair(c2, l2, x2, nxt);
        accumulate_jac_pair<JAC>(q, cold, cur.c, cur.l, cur.x, v.traw, 2 * i, acc);
        cur = ahead;
        i = nxt;
    }
    if ((v.n & 1) && tid == 0) {
        const long j = v.n - 1;
        accumulate_jac<JAC>(cold, v.c[j], v.L[j], v.x[j], v.traw, j, acc);
    }
}

template <class Q>
__device__ __forceinline__ void stream_cost(const SampleView& v, const Q& q, long first_pair, long tid, long nthreads,
                                            double* acc2 /* [0] = sum e^2 */) {
    const double2* __restrict__ c2 = reinterpret_cast<const double2*>(v.c);
    const double2* __restrict__ l2 = reinterpret_cast<const double2*>(v.L);
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(v.x);
    const long npair = v.n >> 1;
    long i = first_pair + tid;
    Pair cur;
    if (i < npair) cur = load_pair(c2, l2, x2, i);
    while (i < npair) {
        const long nxt = i + nthreads;
        Pair ahead = cur;
        if (nxt < npair) ahead = load_pair(c2, l2, x2, nxt);
        accumulate_cost_pair(q, cur.c, cur.l, cur.x, v.traw, 2 * i, acc2);
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

__device__ __forceinline__ double2 lds_pair(unsigned base, int i) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + 16u * (unsigned)i) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------------
// TMA-staged streaming (large sample sets, HBM regime).  The loads of a register-prefetch loop are
// limited by registers: ~48 B per thread in flight, 24-36 KB per SM, while HBM3e needs ~40 KB per SM
// just to cover its latency.  Here one elected thread per CTA moves whole tiles (512 sample pairs of
// each of the three arrays = 24 KB) from global to shared memory with 1-D bulk async copies
// (cp.async.bulk, the TMA engine) into a 4-deep ring guarded by mbarriers: ~72 KB per CTA in flight,
// no registers spent on it, and the compute warps read their pairs back with LDS.128.
// ------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;
constexpr int kTilePairs = 2 * kTileThreads;  // two pairs (four samples) per thread and tile
constexpr int kTileStages = 4;
constexpr size_t kTileRingBytes = (size_t)kTileStages * 3 * kTilePairs * sizeof(double2);  // 96 KB

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// The ring: STAGES x {c, L, x} x (2 * THREADS) pairs in shared memory + full/empty mbarriers.
// `seq` counts the tiles that went through the ring since init (stage = seq % STAGES, mbarrier phase
// parity = (seq / STAGES) & 1), so a persistent kernel can keep using one ring sweep after sweep.
template <int THREADS, int STAGES>
struct TileRing {
    static constexpr int kPairs = 2 * THREADS;
    static constexpr unsigned kArrayBytes = kPairs * sizeof(double2), kStageBytes = 3 * kArrayBytes;
    static constexpr size_t kBytes = (size_t)STAGES * kStageBytes;
    unsigned ring, bars;  // shared-window addresses

    __device__ __forceinline__ void init(void* ring_smem, unsigned long long* bar_smem /*[2 * STAGES]*/) {
        ring = smem_addr(ring_smem);
        bars = smem_addr(bar_smem);
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(bars + 8 * s, 1);                      // the producer's expect_tx arrive
                mbar_init(bars + 8 * (STAGES + s), THREADS / 32);  // one arrive per consumer warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    // Calls body(c0, l0, x0, i0, c1, l1, x1, i1, n_valid) for every group of two sample pairs of the
    // tiles blockIdx.x, blockIdx.x + gridDim.x, ... of the pair range [first_pair, v.n / 2);
    // n_valid in {0, 1, 2} pairs, i* = index of the pair's first sample.  Returns the new seq.
    template <class Body>
    __device__ __forceinline__ long stream(const SampleView& v, long first_pair, long seq, Body&& body) const {
        const long npair = (v.n >> 1) - first_pair;
        const long ntiles = (npair + kPairs - 1) / kPairs;
        const long mine = (long)blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        auto issue = [&](long j) {  // tile number j of this CTA
            const int s = (int)((seq + j) % STAGES);
            const long first = ((long)blockIdx.x + j * gridDim.x) * kPairs;
            const long left = npair - first;
            const unsigned bytes = (unsigned)((left < kPairs ? left : kPairs) * sizeof(double2));
            const unsigned full = bars + 8 * s, dst = ring + s * kStageBytes;
            mbar_expect_tx(full, 3 * bytes);
            bulk_g2s(dst, v.c + 2 * (first_pair + first), bytes, full);
            bulk_g2s(dst + kArrayBytes, v.L + 2 * (first_pair + first), bytes, full);
            bulk_g2s(dst + 2 * kArrayBytes, v.x + 2 * (first_pair + first), bytes, full);
        };
        if (threadIdx.x == 0)
            for (long j = 0; j < mine && j < STAGES; ++j) issue(j);
        for (long j = 0; j < mine; ++j) {
            const int s = (int)((seq + j) % STAGES);
            if (threadIdx.x == 0 && j >= 1 && j - 1 + STAGES < mine) {  // refill the stage drained one tile ago
                mbar_wait(bars + 8 * (STAGES + (int)((seq + j - 1) % STAGES)), (unsigned)(((seq + j - 1) / STAGES) & 1));
                issue(j - 1 + STAGES);
            }
            mbar_wait(bars + 8 * s, (unsigned)(((seq + j) / STAGES) & 1));
            const long first = ((long)blockIdx.x + j * gridDim.x) * kPairs;
            const long left = npair - first;
            const int here = (int)(left < kPairs ? left : kPairs);
            const unsigned base = ring + s * kStageBytes;
            const int a = threadIdx.x, b = threadIdx.x + THREADS;
            const int valid = (a < here) + (b < here);
            double2 c0 = make_double2(0, 0), l0 = c0, x0 = c0, c1 = c0, l1 = c0, x1 = c0;
            if (valid >= 1) { c0 = lds_pair(base, a); l0 = lds_pair(base + kArrayBytes, a); x0 = lds_pair(base + 2 * kArrayBytes, a); }
            if (valid >= 2) { c1 = lds_pair(base, b); l1 = lds_pair(base + kArrayBytes, b); x1 = lds_pair(base + 2 * kArrayBytes, b); }
            body(c0, l0, x0, 2 * (first_pair + first + a), c1, l1, x1, 2 * (first_pair + first + b), valid);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8 * (STAGES + s));
        }
        return seq + mine;
    }
};

// one-shot form for the streaming kernels: ring in dynamic shared memory
template <class Body>
__device__ __forceinline__ void stream_tiles(const SampleView& v, Body&& body) {
    extern __shared__ __align__(128) unsigned char tile_ring[];
    __shared__ __align__(8) unsigned long long bars[2 * kTileStages];
    TileRing<kTileThreads, kTileStages> ring;
    ring.init(tile_ring, bars);
    ring.stream(v, 0, 0, body);
}

template <int JAC>
__global__ void __launch_bounds__(kTileThreads, 2) k_normal_eq_tma(SampleView v, PassParams q, double* partials,
                                                                   unsigned* ticket, Publish pub) {
    __shared__ double red[(kTileThreads / 32) * NACC];
    BG_EXP_TABLE_LOAD();
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    stream_tiles(v, [&](double2 c0, double2 l0, double2 x0, long i0, double2 c1, double2 l1, double2 x1, long i1, int valid) {
        if (valid == 2) {
            const double cc[4] = {c0.x, c0.y, c1.x, c1.y}, ll[4] = {l0.x, l0.y, l1.x, l1.y}, xx[4] = {x0.x, x0.y, x1.x, x1.y};
            const long idx[4] = {i0, i0 + 1, i1, i1 + 1};
            accumulate_jac_n<JAC, 4>(q, q, cc, ll, xx, v.traw, idx, acc);
        } else if (valid == 1) {
            accumulate_jac_pair<JAC>(q, q, c0, l0, x0, v.traw, i0, acc);
        }
    });
    if ((v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long j = v.n - 1;
        accumulate_jac<JAC>(q, v.c[j], v.L[j], v.x[j], v.traw, j, acc);
    }
    block_reduce_to<NACC>(acc, red, partials + (long)blockIdx.x * NACC);
    last_block_finish<NACC>(partials, ticket, red, pub);
}

__global__ void __launch_bounds__(kTileThreads, 2) k_cost_tma(SampleView v, PassParams q, double* partials, unsigned* ticket,
                                                              Publish pub) {
    __shared__ double red[(kTileThreads / 32)];
    BG_EXP_TABLE_LOAD();
    double a0 = 0.0, a1 = 0.0;
    stream_tiles(v, [&](double2 c0, double2 l0, double2 x0, long i0, double2 c1, double2 l1, double2 x1, long i1, int valid) {
        if (valid == 2) accumulate_cost_2pairs(q, c0, l0, x0, i0, c1, l1, x1, i1, v.traw, &a0, &a1);
        else if (valid == 1) accumulate_cost_pair(q, c0, l0, x0, v.traw, i0, &a0);
    });
    if ((v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long j = v.n - 1;
        accumulate_cost(q, v.c[j], v.L[j], v.x[j], v.traw, j, &a0);
    }
    double acc[1] = {a0 + a1};
    block_reduce_to<1>(acc, red, partials + (long)blockIdx.x);
    last_block_finish<1>(partials, ticket, red, pub);
}

template <int JAC>
__global__ void __launch_bounds__(kPassThreads, 2) k_normal_eq(SampleView v, PassParams q, double* partials,
                                                             unsigned* ticket, Publish pub) {
    __shared__ double red[(kPassThreads / 32) * NACC];
    BG_EXP_TABLE_LOAD();
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    stream_jac<JAC>(v, q, q, 0, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x, acc);
    block_reduce_to<NACC>(acc, red, partials + (long)blockIdx.x * NACC);
    last_block_finish<NACC>(partials, ticket, red, pub);
}

template <bool COUNT_BAD>
__global__ void __launch_bounds__(kPassThreads, 4) k_cost(SampleView v, PassParams q, double* partials, unsigned* ticket,
                                                        Publish pub) {
    __shared__ double red[(kPassThreads / 32)];
    BG_EXP_TABLE_LOAD();
    double acc[1] = {0.0};
    if (COUNT_BAD) stream_count_bad(v, q, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x, acc);
    else stream_cost(v, q, 0, (long)blockIdx.x * blockDim.x + threadIdx.x, (long)gridDim.x * blockDim.x, acc);
    block_reduce_to<1>(acc, red, partials + (long)blockIdx.x);
    last_block_finish<1>(partials, ticket, red, pub);
}

// e_i = x_i - f(p)_i, the vector levmar keeps in `e` (lmbc_core.c:526)
__global__ void k_residuals(SampleView v, PassParams q, double* e) {
    BG_EXP_TABLE_LOAD();
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
int wait_ticket(brdfgpu_ctx* ctx, unsigned long long seq) {
    // the last CTA stores the sums and then the ticket into mapped pinned memory
    for (unsigned spin = 0;; ++spin) {
        if (*ctx->h_seq == seq) {
            // the sums were stored before the ticket (GPU side: __threadfence_system); keep the host's loads of
            // h_result behind the load of the ticket on weakly ordered hosts too (Grace / aarch64)
            std::atomic_thread_fence(std::memory_order_acquire);
            return 0;
        }
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

int fetch_result(brdfgpu_ctx* ctx, int count, bool published) {
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

// The TMA-staged kernel pays off for the Jacobian pass (FP64-heavy, register-hungry) once the sample
// set outgrows on-chip residency.  BRDFGPU_TMA=0/1 in the environment forces never/always (tests).
static bool use_tma(brdfgpu_ctx* ctx, long n, bool cost_pass = false) {
    if (ctx->tma_mode == 0) {
        const char* e = getenv("BRDFGPU_TMA");
        int mode = 1;  // 1 = automatic
        if (e && *e) mode = (*e == '0') ? 2 : 3;  // 2 = never, 3 = always
        bool ok = cudaFuncSetAttribute(k_cost_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileRingBytes) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_normal_eq_tma<kJacForward>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileRingBytes) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_normal_eq_tma<kJacCentral>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileRingBytes) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(k_normal_eq_tma<kJacAnalytic>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileRingBytes) == cudaSuccess;
        ctx->tma_mode = ok ? mode : 2;
    }
    if (ctx->tma_mode == 2) return false;
    if (ctx->tma_mode == 3) return n >= 2 * kTilePairs;
    // measured on B200 (profiles/r01_summary.md): K2 5.2 -> 6.0 TB/s at 10^8 samples, break-even near
    // 10^6; the cost pass is purely HBM-bound and already at 6.9 TB/s with register prefetch
    return !cost_pass && n >= 2000000;
}
static int tma_blocks(const brdfgpu_ctx* ctx, long n) {
    const long ntiles = ((n >> 1) + kTilePairs - 1) / kTilePairs;
    const long cap = (long)ctx->sm_count * 2;
    return (int)(ntiles < cap ? (ntiles < 1 ? 1 : ntiles) : cap);
}

static int launch_normal_eq(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta, int jkind,
                            bool publish) {
    const PassParams q = make_pass_params(p, s->model, delta, jkind);
    Publish pub{ctx->d_result, nullptr, nullptr, 0};
    if (publish) pub = Publish{ctx->d_result, ctx->h_result_dev, ctx->h_seq_dev, ++ctx->seq};
    const SampleView v = view_of(s);
    if (use_tma(ctx, s->n)) {
        const int blocks = tma_blocks(ctx, s->n);
        switch (jkind) {
            case kJacForward:
                k_normal_eq_tma<kJacForward><<<blocks, kTileThreads, kTileRingBytes, ctx->stream>>>(v, q, ctx->d_partials, ctx->d_sync, pub);
                break;
            case kJacCentral:
                k_normal_eq_tma<kJacCentral><<<blocks, kTileThreads, kTileRingBytes, ctx->stream>>>(v, q, ctx->d_partials, ctx->d_sync, pub);
                break;
            default:
                k_normal_eq_tma<kJacAnalytic><<<blocks, kTileThreads, kTileRingBytes, ctx->stream>>>(v, q, ctx->d_partials, ctx->d_sync, pub);
                break;
        }
        ++ctx->launches;
        BG_CUDA_OK(ctx, cudaGetLastError());
        return 0;
    }
    const int blocks = pass_blocks(ctx, s->n, 2);
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

int launch_cost(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, bool publish, bool count_bad) {
    const PassParams q = make_pass_params(p, s->model, 1.0, kJacAnalytic);
    const int blocks = pass_blocks(ctx, s->n, 4);
    Publish pub{ctx->d_result, nullptr, nullptr, 0};
    if (publish) pub = Publish{ctx->d_result, ctx->h_result_dev, ctx->h_seq_dev, ++ctx->seq};
    if (!count_bad && use_tma(ctx, s->n, true)) {
        k_cost_tma<<<tma_blocks(ctx, s->n), kTileThreads, kTileRingBytes, ctx->stream>>>(view_of(s), q, ctx->d_partials, ctx->d_sync, pub);
        ++ctx->launches;
        BG_CUDA_OK(ctx, cudaGetLastError());
        return 0;
    }
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

// Memory that came from cudaMallocAsync goes back to the pool in stream order while its context is alive
// (no device-wide synchronisation, unlike cudaFree); handles that outlive their context fall back to cudaFree.
void free_block(brdfgpu_ctx* ctx, void* block, cudaStream_t stream) {
    if (!block) return;
    if (ctx && ctx->stream == stream && stream) cudaFreeAsync(block, stream);
    else cudaFree(block);
}

int samples_alloc(brdfgpu_ctx* ctx, long n, int model, brdfgpu_samples** out) {
    if (n < 0 || (model != 0 && model != 1)) {
        set_error(ctx, "samples: bad size or model");
        return BRDFGPU_LM_ERROR;
    }
    brdfgpu_samples* s = new brdfgpu_samples;
    s->n = n;
    s->capacity = n;
    s->model = model;
    s->stream = ctx->stream;
    // every array starts on a 128-byte boundary (16-byte vector loads, bulk async copies) and holds an even count
    const size_t stride = ((size_t)(n > 0 ? n : 2) + 15) & ~(size_t)15;
    cudaError_t e = cudaMallocAsync(&s->block, sizeof(double) * 4 * stride, ctx->stream);
    if (e != cudaSuccess) {
        delete s;
        set_error(ctx, std::string("samples: cudaMallocAsync: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    double* base = static_cast<double*>(s->block);
    s->c = base; s->L = base + stride; s->x = base + 2 * stride; s->traw = base + 3 * stride;
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
    // one stream-ordered block (a maintainer may well leave this spot check inside a host loop: no cudaMalloc / cudaFree,
    // which synchronise the device, on a per-call path)
    const size_t nb = sizeof(double) * (size_t)n, ob = jac ? nb * m : nb;
    char* block = nullptr;
    BG_CUDA_OK(ctx, cudaMallocAsync(&block, 2 * nb + ob, ctx->stream));
    double *d_c = reinterpret_cast<double*>(block), *d_t = reinterpret_cast<double*>(block + nb),
           *d_o = reinterpret_cast<double*>(block + 2 * nb);
    cudaError_t e = cudaMemcpyAsync(d_c, angles_host, nb, cudaMemcpyHostToDevice, ctx->stream);
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
    cudaFreeAsync(block, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
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
    static constexpr int kCostBatch = 1;
    static constexpr bool kLanePgWalk = false;
    // Speculative Jacobians as in the persistent kernel (lm_engine.cuh "sites"): a K2 pass costs ~15 % more
    // than a K3 pass and saves a whole pass (plus its launch and ticket round trip) when the point is taken.
    // K2 and K3 sum ||e||^2 in different orders, so here the value may differ from K3's in the last bits.
    static constexpr bool kSpecJac = true;
    brdfgpu_ctx* ctx;
    const brdfgpu_samples* s;
    double delta;
    int jkind;
    bool failed;
    bool spec_on = true, memo_valid = false, sp_trial = false, sp_pg = false;
    int sp_ls = 0;
    double memo_p[3] = {0, 0, 0}, memo[NACC] = {0};
    unsigned long long spec_issued = 0, spec_hits = 0, jac_launched = 0, cost_launched = 0;

    bool run_jac(const double* p) {
        const bool pub = ctx->nranks == 1;
        ++jac_launched;
        if (failed || launch_normal_eq(ctx, s, p, delta, jkind, pub) != 0 || fetch_result(ctx, NACC, pub) != 0) failed = true;
        return !failed;
    }
    static void unpack(const double* r, double* JtJ, double* Jte) {
        JtJ[0] = r[A00]; JtJ[1] = r[A01]; JtJ[2] = r[A02];
        JtJ[3] = r[A01]; JtJ[4] = r[A11]; JtJ[5] = r[A12];
        JtJ[6] = r[A02]; JtJ[7] = r[A12]; JtJ[8] = r[A22];
        Jte[0] = r[G0]; Jte[1] = r[G1]; Jte[2] = r[G2];
    }
    void jac(const double* p, double* JtJ, double* Jte) {
        if (memo_valid && memcmp(p, memo_p, sizeof(memo_p)) == 0) {
            memo_valid = false;
            ++spec_hits;
            unpack(memo, JtJ, Jte);
            return;
        }
        memo_valid = false;
        if (!run_jac(p)) {
            for (int i = 0; i < 9; ++i) JtJ[i] = NAN;
            for (int i = 0; i < 3; ++i) Jte[i] = NAN;
            return;
        }
        unpack(ctx->h_result, JtJ, Jte);
    }
    double cost_site(int site, const double* p, bool& bad) {
        const bool spec = spec_on && (site == kSiteTrial ? sp_trial : site == kSitePgFirst ? sp_pg : site == sp_ls);
        if (!spec) return cost(p, bad);
        memo_valid = false;
        if (!run_jac(p)) {
            bad = true;
            return NAN;
        }
        ++spec_issued;
        for (int k = 0; k < NACC; ++k) memo[k] = ctx->h_result[k];
        for (int i = 0; i < 3; ++i) memo_p[i] = p[i];
        memo_valid = true;
        const double esq = memo[ESQ];
        bad = false;
        if (!lm_finite(esq)) {
            double cnt = 0.0;
            if (count_bad(ctx, s, p, &cnt) != 0) failed = true;
            bad = failed || cnt != 0.0;
        }
        return esq;
    }
    void trial_outcome(bool accepted) { sp_trial = accepted; }
    void ls_outcome(int accepted_probe) { sp_ls = accepted_probe; }
    void pg_outcome(bool took_first) { sp_pg = took_first; }
    void ls_fallback(const double*, const double*, double, const double*, const double*) {}
    void probe_hint(const double*) {}
    bool wants_candidate_hint() const { return false; }
    void candidate_hint(const double*, const double*, double, const double*, const double*) {}

    double cost(const double* p, bool& bad) {
        const bool pub = ctx->nranks == 1;
        ++cost_launched;
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
//
//   * one CTA of 512 threads per SM; every thread runs the same levmar control code (lm_engine.cuh)
//     on the same reduced sums, so the grid takes identical decisions with no broadcast;
//   * the CTA's slice of the sample set is copied ONCE into its shared memory (up to ~210 KB per SM,
//     148 SMs -> ~1.3 * 10^6 samples stay on chip for the whole fit: BASELINE configs[1] never
//     touches HBM or L2 again after the first microseconds); samples beyond that capacity are
//     streamed from global memory every pass (register double buffer);
//   * one grid-wide exchange per evaluation: every CTA publishes its partial sums as 16-byte
//     flagged cells {lo, tag, hi, tag} and polls the cells of all CTAs -- the data IS the barrier
//     (no counter, no second round trip); the sums are then added in a fixed order, so every CTA
//     (and every rank, below) holds identical bits;
//   * the projected-gradient walk hands its candidate points to cost_many() eight at a time: one
//     sweep over the samples and one exchange for eight levmar function evaluations.
// ------------------------------------------------------------------------------------------------
#ifndef BG_PERSIST_THREADS
#define BG_PERSIST_THREADS 512
#endif
constexpr int kPersistThreads = BG_PERSIST_THREADS;  // 512 WORKER threads (16 warps) + one control warp: <= 120 registers per thread
constexpr int kMaxPersistBlocks = 160;  // >= SM count (148 on B200)
constexpr int kGridCostBatch = 8;       // trial points per cost_many() sweep of the sequential walk (<= NACC)
// The lane-parallel projected-gradient walk takes up to 32 candidates per sweep when the whole shard is on chip
// (one candidate per lane of the control warp).  levmar's futile walks try 393 step lengths: at 8 per sweep that is
// 52 sweeps, each paying ~8 k cycles of exchange + control on top of its arithmetic.  Widths 1, 2, 4, 8, then 16 and 32
// once the walk has consumed four times as much (GridEval::pg_walk) need 24.  Candidates past the deciding one are
// discarded uncounted.
constexpr int kWalkMaxBatch = 32;
constexpr int NSUM = kWalkMaxBatch;     // sums of the widest sweep (>= NACC + 2: a Jacobian at one point + the cost at two others)
constexpr long long kSpinCycles = 20000000000LL;  // ~10 s at 2 GHz, then the fit is abandoned (ranks may enter seconds apart on a cold box)

// A cell is two 64-bit words {value.lo | tag << 32, value.hi | tag << 32}; each word is one scalar
// access, so a reader that finds the current tag in both has the whole double -- the data is its own
// flag (no fence, no counter).  SYS = cells written by a peer GPU over NVLink, else GPU scope.
template <bool SYS>
__device__ __forceinline__ void store_cell(uint4* dst, double v, unsigned tag) {
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long w0 = t | (unsigned)__double2loint(v), w1 = t | (unsigned)__double2hiint(v);
    if (SYS) asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
    else asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
}
struct Cell {
    unsigned long long w0, w1;
    __device__ __forceinline__ bool has(unsigned tag) const { return (unsigned)(w0 >> 32) == tag && (unsigned)(w1 >> 32) == tag; }
    __device__ __forceinline__ double value() const { return __hiloint2double((int)(unsigned)w1, (int)(unsigned)w0); }
};
template <bool SYS>
__device__ __forceinline__ Cell load_cell(const uint4* src) {
    Cell c;
    if (SYS) asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(c.w0), "=l"(c.w1) : "l"(src) : "memory");
    else asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(c.w0), "=l"(c.w1) : "l"(src) : "memory");
    return c;
}
// Spin until the cell carries `tag`.
template <bool SYS>
__device__ __forceinline__ double wait_cell(const uint4* src, unsigned tag, int* abort_flag) {
    Cell c = load_cell<SYS>(src);
    if (!c.has(tag)) {
        const long long t0 = clock64();
        unsigned spins = 0;
        do {
            if ((++spins & 0xffu) == 0 && (clock64() - t0 > kSpinCycles || *(volatile int*)abort_flag)) {
                *abort_flag = 1;  // somebody (a peer GPU) never delivered: everybody gives up
                return __longlong_as_double(0x7ff8000000000000LL);  // NaN: the control loop stops with reason 7
            }
            c = load_cell<SYS>(src);
        } while (!c.has(tag));
    }
    return c.value();
}
__device__ __forceinline__ unsigned next_tag(unsigned e) { return e + 1u ? e + 1u : 1u; }  // never 0: buffers start zeroed

// What the control warp asks the CTA to do next (shared memory).
enum SweepKind { kQuit = 0, kSweepJacForward, kSweepJacCentral, kSweepJacAnalytic, kSweepCost, kSweepMany, kSweepBad, kSweepMany16, kSweepMany32 };
struct SweepRequest {
    int kind, cnt;  // cnt: trial points of kSweepMany / index of the point of kSweepBad
    int extra;      // Jacobian sweeps: also ||x - f||^2 at pts[0] (.. pts[extra-1]), sums number NACC (, NACC+1)
    PassParams q;   // Jacobian sweeps
    CostPoint pts[kWalkMaxBatch];
};

// Everything a sweep needs, written once at kernel start.  Shared memory on purpose: the sweeps are
// separate (noinline) functions and local memory -- where a context object passed by pointer would
// live -- costs ~350 cycles per dependent access on B200 (profiles/micro/lat.cu); LDS costs ~30.
struct FitContext {
    SampleView v;
    int model;
    unsigned sc, sl, sx;  // resident slice (shared-window addresses), in sample PAIRS
    int res_pairs;        // pairs held by this CTA
    long res_first;       // global index of its first pair
    long stream_first;    // pairs >= stream_first are streamed from global memory by the whole grid
    uint4* cells;         // [2 parities][kMaxPersistBlocks][NSUM], cell_index()
    double* stage;        // transposition area of the exchange: stage_q quantities x kStagePitch CTAs
    int stage_q;          // quantities per round of the transposition (kStageWide or kStageNarrow)
    PeerView peer;        // nranks == 1: no cross-GPU step
    int* abort_flag;      // global
};

__shared__ FitContext s_ctx;
__shared__ SweepRequest s_req;
__shared__ unsigned s_epoch, s_peer_epoch;  // tags of the last grid / peer exchange (uniform in the CTA)
__shared__ double s_red[(kPersistThreads / 32) * NSUM];
__shared__ double s_res[NSUM];
__shared__ double s_pstage[kMaxRanks * NSUM];
__shared__ double s_cand[kGridCostBatch * 3];  // candidate points of the projected-gradient walk
__shared__ double s_cand_cost[kGridCostBatch];
__shared__ int s_cand_bad[kGridCostBatch];
__shared__ double s_memo[3 + NACC];  // speculative Jacobian: the point, then A00..A22, G0..G2, ||e||^2 (GridEval::cost_site)
#ifdef BG_CTL_TICKS
// debug variant (make VARIANT=_ticks EXTRA=-DBG_CTL_TICKS): control-code cycles between the evaluator's hooks, sweeps
// and exchanges excluded, printed by CTA 0 at the end of the fit
__shared__ long long s_tick[12], s_tick_last, s_tick_busy;
__shared__ int s_tick_prev;
#define BG_TICK(id) tick(id)
#else
#define BG_TICK(id) ((void)0)
#endif
__shared__ double s_hint[3];         // first projected-gradient candidate announced by the line search (GridEval::ls_fallback)
__shared__ double s_ahead[6 + 4];    // announced before the trial: the lambda = 0.1 probe, the walk's first candidate;
                                     // then the probe's evaluated point and ||e||^2 (GridEval::probe_hint, candidate_hint)
__shared__ long long s_cyc[6];  // thread 0: cycles in sweeps, exchanges, and the 4 exchange phases
__shared__ long long s_ctl[8];  // thread 0: control-code cycles by the kind of request they led to (wide batches count as kSweepMany); [7] = time of the last exchange end
// TMA ring of the streamed part (sample sets beyond on-chip residency): 3 stages x 48 KB
using PersistRing = TileRing<kPersistThreads, 3>;
__shared__ PersistRing s_ring;
__shared__ __align__(8) unsigned long long s_ring_bars[2 * 3];
__shared__ long s_ring_seq;  // tiles that went through the ring so far (uniform in the CTA)

// Roles.  Threads 0..511 (16 warps) are the WORKERS: they own the samples and run the sweeps and the exchanges; warp
// 16 is the CONTROL warp running lm_engine.cuh and never enters a sweep.  Kept apart, the control code no longer
// carries the sweeps' code and registers through every evaluation site of the engine (kernel spills 4.0 -> 0.9 KB,
// control cycles per sweep 12.5 k -> 9.5 k at configs[1]); 17 warps are allocated like 20, which caps the kernel at
// 96 registers per thread (the Jacobian sweeps pay ~1 k cycles for it).
// worker_sync(): the 512 workers (named barrier 1); __syncthreads(): all 544 threads, the two hand-overs of a sweep.
#ifndef BG_SWEEP_FN
#define BG_SWEEP_FN __noinline__
#endif
constexpr int kControlWarpThreads = 32;
constexpr int kPersistBlockThreads = kPersistThreads + kControlWarpThreads;
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kPersistThreads) : "memory"); }
__device__ __forceinline__ int ctl_lane() { return (int)threadIdx.x - kPersistThreads; }

// Cross-GPU step, fused into the same kernel: CTA 0 of every rank stores its rank's sums as flagged
// cells into slot [parity][rank] of EVERY rank's exchange buffer (peer stores over NVLink /
// NVSwitch); every CTA polls its own rank's buffer until all slots carry the tag and adds them in
// rank order -- identical bits on all ranks, so all ranks take identical LM decisions.  Two
// parities: a rank can be at most one exchange ahead of any CTA of any peer.
template <int NV>
__device__ __forceinline__ void peer_exchange() {
    const unsigned tag = next_tag(s_peer_epoch);
    const int par = (int)(tag & 1u);
    const int nranks = s_ctx.peer.nranks;
    const int r = threadIdx.x / NV, k = threadIdx.x - r * NV;
    if (threadIdx.x < NV * nranks) {
        if (blockIdx.x == 0)
            store_cell<true>(s_ctx.peer.remote[r] + ((long)(par * kMaxRanks + s_ctx.peer.rank) * kPeerCellsPerRank + k), s_res[k], tag);
        s_pstage[r * NV + k] =
            wait_cell<true>(s_ctx.peer.local + ((long)(par * kMaxRanks + r) * kPeerCellsPerRank + k), tag, s_ctx.abort_flag);
    }
    worker_sync();
    if (threadIdx.x < NV) {
        double sum = 0.0;
        for (int q = 0; q < nranks; ++q) sum += s_pstage[q * NV + threadIdx.x];
        s_res[threadIdx.x] = sum;
    }
    if (threadIdx.x == 0) s_peer_epoch = tag;
    worker_sync();
}

// The sums of all CTAs (and all ranks) in s_res[0..NV), identical bits everywhere.  Latency is
// everything here (one exchange per evaluation): every step is written so that the NV quantities
// advance together instead of one after the other.
//
// Stage 1 (warp_sums_to_smem): warp butterflies, step-major; lane 0 of every warp leaves its NV sums in s_red.
// Stage 2 (grid_exchange): 16 lanes per quantity add the 16 warps and publish the CTA's sum as a flagged cell; then
// all threads collect the cells of ALL CTAs into shared memory (all loads go out together, late ones are re-polled)
// and one warp per quantity adds them (lane l takes CTAs l, l + 32, ... in ascending order, then a butterfly) -- a
// fixed order, so every CTA (and every rank, below) holds identical bits.  The transposition takes 16 (or 8, when the
// shard needs the shared memory) quantities per round; the loads behind it all go out at once.
template <int NV>
__device__ __forceinline__ void warp_sums_to_smem(const double* acc, int first = 0, int stride = NV) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) s[k] = acc[k];
#pragma unroll
    for (int off = 16; off; off >>= 1) {
#pragma unroll
        for (int k = 0; k < NV; ++k) s[k] += __shfl_down_sync(0xffffffffu, s[k], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) s_red[warp * stride + first + k] = s[k];
    }
}

// cell of quantity k of CTA b.  CTA-major on purpose: a CTA's cells share lines with at most one neighbour.  With
// the quantity-major order (8 CTAs writing into every 128-byte line while 148 poll it) an exchange took 2x longer.
__device__ __forceinline__ int cell_index(int b, int k) { return b * NSUM + k; }
constexpr int kStagePitch = kMaxPersistBlocks + 1;  // quantity-major with an odd pitch: conflict-free both ways
// The transposition area of the exchange lives in dynamic shared memory behind the shard (and the TMA ring): 16
// quantities per round when that fits, 8 when the shard needs the room (persistent_plan()).  Shared memory is what the
// resident shard lives in -- and what is left of the 256 KB is the L1 that catches the spills of the Jacobian sweeps
// and of the control warp: configs[1] with 16 quantities stays within the 196 KB carve-out (60 KB of L1); with the
// full 32 it crossed into the 228 KB one and the Jacobian sweeps took 11 % longer.
constexpr int kStageWide = 16, kStageNarrow = 8;
constexpr size_t stage_bytes(int q) { return sizeof(double) * kStagePitch * (size_t)q; }
// Not inlined: with its own register allocation the cells in flight do not push the sweeps' accumulators into local
// memory (inlined into the Jacobian sweeps it cost them up to 1.3 KB of spills and ~3 k cycles per sweep).
template <int NV>
__device__ __noinline__ void grid_exchange(long long t_sweep_start, long long t_in) {
    constexpr int kWarps = kPersistThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grid = gridDim.x;
    worker_sync();
    const long long t_a = clock64();
    // across the 16 warps: 16 lanes per quantity, 4 butterfly steps; publish
    const unsigned tag = next_tag(s_epoch);
    uint4* base = s_ctx.cells + (long)(tag & 1u) * NSUM * kMaxPersistBlocks;
    if (threadIdx.x < ((NV * kWarps + 31) & ~31)) {
        const int k = threadIdx.x / kWarps, w = threadIdx.x % kWarps;
        double t = (k < NV) ? s_red[w * NV + k] : 0.0;
#pragma unroll
        for (int off = kWarps / 2; off; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off, kWarps);
        if (k < NV && w == 0) store_cell<false>(base + cell_index(blockIdx.x, k), t, tag);
    }
    const long long t_b = clock64();
    // every thread collects a few cells (consecutive threads = consecutive quantities of one CTA): ALL loads go out
    // together, late ones are re-polled; then, stage_q quantities per round, the values are transposed through shared
    // memory and one warp per quantity adds them in ascending CTA order
    constexpr int kPerThread = (kMaxPersistBlocks * NV + kPersistThreads - 1) / kPersistThreads;
    double val[kPerThread];
    {
        // (re-)load every cell that is still missing, all in one go, until none is: a round costs one L2 round trip
        // whatever the number of late cells (this thread polls right after publishing, so most cells ARE late)
        unsigned pending = 0u;
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            val[j] = 0.0;
            if ((threadIdx.x + j * kPersistThreads) / NV < grid) pending |= 1u << j;
        }
        const long long t_poll = clock64();
        for (unsigned spins = 0; pending; ++spins) {
            Cell c[kPerThread];
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const int idx = threadIdx.x + j * kPersistThreads;
                const int b = idx / NV, k = idx - b * NV;
                if (pending & (1u << j)) c[j] = load_cell<false>(base + cell_index(b, k));
            }
#pragma unroll
            for (int j = 0; j < kPerThread; ++j)
                if ((pending & (1u << j)) && c[j].has(tag)) {
                    val[j] = c[j].value();
                    pending &= ~(1u << j);
                }
            if (pending && (spins & 0xffu) == 0xffu && (clock64() - t_poll > kSpinCycles || *(volatile int*)s_ctx.abort_flag)) {
                *s_ctx.abort_flag = 1;  // somebody never delivered: everybody gives up (the control loop stops on the NaN)
                val[0] = __longlong_as_double(0x7ff8000000000000LL);
                pending = 0u;
            }
        }
    }
    double* const stage = s_ctx.stage;
    const int stage_q = s_ctx.stage_q;
    for (int k0 = 0; k0 < NV; k0 += stage_q) {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            const int idx = threadIdx.x + j * kPersistThreads;
            const int b = idx / NV, k = idx - b * NV;
            if (b < grid && k >= k0 && k < k0 + stage_q) stage[(k - k0) * kStagePitch + b] = val[j];
        }
        worker_sync();
        for (int k = k0 + warp; k < NV && k < k0 + stage_q; k += kWarps) {
            double t = 0.0;
            for (int b = lane; b < grid; b += 32) t += stage[(k - k0) * kStagePitch + b];
#pragma unroll
            for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) s_res[k] = t;
        }
        if (k0 + stage_q < NV) worker_sync();
    }
    const long long t_c = clock64();
    worker_sync();  // (everybody read s_epoch before this barrier)
    if (threadIdx.x == 0) s_epoch = tag;
    if (s_ctx.peer.nranks > 1) peer_exchange<NV>();
    if (threadIdx.x == 0) {
        const long long t_d = clock64();
        s_cyc[0] += t_in - t_sweep_start; s_cyc[1] += t_d - t_in;
        s_cyc[2] += t_a - t_in; s_cyc[3] += t_b - t_a; s_cyc[4] += t_c - t_b; s_cyc[5] += t_d - t_c;
        s_ctl[7] = t_d;
    }
}

template <int NV>
__device__ __forceinline__ void all_reduce(const double* acc, long long t_sweep_start) {
    const long long t_in = clock64();
    warp_sums_to_smem<NV>(acc);
    grid_exchange<NV>(t_sweep_start, t_in);
}

// ---- sweeps (all threads of the CTA) ----
// ||e||^2 of a Jacobian sweep is accumulated in exactly the order the cost sweeps use (two running sums
// over alternate pairs, resident part first, then the streamed part), so a trial point evaluated by a
// Jacobian sweep (speculation, GridEval::cost_site) gets the very bits a cost sweep would give it.
// EXTRA (0, 1, 2): the same sweep also sums ||x - f||^2 at the trial points s_req.pts[0 .. EXTRA-1] (sums number
// NACC, NACC+1), again in the order of the cost sweeps.
template <int JAC, int EXTRA>
__device__ BG_SWEEP_FN void jac_sweep() {
    const long long t0 = clock64();
    const PassParams q = s_req.q;
    const CostPoint xq = s_req.pts[0], zq = s_req.pts[EXTRA > 1 ? 1 : 0];
    double xa = 0.0, xb = 0.0, za = 0.0, zb = 0.0;
    const unsigned sc = s_ctx.sc, sl = s_ctx.sl, sx = s_ctx.sx;
    const int res_pairs = s_ctx.res_pairs;
    const long res_first = s_ctx.res_first;
    const double* traw = s_ctx.v.traw;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    double esq_a = 0.0, esq_b = 0.0;
    int i = threadIdx.x;
    for (; i + kPersistThreads < res_pairs; i += 2 * kPersistThreads) {
        const int i2 = i + kPersistThreads;
        const double2 c = lds_pair(sc, i), l = lds_pair(sl, i), x = lds_pair(sx, i);
        const double2 d = lds_pair(sc, i2), m = lds_pair(sl, i2), y = lds_pair(sx, i2);
        acc[ESQ] = esq_a;
        accumulate_jac_pair<JAC>(q, s_req.q, c, l, x, traw, 2 * (res_first + i), acc);
        esq_a = acc[ESQ];
        acc[ESQ] = esq_b;
        accumulate_jac_pair<JAC>(q, s_req.q, d, m, y, traw, 2 * (res_first + i2), acc);
        esq_b = acc[ESQ];
        if (EXTRA >= 1) accumulate_cost_2pairs(xq, c, l, x, 2 * (res_first + i), d, m, y, 2 * (res_first + i2), traw, &xa, &xb);
        if (EXTRA >= 2) accumulate_cost_2pairs(zq, c, l, x, 2 * (res_first + i), d, m, y, 2 * (res_first + i2), traw, &za, &zb);
    }
    if (i < res_pairs) {
        const double2 c = lds_pair(sc, i), l = lds_pair(sl, i), x = lds_pair(sx, i);
        acc[ESQ] = esq_a;
        accumulate_jac_pair<JAC>(q, s_req.q, c, l, x, traw, 2 * (res_first + i), acc);
        esq_a = acc[ESQ];
        if (EXTRA >= 1) accumulate_cost_pair(xq, c, l, x, traw, 2 * (res_first + i), &xa);
        if (EXTRA >= 2) accumulate_cost_pair(zq, c, l, x, traw, 2 * (res_first + i), &za);
    }
    acc[ESQ] = esq_a + esq_b;
    double extra[2] = {xa + xb, za + zb};
    if (s_ctx.stream_first < (s_ctx.v.n >> 1)) {
        const SampleView v = s_ctx.v;
        // streamed part: every group of two pairs adds (pair 0 + pair 1) to the running sums -- the order of
        // many_sweep and cost_sweep, so a point gets the same ||e||^2 bits from any kind of sweep
        double esq_run = acc[ESQ], x_run = extra[0], z_run = extra[1];
        const long seq = s_ring.stream(v, s_ctx.stream_first, s_ring_seq,
            [&](double2 c0, double2 l0, double2 x0, long i0, double2 c1, double2 l1, double2 x1, long i1, int valid) {
                double sa = 0.0, sb = 0.0;
                if (valid >= 1) {
                    acc[ESQ] = sa;
                    accumulate_jac_pair<JAC>(q, s_req.q, c0, l0, x0, v.traw, i0, acc);
                    sa = acc[ESQ];
                }
                if (valid >= 2) {
                    acc[ESQ] = sb;
                    accumulate_jac_pair<JAC>(q, s_req.q, c1, l1, x1, v.traw, i1, acc);
                    sb = acc[ESQ];
                }
                esq_run += sa + sb;
                if (EXTRA >= 1) {
                    double ya = 0.0, yb = 0.0;
                    if (valid == 2) accumulate_cost_2pairs(xq, c0, l0, x0, i0, c1, l1, x1, i1, v.traw, &ya, &yb);
                    else if (valid == 1) accumulate_cost_pair(xq, c0, l0, x0, v.traw, i0, &ya);
                    x_run += ya + yb;
                }
                if (EXTRA >= 2) {
                    double wa = 0.0, wb = 0.0;
                    if (valid == 2) accumulate_cost_2pairs(zq, c0, l0, x0, i0, c1, l1, x1, i1, v.traw, &wa, &wb);
                    else if (valid == 1) accumulate_cost_pair(zq, c0, l0, x0, v.traw, i0, &wa);
                    z_run += wa + wb;
                }
            });
        acc[ESQ] = esq_run;
        extra[0] = x_run;
        extra[1] = z_run;
        worker_sync();  // everybody has read s_ring_seq
        if (threadIdx.x == 0) s_ring_seq = seq;
    }
    if ((s_ctx.v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long j = s_ctx.v.n - 1;
        accumulate_jac<JAC>(s_req.q, s_ctx.v.c[j], s_ctx.v.L[j], s_ctx.v.x[j], traw, j, acc);
        if (EXTRA >= 1) accumulate_cost(xq, s_ctx.v.c[j], s_ctx.v.L[j], s_ctx.v.x[j], traw, j, &extra[0]);
        if (EXTRA >= 2) accumulate_cost(zq, s_ctx.v.c[j], s_ctx.v.L[j], s_ctx.v.x[j], traw, j, &extra[1]);
    }
    if (EXTRA >= 1) {
        double all[NACC + EXTRA];
#pragma unroll
        for (int k = 0; k < NACC; ++k) all[k] = acc[k];
        all[NACC] = extra[0];
        if (EXTRA >= 2) all[NACC + EXTRA - 1] = extra[1];
        all_reduce<NACC + EXTRA>(all, t0);
    } else {
        all_reduce<NACC>(acc, t0);
    }
}

// sum of squared residuals at one point over this thread's share of the resident samples
__device__ __forceinline__ double resident_cost(const CostPoint& q, unsigned sc, unsigned sl, unsigned sx, int res_pairs,
                                                long res_first, const double* traw) {
    double a0 = 0.0, a1 = 0.0;  // two pairs in flight: four independent exp chains per thread
    int i = threadIdx.x;
    for (; i + kPersistThreads < res_pairs; i += 2 * kPersistThreads) {
        const int i2 = i + kPersistThreads;
        const double2 c = lds_pair(sc, i), l = lds_pair(sl, i), x = lds_pair(sx, i);
        const double2 d = lds_pair(sc, i2), m = lds_pair(sl, i2), y = lds_pair(sx, i2);
        accumulate_cost_2pairs(q, c, l, x, 2 * (res_first + i), d, m, y, 2 * (res_first + i2), traw, &a0, &a1);
    }
    if (i < res_pairs) {
        const double2 c = lds_pair(sc, i), l = lds_pair(sl, i), x = lds_pair(sx, i);
        accumulate_cost_pair(q, c, l, x, traw, 2 * (res_first + i), &a0);
    }
    return a0 + a1;
}

// two trial points at once over the resident samples (eight exp chains per thread; each point's sum is
// accumulated in exactly the order resident_cost uses, so batching never changes a value)
__device__ __forceinline__ void resident_cost_x2(const CostPoint& qa_in, const CostPoint& qb_in, unsigned sc, unsigned sl, unsigned sx,
                                                 int res_pairs, long res_first, const double* traw, double* out_a, double* out_b) {
    const CostPoint qa = qa_in, qb = qb_in;  // (callers pass shared memory: read once, not once per trip)
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    int i = threadIdx.x;
    for (; i + kPersistThreads < res_pairs; i += 2 * kPersistThreads) {
        const int i2 = i + kPersistThreads;
        const double2 c = lds_pair(sc, i), l = lds_pair(sl, i), x = lds_pair(sx, i);
        const double2 d = lds_pair(sc, i2), m = lds_pair(sl, i2), y = lds_pair(sx, i2);
        const long g = 2 * (res_first + i), h = 2 * (res_first + i2);
        double ea[4], eb[4];
        const double cc[4] = {c.x, c.y, d.x, d.y}, ll[4] = {l.x, l.y, m.x, m.y}, xx[4] = {x.x, x.y, y.x, y.y};
        const long idx[4] = {g, g + 1, h, h + 1};
        residuals_n_x2<4>(qa, qb, cc, ll, xx, traw, idx, ea, eb);
        a0 = __fma_rn(ea[0], ea[0], a0); a0 = __fma_rn(ea[1], ea[1], a0);
        a1 = __fma_rn(ea[2], ea[2], a1); a1 = __fma_rn(ea[3], ea[3], a1);
        b0 = __fma_rn(eb[0], eb[0], b0); b0 = __fma_rn(eb[1], eb[1], b0);
        b1 = __fma_rn(eb[2], eb[2], b1); b1 = __fma_rn(eb[3], eb[3], b1);
    }
    if (i < res_pairs) {
        const double2 c = lds_pair(sc, i), l = lds_pair(sl, i), x = lds_pair(sx, i);
        accumulate_cost_pair(qa, c, l, x, traw, 2 * (res_first + i), &a0);
        accumulate_cost_pair(qb, c, l, x, traw, 2 * (res_first + i), &b0);
    }
    *out_a = a0 + a1;
    *out_b = b0 + b1;
}

__device__ BG_SWEEP_FN void cost_sweep() {
    const long long t0 = clock64();
    const CostPoint q = s_req.pts[0];
    double acc[1];
    acc[0] = resident_cost(q, s_ctx.sc, s_ctx.sl, s_ctx.sx, s_ctx.res_pairs, s_ctx.res_first, s_ctx.v.traw);
    if (s_ctx.stream_first < (s_ctx.v.n >> 1)) {
        const SampleView v = s_ctx.v;
        double run = acc[0];  // (same order as many_sweep and jac_sweep: pair 0 + pair 1 of every group, in sequence)
        const long seq = s_ring.stream(v, s_ctx.stream_first, s_ring_seq,
            [&](double2 c0, double2 l0, double2 x0, long i0, double2 c1, double2 l1, double2 x1, long i1, int valid) {
                double a0 = 0.0, a1 = 0.0;
                if (valid == 2) accumulate_cost_2pairs(q, c0, l0, x0, i0, c1, l1, x1, i1, v.traw, &a0, &a1);
                else if (valid == 1) accumulate_cost_pair(q, c0, l0, x0, v.traw, i0, &a0);
                run += a0 + a1;
            });
        acc[0] = run;
        worker_sync();
        if (threadIdx.x == 0) s_ring_seq = seq;
    }
    if ((s_ctx.v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long j = s_ctx.v.n - 1;
        accumulate_cost(q, s_ctx.v.c[j], s_ctx.v.L[j], s_ctx.v.x[j], s_ctx.v.traw, j, acc);
    }
    all_reduce<1>(acc, t0);
}

// number of non-finite residuals at point s_req.pts[s_req.cnt] (rare second sweep)
__device__ BG_SWEEP_FN void bad_sweep() {
    const long long t0 = clock64();
    const CostPoint q = s_req.pts[s_req.cnt];
    const SampleView v = s_ctx.v;
    double cnt[1] = {0.0};
    for (int i = threadIdx.x; i < s_ctx.res_pairs; i += kPersistThreads) {
        const double2 c = lds_pair(s_ctx.sc, i), l = lds_pair(s_ctx.sl, i), x = lds_pair(s_ctx.sx, i);
        const long g = 2 * (s_ctx.res_first + i);
        cnt[0] += lm_finite(residual_of(q, c.x, l.x, x.x, v.traw, g)) ? 0.0 : 1.0;
        cnt[0] += lm_finite(residual_of(q, c.y, l.y, x.y, v.traw, g + 1)) ? 0.0 : 1.0;
    }
    const long nth = (long)gridDim.x * kPersistThreads;
    for (long i = 2 * s_ctx.stream_first + (long)blockIdx.x * kPersistThreads + threadIdx.x; i < v.n; i += nth)
        cnt[0] += lm_finite(residual_of(q, v.c[i], v.L[i], v.x[i], v.traw, i)) ? 0.0 : 1.0;
    all_reduce<1>(cnt, t0);
}

// up to kGridCostBatch trial points in ONE sweep + ONE exchange
__device__ BG_SWEEP_FN void many_sweep() {
    const long long t0 = clock64();
    const int cnt = s_req.cnt;
    const unsigned sc = s_ctx.sc, sl = s_ctx.sl, sx = s_ctx.sx;
    const int res_pairs = s_ctx.res_pairs;
    const long res_first = s_ctx.res_first;
    const double* traw = s_ctx.v.traw;
    double acc[kGridCostBatch];
#pragma unroll
    for (int k = 0; k < kGridCostBatch; ++k) acc[k] = 0.0;
    // resident slice: candidate-outer, parameters in registers, samples re-read from shared memory
#pragma unroll
    for (int k = 0; k < kGridCostBatch; k += 2) {
        if (k + 1 < cnt) resident_cost_x2(s_req.pts[k], s_req.pts[k + 1], sc, sl, sx, res_pairs, res_first, traw, &acc[k], &acc[k + 1]);
        else if (k < cnt) acc[k] = resident_cost(s_req.pts[k], sc, sl, sx, res_pairs, res_first, traw);
    }
    // streamed remainder: sample-outer (one trip through the ring for all candidates)
    const SampleView v = s_ctx.v;
    if (s_ctx.stream_first < (v.n >> 1)) {
        const long seq = s_ring.stream(v, s_ctx.stream_first, s_ring_seq,
            [&](double2 c0, double2 l0, double2 x0, long i0, double2 c1, double2 l1, double2 x1, long i1, int valid) {
#pragma unroll
                for (int k = 0; k < kGridCostBatch; ++k) {
                    if (k < cnt) {
                        const CostPoint q = s_req.pts[k];
                        double a0 = 0.0, a1 = 0.0;
                        if (valid == 2) accumulate_cost_2pairs(q, c0, l0, x0, i0, c1, l1, x1, i1, v.traw, &a0, &a1);
                        else if (valid == 1) accumulate_cost_pair(q, c0, l0, x0, v.traw, i0, &a0);
                        acc[k] += a0 + a1;
                    }
                }
            });
        worker_sync();
        if (threadIdx.x == 0) s_ring_seq = seq;
    }
    if ((v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long j = v.n - 1;
#pragma unroll
        for (int k = 0; k < kGridCostBatch; ++k)
            if (k < cnt) accumulate_cost(s_req.pts[k], v.c[j], v.L[j], v.x[j], v.traw, j, &acc[k]);
    }
    all_reduce<kGridCostBatch>(acc, t0);
}

#if defined(BG_RUN_SWEEP_NOINLINE)
__device__ __noinline__ void run_sweep(int kind) {
#else
// up to 16 / 32 trial points in ONE sweep + ONE exchange, whole shard on chip (the lane-parallel walk's wide batches).
// The candidates are taken two at a time and each pair's sums are warp-reduced at once, so no thread ever holds 32
// accumulators; per candidate the arithmetic -- and with it the value -- is exactly many_sweep's and cost_sweep's.
template <int NV>
__device__ BG_SWEEP_FN void many_sweep_wide() {
    const long long t0 = clock64();
    const int cnt = s_req.cnt;
    const unsigned sc = s_ctx.sc, sl = s_ctx.sl, sx = s_ctx.sx;
    const int res_pairs = s_ctx.res_pairs;
    const long res_first = s_ctx.res_first;
    const SampleView v = s_ctx.v;
    const bool odd_tail = (v.n & 1) && blockIdx.x == 0 && threadIdx.x == 0;
#pragma unroll 1
    for (int k = 0; k < NV; k += 2) {
        double a[2] = {0.0, 0.0};
        if (k + 1 < cnt) resident_cost_x2(s_req.pts[k], s_req.pts[k + 1], sc, sl, sx, res_pairs, res_first, v.traw, &a[0], &a[1]);
        else if (k < cnt) a[0] = resident_cost(s_req.pts[k], sc, sl, sx, res_pairs, res_first, v.traw);
        if (odd_tail) {
            const long j = v.n - 1;
            if (k < cnt) accumulate_cost(s_req.pts[k], v.c[j], v.L[j], v.x[j], v.traw, j, &a[0]);
            if (k + 1 < cnt) accumulate_cost(s_req.pts[k + 1], v.c[j], v.L[j], v.x[j], v.traw, j, &a[1]);
        }
        warp_sums_to_smem<2>(a, k, NV);
    }
    grid_exchange<NV>(t0, clock64());
}

__device__ __forceinline__ void run_sweep(int kind) {
#endif
    switch (kind) {
        case kSweepJacForward:
            if (s_req.extra == 2) jac_sweep<kJacForward, 2>(); else if (s_req.extra) jac_sweep<kJacForward, 1>(); else jac_sweep<kJacForward, 0>();
            break;
        case kSweepJacCentral:
            if (s_req.extra == 2) jac_sweep<kJacCentral, 2>(); else if (s_req.extra) jac_sweep<kJacCentral, 1>(); else jac_sweep<kJacCentral, 0>();
            break;
        case kSweepJacAnalytic:
            if (s_req.extra == 2) jac_sweep<kJacAnalytic, 2>(); else if (s_req.extra) jac_sweep<kJacAnalytic, 1>(); else jac_sweep<kJacAnalytic, 0>();
            break;
        case kSweepCost: cost_sweep(); break;
        case kSweepMany: many_sweep(); break;
        case kSweepMany16: many_sweep_wide<16>(); break;
        case kSweepMany32: many_sweep_wide<32>(); break;
        default: bad_sweep(); break;
    }
}

// the workers: serve sweeps until the control warp quits
__device__ __forceinline__ void serve_sweeps() {
    for (;;) {
        __syncthreads();  // the request is posted
        const int kind = s_req.kind;
        if (kind == kQuit) return;
        run_sweep(kind);
        __syncthreads();  // the sums are in s_res
    }
}

// The Evaluator of lm_engine.cuh as seen by the control warp (warp 16 of every CTA): each call posts
// a SweepRequest, waits for the workers and reads the reduced sums.  All 512 threads running the control
// code redundantly cost more than the sweeps themselves.
struct GridEval {
    static constexpr int kCostBatch = kGridCostBatch;
#ifndef BG_LANE_PG_WALK
#define BG_LANE_PG_WALK 1
#endif
    static constexpr bool kLanePgWalk = BG_LANE_PG_WALK != 0;
    // Speculative Jacobians (lm_engine.cuh): a cost request at a point that is likely to become the next
    // iterate is served by a Jacobian sweep (same ||e||^2 bits, see jac_sweep) and the sums are kept in
    // s_memo; jac() at a bit-identical point then needs no sweep, no exchange and no control round.
    // Prediction per site is "what happened last time": the LM trial was accepted / the line search
    // accepted probe number k / the projected-gradient walk took its first candidate.  Every CTA (and
    // rank) sees identical sums, so all of them predict alike.
    static constexpr bool kSpecJac = true;
    int model, jkind;
    double delta;
    unsigned jac_passes, cost_passes, cost_points, spec_issued, spec_hits;
    bool narrow_walk;  // BRDFGPU_SPEC_JAC & 64 (tests): at most 8 candidates per sweep, as for a streamed shard
    bool spec_on, fuse_on, width_on, memo_valid, sp_trial, sp_pg, hint_valid, ahead_valid, cand_valid, probe_known, sp_clip;
    unsigned creep_fused;
    int sp_ls;    // probe number the last line search accepted (0: it failed)
    int pg_last;  // candidates the last projected-gradient walk consumed

#ifdef BG_CTL_TICKS
    // segment ids: the hook that ENDS the segment -- 0 jac() entry (end of iteration + loop head), 1 probe_hint
    // (after the Jacobian: gradient tests, LU solve, projection, probe), 2 trial cost_site (candidate hint, set-up),
    // 3 trial_outcome, 4 line-search probe cost_site (rejection algebra, pow, search set-up, first backtrack),
    // 5 ls_outcome, 6 pg_walk entry, 7 pg_walk exit (walk control without its sweeps), 8 other cost_site
    __device__ __forceinline__ void tick(int id) {
        if (ctl_lane() == 0) {
            const long long now = clock64(), busy = s_cyc[0] + s_cyc[1];
            s_tick[id] += (now - s_tick_last) - (busy - s_tick_busy);
            s_tick_last = now;
            s_tick_busy = busy;
        }
    }
#endif
#if defined(BG_POST_NOINLINE)
    __device__ __noinline__ void post(int kind) {
#else
    __device__ __forceinline__ void post(int kind) {
#endif
        if (ctl_lane() == 0) {
            s_req.kind = kind;
            s_ctl[kind >= kSweepMany16 ? (int)kSweepMany : kind] += clock64() - s_ctl[7];
        }
        __syncthreads();                     // hand the request to the workers ...
        if (kind != kQuit) __syncthreads();  // ... and wait for the sums
    }

    __device__ __forceinline__ int jac_sweep_kind() const {
        return jkind == kJacForward ? kSweepJacForward : jkind == kJacCentral ? kSweepJacCentral : kSweepJacAnalytic;
    }
    static __device__ __forceinline__ bool same_bits(double a, double b) {
        return __double_as_longlong(a) == __double_as_longlong(b);
    }

    // post a Jacobian sweep at q; extra: also the cost at s_req.pts[0]
    __device__ __forceinline__ void post_jac(const PassParams& q, int extra) {
        if (ctl_lane() == 0) {
            s_req.q = q;
            s_req.extra = extra;
        }
        post(jac_sweep_kind());
        ++jac_passes;
    }
    // keep the sums of the Jacobian sweep that just ended, as the Jacobian (and cost) of `pt`
    __device__ __forceinline__ void keep_jac(const double* pt) {
        __syncwarp();  // nobody still reads the previous memo
        if (ctl_lane() < NACC) s_memo[3 + ctl_lane()] = s_res[ctl_lane()];  // A00..A22, G0..G2, ESQ are 0..9
        if (ctl_lane() == 0) { s_memo[0] = pt[0]; s_memo[1] = pt[1]; s_memo[2] = pt[2]; }
        __syncwarp();
        memo_valid = true;
        ++spec_issued;
    }
    __device__ __forceinline__ bool memo_is(const double* pt) const {
        return memo_valid && same_bits(pt[0], s_memo[0]) && same_bits(pt[1], s_memo[1]) && same_bits(pt[2], s_memo[2]);
    }

    __device__ __forceinline__ void jac(const double* p, double* JtJ, double* Jte) {
        BG_TICK(0);
        const double* r = s_res;
        if (memo_is(p)) {
            ++spec_hits;  // the sums of this very point are already here
            r = s_memo + 3;
        } else {
            post_jac(make_pass_params(p, model, delta, jkind), 0);
        }
        memo_valid = false;
        JtJ[0] = r[A00]; JtJ[1] = r[A01]; JtJ[2] = r[A02];
        JtJ[3] = r[A01]; JtJ[4] = r[A11]; JtJ[5] = r[A12];
        JtJ[6] = r[A02]; JtJ[7] = r[A12]; JtJ[8] = r[A22];
        Jte[0] = r[G0]; Jte[1] = r[G1]; Jte[2] = r[G2];
    }

    // ||x - f(p)||^2 through a Jacobian sweep at p; the normal-equation sums are kept for jac(p)
    __device__ __forceinline__ double cost_with_jac(const double* p, bool& bad) {
        if (ctl_lane() == 0) s_req.pts[0] = make_cost_point(p, model);  // count_bad(0), should the sum come out non-finite
        post_jac(make_pass_params(p, model, delta, jkind), 0);
        const double esq = s_res[ESQ];
        keep_jac(p);
        bad = false;
        if (!lm_finite(esq)) bad = count_bad(0) != 0.0;
        return esq;
    }

    // ||x - f(p)||^2 and, in the same sweep, the Jacobian (and cost) at the announced first candidate h of the
    // projected-gradient walk that follows if this last line-search probe is not accepted
    __device__ __forceinline__ double cost_and_jac_at(const double* p, const double* h, bool& bad) {
        if (ctl_lane() == 0) s_req.pts[0] = make_cost_point(p, model);
        post_jac(make_pass_params(h, model, delta, jkind), 1);
        const double esq = s_res[NACC];
        keep_jac(h);
        bad = false;
        if (!lm_finite(esq)) bad = count_bad(0) != 0.0;
        return esq;
    }

    // first candidate of a projected-gradient walk from p along -g with step t (lmbc_core.c:886-889);
    // one spelling for the announcement and the walk itself, so the bits agree
    static __device__ __forceinline__ void pg_candidate(const double* p, const double* g, double tc, const Box& box, double* cand) {
#pragma unroll
        for (int i = 0; i < 3; ++i) cand[i] = p[i] - tc * g[i];
        box_project<3>(cand, box, 3);
    }
    __device__ __forceinline__ void ls_fallback(const double* p, const double* g, double t, const double* lb, const double* ub) {
        if (!fuse_on || !sp_pg) return;  // only while the walks keep taking their first candidate
        double cand[3];
        pg_candidate(p, g, t, Box{lb, ub}, cand);
        __syncwarp();
        if (ctl_lane() == 0) { s_hint[0] = cand[0]; s_hint[1] = cand[1]; s_hint[2] = cand[2]; }
        __syncwarp();
        hint_valid = true;
    }

    // Before every trial evaluation the engine announces what the iteration would evaluate next if the trial is
    // rejected: the line search's probe at lambda = 0.1 and the first candidate p - t g of the projected-gradient
    // walk (g = -J^T e).  A fit that creeps repeats one pattern -- LM step rejected, first backtrack clipped to
    // lambda = 0.1, that probe already below the minimum step and not accepted, first candidate of the walk
    // taken -- and all three points of it are known here, so ONE sweep evaluates the costs of the trial point
    // and of the probe and the Jacobian (and cost) of the candidate: the iteration needs no second sweep.
    __device__ __forceinline__ void probe_hint(const double* probe) {
        BG_TICK(1);
        __syncwarp();
        if (ctl_lane() == 0) { s_ahead[0] = probe[0]; s_ahead[1] = probe[1]; s_ahead[2] = probe[2]; }
        __syncwarp();
        ahead_valid = true;
        cand_valid = false;
    }
    // the pattern held last time: the candidate is worth the square root and the division its step length costs
    __device__ __forceinline__ bool wants_candidate_hint() const { return fuse_on && sp_ls == 0 && sp_pg && sp_clip && !(spec_on && sp_trial); }
    __device__ __forceinline__ void candidate_hint(const double* p, const double* Jte, double t, const double* lb, const double* ub) {
        const double g[3] = {-Jte[0], -Jte[1], -Jte[2]};
        double cand[3];
        pg_candidate(p, g, t, Box{lb, ub}, cand);
        __syncwarp();
        if (ctl_lane() == 0) { s_ahead[3] = cand[0]; s_ahead[4] = cand[1]; s_ahead[5] = cand[2]; }
        __syncwarp();
        cand_valid = true;
    }

    // costs at the trial point p and at the announced probe, Jacobian (and cost) at the announced candidate
    __device__ __forceinline__ double cost_probe_and_candidate(const double* p, bool& bad) {
        const double probe[3] = {s_ahead[0], s_ahead[1], s_ahead[2]}, cand[3] = {s_ahead[3], s_ahead[4], s_ahead[5]};
        if (ctl_lane() == 0) {
            s_req.pts[0] = make_cost_point(p, model);
            s_req.pts[1] = make_cost_point(probe, model);
        }
        post_jac(make_pass_params(cand, model, delta, jkind), 2);
        const double esq = s_res[NACC], esq_probe = s_res[NACC + 1];
        keep_jac(cand);
        __syncwarp();
        if (ctl_lane() == 0) { s_ahead[6] = probe[0]; s_ahead[7] = probe[1]; s_ahead[8] = probe[2]; s_ahead[9] = esq_probe; }
        __syncwarp();
        probe_known = lm_finite(esq_probe);
        ++creep_fused;
        bad = false;
        if (!lm_finite(esq)) bad = count_bad(0) != 0.0;
        return esq;
    }

    // lm_engine.cuh sites: 0 = the LM trial point, k >= 1 = line-search probe number k
    __device__ __forceinline__ double cost_site(int site, const double* p, bool& bad) {
        BG_TICK(site == kSiteTrial ? 2 : site >= 1 ? 4 : 8);
        const bool spec = spec_on && (site == kSiteTrial ? sp_trial : site == kSitePgFirst ? sp_pg : site == sp_ls);
        const bool fallback = hint_valid && !spec;
        hint_valid = false;
        if (site == kSiteTrial) {
            probe_known = false;
            // the last iteration crept (search failed at its lambda = 0.1 probe, walk took its first candidate)
            if (!spec && ahead_valid && cand_valid) return cost_probe_and_candidate(p, bad);
        } else if (site == 2) {
            sp_clip = ahead_valid && same_bits(p[0], s_ahead[0]) && same_bits(p[1], s_ahead[1]) && same_bits(p[2], s_ahead[2]);
            ahead_valid = false;
            if (probe_known && same_bits(p[0], s_ahead[6]) && same_bits(p[1], s_ahead[7]) && same_bits(p[2], s_ahead[8])) {
                probe_known = false;  // evaluated together with the trial point
                bad = false;
                return s_ahead[9];
            }
        }
        if (spec) return cost_with_jac(p, bad);
        if (fallback && !memo_is(s_hint)) {
            const double h[3] = {s_hint[0], s_hint[1], s_hint[2]};
            return cost_and_jac_at(p, h, bad);
        }
        return cost(p, bad);
    }
    __device__ __forceinline__ void trial_outcome(bool accepted) { BG_TICK(3); sp_trial = accepted; }
    __device__ __forceinline__ void ls_outcome(int accepted_probe) { BG_TICK(5); sp_ls = accepted_probe; }
    __device__ __forceinline__ void pg_outcome(bool took_first) { sp_pg = took_first; }  // sequential form of the walk only

    __device__ __forceinline__ double count_bad(int k) {
        if (ctl_lane() == 0) s_req.cnt = k;
        post(kSweepBad);
        ++cost_passes;
        return s_res[0];
    }

    __device__ __forceinline__ double cost(const double* p, bool& bad) {
        const CostPoint q = make_cost_point(p, model);
        if (ctl_lane() == 0) s_req.pts[0] = q;
        post(kSweepCost);
        ++cost_passes;
        ++cost_points;
        const double esq = s_res[0];
        bad = false;
        if (!lm_finite(esq)) bad = count_bad(0) != 0.0;  // uniform across the grid: same sums everywhere
        return esq;
    }

    // projected-gradient candidates (lm_engine.cuh PgBatch): the engine writes the points into
    // shared memory, lanes 0..cnt-1 turn them into CostPoints in parallel
    __device__ __forceinline__ double* batch_points() { return s_cand; }
    __device__ __forceinline__ void cost_many(int cnt, const double* dscl, int) {
        __syncwarp();
        if (ctl_lane() < cnt) {
            double q[3];
            for (int i = 0; i < 3; ++i) q[i] = dscl ? s_cand[3 * ctl_lane() + i] * dscl[i] : s_cand[3 * ctl_lane() + i];
            s_req.pts[ctl_lane()] = make_cost_point(q, model);
        }
        if (ctl_lane() == 0) s_req.cnt = cnt;
        post(kSweepMany);
        ++cost_passes;
        cost_points += cnt;
        if (ctl_lane() < cnt) {
            s_cand_cost[ctl_lane()] = s_res[ctl_lane()];
            s_cand_bad[ctl_lane()] = 0;
        }
        __syncwarp();
        for (int k = 0; k < cnt; ++k) {
            if (!lm_finite(s_cand_cost[k])) {  // uniform: every lane reads the same value
                const double nbad = count_bad(k);
                if (ctl_lane() == 0) s_cand_bad[k] = nbad != 0.0;
                __syncwarp();
            }
        }
    }
    __device__ __forceinline__ double batch_cost(int c) const { return s_cand_cost[c]; }
    __device__ __forceinline__ bool batch_bad(int c) const { return s_cand_bad[c] != 0; }

    // levmar's projected-gradient walk (lmbc_core.c:885-934) with one candidate per LANE of the control
    // warp: lane c of a batch builds p - t*beta^c * g, projects it and makes its CostPoint; after the
    // sweep it forms its own Dp, ||Dp||^2, g.Dp and the three tests; the first lane with an event
    // (fatal / restart / found, in levmar's order of checks) decides and its state is broadcast.  Same
    // candidates, same sweeps (batches of 1, 2, 4, 8 ...), same numbers as the sequential form in
    // lm_engine.cuh -- only the ~1000 serial control instructions per batch are spread over lanes.
    // Returns 0 = nothing found, 1 = found (point in pDp), 2 = non-finite residuals (stop 7).
    __device__ __forceinline__ int pg_walk(const double* p, const double* g, double e_cur, const double* lb, const double* ub,
                                           double& t, double t0, int& gprevtaken, double* pDp, double* Dp, double& Dp_L2,
                                           double& e_new, int& nfev) {
        const double alpha = 1e-4, beta = 0.9, tming = 1e-18;
        const int lane = ctl_lane();  // control warp: 0..31
        const Box box{lb, ub};
        BG_TICK(6);
        // wide batches need the whole shard on chip (many_sweep_wide); a streamed shard keeps 8 per sweep
        const int wmax = (narrow_walk || s_ctx.stream_first < (s_ctx.v.n >> 1)) ? kGridCostBatch : kWalkMaxBatch;
        // walks tend to repeat: start with the batch width the last walk needed (1, 2, 4 or 8 candidates; the wide
        // batches are left to the walks that get that far -- started wide, a short walk evaluates candidates for nothing)
        int width = 1;
        if (width_on) while (width < pg_last && width < kGridCostBatch) width *= 2;
        bool first_batch = true;
        int consumed = 0;
        while (t > tming) {
            // lane c: t_c = t * beta^c by the same repeated multiplication the sequential walk makes
            double tc = t;
            for (int c = 0; c < lane && c < width; ++c) tc *= beta;
            const bool mine = lane < width && tc > tming;
            const int nc = __popc(__ballot_sync(0xffffffffu, mine));  // candidates are a prefix of the lanes
            double cand[3] = {0.0, 0.0, 0.0};
            if (mine) {
                pg_candidate(p, g, tc, box, cand);
                s_req.pts[lane] = make_cost_point(cand, model);
            }
            // (a) the line search announced this walk's first candidate and it was evaluated with the last
            //     probe: nothing to do; (b) the walk took its first candidate last time: evaluate this one by
            //     a Jacobian sweep and keep the sums for the next iteration; (c) a plain batch of cost points
            bool known = false, spec = false;
            if (first_batch && nc == 1) {
                known = __shfl_sync(0xffffffffu, (int)(memo_is(cand) && lm_finite(s_memo[3 + ESQ])), 0) != 0;
                spec = !known && spec_on && sp_pg;
            }
            if (known) {
                // value in s_memo[3 + ESQ]
            } else if (spec) {
                post_jac(make_pass_params(cand, model, delta, jkind), 0);  // lane 0's values are the ones that count
                double c0[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) c0[i] = __shfl_sync(0xffffffffu, cand[i], 0);
                keep_jac(c0);
                cost_points += nc;
            } else {
                if (lane == 0) s_req.cnt = nc;
                post(nc <= kGridCostBatch ? kSweepMany : nc <= 16 ? kSweepMany16 : kSweepMany32);
                ++cost_passes;
                cost_points += nc;
            }
            double e = mine ? (known ? s_memo[3 + ESQ] : s_res[spec ? (int)ESQ : lane]) : 0.0;
            bool bad = false;
            unsigned need = __ballot_sync(0xffffffffu, mine && !lm_finite(e));
            while (need) {  // rare: which of the non-finite sums come from non-finite residuals?
                const int k = __ffs(need) - 1;
                need &= need - 1;
                const double nbad = count_bad(k);
                if (lane == k) bad = nbad != 0.0;
            }
            // the tests of lmbc_core.c:905-932 for this lane's candidate
            double d[3], dl2 = 0.0, gTd = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                d[i] = cand[i] - p[i];
                dl2 += d[i] * d[i];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) gTd += g[i] * d[i];
            const bool fatal = mine && !lm_finite(e) && bad;
            const bool restart = mine && !fatal && gprevtaken && e <= e_cur + 2.0 * 0.99999 * gTd;
            const bool found = mine && !fatal && !restart && e <= e_cur + 2.0 * alpha * gTd;
            const unsigned events = __ballot_sync(0xffffffffu, fatal || restart || found);
            const int src = events ? __ffs(events) - 1 : nc - 1;  // deciding lane, else the last candidate
            nfev += src + 1;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                pDp[i] = __shfl_sync(0xffffffffu, cand[i], src);
                Dp[i] = __shfl_sync(0xffffffffu, d[i], src);
            }
            Dp_L2 = __shfl_sync(0xffffffffu, dl2, src);
            e_new = __shfl_sync(0xffffffffu, e, src);
            const double t_src = __shfl_sync(0xffffffffu, tc, src);
            if (events) {
                const int kind = __shfl_sync(0xffffffffu, fatal ? 2 : (restart ? 3 : 1), src);
                consumed += src + 1;
                if (kind == 2) { t = t_src; sp_pg = false; BG_TICK(7); return 2; }
                if (kind == 1) { t = t_src; sp_pg = first_batch && src == 0; pg_last = consumed; BG_TICK(7); return 1; }
                t = t0 * beta;  // restart: t = t0, then the loop increment still applies (:926-930)
                gprevtaken = 0;
            } else {
                consumed += nc;
                t = t_src * beta;
            }
            first_batch = false;
            // 1, 2, 4, 8, then wider only as the walk proves long: a batch never exceeds a quarter of what the walk has
            // consumed, so a walk that ends wastes at most a quarter of its candidates (a trial point costs a third of
            // what a sweep's exchange does at 10^6 samples) while the futile 393-candidate walk still takes 24 sweeps
            // instead of 52
            if (2 * width <= kGridCostBatch || (2 * width <= wmax && 8 * width <= consumed)) width *= 2;
        }
        sp_pg = false;
        pg_last = consumed;
        BG_TICK(7);
        return 0;
    }
};

#ifdef BG_PERSIST_MAXNREG
__global__ void __maxnreg__(BG_PERSIST_MAXNREG) k_persistent_fit(
#else
__global__ void __launch_bounds__(kPersistBlockThreads, 1) k_persistent_fit(
#endif
    SampleView v, int model, GlobalFitSpec spec, uint4* cells, long resident_pairs, int stage_q, int stage_off, PeerView peer,
    GlobalFitOut* out) {
    extern __shared__ double2 smem_dyn[];
    // this CTA's resident slice: pairs [first, last) of the first `resident_pairs` pairs, balanced
    BG_EXP_TABLE_LOAD();
    const long first = resident_pairs * blockIdx.x / gridDim.x, last = resident_pairs * (blockIdx.x + 1) / gridDim.x;
    const int cap = (int)((resident_pairs + gridDim.x - 1) / gridDim.x);
    const int mine = (int)(last - first);
    double2 *sc = smem_dyn, *sl = smem_dyn + cap, *sx = smem_dyn + 2 * cap;
    // the TMA ring of the streamed part sits behind the resident slice (128-byte aligned)
    const size_t ring_off = (((size_t)3 * cap * sizeof(double2)) + 127) & ~(size_t)127;
    if (resident_pairs < (v.n >> 1)) s_ring.init(reinterpret_cast<unsigned char*>(smem_dyn) + ring_off, s_ring_bars);
    {
        const double2* __restrict__ c2 = reinterpret_cast<const double2*>(v.c);
        const double2* __restrict__ l2 = reinterpret_cast<const double2*>(v.L);
        const double2* __restrict__ x2 = reinterpret_cast<const double2*>(v.x);
        for (int i = threadIdx.x; i < mine; i += kPersistBlockThreads) {
            sc[i] = __ldg(c2 + first + i);
            sl[i] = __ldg(l2 + first + i);
            sx[i] = __ldg(x2 + first + i);
        }
    }
    if (threadIdx.x == 0) {
        s_ctx.v = v; s_ctx.model = model;
        s_ctx.sc = (unsigned)__cvta_generic_to_shared(sc); s_ctx.sl = (unsigned)__cvta_generic_to_shared(sl);
        s_ctx.sx = (unsigned)__cvta_generic_to_shared(sx);
        s_ctx.res_pairs = mine; s_ctx.res_first = first; s_ctx.stream_first = resident_pairs;
        s_ctx.cells = cells; s_ctx.peer = peer; s_ctx.abort_flag = &out->aborted;
        s_ctx.stage = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(smem_dyn) + stage_off);  // persistent_plan()
        s_ctx.stage_q = stage_q;
        s_epoch = 0u; s_peer_epoch = peer.epoch; s_ring_seq = 0;
        for (int i = 0; i < 6; ++i) s_cyc[i] = 0;
        for (int i = 0; i < 7; ++i) s_ctl[i] = 0;
        s_ctl[7] = clock64();
#ifdef BG_CTL_TICKS
        for (int i = 0; i < 12; ++i) s_tick[i] = 0;
        s_tick_last = clock64();
        s_tick_busy = 0;
#endif
    }
    __syncthreads();

    // covariance on several GPUs needs the sample count of ALL ranks (lmbc_core.c:994-1002: sumsq / (n - m)); it
    // travels through the kernel's own exchange, so a context that only attached peer buffers (no NCCL
    // communicator) can still return it.  Exact: counts are integers far below 2^53.
    if (threadIdx.x < kPersistThreads) {
        if (spec.want_n_all) {
            double cnt[1] = {(blockIdx.x == 0 && threadIdx.x == 0) ? (double)v.n : 0.0};
            all_reduce<1>(cnt, clock64());
            if (blockIdx.x == 0 && threadIdx.x == 0) out->n_all = s_res[0];
        }
        serve_sweeps();
        return;
    }
    const long long t_start = clock64();
    GridEval ev;
    ev.model = model; ev.jkind = spec.jac_mode; ev.delta = spec.delta;
    ev.jac_passes = ev.cost_passes = ev.cost_points = ev.spec_issued = ev.spec_hits = 0u;
    ev.spec_on = (spec.spec_jac & 1) != 0;
    ev.fuse_on = (spec.spec_jac & 2) != 0;
    ev.width_on = (spec.spec_jac & 4) != 0;
    ev.narrow_walk = (spec.spec_jac & 64) != 0;
    ev.memo_valid = ev.sp_trial = ev.sp_pg = ev.hint_valid = ev.ahead_valid = ev.cand_valid = ev.probe_known = ev.sp_clip = false;
    ev.creep_fused = 0u;
    ev.sp_ls = ev.pg_last = 0;
    double p[3], info[10], JtJ[9];
    for (int i = 0; i < 3; ++i) p[i] = spec.p[i];
    int ret;
    if (spec.spec_jac & 8) {
        // self-check (BRDFGPU_SPEC_JAC=8, tests only): ||x - f(p)||^2 at the start point from a cost sweep, from a
        // Jacobian sweep, from a fused Jacobian-at-p + cost-at-p sweep and from a batch of two identical points --
        // the speculation relies on all of them having the same bits
        bool bad;
        for (int i = 0; i < 10; ++i) info[i] = 0.0;
        for (int i = 0; i < 9; ++i) JtJ[i] = 0.0;
        info[0] = ev.cost(p, bad);
        info[1] = ev.cost_with_jac(p, bad);
        info[2] = ev.cost_and_jac_at(p, p, bad);
        info[3] = s_memo[3 + ESQ];
        __syncwarp();
        if (ctl_lane() < 6) s_cand[ctl_lane()] = p[ctl_lane() % 3];
        __syncwarp();
        ev.cost_many(2, nullptr, 3);
        info[4] = ev.batch_cost(0);
        info[5] = ev.batch_cost(1);
        // the wide batches of the lane-parallel walk (resident shards only): the same point in the first and the last
        // slot of a 13-candidate sweep, in an odd and the last slot of a 32-candidate one.  info[6..9] are rewritten by
        // the host's accounting, so a value that differs from the cost sweep's takes the place of info[5].
        if (s_ctx.stream_first >= (s_ctx.v.n >> 1)) {
            __syncwarp();
            s_req.pts[ctl_lane()] = make_cost_point(p, model);
            __syncwarp();
            if (ctl_lane() == 0) s_req.cnt = 13;
            ev.post(kSweepMany16);
            const double w0 = s_res[0], w1 = s_res[12];
            if (ctl_lane() == 0) s_req.cnt = kWalkMaxBatch;
            ev.post(kSweepMany32);
            const double w2 = s_res[17], w3 = s_res[kWalkMaxBatch - 1];
            if (!GridEval::same_bits(w0, info[0])) info[5] = w0;
            if (!GridEval::same_bits(w1, info[0])) info[5] = w1;
            if (!GridEval::same_bits(w2, info[0])) info[5] = w2;
            if (!GridEval::same_bits(w3, info[0])) info[5] = w3;
        }
        ret = 0;
    } else if (spec.spec_jac & 16) {
        // scripted sequence (BRDFGPU_SPEC_JAC=16, bench.py's trajectory-invariant scaling figure): itmax rounds of one
        // Jacobian sweep + one cost sweep at the start point, each with its grid-wide (and cross-GPU) exchange and no LM
        // control code in between -- the same work on every rank count, whatever the data
        bool bad;
        double Jte[3];
        for (int i = 0; i < 10; ++i) info[i] = 0.0;
        // (bits 8..13 = W > 0, profiling only: rounds of ONE cost sweep over W trial points instead)
        const int width = (spec.spec_jac >> 8) & 63;
        Jte[0] = 0.0;
        for (int it = 0; it < spec.itmax; ++it) {
            if (width == 0) {
                ev.jac(p, JtJ, Jte);
                info[1] = ev.cost(p, bad);
            } else {
                if (ctl_lane() < width && ctl_lane() < kWalkMaxBatch) s_req.pts[ctl_lane()] = make_cost_point(p, model);
                if (ctl_lane() == 0) s_req.cnt = width;
                ev.post(width == 1 ? kSweepCost : width <= kGridCostBatch ? kSweepMany : width <= 16 ? kSweepMany16 : kSweepMany32);
                info[1] = s_res[0];
                ++ev.cost_passes;
                ev.cost_points += width;
            }
        }
        info[5] = (double)spec.itmax;
        info[2] = Jte[0];
        ret = spec.itmax;
    } else if (spec.unconstrained)
        ret = lm_der<3>(ev, 3, p, spec.opt, info, JtJ);
    else
        ret = lm_bc_der<3>(ev, 3, p, spec.has_lb ? spec.lb : nullptr, spec.has_ub ? spec.ub : nullptr,
                           spec.has_dscl ? spec.dscl : nullptr, spec.opt, info, JtJ);
    ev.post(kQuit);
#ifdef BG_CTL_TICKS
    if (blockIdx.x == 0 && ctl_lane() == 0)
        printf("ticks: iter-end %lld | post-jac+LU+probe %lld | cand-hint %lld | trial-test %lld | reject+LS-setup %lld | LS-tests %lld | to-PG %lld | PG-control %lld | other %lld  (iterations %d)\n",
               s_tick[0], s_tick[1], s_tick[2], s_tick[3], s_tick[4], s_tick[5], s_tick[6], s_tick[7], s_tick[8], (int)info[5]);
#endif
    if (blockIdx.x == 0 && ctl_lane() == 0) {
        out->ret = ret;
        out->peer_epoch = s_peer_epoch;
        out->jac_passes = ev.jac_passes;
        out->cost_passes = ev.cost_passes;
        out->cost_points = ev.cost_points;
        out->spec_issued = ev.spec_issued;
        out->spec_hits = ev.spec_hits;
        out->creep_fused = ev.creep_fused;
        out->cyc_sweep = s_cyc[0];
        out->cyc_exchange = s_cyc[1];
        out->cyc_total = clock64() - t_start;
        for (int i = 0; i < 4; ++i) out->cyc_x[i] = s_cyc[2 + i];
        for (int i = 0; i < 7; ++i) out->cyc_ctl[i] = s_ctl[i];
        for (int i = 0; i < 3; ++i) out->p[i] = p[i];
        for (int i = 0; i < 10; ++i) out->info[i] = info[i];
        for (int i = 0; i < 9; ++i) out->JtJ[i] = JtJ[i];
    }
}

// launch geometry of the persistent fit: grid (<= one CTA per SM), resident pairs, dynamic smem bytes
struct PersistPlan {
    int grid;
    long resident_pairs;
    size_t smem;
    int stage_q;     // quantities per round of the exchange's transposition area ...
    size_t stage_off;  // ... and where it starts in the dynamic shared memory
};

static bool persistent_plan(brdfgpu_ctx* ctx, long n, PersistPlan* plan) {
    if (ctx->persist_smem_max < 0) return false;
    if (ctx->persist_smem_max == 0) {  // first use: how much dynamic shared memory one CTA can have
        cudaFuncAttributes fa;
        int optin = 0;
        if (cudaFuncGetAttributes(&fa, k_persistent_fit) != cudaSuccess ||
            cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device) != cudaSuccess) {
            ctx->persist_smem_max = -1;
            return false;
        }
        long dyn = (long)optin - (long)fa.sharedSizeBytes - 1024;  // 1 KB per CTA is reserved by the system
        if (dyn < 0) dyn = 0;
        if (cudaFuncSetAttribute(k_persistent_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) {
            ctx->persist_smem_max = -1;
            return false;
        }
        ctx->persist_smem_max = dyn > 0 ? dyn : 1;
    }
    const long npair = n >> 1;
    long grid = (npair + kPersistThreads - 1) / kPersistThreads;  // small problems: fewer CTAs, shorter exchange
    long cap = ctx->sm_count < kMaxPersistBlocks ? ctx->sm_count : kMaxPersistBlocks;
    if (grid < 1) grid = 1;
    if (grid > cap) grid = cap;
    // dynamic shared memory = [shard slice | TMA ring if the shard does not fit | transposition area], 128-byte aligned
    const long pair_bytes = 3 * (long)sizeof(double2), room = ctx->persist_smem_max;
    int stage_q = kStageWide;
    long cap_pairs = (room - (long)stage_bytes(stage_q) - 128) / pair_bytes;
    if (npair > grid * cap_pairs) {  // give the shard the room of the wide transposition area first
        stage_q = kStageNarrow;
        cap_pairs = (room - (long)stage_bytes(stage_q) - 128) / pair_bytes;
    }
    long resident = npair;
    size_t ring_bytes = 0;
    if (npair > grid * cap_pairs) {  // not everything fits on chip: part of the shared memory becomes the TMA ring
        ring_bytes = PersistRing::kBytes + 128;
        cap_pairs = (room - (long)ring_bytes - (long)stage_bytes(stage_q) - 128) / pair_bytes;
        if (cap_pairs < 0) return false;
        resident = grid * cap_pairs;
    }
    // a slice is ceil(resident / grid) pairs at most
    while (resident > 0 && (resident + grid - 1) / grid > cap_pairs) --resident;
    plan->grid = (int)grid;
    plan->resident_pairs = resident;
    const size_t slice = (size_t)((resident + grid - 1) / grid) * (size_t)pair_bytes;
    plan->stage_q = stage_q;
    plan->stage_off = ((slice + 127) & ~(size_t)127) + ring_bytes;  // (the kernel puts the ring at the same aligned offset)
    plan->smem = plan->stage_off + stage_bytes(stage_q);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_persistent_fit, kPersistBlockThreads, plan->smem) != cudaSuccess ||
        per_sm < 1)
        return false;
    return true;
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
    long n_all_kernel = 0;
    const bool can_persist = ctx->coop && (ctx->nranks == 1 || ctx->peer_attached);
    PersistPlan plan{0, 0, 0};
    // Far beyond on-chip residency (> 4e7 samples per GPU, ~1 GB) the kernel-per-evaluation driver is
    // the faster one: its streaming kernels run at 6-7 TB/s, the launch + ticket round trip (~10 us) no
    // longer matters next to a 350 us pass (measured crossover between 1e7 and 1e8, profiles/r01_summary.md).
    const bool host_is_better = s->n > 40000000 && (ctx->nranks == 1 || ctx->nccl_comm != nullptr);
    const bool persist = drive == BRDFGPU_DRIVE_PERSISTENT && can_persist && !host_is_better && persistent_plan(ctx, s->n, &plan);
    if (persist) {
        GlobalFitSpec spec;
        memset(&spec, 0, sizeof(spec));
        spec.m = m; spec.itmax = itmax; spec.jac_mode = jkind; spec.delta = delta; spec.opt = o;
        spec.has_lb = lb != nullptr; spec.has_ub = ub != nullptr; spec.has_dscl = dscl != nullptr;
        spec.unconstrained = unconstrained;
        spec.want_n_all = (covar != nullptr && ctx->nranks > 1) ? 1 : 0;
        // BRDFGPU_SPEC_JAC=0 switches the speculative Jacobians off (A/B tests: results must not change)
        // (a bit mask for experiments: 1 = speculate at the trial / line-search / first-candidate sites, 2 = fuse the
        // announced first candidate into the last line-search probe, 4 = start a walk at the last walk's width,
        // 8 = self-check of the sweep kinds, 16 = scripted Jacobian + cost sweeps instead of a fit, 64 = walks of at most
        // 8 candidates per sweep even when the shard is resident)
        const char* sj = getenv("BRDFGPU_SPEC_JAC");
        spec.spec_jac = sj ? atoi(sj) : 7;
        for (int i = 0; i < 3; ++i) {
            spec.p[i] = p[i];
            if (lb) spec.lb[i] = lbs[i];
            if (ub) spec.ub[i] = ubs[i];
            if (dscl) spec.dscl[i] = dscl[i];
        }
        SampleView v = view_of(s);
        int model = s->model;
        GlobalFitOut* d_out = static_cast<GlobalFitOut*>(ctx->d_fitio);
        uint4* cells = ctx->d_cells;
        long resident_pairs = plan.resident_pairs;
        // tags restart at 1 every launch: clear the cells this grid will use and the abort flag
        BG_CUDA_OK(ctx, cudaMemsetAsync(cells, 0, sizeof(uint4) * 2 * (size_t)kMaxPersistBlocks * NSUM, ctx->stream));
        BG_CUDA_OK(ctx, cudaMemsetAsync(d_out, 0, sizeof(GlobalFitOut), ctx->stream));
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
        int stage_q = plan.stage_q, stage_off = (int)plan.stage_off;
        void* args[] = {&v, &model, &spec, &cells, &resident_pairs, &stage_q, &stage_off, &peer, &d_out};
        BG_CUDA_OK(ctx, cudaLaunchCooperativeKernel((const void*)k_persistent_fit, dim3(plan.grid), dim3(kPersistBlockThreads), args,
                                                     plan.smem, ctx->stream));
        ++ctx->launches;
        BG_CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_fitio, d_out, sizeof(GlobalFitOut), cudaMemcpyDeviceToHost, ctx->stream));
        BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        const GlobalFitOut* h = static_cast<const GlobalFitOut*>(ctx->h_fitio);
        if (ctx->nranks > 1) ctx->peer_epoch = h->peer_epoch;
        if (h->aborted) {
            ctx->peer_attached = false;  // exchange tags are out of step now: every rank exports again and re-attaches
            set_error(ctx, "fit abandoned: an exchange partner (CTA or peer GPU) never delivered its sums");
            return BRDFGPU_LM_ERROR;
        }
        ctx->fit_stats[0] = h->jac_passes; ctx->fit_stats[1] = h->cost_passes; ctx->fit_stats[2] = h->cost_points;
        ctx->fit_stats[3] = (unsigned long long)plan.resident_pairs * 2; ctx->fit_stats[4] = (unsigned long long)plan.grid;
        ctx->fit_stats[5] = (unsigned long long)h->cyc_sweep; ctx->fit_stats[6] = (unsigned long long)h->cyc_exchange;
        ctx->fit_stats[7] = (unsigned long long)h->cyc_total;
        for (int i = 0; i < 4; ++i) ctx->fit_stats[8 + i] = (unsigned long long)h->cyc_x[i];
        for (int i = 0; i < 7; ++i) ctx->fit_stats[12 + i] = (unsigned long long)h->cyc_ctl[i];
        ctx->fit_stats[19] = h->spec_issued; ctx->fit_stats[20] = h->spec_hits; ctx->fit_stats[21] = h->creep_fused;
        ret = h->ret;
        n_all_kernel = (long)h->n_all;
        for (int i = 0; i < 3; ++i) p[i] = h->p[i];
        for (int i = 0; i < 10; ++i) fit_info[i] = h->info[i];
        for (int i = 0; i < 9; ++i) JtJ[i] = h->JtJ[i];
    } else {
        HostEval ev{ctx, s, delta, jkind, false};
        {
            const char* sj = getenv("BRDFGPU_SPEC_JAC");
            ev.spec_on = sj ? (atoi(sj) & 1) != 0 : true;
        }
        if (unconstrained) ret = lm_der<3>(ev, 3, p, o, fit_info, JtJ);
        else ret = lm_bc_der<3>(ev, 3, p, lb ? lbs : nullptr, ub ? ubs : nullptr, dscl, o, fit_info, JtJ);
        if (ev.failed) return BRDFGPU_LM_ERROR;
        ctx->fit_stats[0] = ev.jac_launched; ctx->fit_stats[1] = ev.cost_launched;  // passes over the samples
        ctx->fit_stats[2] = ev.cost_launched;
        for (int i = 3; i < 19; ++i) ctx->fit_stats[i] = 0;
        ctx->fit_stats[19] = ev.spec_issued; ctx->fit_stats[20] = ev.spec_hits;
    }
    finish_info(info, fit_info, m, jkind, jac_mode == BRDFGPU_JAC_FD);
    if (covar) {  // lmbc_core.c:994-1002
        // ||e||^2 over ALL samples of all ranks went into fit_info[1]; n is the global count
        long n_all = s->n;
        if (ctx->nranks > 1 && persist) {
            n_all = n_all_kernel;  // summed by the fit kernel's own exchange
        } else if (ctx->nranks > 1) {
            double cnt = (double)s->n;  // host-driven fit: the communicator that carried its sums
            double* tmp = ctx->d_result;
            BG_CUDA_OK(ctx, cudaMemcpyAsync(tmp, &cnt, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            if (comm_allreduce_device(ctx, tmp, 1) != 0) return BRDFGPU_LM_ERROR;
            BG_CUDA_OK(ctx, cudaMemcpyAsync(&cnt, tmp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
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
