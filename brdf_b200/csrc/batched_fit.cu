// batched_fit.cu -- batched mode: thousands of independent small BRDF fits in one launch.
//
// Replaces the per-pixel loop of CBRDFdata::CalcBRDFEquation (brdfdata.cpp:1195-1221), which calls
// SolveEquation -> dlevmar_bc_dif (brdfdata.cpp:1119) once per mapped face and colour channel.
// One lane group (a full warp, or a half warp for <= 16 samples) owns one fit: its samples live in
// registers for the whole solve, every evaluation is a few model evaluations per lane plus a
// butterfly reduction over the group, and every lane of the group runs the same levmar control
// code (lm_engine.cuh) on the same sums, so there is no divergence inside a group and no
// communication between groups.  Fits diverge freely from each other (different iteration counts,
// line searches, projected-gradient probes).
#include <cstdlib>

#include "brdf_model.cuh"
#include "common.cuh"

namespace brdfgpu {

constexpr int kBatchThreads = 128;
// resident CTAs per SM the register allocation is held to: 8 (64 registers) for half-warp fits,
// 6 (80 registers) for warp fits -- measured best on B200 (profiles/r01_summary.md); the kernel is
// latency-bound (divergent control flow between fits), so occupancy beats a spill-free allocation
#ifndef BG_BATCH_MIN_BLOCKS
#define BG_BATCH_MIN_BLOCKS(G) ((G) == 32 ? 6 : 8)
#endif

struct BatchSpec {
    int itmax, jkind, has_lb, has_ub, dif_accounting;
    double delta;
    double p0[3], lb[3], ub[3];
    LmOptions opt;
};

// G lanes per fit, S samples per lane held in registers (S == 0: samples re-read from memory)
template <int G, int S>
struct GroupEval {
    // Candidates of the projected-gradient walk evaluated together (lm_engine.cuh PgBatch).  Measured
    // on B200 for 65 536 x 64: 1 -> 16.8 ms, 2 -> 18.7 ms, 4 -> 19.8 ms, 8 -> 26.2 ms: with no exchange
    // to amortise, the discarded speculative evaluations cost more than the extra ILP gains, so the
    // per-fit kernel walks one candidate at a time (the persistent global fit uses 8).
#ifndef BG_BATCH_PG
#define BG_BATCH_PG 1
#endif
    static constexpr int kCostBatch = (S > 0) ? BG_BATCH_PG : 1;
#ifndef BG_BATCH_LANE_WALK
#define BG_BATCH_LANE_WALK 0
#endif
    // Opt-in projected-gradient walk with one candidate per LANE of the group and a batch width that
    // doubles 1, 2, 4, ... up to min(BG_BATCH_LANE_WALK, G) (0 = the sequential walk of lm_engine.cuh, the
    // default); needs the samples in registers.  Motivation: 94 % of the cost evaluations of a batch are
    // walk candidates (9.6 walks per fit, median length 3, but 2/3 of the candidates sit in walks longer
    // than 100 -- levmar tries t, 0.9 t, ... down to 1e-18, 393 evaluations, when nothing is to be found).
    // A round evaluates its candidates BG_BATCH_WALK_CHUNK at a time with the control code of the round spread
    // over the lanes.  Bit-identical to the sequential walk (checked on the GPU for 64 / 40 / 16 samples per
    // fit).  Measured on B200, 65 536 x 64 (profiles/r01_summary.md): warp instructions 9.7e9 -> 6.2e9, but
    // the issue rate falls from 0.56 to 0.35 per cycle (the longer register-hungry rounds cost occupancy or
    // spill: chunk 4 @ 80 regs 22.4 ms, chunk 4 @ 128 regs 21.8 ms, chunk 2 @ 80 regs 16.7 ms, chunk 2 @ 128
    // regs 16.2 ms) against 16.5 ms sequential -- no gain worth a second code path, hence off by default.
    // (A fixed width of 2 / 4 / 8 from the first round on: 17.9 / 19.0 / 23.9 ms.)
    static constexpr int kWalk = (S > 0) ? ((BG_BATCH_LANE_WALK) < G ? (BG_BATCH_LANE_WALK) : G) : 0;
    static constexpr bool kLanePgWalk = kWalk > 0;
#ifndef BG_BATCH_WALK_CHUNK
#define BG_BATCH_WALK_CHUNK 2
#endif
    static constexpr int kWalkChunk = BG_BATCH_WALK_CHUNK;  // candidates evaluated together (2 x this many exp chains per lane)
    static constexpr int KB = kCostBatch;
    static constexpr int SR = S > 0 ? S : 1;
    double* s_pts;   // shared memory of this lane group: KB x 3 candidate points
    double* s_out;   // KB costs + KB "some residual non-finite" flags
    double c[SR], L[SR], x[SR];
    const double *gc, *gL, *gx, *traw;  // this fit's rows
    int nper, lane, model, jkind;
    unsigned mask;
    double delta;

    __device__ __forceinline__ double group_sum(double v) const {
#pragma unroll
        for (int off = G / 2; off; off >>= 1) v += __shfl_xor_sync(mask, v, off, G);
        return v;
    }

    template <int JAC>
    __device__ __forceinline__ void jac_body(const PassParams& q, double* acc) const {
        if (S > 0) {
#pragma unroll
            for (int s = 0; s < SR; ++s) {
                const int idx = s * G + lane;
                if (idx < nper) accumulate_jac<JAC>(q, c[s], L[s], x[s], traw, idx, acc);
            }
        } else {
            for (int idx = lane; idx < nper; idx += G) accumulate_jac<JAC>(q, gc[idx], gL[idx], gx[idx], traw, idx, acc);
        }
    }

    __device__ __forceinline__ void jac(const double* p, double* JtJ, double* Jte) const {
        const PassParams q = make_pass_params(p, model, delta, jkind);
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        if (jkind == kJacForward) jac_body<kJacForward>(q, acc);
        else if (jkind == kJacCentral) jac_body<kJacCentral>(q, acc);
        else jac_body<kJacAnalytic>(q, acc);
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = group_sum(acc[k]);
        JtJ[0] = acc[A00]; JtJ[1] = acc[A01]; JtJ[2] = acc[A02];
        JtJ[3] = acc[A01]; JtJ[4] = acc[A11]; JtJ[5] = acc[A12];
        JtJ[6] = acc[A02]; JtJ[7] = acc[A12]; JtJ[8] = acc[A22];
        Jte[0] = acc[G0]; Jte[1] = acc[G1]; Jte[2] = acc[G2];
    }

    __device__ __forceinline__ double cost(const double* p, bool& bad) const {
        const CostPoint q = make_cost_point(p, model);
        double esq = 0.0, nbad = 0.0;
        if (S > 0) {
#pragma unroll
            for (int s = 0; s < SR; ++s) {
                const int idx = s * G + lane;
                if (idx < nper) {
                    const double e = residual_of(q, c[s], L[s], x[s], traw, idx);
                    esq = __fma_rn(e, e, esq);
                    nbad += lm_finite(e) ? 0.0 : 1.0;
                }
            }
        } else {
            for (int idx = lane; idx < nper; idx += G) {
                const double e = residual_of(q, gc[idx], gL[idx], gx[idx], traw, idx);
                esq = __fma_rn(e, e, esq);
                nbad += lm_finite(e) ? 0.0 : 1.0;
            }
        }
        esq = group_sum(esq);
        bad = false;
        if (!lm_finite(esq)) bad = group_sum(nbad) != 0.0;  // uniform within the group
        return esq;
    }

    // residuals-squared sums of CH candidates (lanes c0 .. c0+CH-1 of the group hold them in my_q) over this
    // lane's samples, reduced over the group; per candidate exactly the arithmetic of cost()
    template <int CH>
    __device__ __forceinline__ void walk_chunk(const CostPoint& my_q, int c0, int nc, double& e, bool& bad) const {
        CostPoint q[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            q[k].kd = __shfl_sync(mask, my_q.kd, c0 + k, G);
            q[k].cks = __shfl_sync(mask, my_q.cks, c0 + k, G);
            q[k].n = __shfl_sync(mask, my_q.n, c0 + k, G);
        }
        double esq[CH], nbad[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) esq[k] = nbad[k] = 0.0;
#pragma unroll
        for (int s = 0; s < SR; ++s) {
            const int idx = s * G + lane;
            if (idx < nper) {
                double y[CH], pw[CH];
                bool slow = false;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    y[k] = q[k].n * L[s];
                    slow |= (c0 + k < nc) && needs_care(y[k]);
                }
                exp_core_n<CH>(y, pw);
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    double r = x[s] - __fma_rn(q[k].kd, c[s], q[k].cks * pw[k]);
                    if (slow && c0 + k < nc && needs_care(y[k])) r = residual_careful(q[k], c[s], traw[idx], x[s]);
                    esq[k] = __fma_rn(r, r, esq[k]);
                    nbad[k] += lm_finite(r) ? 0.0 : 1.0;
                }
            }
        }
#pragma unroll
        for (int off = G / 2; off; off >>= 1) {
#pragma unroll
            for (int k = 0; k < CH; ++k) esq[k] += __shfl_xor_sync(mask, esq[k], off, G);
        }
        bool any_nonfinite = false;
#pragma unroll
        for (int k = 0; k < CH; ++k) any_nonfinite |= (c0 + k < nc) && !lm_finite(esq[k]);
        if (any_nonfinite) {  // uniform within the group
#pragma unroll
            for (int k = 0; k < CH; ++k) nbad[k] = group_sum(nbad[k]);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k)
            if (lane == c0 + k) { e = esq[k]; bad = any_nonfinite && nbad[k] != 0.0; }
    }

    // levmar's projected-gradient walk (lmbc_core.c:885-934) with one candidate per lane of the group:
    // lane c of a round builds and projects candidate c and later applies levmar's tests to it; the first
    // lane with an event, in levmar's order of checks, decides.  Same candidates and same arithmetic per
    // candidate as the sequential walk; candidates past the deciding one are discarded and not counted.
    // Returns 0 = nothing found, 1 = found (pDp), 2 = non-finite residuals.
    __device__ __forceinline__ int pg_walk(const double* p, const double* g, double e_cur, const double* lb, const double* ub,
                                           double& t, double t0, int& gprevtaken, double* pDp, double* Dp, double& Dp_L2,
                                           double& e_new, int& nfev) const {
        constexpr int W = kWalk > 0 ? kWalk : 1;
        const double alpha = 1e-4, beta = 0.9, tming = 1e-18;
        const Box box{lb, ub};
        const int shift = ((threadIdx.x & 31) / G) * G;  // first lane of this group inside the warp
        int width = 1;
        while (t > tming) {
            double tc = t;
            for (int c = 0; c < lane && c < width; ++c) tc *= beta;
            const bool mine = lane < width && tc > tming;
            const int nc = __popc(__ballot_sync(mask, mine));
            double cand[3] = {0.0, 0.0, 0.0};
            CostPoint my_q = {0.0, 0.0, 0.0};
            if (mine) {
#pragma unroll
                for (int i = 0; i < 3; ++i) cand[i] = p[i] - tc * g[i];
                box_project<3>(cand, box, 3);
                my_q = make_cost_point(cand, model);
            }
            double e = 0.0;
            bool bad = false;
            if (nc == 1) {
                walk_chunk<1>(my_q, 0, 1, e, bad);
            } else if (nc == 2) {
                walk_chunk<2>(my_q, 0, 2, e, bad);
            } else {
                for (int c0 = 0; c0 < nc; c0 += kWalkChunk) walk_chunk<kWalkChunk>(my_q, c0, nc, e, bad);
            }
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
            const unsigned events = (__ballot_sync(mask, fatal || restart || found) & mask) >> shift;
            const int src = events ? __ffs(events) - 1 : nc - 1;
            nfev += src + 1;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                pDp[i] = __shfl_sync(mask, cand[i], src, G);
                Dp[i] = __shfl_sync(mask, d[i], src, G);
            }
            Dp_L2 = __shfl_sync(mask, dl2, src, G);
            e_new = __shfl_sync(mask, e, src, G);
            const double t_src = __shfl_sync(mask, tc, src, G);
            if (events) {
                const int kind = __shfl_sync(mask, fatal ? 2 : (restart ? 3 : 1), src, G);
                if (kind == 2) { t = t_src; return 2; }
                if (kind == 1) { t = t_src; return 1; }
                t = t0 * beta;
                gprevtaken = 0;
            } else {
                t = t_src * beta;
            }
            width = (2 * width < W) ? 2 * width : W;
        }
        return 0;
    }

    __device__ __forceinline__ double* batch_points() const { return s_pts; }
    __device__ __forceinline__ double batch_cost(int k) const { return s_out[k]; }
    __device__ __forceinline__ bool batch_bad(int k) const { return s_out[KB + k] != 0.0; }

    // up to KB trial points at once; the engine wrote them (every lane the same values) into s_pts
    __device__ __forceinline__ void cost_many(int cnt, const double* /*dscl: batched fits are unscaled*/, int) const {
        __syncwarp(mask);
        CostPoint q[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const int kk = k < cnt ? k : 0;
            const double pt[3] = {s_pts[3 * kk], s_pts[3 * kk + 1], s_pts[3 * kk + 2]};
            q[k] = make_cost_point(pt, model);
        }
        double esq[KB], nbad[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) esq[k] = nbad[k] = 0.0;
#pragma unroll
        for (int s = 0; s < SR; ++s) {
            const int idx = s * G + lane;
            if (idx < nper) {
                double y[KB], pw[KB];
                bool slow = false;
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    y[k] = q[k].n * L[s];
                    slow |= needs_care(y[k]);
                }
                exp_core_n<KB>(y, pw);
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    double e = x[s] - __fma_rn(q[k].kd, c[s], q[k].cks * pw[k]);
                    if (slow && needs_care(y[k])) e = residual_careful(q[k], c[s], traw[idx], x[s]);
                    esq[k] = __fma_rn(e, e, esq[k]);
                    nbad[k] += lm_finite(e) ? 0.0 : 1.0;
                }
            }
        }
#pragma unroll
        for (int off = G / 2; off; off >>= 1) {
#pragma unroll
            for (int k = 0; k < KB; ++k) esq[k] += __shfl_xor_sync(mask, esq[k], off, G);
        }
        bool any_bad = false;
#pragma unroll
        for (int k = 0; k < KB; ++k) any_bad |= !lm_finite(esq[k]);
        if (any_bad) {  // uniform within the group
#pragma unroll
            for (int k = 0; k < KB; ++k) nbad[k] = group_sum(nbad[k]);
        }
        __syncwarp(mask);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                s_out[k] = esq[k];
                s_out[KB + k] = (!lm_finite(esq[k]) && nbad[k] != 0.0) ? 1.0 : 0.0;
            }
        }
        __syncwarp(mask);
    }
};

template <int G, int S>
__global__ void __launch_bounds__(kBatchThreads, BG_BATCH_MIN_BLOCKS(G)) k_batched_fit(const double* __restrict__ c, const double* __restrict__ L,
                                                                const double* __restrict__ x,
                                                                const double* __restrict__ traw, long nfit, int nper,
                                                                int model, BatchSpec spec, double* __restrict__ p_out,
                                                                double* __restrict__ info_out, int* __restrict__ ret_out) {
    const long fit = ((long)blockIdx.x * kBatchThreads + threadIdx.x) / G;
    if (fit >= nfit) return;
    const int lane = threadIdx.x % G;
    const long base = fit * nper;

    constexpr int kGroups = kBatchThreads / G;
    __shared__ double s_scratch[kGroups * (3 + 2) * (GroupEval<G, S>::kCostBatch)];
    GroupEval<G, S> ev;
    ev.s_pts = s_scratch + (threadIdx.x / G) * 5 * GroupEval<G, S>::kCostBatch;
    ev.s_out = ev.s_pts + 3 * GroupEval<G, S>::kCostBatch;
    ev.gc = c + base; ev.gL = L + base; ev.gx = x + base; ev.traw = traw + base;
    ev.nper = nper; ev.lane = lane; ev.model = model; ev.jkind = spec.jkind; ev.delta = spec.delta;
    ev.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    if (S > 0) {
#pragma unroll
        for (int s = 0; s < (S > 0 ? S : 1); ++s) {
            const int idx = s * G + lane;
            const bool in = idx < nper;
            ev.c[s] = in ? c[base + idx] : 0.0;
            ev.L[s] = in ? L[base + idx] : 0.0;
            ev.x[s] = in ? x[base + idx] : 0.0;
        }
    }

#ifdef BG_FIT_CYCLES
    const long long t_fit0 = clock64();
#endif
    double p[3] = {spec.p0[0], spec.p0[1], spec.p0[2]};
    double info[10];
    const double* lb = spec.has_lb ? spec.lb : nullptr;
    const double* ub = spec.has_ub ? spec.ub : nullptr;
    const Box box{lb, ub};
    box_project(p, box, 3);  // lmbc_core.c:516
    const int ret = lm_bc_der<3>(ev, 3, p, lb, ub, nullptr, spec.opt, info, nullptr);
    if (lane == 0) {
        if (spec.dif_accounting) info[7] += info[8] * (spec.jkind == kJacCentral ? 6.0 : 4.0);  // lmbc_core.c:1119-1124
        for (int i = 0; i < 3; ++i) p_out[fit * 3 + i] = p[i];
        if (info_out)
            for (int i = 0; i < 10; ++i) info_out[fit * 10 + i] = info[i];
        if (ret_out) ret_out[fit] = ret;
#ifdef BG_FIT_CYCLES  // debug variant (profiles/fit_cycles.py): SM cycles this fit took, in units of 64, instead of the return value
        if (ret_out) ret_out[fit] = (int)((clock64() - t_fit0) >> 6);
#endif
    }
}

__global__ void k_prepare_batch(const double* __restrict__ traw, double* __restrict__ L, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        L[i] = log_or_flag(traw[i]);
}

int batch_alloc(brdfgpu_ctx* ctx, long nfit, int nper, int model, brdfgpu_batch** out) {
    if (nfit < 0 || nper < 3 || (model != 0 && model != 1)) {
        set_error(ctx, "batch: bad sizes or model (need nper >= 3 samples per fit)");
        return BRDFGPU_LM_ERROR;
    }
    brdfgpu_batch* b = new brdfgpu_batch;
    b->nfit = nfit; b->capacity = nfit; b->nper = nper; b->model = model;
    b->stream = ctx->stream;
    // one stream-ordered allocation, every array on a 256-byte boundary
    const size_t nf = (size_t)(nfit > 0 ? nfit : 1);
    const size_t na = (sizeof(double) * nf * nper + 255) & ~(size_t)255, np = (sizeof(double) * 3 * nf + 255) & ~(size_t)255,
                 ni = (sizeof(double) * 10 * nf + 255) & ~(size_t)255, nr = (sizeof(int) * nf + 255) & ~(size_t)255;
    cudaError_t e = cudaMallocAsync(&b->block, 4 * na + np + ni + nr, ctx->stream);
    if (e != cudaSuccess) {
        delete b;
        set_error(ctx, std::string("batch: cudaMallocAsync: ") + cudaGetErrorString(e));
        return BRDFGPU_LM_ERROR;
    }
    char* q = static_cast<char*>(b->block);
    b->c = reinterpret_cast<double*>(q); b->L = reinterpret_cast<double*>(q + na); b->x = reinterpret_cast<double*>(q + 2 * na);
    b->traw = reinterpret_cast<double*>(q + 3 * na); b->p = reinterpret_cast<double*>(q + 4 * na);
    b->info = reinterpret_cast<double*>(q + 4 * na + np); b->ret = reinterpret_cast<int*>(q + 4 * na + np + ni);
    *out = b;
    return 0;
}

int batch_prepare(brdfgpu_ctx* ctx, brdfgpu_batch* b) {
    const long n = b->nfit * b->nper;
    if (n == 0) return 0;
    long blocks = (n + 255) / 256;
    if (blocks > (long)ctx->sm_count * 16) blocks = (long)ctx->sm_count * 16;
    k_prepare_batch<<<(int)blocks, 256, 0, ctx->stream>>>(b->traw, b->L, n);
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

template <int G, int S>
static void launch_batch(brdfgpu_ctx* ctx, brdfgpu_batch* b, const BatchSpec& spec) {
    const long threads = b->nfit * G;
    const long blocks = (threads + kBatchThreads - 1) / kBatchThreads;
    k_batched_fit<G, S><<<(unsigned)blocks, kBatchThreads, 0, ctx->stream>>>(b->c, b->L, b->x, b->traw, b->nfit, b->nper,
                                                                            b->model, spec, b->p, b->info, b->ret);
}

int batch_fit(brdfgpu_ctx* ctx, brdfgpu_batch* b, const double* p0, const double* lb, const double* ub, int itmax,
              const double* opts, int jac_mode) {
    if (b->nfit == 0) return 0;
    if (lb && ub)
        for (int i = 0; i < 3; ++i)
            if (lb[i] > ub[i]) {
                fprintf(stderr, "brdfgpu batch fit: at least one lower bound exceeds the upper one\n");
                return BRDFGPU_LM_ERROR;
            }
    BatchSpec spec;
    const double delta_signed = opts ? opts[4] : kDiffDelta;
    spec.itmax = itmax;
    spec.jkind = jac_mode == BRDFGPU_JAC_ANALYTIC ? kJacAnalytic : (delta_signed < 0.0 ? kJacCentral : kJacForward);
    spec.dif_accounting = jac_mode == BRDFGPU_JAC_FD;
    spec.delta = lm_abs(delta_signed);
    spec.has_lb = lb != nullptr;
    spec.has_ub = ub != nullptr;
    for (int i = 0; i < 3; ++i) {
        spec.p0[i] = p0[i];
        spec.lb[i] = lb ? lb[i] : 0.0;
        spec.ub[i] = ub ? ub[i] : 0.0;
    }
    spec.opt = lm_options(opts, itmax);

    // Lanes per fit G and samples per lane S (S * G >= nper).  Several fits share a warp when G < 32:
    // their control flow runs in the same instruction stream wherever it coincides (loops reconverge
    // at their exits), which matters because the control code, not the model, dominates the
    // instruction count of a small fit.  BRDFGPU_BATCH_G overrides the choice (experiments).
    const int n = b->nper;
    int G = n <= 16 ? 16 : 32;
    if (const char* e = getenv("BRDFGPU_BATCH_G")) G = atoi(e);
    const int S = (n + G - 1) / G;
    bool ok = true;
    if (G == 8) {
        if (S <= 1) launch_batch<8, 1>(ctx, b, spec);
        else if (S <= 2) launch_batch<8, 2>(ctx, b, spec);
        else if (S <= 4) launch_batch<8, 4>(ctx, b, spec);
        else if (S <= 8) launch_batch<8, 8>(ctx, b, spec);
        else ok = false;
    } else if (G == 16) {
        if (S <= 1) launch_batch<16, 1>(ctx, b, spec);
        else if (S <= 2) launch_batch<16, 2>(ctx, b, spec);
        else if (S <= 4) launch_batch<16, 4>(ctx, b, spec);
        else ok = false;
    } else {
        ok = false;
    }
    if (!ok) {
        if (n <= 32) launch_batch<32, 1>(ctx, b, spec);
        else if (n <= 64) launch_batch<32, 2>(ctx, b, spec);
        else if (n <= 128) launch_batch<32, 4>(ctx, b, spec);
        else launch_batch<32, 0>(ctx, b, spec);
    }
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

}  // namespace brdfgpu
