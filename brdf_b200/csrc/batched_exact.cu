// batched_exact.cu -- batched mode, levmar-exact: every fit reproduces dlevmar_bc_dif + BRDFFunc bit for bit.
//
// The default batched kernel (batched_fit.cu) evaluates t**n as exp(n ln t), forms the difference
// quotients algebraically and sums with butterflies; its converged fits agree with the reference to the
// parity bars (1e-4 / 1e-6), but levmar's trajectory is sensitive to the last bit of every sum
// (SURVEY.md Q13), so fits that levmar itself abandons (itmax, "no further reduction": 86 % of the
// per-face fits on img/cup) end somewhere else.  This kernel removes every source of difference for
// the problem sizes of the batched mode (n m < 1024: levmar's small-problem branch, lmbc_core.c:573):
//
//   * the model value is BRDFFunc's expression with its roundings (brdfdata.cpp:981, 986: products and sum
//     rounded separately, the host has no FMA contraction) around glibc's pow() reproduced bit for bit
//     (glibc_pow.cuh);
//   * the Jacobian is levmar's literal forward difference jac[i][j] = (f(p + d_j e_j)[i] - f(p)[i]) * (1/d_j)
//     with d_j = max(|1e-4 p_j|, delta) (misc_core.c:153-170) -- t**n is shared by the kd / ks columns
//     because the reference's own calls pass identical arguments there;
//   * J^T J and J^T e are accumulated in the order of levmar's small-problem loop (lmbc_core.c:592-616:
//     samples downwards, lower triangle, product and sum rounded separately), ||e||^2 in the order of
//     dlevmar_L2nrmxmy (misc_core.c:721-807: four interleaved partial sums walking downwards in blocks of
//     eight, the remainder in its switch order, then ((s0 + s1) + s2) + s3);
//   * the control loop is lm_engine.cuh compiled WITHOUT multiply-add contraction (this file is built with
//     --fmad=false and BG_LM_LEVMAR_ARITH), i.e. the instantiation that is bit-identical to levmar on the host.
//
// One lane group (16 lanes for <= 16 samples, else a warp) owns one fit: lanes evaluate their samples in
// parallel, the per-sample terms go through shared memory, and a few lanes add them up in levmar's order.
// Result: p, all ten info[] entries and the return value equal the reference's (tests/test_gpu_exact.py:
// every fit of BASELINE configs[3] and of the per-face path on img/cup).
#define BG_LM_LEVMAR_ARITH 1
#include "glibc_pow.cuh"

#include "brdf_model.cuh"
#include "common.cuh"

namespace brdfgpu {

constexpr int kExactThreads = 128;

struct ExactSpec {
    int itmax, has_lb, has_ub, model, nper;
    double delta;
    double p0[3], lb[3], ub[3];
    LmOptions opt;
};

static __device__ __noinline__ double pow_libm(double x, double y) { return glibc_pow(x, y); }

// model value with BRDFFunc's roundings: kd*c + ks'*pw, ks' = ks (Blinn-Phong) or ((n+2)/2*pi)*ks (Phong)
__device__ __forceinline__ double model_ks(int model, double ks, double n) {
    return model == 1 ? ks : __dmul_rn(__dmul_rn(__ddiv_rn(__dadd_rn(n, 2.0), 2.0), kPi), ks);
}
__device__ __forceinline__ double model_hx(double kd, double c, double ksp, double pw) {
    return __dadd_rn(__dmul_rn(kd, c), __dmul_rn(ksp, pw));
}

#ifndef BG_EXACT_PG
#define BG_EXACT_PG 4
#endif

template <int G, int S>
struct ExactEval {
    // Candidates of the projected-gradient walk evaluated together (lm_engine.cuh PgBatch): levmar walks t, 0.9 t,
    // 0.81 t ... one function call at a time and most evaluations of a batch of fits are such candidates (the median
    // fit makes one 393-step walk that finds nothing).  The candidates depend only on p and J^T e, every evaluation is
    // deterministic, and they are consumed strictly in levmar's order, so evaluating KB of them per call changes no
    // value and no count; it gives each lane KB independent exp chains and puts four lanes per candidate to work in
    // the ordered summation instead of four in all.
    static constexpr int KB = (BG_EXACT_PG) < G / 4 ? (BG_EXACT_PG) : G / 4;
    static constexpr int kCostBatch = KB;
    static constexpr bool kLanePgWalk = false;
    // The engine's J^T J, J^T e, Dp, diag, pDp are identical in every lane: half-warp fits keep one copy per group in
    // shared memory and run at 64 registers (8 CTAs per SM); measured on B200 (profiles/r02_batched.md): 113 007 x 16
    // 30.7 -> 27.2 ms, 10^6 x 16 258 -> 219 ms.  For warp fits (64 samples) the same trade is neutral in throughput
    // (24.4 -> 23.9 ms) and lengthens the tail of small batches (8 192 fits: 4.15 -> 4.56 ms), so they keep registers.
    static constexpr bool kSharedState = G == 16;
    double* s_ws;
    __device__ __forceinline__ double* workspace() const { return s_ws; }
    static_assert(4 * KB <= G, "four summing lanes per candidate");
    double c[S], t[S], x[S];  // this lane's samples: index s * G + lane
    double lhi[S], llo[S];    // log(t) = lhi + llo as glibc's pow computes it: depends on t alone, so once per fit
    unsigned ordinary;        // bit s: t[s] is positive, normal and finite (the pow() main path)
    double* scratch;          // shared, this group's: max(4, KB) doubles per sample
    double* s_pts;            // shared, this group's: KB candidate points + KB costs + KB flags
    int nper, lane, model;
    unsigned mask;
    double delta;

    __device__ __forceinline__ double from_lane(double v, int src) const { return __shfl_sync(mask, v, src, G); }

    __device__ __forceinline__ void prepare() {
        ordinary = 0u;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            lhi[s] = llo[s] = 0.0;
            if (pow_base_is_ordinary(t[s])) {
                ordinary |= 1u << s;
                pow_log_of_base(t[s], &lhi[s], &llo[s]);
            }
        }
    }

    // pw[s][k] = pow(t[s], y[k]) with libm's bits.  The main path of all S x K values is straight-line code (S x K
    // independent chains); values that leave it -- a base that is not an ordinary number, an exponent below 2^-65, a
    // result that under- or overflows -- are redone through the complete function under one rarely taken branch.
    template <int K>
    __device__ __forceinline__ void powers(const double* y, double (*pw)[K]) const {
        bool care = false;
        double ehi[S][K], elo[S][K];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                pow_scaled_log(lhi[s], llo[s], y[k], &ehi[s][k], &elo[s][k]);
                care |= glibcpow::exp_needs_care(ehi[s][k]);
            }
#pragma unroll
        for (int k = 0; k < K; ++k) care |= !pow_exponent_is_ordinary(y[k]);
        care |= ordinary != ((1u << S) - 1u);
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int k = 0; k < K; ++k) pw[s][k] = glibcpow::exp_ordinary(ehi[s][k], elo[s][k]);
        if (care) {
#pragma unroll
            for (int s = 0; s < S; ++s)
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (!((ordinary >> s) & 1u) || !pow_exponent_is_ordinary(y[k]) || glibcpow::exp_needs_care(ehi[s][k]))
                        pw[s][k] = pow_libm(t[s], y[k]);
        }
    }

    // dlevmar_L2nrmxmy's sum of sq[0 .. nper) (misc_core.c:721-807) by the four lanes q4 .. q4+3 of a quad: lane part p
    // walks partial sum p downwards through the blocks of eight, then the remainder in its switch order
    __device__ __forceinline__ double l2_partial(const double* sq, int part) const {
        const int blockn = (nper >> 3) << 3;
        double sum = 0.0;
        for (int i = blockn - 1 - part; i >= 0; i -= 4) sum = __dadd_rn(sum, sq[i]);
        const int r = nper - blockn;  // sums 0,1,2,3,0,1,2 entered at case r
        for (int j = 0; j < r; ++j)
            if (((7 - r + j) & 3) == part) sum = __dadd_rn(sum, sq[blockn + j]);
        return sum;
    }
    __device__ __forceinline__ double l2_combine(double partial, int quad_first) const {
        const double s0 = from_lane(partial, quad_first), s1 = from_lane(partial, quad_first + 1),
                     s2 = from_lane(partial, quad_first + 2), s3 = from_lane(partial, quad_first + 3);
        return __dadd_rn(__dadd_rn(__dadd_rn(s0, s1), s2), s3);
    }

    // VECNORM(e) is not finite (lmbc_core.c:146-170, 748), literally; e in sq[0 .. nper).  Rare.
    __device__ __forceinline__ bool vecnorm_not_finite(const double* e) const {
        double mx = 0.0;
        for (int i = nper; i-- > 0;) {
            const double v = e[i];
            if (v > mx) mx = v;
            else if (v < -mx) mx = -v;
        }
        double sm = 0.0;
        for (int i = nper; i-- > 0;) {
            const double q = e[i] / mx;
            sm += q * q;
        }
        return !lm_finite(mx * sqrt(sm));
    }

    // ||x - f(p)||^2 in dlevmar_L2nrmxmy's order; bad = VECNORM(e) is not finite
    __device__ __forceinline__ double cost(const double* p, bool& bad) const {
        const double kd = p[0], n = p[2], ksp = model_ks(model, p[1], n);
        double pw[S][1];
        powers<1>(&n, pw);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int idx = s * G + lane;
            if (idx < nper) {
                const double e = __dsub_rn(x[s], model_hx(kd, c[s], ksp, pw[s][0]));
                scratch[idx] = __dmul_rn(e, e);
                scratch[nper + idx] = e;
            }
        }
        __syncwarp(mask);
        const double esq = l2_combine(lane < 4 ? l2_partial(scratch, lane) : 0.0, 0);
        bad = false;
        if (!lm_finite(esq)) bad = vecnorm_not_finite(scratch + nper);  // (every lane redundantly: the values are uniform)
        __syncwarp(mask);
        return esq;
    }

    // up to KB trial points at once; the engine wrote them (every lane the same values) into s_pts
    __device__ __forceinline__ double* batch_points() const { return s_pts; }
    __device__ __forceinline__ double batch_cost(int k) const { return s_pts[3 * KB + k]; }
    __device__ __forceinline__ bool batch_bad(int k) const { return s_pts[4 * KB + k] != 0.0; }
    __device__ __forceinline__ void cost_many(int cnt, const double* /*dscl: batched fits are unscaled*/, int) const {
        __syncwarp(mask);
        if (cnt == 1) {  // (the first candidate of every walk: most walks end there)
            const double pt[3] = {s_pts[0], s_pts[1], s_pts[2]};
            bool bad;
            const double e = cost(pt, bad);
            if (lane == 0) { s_pts[3 * KB] = e; s_pts[4 * KB] = bad ? 1.0 : 0.0; }
            __syncwarp(mask);
            return;
        }
        double kd[KB], ksp[KB], nn[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const int kk = k < cnt ? k : 0;
            kd[k] = s_pts[3 * kk];
            nn[k] = s_pts[3 * kk + 2];
            ksp[k] = model_ks(model, s_pts[3 * kk + 1], nn[k]);
        }
        double pw[S][KB];
        powers<KB>(nn, pw);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int idx = s * G + lane;
            if (idx < nper) {
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    const double e = __dsub_rn(x[s], model_hx(kd[k], c[s], ksp[k], pw[s][k]));
                    scratch[k * nper + idx] = __dmul_rn(e, e);
                }
            }
        }
        __syncwarp(mask);
        const int quad = lane >> 2;
        const double part = quad < cnt ? l2_partial(scratch + quad * nper, lane & 3) : 0.0;
        const double esq = l2_combine(part, lane & ~3);
        if (quad < cnt && (lane & 3) == 0) { s_pts[3 * KB + quad] = esq; s_pts[4 * KB + quad] = 0.0; }
        __syncwarp(mask);
        for (int k = 0; k < cnt; ++k) {
            if (!lm_finite(s_pts[3 * KB + k])) {  // rare, uniform: the scaled norm needs the residuals themselves
                const double pt[3] = {s_pts[3 * k], s_pts[3 * k + 1], s_pts[3 * k + 2]};
                bool bad;
                cost(pt, bad);
                if (lane == 0) s_pts[4 * KB + k] = bad ? 1.0 : 0.0;
                __syncwarp(mask);
            }
        }
    }

    // levmar's forward-difference Jacobian and its small-problem normal equations
    __device__ __forceinline__ void jac(const double* p, double* JtJ, double* Jte) const {
        double d[3], ph[3], inv[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // misc_core.c:153-170
            double dj = 1E-04 * p[j];
            dj = lm_abs(dj);
            if (dj < delta) dj = delta;
            d[j] = dj;
            ph[j] = p[j] + dj;
            inv[j] = 1.0 / dj;
        }
        const double kd = p[0], n = p[2];
        const double ksp = model_ks(model, p[1], n);            // f(p), f(p + d0 e0)
        const double ksp_hi1 = model_ks(model, ph[1], n);       // f(p + d1 e1)
        const double ksp_hi2 = model_ks(model, p[1], ph[2]);    // f(p + d2 e2): the Phong factor follows n
        const double y2[2] = {n, ph[2]};
        double pw[S][2];
        powers<2>(y2, pw);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int idx = s * G + lane;
            if (idx < nper) {
                const double hx = model_hx(kd, c[s], ksp, pw[s][0]);
                double* row = scratch + 4 * idx;
                row[0] = __dmul_rn(__dsub_rn(model_hx(ph[0], c[s], ksp, pw[s][0]), hx), inv[0]);
                row[1] = __dmul_rn(__dsub_rn(model_hx(kd, c[s], ksp_hi1, pw[s][0]), hx), inv[1]);
                row[2] = __dmul_rn(__dsub_rn(model_hx(kd, c[s], ksp_hi2, pw[s][1]), hx), inv[2]);
                row[3] = __dsub_rn(x[s], hx);  // e, lmbc_core.c:526 / 779
            }
        }
        __syncwarp(mask);
        // lmbc_core.c:603-612: for l = n-1 .. 0: JtJ[i][j] += J[l][j] * J[l][i] (j <= i), Jte[i] += J[l][i] * e[l].
        // Nine independent accumulators, one lane each: (2,2) (2,1) (2,0) (1,1) (1,0) (0,0), then Jte 2, 1, 0.
        double acc = 0.0;
        if (lane < 9) {
            const int a = lane < 3 ? 2 : lane < 5 ? 1 : lane < 6 ? 0 : 8 - lane;             // i (alpha = J[l][i])
            const int b = lane < 3 ? 2 - lane : lane < 5 ? 4 - lane : lane < 6 ? 0 : 3;       // j, or 3 = e
            for (int l = nper; l-- > 0;) acc = __dadd_rn(acc, __dmul_rn(scratch[4 * l + b], scratch[4 * l + a]));
        }
        const double a22 = from_lane(acc, 0), a21 = from_lane(acc, 1), a20 = from_lane(acc, 2), a11 = from_lane(acc, 3),
                     a10 = from_lane(acc, 4), a00 = from_lane(acc, 5);
        Jte[2] = from_lane(acc, 6); Jte[1] = from_lane(acc, 7); Jte[0] = from_lane(acc, 8);
        JtJ[0] = a00; JtJ[1] = a10; JtJ[2] = a20;
        JtJ[3] = a10; JtJ[4] = a11; JtJ[5] = a21;
        JtJ[6] = a20; JtJ[7] = a21; JtJ[8] = a22;
        __syncwarp(mask);
    }
};

#ifndef BG_EXACT_MIN_BLOCKS
#define BG_EXACT_MIN_BLOCKS(G) ((G) == 16 ? 8 : 4)  // 64 / 128 registers (profiles/r02_batched.md)
#endif

template <int G, int S>
__global__ void __launch_bounds__(kExactThreads, BG_EXACT_MIN_BLOCKS(G)) k_batched_fit_exact(const double* __restrict__ c, const double* __restrict__ traw,
                                                                     const double* __restrict__ x, long nfit, ExactSpec spec,
                                                                     double* __restrict__ p_out, double* __restrict__ info_out,
                                                                     int* __restrict__ ret_out) {
    extern __shared__ double exact_smem[];
    const long fit = ((long)blockIdx.x * kExactThreads + threadIdx.x) / G;
    if (fit >= nfit) return;
    const int lane = threadIdx.x % G, nper = spec.nper;
    const long base = fit * nper;
    constexpr int KB = ExactEval<G, S>::KB;
    constexpr int kPerSample = KB > 4 ? KB : 4;
    ExactEval<G, S> ev;
    double* mine = exact_smem + (size_t)(threadIdx.x / G) * ((size_t)kPerSample * nper + 5 * KB + 24);
    ev.scratch = mine;
    ev.s_pts = mine + (size_t)kPerSample * nper;
    ev.s_ws = ev.s_pts + 5 * KB;
    ev.nper = nper; ev.lane = lane; ev.model = spec.model; ev.delta = spec.delta;
    ev.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int idx = s * G + lane;
        const bool in = idx < nper;
        ev.c[s] = in ? c[base + idx] : 0.0;
        ev.t[s] = in ? traw[base + idx] : 0.5;
        ev.x[s] = in ? x[base + idx] : 0.0;
    }
    ev.prepare();
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
        info[7] += info[8] * 4.0;  // dlevmar_bc_dif charges every forward-difference Jacobian m + 1 calls, lmbc_core.c:1119-1124
        for (int i = 0; i < 3; ++i) p_out[fit * 3 + i] = p[i];
        if (info_out)
            for (int i = 0; i < 10; ++i) info_out[fit * 10 + i] = info[i];
        if (ret_out) ret_out[fit] = ret;
#ifdef BG_FIT_CYCLES  // debug variant (profiles/fit_cycles.py): SM cycles this fit took, in units of 64, instead of the return value
        if (ret_out) ret_out[fit] = (int)((clock64() - t_fit0) >> 6);
#endif
    }
}

template <int G, int S>
static int launch_exact(brdfgpu_ctx* ctx, brdfgpu_batch* b, const ExactSpec& spec) {
    constexpr int KB = ExactEval<G, S>::KB;
    const size_t smem = sizeof(double) * ((size_t)(KB > 4 ? KB : 4) * b->nper + 5 * KB + 24) * (kExactThreads / G);
    if (smem > 48 * 1024)
        BG_CUDA_OK(ctx, cudaFuncSetAttribute(k_batched_fit_exact<G, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long blocks = (b->nfit * G + kExactThreads - 1) / kExactThreads;
    k_batched_fit_exact<G, S><<<(unsigned)blocks, kExactThreads, smem, ctx->stream>>>(b->c, b->traw, b->x, b->nfit, spec, b->p, b->info,
                                                                                   b->ret);
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// BRDFGPU_JAC_FD_EXACT of brdfgpu_batch_fit
int batch_fit_exact(brdfgpu_ctx* ctx, brdfgpu_batch* b, const double* p0, const double* lb, const double* ub, int itmax,
                    const double* opts) {
    if (b->nfit == 0) return 0;
    const double delta_signed = opts ? opts[4] : kDiffDelta;
    if (delta_signed < 0.0) {
        set_error(ctx, "batch fit, levmar-exact mode: forward differences only (opts[4] >= 0, as both reference presets use)");
        return BRDFGPU_LM_ERROR;
    }
    if (3 * b->nper >= 1024 || b->nper > 128) {  // lmbc_core.c:573: beyond n m = 1024 levmar sums in 32-row blocks
        set_error(ctx, "batch fit, levmar-exact mode: at most 128 samples per fit (levmar's small-problem summation order)");
        return BRDFGPU_LM_ERROR;
    }
    if (lb && ub)
        for (int i = 0; i < 3; ++i)
            if (lb[i] > ub[i]) {
                fprintf(stderr, "brdfgpu batch fit: at least one lower bound exceeds the upper one\n");
                return BRDFGPU_LM_ERROR;
            }
    ExactSpec spec;
    spec.itmax = itmax; spec.model = b->model; spec.nper = b->nper;
    spec.delta = delta_signed;
    spec.has_lb = lb != nullptr; spec.has_ub = ub != nullptr;
    for (int i = 0; i < 3; ++i) {
        spec.p0[i] = p0[i];
        spec.lb[i] = lb ? lb[i] : 0.0;
        spec.ub[i] = ub ? ub[i] : 0.0;
    }
    spec.opt = lm_options(opts, itmax);
    const int n = b->nper;
    if (n <= 16) return launch_exact<16, 1>(ctx, b, spec);
    if (n <= 32) return launch_exact<32, 1>(ctx, b, spec);
    if (n <= 64) return launch_exact<32, 2>(ctx, b, spec);
    return launch_exact<32, 4>(ctx, b, spec);
}

}  // namespace brdfgpu
