// reduce.cuh -- sample views and the cross-CTA reduction shared by the streaming kernels
// (global_fit.cu, secant_fit.cu).
#pragma once

#include "brdf_model.cuh"
#include "common.cuh"

namespace brdfgpu {

struct SampleView {
    const double *c, *L, *x, *traw;
    long n;
};

inline SampleView view_of(const brdfgpu_samples* s) { return SampleView{s->c, s->L, s->x, s->traw, s->n}; }

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


// host side (global_fit.cu)
int wait_ticket(brdfgpu_ctx* ctx, unsigned long long seq);
int fetch_result(brdfgpu_ctx* ctx, int count, bool published);  // sums of all ranks in ctx->h_result
int launch_cost(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, bool publish, bool count_bad = false);

}  // namespace brdfgpu
