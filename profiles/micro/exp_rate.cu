// What bounds the resident cost sweeps of k_persistent_fit?  One CTA of 512 threads per SM, a configs[1]-sized shard
// (3379 sample pairs) in shared memory, the same per-sample arithmetic (brdf_model.cuh), no control code around it.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a [-DBG_EXP_TABLE=1] -I../../brdf_b200/csrc -o exp_rate exp_rate.cu
// Prints cycles per trial point and CTA for several loop shapes, and the raw DFMA rate / latency of one SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "brdf_model.cuh"
using namespace brdfgpu;

constexpr int kThreads = 512;
__device__ __forceinline__ double2 lds_pair(const double2* a, int i) { return a[i]; }

template <int FORM>
__global__ void __launch_bounds__(kThreads, 1) k_rate(int pairs, int reps, CostPoint q0, CostPoint q1, double* out, long long* cyc) {
    extern __shared__ double2 sm[];
    double2 *sc = sm, *sl = sm + pairs, *sx = sm + 2 * pairs;
    BG_EXP_TABLE_LOAD();
    for (int i = threadIdx.x; i < pairs; i += kThreads) {
        const double t = 0.05 + 0.9 * ((i * 37 + blockIdx.x) % 1000) / 1000.0;
        sc[i] = make_double2(0.3 + 0.001 * (i % 97), 0.4);
        sl[i] = make_double2(log(t), log(0.5 * t + 0.2));
        sx[i] = make_double2(0.5, 0.6);
    }
    __syncthreads();
    const double* traw = nullptr;
    double total = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        CostPoint qa = q0, qb = q1;
        qa.n += 1e-9 * r; qb.n += 2e-9 * r;
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        int i = threadIdx.x;
        if (FORM == 0) {  // one point, two pairs per trip (resident_cost)
            for (; i + kThreads < pairs; i += 2 * kThreads) {
                const int i2 = i + kThreads;
                accumulate_cost_2pairs(qa, sc[i], sl[i], sx[i], 2L * i, sc[i2], sl[i2], sx[i2], 2L * i2, traw, &a0, &a1);
            }
            if (i < pairs) accumulate_cost_pair(qa, sc[i], sl[i], sx[i], traw, 2L * i, &a0);
        } else if (FORM == 1) {  // two points, two pairs per trip (resident_cost_x2)
            for (; i + kThreads < pairs; i += 2 * kThreads) {
                const int i2 = i + kThreads;
                const double2 c = sc[i], l = sl[i], x = sx[i], d = sc[i2], m = sl[i2], y = sx[i2];
                const double cc[4] = {c.x, c.y, d.x, d.y}, ll[4] = {l.x, l.y, m.x, m.y}, xx[4] = {x.x, x.y, y.x, y.y};
                const long idx[4] = {2L * i, 2L * i + 1, 2L * i2, 2L * i2 + 1};
                double ea[4], eb[4];
                residuals_n_x2<4>(qa, qb, cc, ll, xx, traw, idx, ea, eb);
                a0 = __fma_rn(ea[0], ea[0], a0); a0 = __fma_rn(ea[1], ea[1], a0);
                a1 = __fma_rn(ea[2], ea[2], a1); a1 = __fma_rn(ea[3], ea[3], a1);
                b0 = __fma_rn(eb[0], eb[0], b0); b0 = __fma_rn(eb[1], eb[1], b0);
                b1 = __fma_rn(eb[2], eb[2], b1); b1 = __fma_rn(eb[3], eb[3], b1);
            }
            if (i < pairs) {
                accumulate_cost_pair(qa, sc[i], sl[i], sx[i], traw, 2L * i, &a0);
                accumulate_cost_pair(qb, sc[i], sl[i], sx[i], traw, 2L * i, &b0);
            }
        } else if (FORM == 2) {  // one point, one pair per trip
            for (; i < pairs; i += kThreads) accumulate_cost_pair(qa, sc[i], sl[i], sx[i], traw, 2L * i, &a0);
        } else if (FORM == 3) {  // one point, four pairs per trip
            for (; i + 3 * kThreads < pairs; i += 4 * kThreads) {
                double cc[8], ll[8], xx[8], e[8];
                long idx[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 c = sc[i + k * kThreads], l = sl[i + k * kThreads], x = sx[i + k * kThreads];
                    cc[2 * k] = c.x; cc[2 * k + 1] = c.y; ll[2 * k] = l.x; ll[2 * k + 1] = l.y; xx[2 * k] = x.x; xx[2 * k + 1] = x.y;
                    idx[2 * k] = 2L * (i + k * kThreads); idx[2 * k + 1] = idx[2 * k] + 1;
                }
                residuals_n<8>(qa, cc, ll, xx, traw, idx, e);
#pragma unroll
                for (int k = 0; k < 8; k += 2) { a0 = __fma_rn(e[k], e[k], a0); a1 = __fma_rn(e[k + 1], e[k + 1], a1); }
            }
            for (; i < pairs; i += kThreads) accumulate_cost_pair(qa, sc[i], sl[i], sx[i], traw, 2L * i, &a0);
        }
        total += (a0 + a1) + (b0 + b1);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * kThreads + threadIdx.x] = total;
}

// raw rate: CH independent DFMA chains per thread, 64 steps
template <int CH>
__global__ void __launch_bounds__(kThreads, 1) k_dfma(double* out, long long* cyc, double b) {
    double a[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) a[k] = out[threadIdx.x] + k;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 64; ++r) {
#pragma unroll
        for (int s = 0; s < 16; ++s) {
#pragma unroll
            for (int k = 0; k < CH; ++k) a[k] = __fma_rn(a[k], b, b);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; ++k) s += a[k];
    out[blockIdx.x * kThreads + threadIdx.x] = s;
}

template <int FORM>
static void run(const char* name, int points_per_rep) {
    const int pairs = 3379, reps = 200;
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * kThreads);
    cudaMalloc(&cyc, 64);
    const size_t smem = (size_t)3 * pairs * sizeof(double2);
    cudaFuncSetAttribute(k_rate<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    CostPoint q0{0.6, 0.35, 12.0}, q1{0.61, 0.34, 12.5};
    for (int w = 0; w < 2; ++w) k_rate<FORM><<<148, kThreads, smem>>>(pairs, reps, q0, q1, out, cyc);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-44s %8.0f cycles per point and CTA  (%.1f per warp-sample-point and scheduler)%s\n", name, (double)h / reps / points_per_rep,
           (double)h / reps / points_per_rep / (2.0 * pairs / 32 / 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

template <int CH>
static void run_dfma() {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * kThreads);
    cudaMemset(out, 0, sizeof(double) * 148 * kThreads);
    cudaMalloc(&cyc, 64);
    for (int w = 0; w < 2; ++w) k_dfma<CH><<<148, kThreads>>>(out, cyc, 0.999);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = 64.0 * 16 * CH;  // DFMAs per thread
    printf("DFMA, %d chains per thread, 16 warps: %.2f cycles per warp instruction and scheduler (4 warps each); %.1f cycles per step of one chain\n",
           CH, (double)h / (n * 4), (double)h / (64.0 * 16));
    cudaFree(out); cudaFree(cyc);
}

int main() {
#ifdef BG_EXP_TABLE
    printf("table-driven exponential\n");
#else
    printf("polynomial exponential\n");
#endif
    run<2>("1 point, 1 pair per trip (2 chains)", 1);
    run<0>("1 point, 2 pairs per trip (4 chains)", 1);
    run<3>("1 point, 4 pairs per trip (8 chains)", 1);
    run<1>("2 points, 2 pairs per trip (8 chains)", 2);
    run_dfma<1>(); run_dfma<2>(); run_dfma<4>(); run_dfma<8>();
    return 0;
}
