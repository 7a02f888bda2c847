// Latency probes used to design the persistent fit's exchange (B200, sm_100a).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, unsigned long long* cells) {
    __shared__ double sm[512];
    double a = out[0], b = out[1];
    long long t0, t1;
    int slot = 0;
    sm[threadIdx.x] = a;
    __syncthreads();
    // 1. dependent DFMA chain (64)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = __fma_rn(a, b, b);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 2. dependent DADD chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) a = a + b;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 3. shuffle + DADD chain (32 steps)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1 + (i & 15));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 4. dependent LDS chain (32)
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) { double v = sm[idx]; idx = (idx + (int)v) & 511; a += v; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 5. __syncthreads x 16
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 6. dependent ld.relaxed.gpu chain from L2 (16)
    unsigned long long p = 0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        unsigned long long v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(cells + p) : "memory");
        p = v;  // cells hold 0
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 7. same, volatile (sys)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        unsigned long long v;
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(cells + p) : "memory");
        p = v;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 8. fp64 division chain (16)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) a = b / (a + 3.0);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 9. dependent local-memory (stack) access: dynamic index array
    double loc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) loc[i] = a + i;
    int j = threadIdx.x & 31;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) { double v = loc[j]; j = (j + 1 + (int)(v * 0.0)) & 31; a += v; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[slot] = t1 - t0; slot++;
    // 10. store -> remote CTA sees it: ping-pong between block 0 and block 1 (32 round trips)
    if (gridDim.x >= 2 && threadIdx.x == 0 && blockIdx.x < 2) {
        unsigned long long* flag = cells + 64;
        t0 = clock64();
        for (unsigned long long i = 1; i <= 32; ++i) {
            if (blockIdx.x == 0) {
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(2 * i - 1) : "memory");
                unsigned long long v;
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory"); } while (v != 2 * i);
            } else {
                unsigned long long v;
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory"); } while (v != 2 * i - 1);
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(2 * i) : "memory");
            }
        }
        t1 = clock64();
        if (blockIdx.x == 0) cyc[slot] = t1 - t0;
    }
    slot++;
    out[2 + threadIdx.x % 2] = a + (double)p + idx + j;
}
int main() {
    double* out; long long* cyc; unsigned long long* cells;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&cells, 1024);
    cudaMemset(out, 0, 64); cudaMemset(cells, 0, 1024); cudaMemset(cyc, 0, 128);
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(cells, 0, 1024);
        probe<<<2, 512>>>(out, cyc, cells);
        cudaDeviceSynchronize();
    }
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA dependent /64", "DADD dependent /64", "SHFL+DADD /32", "LDS dependent /32", "__syncthreads(512 thr) /16",
                           "ld.relaxed.gpu L2 /16", "ld.volatile L2 /16", "fp64 div (+add) /16", "local mem dyn-index /32", "CTA<->CTA flag ping-pong (round trip) /32"};
    const int div[] = {64, 64, 32, 32, 16, 16, 16, 16, 32, 32};
    for (int i = 0; i < 10; ++i) printf("%-44s %8.1f cycles\n", names[i], (double)h[i] / div[i]);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
