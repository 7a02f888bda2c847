// synth.cu -- synthetic sample sets of the shape BASELINE.json names, generated on the device so
// that the 10^8-sample sweep point never crosses PCIe.  Same integer recipe as tests/synth.py
// (counter-based splitmix64 finaliser, 53-bit uniforms, truth model + +-0.005 uniform noise,
// 8-bit quantisation like the photographs: SURVEY.md 8d configs 2/4/5).
#include "brdf_model.cuh"
#include "common.cuh"

namespace brdfgpu {

__device__ __forceinline__ double synth_uniform(unsigned long long idx, int stream, unsigned long long seed) {
    unsigned long long z = seed + (idx * 4ull + (unsigned long long)(stream + 1)) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * 0x1.0p-53;
}

__device__ __forceinline__ double synth_measure(double kd, double ks, double nn, int model, double c, double t,
                                                double noise) {
    const double coef = model == 1 ? 1.0 : ((nn + 2.0) / 2.0 * kPi);
    const double val = __dadd_rn(__dadd_rn(__dmul_rn(kd, c), __dmul_rn(__dmul_rn(coef, ks), pow(t, nn))), noise);
    double q = floor(__dmul_rn(255.0, val));
    q = q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q);
    return q / 255.0;
}

__global__ void k_synth_samples(double* c, double* traw, double* x, long n, unsigned long long seed, long start,
                                double kd, double ks, double nn, int model) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const unsigned long long idx = (unsigned long long)(start + i);
        const double cc = synth_uniform(idx, 0, seed);
        const double t = synth_uniform(idx, model == 1 ? 1 : 2, seed);
        const double noise = __dmul_rn(__dsub_rn(synth_uniform(idx, 3, seed), 0.5), 0.01);
        c[i] = cc;
        traw[i] = t;
        x[i] = synth_measure(kd, ks, nn, model, cc, t, noise);
    }
}

// tests/synth.py batched(): truth of fit f from stream seed+1 at index f, samples f*nper..
__global__ void k_synth_batch(double* c, double* traw, double* x, long nfit, int nper, unsigned long long seed,
                              long first_fit, int model) {
    const long total = nfit * nper;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const unsigned long long f = (unsigned long long)(first_fit + i / nper);
        const unsigned long long idx = (unsigned long long)(first_fit * nper + i);
        const double kd = __dadd_rn(0.1, __dmul_rn(0.8, synth_uniform(f, 0, seed + 1)));
        const double ks = __dadd_rn(0.05, __dmul_rn(0.75, synth_uniform(f, 1, seed + 1)));
        const double nn = __dadd_rn(1.0, __dmul_rn(49.0, synth_uniform(f, 2, seed + 1)));
        const double cc = synth_uniform(idx, 0, seed);
        const double t = synth_uniform(idx, model == 1 ? 1 : 2, seed);
        const double noise = __dmul_rn(__dsub_rn(synth_uniform(idx, 3, seed), 0.5), 0.01);
        c[i] = cc;
        traw[i] = t;
        x[i] = synth_measure(kd, ks, nn, model, cc, t, noise);
    }
}

int synth_samples(brdfgpu_ctx* ctx, brdfgpu_samples* s, unsigned long long seed, long start, const double* truth) {
    if (s->n == 0) return 0;
    k_synth_samples<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(s->c, s->traw, s->x, s->n, seed, start, truth[0],
                                                                 truth[1], truth[2], s->model);
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return samples_prepare(ctx, s);
}

int synth_batch(brdfgpu_ctx* ctx, brdfgpu_batch* b, unsigned long long seed, long first_fit) {
    if (b->nfit == 0) return 0;
    k_synth_batch<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(b->c, b->traw, b->x, b->nfit, b->nper, seed, first_fit,
                                                               b->model);
    ++ctx->launches;
    BG_CUDA_OK(ctx, cudaGetLastError());
    return batch_prepare(ctx, b);
}

}  // namespace brdfgpu
