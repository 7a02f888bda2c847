// secant_fit.cu -- dlevmar_dif on resident samples (placeholder until the secant loop lands)
#include "common.cuh"
namespace brdfgpu {
int global_fit_secant(brdfgpu_ctx* ctx, brdfgpu_samples*, double*, int, int, const double*, double*, double*) {
    set_error(ctx, "dlevmar_dif (secant) is not implemented yet");
    return BRDFGPU_LM_ERROR;
}
}  // namespace brdfgpu
