// common.cuh -- context, resident data handles and launch helpers shared by the .cu files.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "../../include/brdfgpu.h"
#include "lm_engine.cuh"

namespace brdfgpu {

constexpr int kMaxM = BRDFGPU_MAX_PARAMS;
constexpr int kPassThreads = 256;    // threads per CTA of the streaming passes
constexpr int kMaxPassBlocks = 2048; // upper bound on CTAs of one pass (partials buffer)
constexpr int kResultDoubles = 16;   // >= NACC
constexpr int kMaxRanks = 8;         // GPUs of one NVSwitch domain
constexpr int kPeerCellsPerRank = 32;  // flagged cells per (parity, rank): the widest sweep carries 32 sums
constexpr int kCountStagingInts = 1024;  // pinned staging for the per-view fit counts of a gather

// Exchange buffer of the fused in-kernel all-reduce: [2 parities][kMaxRanks][kPeerCellsPerRank]
// 16-byte cells {value.lo, tag, value.hi, tag}.  Every 8-byte half is written atomically, so a
// reader that sees the current tag in both halves has the whole double (no fences, no flags).
struct PeerView {
    uint4* local;
    uint4* remote[kMaxRanks];  // remote[rank] == local
    int rank, nranks;
    unsigned epoch;            // tag of the last exchange
};

}  // namespace brdfgpu

// Opaque handles of include/brdfgpu.h
struct brdfgpu_ctx {
    int device = 0;
    int sm_count = 0;
    int coop = 0;  // cooperative launch supported
    cudaStream_t stream = nullptr;
    std::string err;
    unsigned long long launches = 0;

    // reduction scratch of the streaming passes
    double* d_partials = nullptr;  // 2 x kMaxPassBlocks x kResultDoubles (double-buffered)
    unsigned* d_sync = nullptr;    // [0] ticket of the "last block done" reduction
    double* d_result = nullptr;    // kResultDoubles
    double* h_result = nullptr;    // pinned + mapped mirror the last block writes directly
    double* h_result_dev = nullptr;           // device alias of h_result
    volatile unsigned long long* h_seq = nullptr;  // pinned + mapped completion ticket
    unsigned long long* h_seq_dev = nullptr;
    unsigned long long seq = 0;
    // persistent-fit in/out block
    void* d_fitio = nullptr;
    void* h_fitio = nullptr;  // pinned
    int* h_counts = nullptr;      // pinned, kCountStagingInts: per-view fit counts of the running gather
    uint4* d_cells = nullptr;     // flagged exchange cells of the persistent fit: 2 parities x kMaxPersistBlocks x NSUM (32)
    int tma_mode = 0;             // 0: not probed, 1: automatic, 2: never, 3: always (BRDFGPU_TMA)
    long persist_smem_max = 0;    // dynamic shared memory one CTA of the persistent fit may use (0: not probed, <0: unusable)
    // what the last global fit did: sweeps with a Jacobian, cost-only sweeps, trial points evaluated
    // (>= the levmar-counted ones: the projected-gradient walk is evaluated eight points per sweep),
    // samples resident in shared memory, CTAs
    unsigned long long fit_stats[24] = {0};

    // device buffers of the levmar-signature entry points, kept between calls (cudaMalloc/cudaFree
    // per call cost more than the fit itself at 10^6 samples)
    struct brdfgpu_samples* pooled = nullptr;
    long pooled_capacity = 0;
    bool pooled_busy = false;

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory exchange (fused one-shot all-reduce inside the fit kernels)
    uint4* peer_local = nullptr;            // this rank's exchange buffer (cudaMalloc, IPC-exported)
    uint4* peer_remote[brdfgpu::kMaxRanks] = {nullptr};
    bool peer_attached = false;
    unsigned peer_epoch = 0;
};

struct brdfgpu_samples {
    long n = 0;
    long capacity = 0;  // samples the arrays were allocated for (>= n)
    int model = 1;
    // c, L, x, traw are carved from ONE stream-ordered allocation (cudaMallocAsync on the context's stream);
    // block == nullptr: a view onto arrays somebody else owns (the three channel sets of one gather)
    void* block = nullptr;
    cudaStream_t stream = nullptr;
    double* c = nullptr;     // cosphi
    double* L = nullptr;     // log(t), NaN = take the pow() path
    double* x = nullptr;     // measurements
    double* traw = nullptr;  // raw model cosine (read only on the pow() path)
    // secant (dlevmar_dif) state: stored Jacobian, 3 columns SoA, allocated on first use
    double* jac = nullptr;
    long jac_capacity = 0;  // samples the secant state was allocated for
    bool pooled = false;  // owned by the context (brdfgpu_ctx::pooled): free only returns it
};

struct brdfgpu_batch {
    long nfit = 0;
    long capacity = 0;  // fits the arrays were allocated for (>= nfit)
    int nper = 0;
    int model = 1;
    void* block = nullptr;  // one stream-ordered allocation behind all seven arrays
    cudaStream_t stream = nullptr;
    double *c = nullptr, *L = nullptr, *x = nullptr, *traw = nullptr;  // nfit x nper each
    double* p = nullptr;     // nfit x 3
    double* info = nullptr;  // nfit x 10
    int* ret = nullptr;      // nfit
};

namespace brdfgpu {

brdfgpu_ctx* default_ctx();  // process-wide context for the levmar-signature entry points
void set_error(brdfgpu_ctx* ctx, const std::string& msg);

#define BG_CUDA_OK(ctx, call)                                                                    \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            ::brdfgpu::set_error((ctx), std::string(#call) + ": " + cudaGetErrorString(e_));     \
            return BRDFGPU_LM_ERROR;                                                             \
        }                                                                                        \
    } while (0)

inline int pass_blocks(const brdfgpu_ctx* ctx, long n, int ctas_per_sm) {
    // grid sized in multiples of the SM count (148 on B200); small problems get fewer CTAs so the
    // final cross-CTA reduction stays short
    long want = (n / 2 + kPassThreads - 1) / kPassThreads;
    long cap = (long)ctx->sm_count * ctas_per_sm;
    if (cap > kMaxPassBlocks) cap = kMaxPassBlocks;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// ---- global_fit.cu ----
struct GlobalFitSpec {
    int m, itmax, jac_mode, has_lb, has_ub, has_dscl, unconstrained, spec_jac;
    int want_n_all;  // several ranks + covariance: sum the ranks' sample counts through the kernel's exchange
    double delta;  // |opts[4]|
    double p[kMaxM], lb[kMaxM], ub[kMaxM], dscl[kMaxM];
    LmOptions opt;
};
struct GlobalFitOut {
    int ret;
    unsigned peer_epoch;  // exchange tag after the fit (all ranks advance in lock step)
    int aborted;          // an exchange partner never delivered: the fit was abandoned
    unsigned jac_passes, cost_passes, cost_points, spec_issued, spec_hits, creep_fused;
    long long cyc_sweep, cyc_exchange, cyc_total;  // SM cycles of CTA 0 / thread 0
    long long cyc_x[4];
    long long cyc_ctl[7];  // control-code cycles by the kind of sweep they led to (SweepKind)
    double p[kMaxM];
    double info[10];
    double JtJ[kMaxM * kMaxM];
    double n_all;  // samples of all ranks (GlobalFitSpec::want_n_all)
};

int samples_alloc(brdfgpu_ctx* ctx, long n, int model, brdfgpu_samples** out);
void free_block(brdfgpu_ctx* ctx, void* block, cudaStream_t stream);  // stream-ordered when ctx still owns `stream`
int samples_prepare(brdfgpu_ctx* ctx, brdfgpu_samples* s);  // L from traw
int global_normal_eq(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta, int jac_mode,
                     double* out11);
int global_cost(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double* out2);
int global_repeat(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double delta, int kind, int reps);
int global_residuals(brdfgpu_ctx* ctx, const brdfgpu_samples* s, const double* p, double* e_host);
int global_fit(brdfgpu_ctx* ctx, const brdfgpu_samples* s, double* p, int m, const double* lb, const double* ub,
               const double* dscl, int itmax, const double* opts, double* info, double* covar, int drive,
               int jac_mode, int unconstrained);
int global_fit_secant(brdfgpu_ctx* ctx, brdfgpu_samples* s, double* p, int m, int itmax, const double* opts,
                      double* info, double* covar);
int model_predict(brdfgpu_ctx* ctx, const double* p, const double* angles_host, int model, int n, double* hx_host);
int model_jacobian(brdfgpu_ctx* ctx, const double* p, const double* angles_host, int model, int n, int m,
                   double* jac_host);

// ---- synth.cu ----
int synth_samples(brdfgpu_ctx* ctx, brdfgpu_samples* s, unsigned long long seed, long start, const double* truth);
int synth_batch(brdfgpu_ctx* ctx, brdfgpu_batch* b, unsigned long long seed, long first_fit);

// ---- batched_fit.cu ----
int batch_alloc(brdfgpu_ctx* ctx, long nfit, int nper, int model, brdfgpu_batch** out);
int batch_prepare(brdfgpu_ctx* ctx, brdfgpu_batch* b);
int batch_fit(brdfgpu_ctx* ctx, brdfgpu_batch* b, const double* p0, const double* lb, const double* ub, int itmax,
              const double* opts, int jac_mode);

// ---- batched_exact.cu ----
int batch_fit_exact(brdfgpu_ctx* ctx, brdfgpu_batch* b, const double* p0, const double* lb, const double* ub, int itmax,
                    const double* opts);

// ---- comm.cu ----
int comm_allreduce_device(brdfgpu_ctx* ctx, double* d_buf, int count);

}  // namespace brdfgpu
