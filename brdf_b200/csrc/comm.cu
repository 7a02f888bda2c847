// comm.cu -- multi-GPU exchange step of the global fit: one process per GPU, NCCL over
// NVLink 5 / NVSwitch.  The reference has no communication at all; the only
// collective the sharded path needs is a sum of the m(m+1)/2 + m + 2 partial sums (fp64) per
// evaluation -- 88 bytes, latency bound.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that libbrdfgpu.so carries no link-time
// dependency on it: a process that already loaded torch's bundled NCCL shares that copy, a plain
// C host gets the system one.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace brdfgpu {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api(brdfgpu_ctx* ctx) {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
        set_error(ctx, "NCCL (libnccl.so.2) could not be loaded");
        return nullptr;
    }
    return &api;
}

#define BG_NCCL_OK(ctx, api, call)                                                                   \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess) {                                                                     \
            set_error((ctx), std::string(#call) + ": " + ((api)->GetErrorString ? (api)->GetErrorString(r_) : "?")); \
            return BRDFGPU_LM_ERROR;                                                                 \
        }                                                                                            \
    } while (0)

int comm_allreduce_device(brdfgpu_ctx* ctx, double* d_buf, int count) {
    if (ctx->nranks <= 1) return 0;
    NcclApi* api = nccl_api(ctx);
    if (!api || !ctx->nccl_comm) {
        set_error(ctx, "all-reduce without a communicator");
        return BRDFGPU_LM_ERROR;
    }
    BG_NCCL_OK(ctx, api, api->AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm,
                                        ctx->stream));
    return 0;
}

}  // namespace brdfgpu

using namespace brdfgpu;

extern "C" int brdfgpu_comm_unique_id(char* id128) {
    NcclApi* api = nccl_api(default_ctx());
    if (!api) return BRDFGPU_LM_ERROR;
    static_assert(sizeof(ncclUniqueId) == BRDFGPU_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return BRDFGPU_LM_ERROR;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

extern "C" int brdfgpu_comm_init(brdfgpu_ctx* ctx, const char* id128, int rank, int nranks) {
    if (!ctx) ctx = default_ctx();
    if (!ctx) return BRDFGPU_LM_ERROR;
    if (nranks <= 1) {
        ctx->rank = 0;
        ctx->nranks = 1;
        return 0;
    }
    NcclApi* api = nccl_api(ctx);
    if (!api) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    BG_NCCL_OK(ctx, api, api->CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return 0;
}

extern "C" void brdfgpu_comm_destroy(brdfgpu_ctx* ctx) {
    if (!ctx) ctx = default_ctx();
    if (!ctx || !ctx->nccl_comm) return;
    NcclApi* api = nccl_api(ctx);
    if (api) api->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->nranks = 1;
}

extern "C" int brdfgpu_comm_allreduce(brdfgpu_ctx* ctx, double* buf, int count) {
    if (!ctx) ctx = default_ctx();
    if (!ctx) return BRDFGPU_LM_ERROR;
    if (count < 0 || count > kResultDoubles) {
        set_error(ctx, "comm_allreduce: count out of range");
        return BRDFGPU_LM_ERROR;
    }
    if (ctx->nranks <= 1) return 0;
    BG_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_result, buf, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (comm_allreduce_device(ctx, ctx->d_result, count) != 0) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaMemcpyAsync(buf, ctx->d_result, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// peer-memory exchange buffers for the fused in-kernel all-reduce (global_fit.cu: peer_exchange).
// Each rank exports one small cudaMalloc'ed buffer through CUDA IPC; the launcher moves the 64-byte
// handles between the processes (any transport) and every rank maps all the others.  Stores to a
// mapped peer pointer travel over NVLink / NVSwitch.
// ------------------------------------------------------------------------------------------------
static size_t peer_buffer_bytes() { return sizeof(uint4) * 2 * kMaxRanks * kPeerCellsPerRank; }

// exported record: the IPC handle, then the exchange tag this rank's session has reached
static_assert(sizeof(cudaIpcMemHandle_t) + 8 == BRDFGPU_IPC_HANDLE_BYTES, "export record = cudaIpcMemHandle_t + 8 bytes");

extern "C" int brdfgpu_peer_export(brdfgpu_ctx* ctx, char* handle64) {
    if (!ctx) ctx = default_ctx();
    if (!ctx || !handle64) return BRDFGPU_LM_ERROR;
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (!ctx->peer_local) {
        BG_CUDA_OK(ctx, cudaMalloc(&ctx->peer_local, peer_buffer_bytes()));
        BG_CUDA_OK(ctx, cudaMemset(ctx->peer_local, 0, peer_buffer_bytes()));
    }
    cudaIpcMemHandle_t h;
    BG_CUDA_OK(ctx, cudaIpcGetMemHandle(&h, ctx->peer_local));
    memcpy(handle64, &h, sizeof(h));
    const unsigned long long reached = ctx->peer_epoch;
    memcpy(handle64 + sizeof(h), &reached, sizeof(reached));
    return 0;
}

extern "C" int brdfgpu_peer_attach(brdfgpu_ctx* ctx, const char* handles, int rank, int nranks) {
    if (!ctx) ctx = default_ctx();
    if (!ctx || !handles) return BRDFGPU_LM_ERROR;
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) {
        set_error(ctx, "peer_attach: need 1 <= nranks <= 8 and 0 <= rank < nranks");
        return BRDFGPU_LM_ERROR;
    }
    if (!ctx->peer_local) {
        set_error(ctx, "peer_attach: call brdfgpu_peer_export first");
        return BRDFGPU_LM_ERROR;
    }
    if (ctx->nccl_comm && (ctx->rank != rank || ctx->nranks != nranks)) {
        set_error(ctx, "peer_attach: rank/nranks differ from the communicator's");
        return BRDFGPU_LM_ERROR;
    }
    BG_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    brdfgpu_peer_detach(ctx);  // mappings of an earlier session
    // The buffers are zeroed once, at allocation; cells of an earlier session stay in them.  Tags therefore never
    // go back: the new session starts above the highest tag any rank has reached (every rank computes the same
    // number from the same records), so a stale cell {value | tag} can never satisfy a wait of this session.
    unsigned long long reached = ctx->peer_epoch;
    for (int r = 0; r < nranks; ++r) {
        unsigned long long e = 0;
        memcpy(&e, handles + (size_t)r * BRDFGPU_IPC_HANDLE_BYTES + sizeof(cudaIpcMemHandle_t), sizeof(e));
        if (e > reached) reached = e;
    }
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            ctx->peer_remote[r] = ctx->peer_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * BRDFGPU_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        BG_CUDA_OK(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_remote[r] = static_cast<uint4*>(ptr);
    }
    ctx->rank = rank;
    ctx->nranks = nranks;
    ctx->peer_epoch = (unsigned)((reached + 4) & 0xfffffffeull);  // (even; 32-bit tags wrap after 2^32 exchanges, next_tag skips 0)
    ctx->peer_attached = nranks > 1;
    return 0;
}

extern "C" void brdfgpu_peer_detach(brdfgpu_ctx* ctx) {
    if (!ctx) ctx = default_ctx();
    if (!ctx) return;
    for (int r = 0; r < kMaxRanks; ++r) {
        if (ctx->peer_remote[r] && ctx->peer_remote[r] != ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_remote[r]);
        ctx->peer_remote[r] = nullptr;
    }
    ctx->peer_attached = false;
    if (!ctx->nccl_comm) {
        ctx->rank = 0;
        ctx->nranks = 1;
    }
}
