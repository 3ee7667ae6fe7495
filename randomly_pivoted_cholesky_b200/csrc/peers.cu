// Peer-memory exchange of the row-sharded factorisation (include/rpc_b200.h, "peer-memory exchange"): buffers exported
// over CUDA IPC, the flag wait, and the pivot gather whose owners store into every rank's payload through NVLink.
// The diag side of the exchange lives in the Schur kernel's epilogue (dmma_gemm.cuh).
#include <string.h>
#include "rpc_common.cuh"

extern "C" int rpc_peer_alloc(rpc_ctx *ctx, int64_t bytes, void **ptr, unsigned char handle[64]) {
    RPC_CHECK_ARG(ctx && ptr && handle && bytes > 0, "null pointer or bad size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    RPC_CUDA(cudaSetDevice(ctx->device));
    void *p = nullptr;
    RPC_CUDA(cudaMalloc(&p, (size_t)bytes));
    RPC_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        rpc_set_error("rpc_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return (int)e;
    }
    memcpy(handle, &h, 64);
    *ptr = p;
    return 0;
}

extern "C" int rpc_peer_free(rpc_ctx *ctx, void *ptr) {
    RPC_CHECK_ARG(ctx != nullptr, "null context");
    if (ptr) RPC_CUDA(cudaFree(ptr));
    return 0;
}

extern "C" int rpc_peer_open(rpc_ctx *ctx, const unsigned char handle[64], void **ptr) {
    RPC_CHECK_ARG(ctx && handle && ptr, "null pointer");
    RPC_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    RPC_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int rpc_peer_close(rpc_ctx *ctx, void *ptr) {
    RPC_CHECK_ARG(ctx != nullptr, "null context");
    if (ptr) RPC_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}

// lane i spins on flag word i (written by rank i's kernels through NVLink, landing in this GPU's memory)
__global__ void __launch_bounds__(32) peer_wait_kernel(const unsigned long long *flags, int count, unsigned long long epoch,
                                                       unsigned long long timeout_ns, int32_t *status) {
    const int i = threadIdx.x;
    if (i >= count) return;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(flags + i) < epoch) {
        if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > timeout_ns) {
            if (status) atomicExch(status, 1);
            return;
        }
    }
}

extern "C" int rpc_peer_wait(rpc_ctx *ctx, void *stream, const unsigned long long *flags, int count, unsigned long long epoch,
                             int timeout_ms, int32_t *status) {
    RPC_CHECK_ARG(ctx && flags && count >= 1 && count <= 32, "null pointer or bad count");
    if (timeout_ms <= 0) timeout_ms = 20000;
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, count, epoch, (unsigned long long)timeout_ms * 1000000ull, status);
    RPC_LAUNCHED(ctx);
    return 0;
}

struct PeerPtrs {
    int n;
    double *payload[RPC_MAX_PEERS];
    unsigned long long *flag[RPC_MAX_PEERS];      // already offset to this source rank's payload flag word
};

// blocks [0, c): row blk of P = G[blk, idx - lo] for the pivots this rank owns; blocks [c, c + m): pivot point j and its norm
__global__ void __launch_bounds__(128) gather_pivots_peers_kernel(const int64_t *__restrict__ idx, int m, const double *__restrict__ X,
                                                                  const double *__restrict__ xnorm, int d, int64_t ldx, int64_t ldp,
                                                                  int64_t off_norm, const double *__restrict__ G, int64_t ldg, int c,
                                                                  int64_t off_P, int64_t ldP, int64_t lo, int64_t hi, PeerPtrs pp,
                                                                  unsigned long long epoch, unsigned int *done_counter) {
    const int blk = blockIdx.x;
    if (blk < c) {
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const int64_t g = idx[j];
            if (g >= lo && g < hi) {
                const double v = G[(int64_t)blk * ldg + (g - lo)];
                for (int r = 0; r < pp.n; ++r) pp.payload[r][off_P + (int64_t)blk * ldP + j] = v;
            }
        }
    } else {
        const int j = blk - c;
        const int64_t g = idx[j];
        if (g >= lo && g < hi) {
            const int64_t src = g - lo;
            for (int f = threadIdx.x; f < ldp; f += blockDim.x) {
                const double v = f < d ? X[src * ldx + f] : 0.0;
                for (int r = 0; r < pp.n; ++r) pp.payload[r][(int64_t)j * ldp + f] = v;
            }
            if (threadIdx.x == 0 && xnorm) {
                const double v = xnorm[src];
                for (int r = 0; r < pp.n; ++r) pp.payload[r][off_norm + j] = v;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(done_counter, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            *done_counter = 0;
            for (int r = 0; r < pp.n; ++r)
                asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(pp.flag[r]), "l"(epoch) : "memory");
        }
    }
}

extern "C" int rpc_gather_pivots_peers(rpc_ctx *ctx, void *stream, const int64_t *idx, int m, const double *X,
                                       const double *xnorm, int d, int64_t ldx, int64_t ldp, int64_t off_norm, const double *G,
                                       int64_t ldg, int c, int64_t off_P, int64_t ldP, int64_t lo, int64_t hi,
                                       const rpc_peers *peers, unsigned long long epoch) {
    RPC_CHECK_ARG(ctx && idx && peers && m >= 1 && c >= 0, "null pointer or bad sizes");
    RPC_CHECK_ARG(c == 0 || G, "G required when c > 0");
    RPC_CHECK_ARG(peers->world >= 1 && peers->world <= RPC_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
                  "bad peer table");
    RPC_CHECK_ARG(hi >= lo, "empty ownership range");
    int rc = rpc_sync_reserve(ctx);
    if (rc) return rc;
    PeerPtrs pp{};
    pp.n = peers->world;
    for (int r = 0; r < peers->world; ++r) {
        RPC_CHECK_ARG(peers->payload[r] && peers->flags[r], "peer table has null entries");
        pp.payload[r] = peers->payload[r];
        pp.flag[r] = peers->flags[r] + RPC_MAX_PEERS + peers->rank;
    }
    // at least one block so that the epoch is always published, even by a rank that owns none of the pivots
    const int blocks = c + (X ? m : 0) > 0 ? c + (X ? m : 0) : 1;
    if (!X && c == 0) {
        gather_pivots_peers_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(idx, 0, nullptr, nullptr, 0, 0, ldp, off_norm, nullptr, 0, 0,
                                                                       off_P, ldP, 0, 0, pp, epoch, ctx->sync_counter + 1);
    } else {
        gather_pivots_peers_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(idx, m, X, xnorm, d, ldx, ldp, off_norm, G, ldg, c,
                                                                            off_P, ldP, lo, hi, pp, epoch, ctx->sync_counter + 1);
    }
    RPC_LAUNCHED(ctx);
    return 0;
}
