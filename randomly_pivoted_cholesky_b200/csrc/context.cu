// Context, scratch management and error reporting of the C ABI (include/rpc_b200.h).
#include <stdarg.h>
#include <string.h>
#include "rpc_common.cuh"

static thread_local char g_err[512] = "";

// the library never changes the caller's current device: allocations of a context happen on ITS device under this guard
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void rpc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *rpc_last_error(void) { return g_err; }
extern "C" int rpc_abi_version(void) { return RPC_ABI_VERSION; }
extern "C" int rpc_abi_sizeof(int which) {
    return which == 0 ? (int)sizeof(rpc_kernel_spec) : which == 1 ? (int)sizeof(rpc_peers) : -1;
}

extern "C" int rpc_ctx_create(int device, rpc_ctx **out) {
    RPC_CHECK_ARG(out != nullptr, "out is null");
    int count = 0;
    RPC_CUDA(cudaGetDeviceCount(&count));
    RPC_CHECK_ARG(device >= 0 && device < count, "no such CUDA device");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    RPC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        rpc_set_error("rpc_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                      prop.major, prop.minor);
        return -2;
    }
    rpc_ctx *ctx = new rpc_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return 0;
}

extern "C" int rpc_ctx_destroy(rpc_ctx *ctx) {
    if (!ctx) return 0;
    DeviceGuard guard(ctx->device);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->flags) cudaFree(ctx->flags);
    if (ctx->sync_counter) cudaFree(ctx->sync_counter);
    if (ctx->panel_ld_table) cudaFree(ctx->panel_ld_table);
    delete ctx;
    return 0;
}

extern "C" int rpc_ctx_reserve_sms(rpc_ctx *ctx, int count) {
    RPC_CHECK_ARG(ctx && count >= 0 && count < ctx->sm_count, "count must be in [0, SM count)");
    ctx->sm_reserved = count;
    return 0;
}

extern "C" int64_t rpc_ctx_launch_count(const rpc_ctx *ctx) { return ctx ? ctx->launches : -1; }

// cudaFree synchronises the device, so growing a buffer that in-flight kernels still read is safe.
int rpc_ws_reserve(rpc_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return 0;
    DeviceGuard guard(ctx->device);
    size_t want = bytes + bytes / 4 + (1 << 20);
    if (ctx->ws) RPC_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    RPC_CUDA(cudaMalloc((void **)&ctx->ws, want));
    ctx->ws_bytes = want;
    return 0;
}

int rpc_sync_reserve(rpc_ctx *ctx) {
    if (ctx->sync_counter) return 0;
    DeviceGuard guard(ctx->device);
    RPC_CUDA(cudaMalloc((void **)&ctx->sync_counter, 16 * sizeof(unsigned int)));
    RPC_CUDA(cudaMemset(ctx->sync_counter, 0, 16 * sizeof(unsigned int)));
    return 0;
}

int rpc_flags_reserve(rpc_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->flags_bytes) return 0;
    DeviceGuard guard(ctx->device);
    size_t want = bytes * 2 + 4096;
    if (ctx->flags) RPC_CUDA(cudaFree(ctx->flags));
    ctx->flags = nullptr;
    ctx->flags_bytes = 0;
    RPC_CUDA(cudaMalloc((void **)&ctx->flags, want));
    ctx->flags_bytes = want;
    return 0;
}
