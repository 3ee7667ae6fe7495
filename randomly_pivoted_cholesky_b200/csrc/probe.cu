// Measurement aid: the FP64 tensor-pipe peak of THIS device, so that bench.py's roofline denominator is measured on the
// box the bench runs on (MEASURED_PEAKS.json holds HBM and bf16 only).  One CTA of 256 threads per SM, every warp issues a
// stream of DMMA.8x8x4 on 16 independent accumulator pairs: no memory traffic, no dependency stalls (DMMA latency ~26
// cycles, 16 in flight per warp, two warps per SM sub-partition).  flops = CTAs x 8 warps x iters x 16 x 512.
#include "rpc_common.cuh"

__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a0, double b0) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;            // never true: keeps the loop alive
}

extern "C" int rpc_probe_fp64_peak(rpc_ctx *ctx, void *stream, int iters, double *flops_per_launch) {
    RPC_CHECK_ARG(ctx && iters >= 1 && flops_per_launch, "null pointer or bad iteration count");
    int rc = rpc_ws_reserve(ctx, 64);
    if (rc) return rc;
    fp64_peak_kernel<<<ctx->sm_count, 256, 0, (cudaStream_t)stream>>>(ctx->ws, iters, 1.0, 1.0);
    RPC_LAUNCHED(ctx);
    *flops_per_launch = (double)ctx->sm_count * 8.0 * iters * 16.0 * 512.0;
    return 0;
}
