// Are DMMA and DFMA separate pipes on B200?  Half the warps run DMMA.8x8x4, the other half DFMA.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: all DMMA, 1: all DFMA, 2: odd warps DMMA / even warps DFMA, 3: same warp interleaves 1 DMMA : R DFMA
template <int MODE, int R>
__global__ void __launch_bounds__(512) k_mix(double *out, int iters, double a0, double b0) {
    double c[16][2], f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; f[i] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    int w = threadIdx.x >> 5;
    bool do_mma = MODE == 0 || (MODE == 2 && (w & 4));   // warps 4-7,12-15 DMMA: one of the two warps on each SMSP
    if (MODE == 3) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
                for (int r = 0; r < R; ++r) f[(i * R + r) & 15] = fma(f[(i * R + r) & 15], a, b);
            }
        }
    } else if (do_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1] + f[i];
    if (s == 123.456) out[0] = s;
}
template <class F> static float time_ms(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize()); float best = 1e30f;
    for (int r = 0; r < 5; ++r) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, 64));
    int iters = 4000, T = 256, grid = sms; int warps = T / 32;
    float ms;
    printf("{");
    ms = time_ms([&] { k_mix<0, 0><<<grid, T>>>(out, iters, 1.0, 1.0); });
    printf("\"all_dmma_tflops\": %.2f", (double)grid * warps * iters * 16.0 * 512 / ms / 1e9);
    ms = time_ms([&] { k_mix<1, 0><<<grid, T>>>(out, iters, 1.0, 1.0); });
    printf(", \"all_dfma_tflops\": %.2f", (double)grid * warps * iters * 16.0 * 64 / ms / 1e9);
    ms = time_ms([&] { k_mix<2, 0><<<grid, T>>>(out, iters, 1.0, 1.0); });
    printf(", \"split_warps_ms\": %.3f, \"split_dmma_part_tflops\": %.2f, \"split_dfma_part_tflops\": %.2f", ms,
           (double)grid * (warps / 2) * iters * 16.0 * 512 / ms / 1e9, (double)grid * (warps / 2) * iters * 16.0 * 64 / ms / 1e9);
    ms = time_ms([&] { k_mix<3, 2><<<grid, T>>>(out, iters, 1.0, 1.0); });
    printf(", \"interleave_1to2_tflops\": %.2f", (double)grid * warps * iters * 16.0 * (512 + 2 * 64) / ms / 1e9);
    ms = time_ms([&] { k_mix<3, 4><<<grid, T>>>(out, iters, 1.0, 1.0); });
    printf(", \"interleave_1to4_tflops\": %.2f", (double)grid * warps * iters * 16.0 * (512 + 4 * 64) / ms / 1e9);
    ms = time_ms([&] { k_mix<3, 8><<<grid, T>>>(out, iters, 1.0, 1.0); });
    printf(", \"interleave_1to8_tflops\": %.2f", (double)grid * warps * iters * 16.0 * (512 + 8 * 64) / ms / 1e9);
    ms = time_ms([&] { k_mix<3, 16><<<grid, T>>>(out, iters, 1.0, 1.0); });
    printf(", \"interleave_1to16_tflops\": %.2f", (double)grid * warps * iters * 16.0 * (512 + 16 * 64) / ms / 1e9);
    printf("}\n");
    return 0;
}
