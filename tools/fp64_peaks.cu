// FP64 peak micro-benchmark for B200 (sm_100a): DMMA.8x8x4 vs DFMA vs DADD latency.
// Writes one JSON object to stdout. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double *out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double *out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

// dependent DADD chain, one thread: latency in cycles
__global__ void k_dadd_lat(double *out, long long *cyc, int iters, double inc) {
    double c = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 16
        for (int i = 0; i < 16; ++i) c += inc;
    }
    long long t1 = clock64();
    out[0] = c; cyc[0] = t1 - t0;
}
// dependent DMMA chain (one warp): latency in cycles
__global__ void k_dmma_lat(double *out, long long *cyc, int iters, double a, double b) {
    double c0 = 0, c1 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll 16
        for (int i = 0; i < 16; ++i) dmma884(c0, c1, a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = c0 + c1; cyc[0] = t1 - t0; }
}

template <class F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double *out; long long *cyc; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    int iters = 4000;
    const int thr[3] = {128, 256, 512};
    for (int t = 0; t < 3; ++t) {
        for (int occ = 1; occ <= 2; ++occ) {
            int grid = sms * occ;
            float ms = time_ms([&] { k_dmma<16><<<grid, thr[t]>>>(out, iters, 1.0, 1.0); }, 5);
            double flops = (double)grid * (thr[t] / 32) * iters * 16.0 * 512.0;
            printf(", \"dmma_t%d_o%d_tflops\": %.2f", thr[t], occ, flops / ms / 1e9);
            ms = time_ms([&] { k_dfma<16><<<grid, thr[t]>>>(out, iters, 1.0, 1.0); }, 5);
            flops = (double)grid * thr[t] * iters * 16.0 * 2.0;
            printf(", \"dfma_t%d_o%d_tflops\": %.2f", thr[t], occ, flops / ms / 1e9);
        }
    }
    // sustained (about 2 s) DMMA to see the power-capped steady state
    {
        int grid = sms * 2; cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0; for (; n < 400; ++n) k_dmma<16><<<grid, 256>>>(out, iters * 4, 1.0, 1.0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)n * grid * 8 * (iters * 4) * 16.0 * 512.0;
        printf(", \"dmma_sustained_tflops\": %.2f, \"dmma_sustained_s\": %.2f", flops / ms / 1e9, ms / 1e3);
    }
    long long h;
    k_dadd_lat<<<1, 1>>>(out, cyc, 1000, 1e-3); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"dadd_latency_cyc\": %.2f", (double)h / 16000.0);
    k_dmma_lat<<<1, 32>>>(out, cyc, 1000, 1.0, 1.0); CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf(", \"dmma_latency_cyc\": %.2f", (double)h / 16000.0);
    printf("}\n");
    return 0;
}
