// Shared declarations of the sm_100a RPCholesky kernels (internal; the public ABI is include/rpc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/rpc_b200.h"

struct rpc_ctx {
    int device;
    int sm_count;
    int64_t launches;
    // scratch owned by the context (grown on demand, never shrunk)
    double *ws;        // general double workspace
    size_t ws_bytes;
    int32_t *flags;    // small int workspace (sampler flags etc.)
    size_t flags_bytes;
    int32_t *panel_ld_table;      // rpc_panel_ld(s) for s <= 256, for the kernels that read s on the device
    unsigned int *sync_counter;   // [0] Schur, [1] gather: CTA arrival counters of the peer-publishing kernels (self re-arming)
};

void rpc_set_error(const char *fmt, ...);
int rpc_ws_reserve(rpc_ctx *ctx, size_t bytes);
int rpc_flags_reserve(rpc_ctx *ctx, size_t bytes);
int rpc_sync_reserve(rpc_ctx *ctx);
// out (k_pad x ldo) = zero-padded transpose of in (m x d)
int rpc_transpose_pad(rpc_ctx *ctx, cudaStream_t st, const double *in, int64_t ldi, int m, int d, double *out, int64_t ldo,
                      int k_pad);

#define RPC_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            rpc_set_error("%s: bad argument: %s", __func__, msg);  \
            return -1;                                             \
        }                                                          \
    } while (0)

#define RPC_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            rpc_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_));        \
            return (int)e_;                                                                     \
        }                                                                                       \
    } while (0)

// every kernel launch goes through this so the context can count them
#define RPC_LAUNCHED(ctx)                                                                       \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e_ = cudaPeekAtLastError();                                                 \
        if (e_ != cudaSuccess) {                                                                \
            rpc_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_));    \
            return (int)e_;                                                                     \
        }                                                                                       \
    } while (0)

// ---- kernel function on a squared Euclidean distance / L1 distance ---------------------------------
struct KernParams {
    int kind;      // RPC_KERN_*
    int nu2;       // Matern 2*nu
    int stab;      // extra_stability
    double bw;     // bandwidth
    double gscale; // Gaussian: -0.5 / bw^2
};

__host__ inline KernParams make_kern_params(const rpc_kernel_spec *s) {
    KernParams p;
    p.kind = s->kind; p.nu2 = s->nu2; p.stab = s->extra_stability; p.bw = s->bandwidth;
    p.gscale = -0.5 / (s->bandwidth * s->bandwidth);
    return p;
}

// Value of the kernel from the clamped squared distance d2 (expanded or direct form).
// Reference: dsts = sqrt(d2); Gaussian exp(-0.5*dsts**2/bw**2) (kernels.py:39); Matern r = dsts/bw, or
// r = dsts when extra_stability (the reference omits /bw there, kernels.py:74-75).
// The Gaussian skips the sqrt/re-square round trip and the division (exp(d2 * (-0.5/bw^2))): the argument
// differs from the reference's by <= 3 ulp, i.e. the value by <= 3e-16*|arg| relative -- five orders below
// the 1e-10 parity tolerance -- and it removes ~40 FP64-pipe instructions per entry from the epilogue.
__device__ __forceinline__ double kern_from_d2(const KernParams &kp, double d2) {
    if (kp.kind == RPC_KERN_GAUSSIAN) {
        return exp(d2 * kp.gscale);
    }
    double dst = sqrt(d2);
    double r = kp.stab ? dst : dst / kp.bw;
    if (kp.nu2 == 1) return exp(-r);
    if (kp.nu2 == 3) {
        const double s3 = 1.7320508075688772;  // np.sqrt(3)
        return (1.0 + s3 * r) * exp(-s3 * r);
    }
    const double s5 = 2.23606797749979;        // np.sqrt(5)
    return (1.0 + s5 * r + 5.0 / 3.0 * r * r) * exp(-s5 * r);
}

// Out-of-line pair version for the unrolled DMMA epilogues: inlining exp() at all 64 accumulator sites of a
// thread made the epilogue ~60 KB of SASS, larger than the instruction cache, and 3x slower than the math.
static __device__ __noinline__ double2 kern_from_d2_pair(const KernParams kp, double d0, double d1) {
    return make_double2(kern_from_d2(kp, d0), kern_from_d2(kp, d1));
}

__device__ __forceinline__ double kern_from_l1(const KernParams &kp, double l1) { return exp(-l1 / kp.bw); }
static __device__ __noinline__ double2 kern_from_l1_pair(const KernParams kp, double a, double b) {
    return make_double2(exp(-a / kp.bw), exp(-b / kp.bw));
}

struct rpc_ctx;
// csrc/direct_eval.cu: persistent TMA-staged evaluator for the non-contraction kernels (returns -100 if the shape
// does not fit in shared memory; the caller then uses the chunked kernel)
int rpc_direct_rows_tma(rpc_ctx *ctx, cudaStream_t st, const KernParams &kp, bool l1, const double *X, int64_t ldx, int d,
                        const double *Pt, int ldpt, int s, double *out, int64_t ldo, int64_t n_pad);
int rpc_direct_pi_for(int s, bool l1);
