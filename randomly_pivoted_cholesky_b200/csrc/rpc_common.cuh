// Shared declarations of the sm_100a RPCholesky kernels (internal; the public ABI is include/rpc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/rpc_b200.h"

struct rpc_ctx {
    int device;
    int sm_count;
    int sm_reserved;   // SMs the persistent kernels leave free (rpc_ctx_reserve_sms): room for a concurrent single-CTA kernel
    int64_t launches;
    // scratch owned by the context (grown on demand, never shrunk)
    double *ws;        // general double workspace
    size_t ws_bytes;
    int32_t *flags;    // small int workspace (sampler flags etc.)
    size_t flags_bytes;
    int32_t *panel_ld_table;      // rpc_panel_ld(s) for s <= 256, for the kernels that read s on the device
    unsigned int *sync_counter;   // [0] Schur, [1] gather: CTA arrival counters of the peer-publishing kernels (self re-arming)
};

static inline int rpc_persistent_ctas(const rpc_ctx *ctx) {
    int n = ctx->sm_count - ctx->sm_reserved;
    return n < 1 ? 1 : n;
}

// "was this function configured on the context's device?" for the per-function static bit masks that guard
// cudaFuncSetAttribute (a per-device setting: a second GPU in the same process needs its own opt-in)
static inline bool rpc_dev_bit(unsigned long long mask, const rpc_ctx *ctx) { return (mask >> (ctx->device & 63)) & 1ull; }

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
        cudaError_t e_ = cudaGetLastError();                                                    \
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

// Gaussian epilogue of the DMMA evaluator: exp(x) for x = d2 * gscale <= 0, N values at a time so that the N polynomial
// chains interleave (the pair version above leaves the FP64 pipe waiting on two dependent Horner chains).  Cody-Waite
// reduction x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (remainder 4e-18 relative), scaling by 2^n through the
// exponent bits.  No special cases are needed on this domain: x <= 0, and below -708 (value < 3.3e-308, where CUDA's exp
// returns denormals) the result is flushed to 0.  <= 2 ulp from the correctly rounded value.
template <int N>
__device__ __forceinline__ void exp_neg_batch(double (&x)[N]) {
    double nf[N], r[N], pl[N];
    int ni[N];
#pragma unroll
    for (int t = 0; t < N; ++t) {
        const double xc = x[t] > -708.0 ? x[t] : -708.0;
        const double tm = fma(xc, 1.4426950408889634, 6755399441055744.0);      // round(x log2 e) in the low mantissa bits
        ni[t] = __double2loint(tm);
        nf[t] = tm - 6755399441055744.0;
        r[t] = fma(nf[t], -6.93147180369123816490e-01, xc);
        r[t] = fma(nf[t], -1.90821492927058770002e-10, r[t]);
        pl[t] = 1.6059043836821613e-10;                                           // 1/13!
    }
    const double cf[12] = {2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
                           2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03,
                           4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0};
#pragma unroll
    for (int c = 0; c < 12; ++c)
#pragma unroll
        for (int t = 0; t < N; ++t) pl[t] = fma(pl[t], r[t], cf[c]);
#pragma unroll
    for (int t = 0; t < N; ++t) {
        pl[t] = fma(pl[t], r[t], 1.0);
        const double sc = __hiloint2double((ni[t] + 1023) << 20, 0);              // 2^n, n in [-1022, 0]
        x[t] = x[t] > -708.0 ? pl[t] * sc : 0.0;
    }
}

// one row of a warp's C fragment block: 2 * NI squared distances -> Gaussian kernel values, stored as NI double2
static __device__ __noinline__ void gauss_store8(double gscale, double *dst, double d0, double d1, double d2, double d3,
                                                 double d4, double d5, double d6, double d7) {
    double x[8] = {d0 * gscale, d1 * gscale, d2 * gscale, d3 * gscale, d4 * gscale, d5 * gscale, d6 * gscale, d7 * gscale};
    exp_neg_batch<8>(x);
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<double2 *>(dst + j * 8) = make_double2(x[2 * j], x[2 * j + 1]);
}
static __device__ __noinline__ void gauss_store4(double gscale, double *dst, double d0, double d1, double d2, double d3) {
    double x[4] = {d0 * gscale, d1 * gscale, d2 * gscale, d3 * gscale};
    exp_neg_batch<4>(x);
#pragma unroll
    for (int j = 0; j < 2; ++j) *reinterpret_cast<double2 *>(dst + j * 8) = make_double2(x[2 * j], x[2 * j + 1]);
}

// ---- table-based exponential shared by the DMMA evaluator (Gaussian) and the L1 evaluator (Laplace) ---------------------
// 2^(j/32), correctly rounded
static __constant__ double c_exp2_tab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};

// exp(x) for x <= 0, N values at a time (interleaved chains), all 32 lanes of the warp converged.
//   k = round(32 x / ln2),  x = k ln2/32 + r,  |r| <= ln2/64  (two-term Cody-Waite: k * HI is exact, HI has 36 significant bits)
//   exp(x) = 2^(k >> 5) * T[k & 31] * (1 + p(r)),  p = expm1 to degree 6 (remainder < 2^-57)
// T[j] sits in lane j's register and is fetched with a warp shuffle (no shared-memory table, no bank conflicts); the power of
// two is added to the exponent field on the integer pipe.  Results below 2^-1020 are flushed to zero.  <= 1.2 ulp from the
// correctly rounded value (4e7 random arguments in [-700, 0], checked on the host: tools/exp_check.c).
template <int N>
__device__ __forceinline__ void exp_tab_batch(double (&x)[N], double tab_lane) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: the rounded integer lands in the low mantissa bits
    const double INV = 0x1.71547652b82fep+5;                  // 32 / ln2
    const double HI = 0x1.62e42fefa0000p-6, LO = 0x1.cf79abc9e3b3ap-45;   // ln2 / 32
    double r[N], t[N], q[N];
    int k[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double tm = fma(x[i], INV, MAGIC);
        k[i] = __double2loint(tm);
        const double kf = tm - MAGIC;
        r[i] = fma(kf, -HI, x[i]);
        r[i] = fma(kf, -LO, r[i]);
        q[i] = 0x1.6c16c16c16c17p-10;                        // 1/720
    }
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] = __shfl_sync(0xffffffffu, tab_lane, k[i] & 31);
    const double cf[4] = {0x1.1111111111111p-7, 0x1.5555555555555p-5, 0x1.5555555555555p-3, 0.5};   // 1/120, 1/24, 1/6, 1/2
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < N; ++i) q[i] = fma(q[i], r[i], cf[c]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double r2 = r[i] * r[i];
        const double p = fma(q[i], r2, r[i]);
        const double v = fma(t[i], p, t[i]);
        const int m = k[i] >> 5;
        const double sc = __hiloint2double(__double2hiint(v) + (m << 20), __double2loint(v));
        x[i] = m < -1020 ? 0.0 : sc;
    }
}

__device__ __forceinline__ double kern_from_l1(const KernParams &kp, double l1) { return exp(-l1 / kp.bw); }
static __device__ __noinline__ double2 kern_from_l1_pair(const KernParams kp, double a, double b) {
    return make_double2(exp(-a / kp.bw), exp(-b / kp.bw));
}

struct rpc_ctx;
// csrc/direct_eval.cu: persistent TMA-staged evaluator for the non-contraction kernels (returns -100 if the shape
// does not fit in shared memory; the caller then uses the chunked kernel)
int rpc_direct_rows_tma(rpc_ctx *ctx, cudaStream_t st, const KernParams &kp, bool l1, const double *X, int64_t ldx, int d,
                        const double *Pt, int ldpt, int s, double *out, int64_t ldo, int64_t n_pad, const double *w = nullptr,
                        int accumulate = 0);
int rpc_direct_pi_for(int s, bool l1);
// csrc/eval_panel.cu: DMMA evaluator with the pivot panel resident in shared memory (short contractions).
// rpc_eval_panel_ld: leading dimension to build the transposed panel with for m <= 256 rows, 0 if the shape does not fit.
int rpc_eval_panel_ld(int m, int kp, int64_t ldx);
int rpc_eval_panel_rows(rpc_ctx *ctx, cudaStream_t st, const KernParams &kparm, const double *X, const double *xnorm, int64_t n_pad,
                        int kp, int64_t ldx, const double *Pan, const double *pnorm, int m, double *out, int64_t ldo);
// fused kernel-times-vector variant of the panel-resident evaluator (tiles of up to 248 rows)
int rpc_eval_panel_wsum_ld(int m, int kp, int64_t ldx);
int rpc_eval_panel_wsum(rpc_ctx *ctx, cudaStream_t st, const KernParams &kparm, const double *X, const double *xnorm, int64_t n_pad,
                        int kp, int64_t ldx, const double *Pan, const double *pnorm, int m, const double *w, double *out, int accumulate);
