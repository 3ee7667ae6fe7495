// Per-round small dense work and operator helpers: row norms, diagonal, dense-operator gathers,
// pivot gather, b x b kernel block, residual Gram, rejection / shifted Cholesky, panel assembly.
#include "rpc_common.cuh"

// ------------------------------------------------------------------------------------------------
// row norms (warp per point), diagonal fill
// ------------------------------------------------------------------------------------------------
__global__ void row_norms_kernel(const double *__restrict__ X, int64_t n, int d, int64_t ldx, double *__restrict__ out) {
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    double acc = 0.0;
    for (int f = lane; f < d; f += 32) { double v = X[row * ldx + f]; acc = fma(v, v, acc); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc;
}

extern "C" int rpc_row_norms(rpc_ctx *ctx, void *stream, const double *X, int64_t n, int d, int64_t ldx, double *xnorm) {
    RPC_CHECK_ARG(ctx && X && xnorm && n > 0 && d > 0, "null pointer or empty");
    unsigned grid = (unsigned)((n + 7) / 8);
    row_norms_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, n, d, ldx, xnorm);
    RPC_LAUNCHED(ctx);
    return 0;
}

__global__ void fill_diag_kernel(double *diag, int64_t n, int64_t n_pad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) diag[i] = i < n ? 1.0 : 0.0;
}

extern "C" int rpc_kernel_diag(rpc_ctx *ctx, void *stream, double *diag, int64_t n, int64_t n_pad) {
    RPC_CHECK_ARG(ctx && diag && n > 0 && n_pad >= n, "null pointer or bad sizes");
    fill_diag_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(diag, n, n_pad);
    RPC_LAUNCHED(ctx);
    return 0;
}

// The padding contract of include/rpc_b200.h (pad points all-zero, pad features zero, pad diagonal entries 0) is what lets the
// N-proportional kernels run over n_pad columns without edge checks.  This counts its violations (NaN counts as one).
__global__ void __launch_bounds__(256) check_padding_kernel(const double *__restrict__ X, int64_t n, int64_t n_pad, int d, int64_t ldx,
                                                            const double *__restrict__ diag, int32_t *__restrict__ bad) {
    int local = 0;
    const int64_t fp = ldx - d;                                       // pad features per point
    const int64_t nfeat = n * fp, nrow = (n_pad - n) * ldx, ndiag = diag ? n_pad - n : 0;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < nfeat + nrow + ndiag; q += (int64_t)gridDim.x * 256) {
        double v;
        if (q < nfeat) v = X[(q / fp) * ldx + d + q % fp];           // feature pad of a real point
        else if (q < nfeat + nrow) v = X[n * ldx + (q - nfeat)];     // a pad point
        else v = diag[n + (q - nfeat - nrow)];                       // pad entries of the residual diagonal
        local += !(v == 0.0);
    }
    if (local) atomicAdd(bad, local);
}

extern "C" int rpc_check_padding(rpc_ctx *ctx, void *stream, const double *X, int64_t n, int64_t n_pad, int d, int64_t ldx,
                                 const double *diag, int32_t *violations_host) {
    RPC_CHECK_ARG(ctx && X && violations_host, "null pointer");
    RPC_CHECK_ARG(n >= 0 && n_pad >= n && d >= 1 && ldx >= d, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = rpc_flags_reserve(ctx, sizeof(int32_t));
    if (rc) return rc;
    RPC_CUDA(cudaMemsetAsync(ctx->flags, 0, sizeof(int32_t), st));
    const int64_t total = n * (ldx - d) + (n_pad - n) * ldx + (diag ? n_pad - n : 0);
    if (total > 0) {
        int grid = (int)((total + 255) / 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        check_padding_kernel<<<grid, 256, 0, st>>>(X, n, n_pad, d, ldx, diag, ctx->flags);
        RPC_LAUNCHED(ctx);
    }
    RPC_CUDA(cudaMemcpyAsync(violations_host, ctx->flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    RPC_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// out (k_pad x ldo) = zero-padded transpose of in (m x d): out[f][j] = in[j][f]
__global__ void transpose_pad_kernel(const double *__restrict__ in, int64_t ldi, int m, int d, double *__restrict__ out,
                                     int64_t ldo, int k_pad) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (int64_t)k_pad * ldo) return;
    int f = (int)(q / ldo), j = (int)(q - (int64_t)f * ldo);
    out[q] = (f < d && j < m) ? in[(int64_t)j * ldi + f] : 0.0;
}

int rpc_transpose_pad(rpc_ctx *ctx, cudaStream_t st, const double *in, int64_t ldi, int m, int d, double *out, int64_t ldo,
                      int k_pad) {
    int64_t tot = (int64_t)k_pad * ldo;
    transpose_pad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(in, ldi, m, d, out, ldo, k_pad);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// dense PSDMatrix operator (matrix.py:86-102)
// ------------------------------------------------------------------------------------------------
__global__ void psd_diag_kernel(const double *__restrict__ A, int64_t n, int64_t lda, double *__restrict__ diag, int64_t n_pad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) diag[i] = i < n ? A[i * lda + i] : 0.0;
}
extern "C" int rpc_psd_diag(rpc_ctx *ctx, void *stream, const double *A, int64_t n, int64_t lda, double *diag, int64_t n_pad) {
    RPC_CHECK_ARG(ctx && A && diag && n > 0 && n_pad >= n, "null pointer or bad sizes");
    psd_diag_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, n, lda, diag, n_pad);
    RPC_LAUNCHED(ctx);
    return 0;
}

__global__ void psd_rows_kernel(const double *__restrict__ A, int64_t n, int64_t lda, const int64_t *__restrict__ idx,
                                double *__restrict__ out, int64_t ldo, int64_t n_pad) {
    const int64_t src = idx[blockIdx.y];
    for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < n_pad; col += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)blockIdx.y * ldo + col] = col < n ? A[src * lda + col] : 0.0;
}
extern "C" int rpc_psd_rows(rpc_ctx *ctx, void *stream, const double *A, int64_t n, int64_t lda, const int64_t *idx, int s,
                            double *out, int64_t ldo, int64_t n_pad) {
    RPC_CHECK_ARG(ctx && A && idx && out && s >= 1 && n_pad >= n, "null pointer or bad sizes");
    unsigned gx = (unsigned)((n_pad + 255) / 256);
    if (gx > 1024) gx = 1024;
    psd_rows_kernel<<<dim3(gx, s), 256, 0, (cudaStream_t)stream>>>(A, n, lda, idx, out, ldo, n_pad);
    RPC_LAUNCHED(ctx);
    return 0;
}

__global__ void psd_block_kernel(const double *__restrict__ A, int64_t lda, const int64_t *__restrict__ idx, int b,
                                 double *__restrict__ H, int64_t ldh) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= b * b) return;
    int i = q / b, j = q - i * b;
    H[(int64_t)i * ldh + j] = A[idx[i] * lda + idx[j]];
}
extern "C" int rpc_psd_block(rpc_ctx *ctx, void *stream, const double *A, int64_t lda, const int64_t *idx, int b, double *H,
                             int64_t ldh) {
    RPC_CHECK_ARG(ctx && A && idx && H && b >= 1, "null pointer or bad sizes");
    psd_block_kernel<<<(b * b + 255) / 256, 256, 0, (cudaStream_t)stream>>>(A, lda, idx, b, H, ldh);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// b x b kernel block A[idx,idx] (same formulas as the row evaluator)
// ------------------------------------------------------------------------------------------------
__global__ void kernel_block_kernel(KernParams kp, const double *__restrict__ Xp, const double *__restrict__ pn, int b, int d,
                                    int64_t ldp, double *__restrict__ H, int64_t ldh) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= b * b) return;
    int i = q / b, j = q - i * b;
    const double *x = Xp + (int64_t)i * ldp, *y = Xp + (int64_t)j * ldp;
    double v;
    if (kp.kind == RPC_KERN_LAPLACE) {
        double l1 = 0.0;
        for (int f = 0; f < d; ++f) l1 += fabs(x[f] - y[f]);
        v = kern_from_l1(kp, l1);
    } else if (kp.stab) {
        double d2 = 0.0;
        for (int f = 0; f < d; ++f) { double df = x[f] - y[f]; d2 = fma(df, df, d2); }
        v = kern_from_d2(kp, d2);
    } else {
        double dot = 0.0;
        for (int f = 0; f < d; ++f) dot = fma(x[f], y[f], dot);
        double d2 = (-2.0 * dot + pn[i]) + pn[j];
        v = kern_from_d2(kp, d2 > 0.0 ? d2 : 0.0);
    }
    H[(int64_t)i * ldh + j] = v;
}

extern "C" int rpc_kernel_block(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *Xpiv,
                                const double *pnorm, int b, int d, int64_t ldp, double *H, int64_t ldh) {
    RPC_CHECK_ARG(ctx && spec && Xpiv && H && b >= 1 && d >= 1, "null pointer or bad sizes");
    RPC_CHECK_ARG(spec->kind == RPC_KERN_LAPLACE || spec->extra_stability || pnorm, "pivot norms required");
    kernel_block_kernel<<<(b * b + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_kern_params(spec), Xpiv, pnorm, b, d, ldp,
                                                                             H, ldh);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// pivot gather: Xpiv = X[idx], pnorm = xnorm[idx], P = G[:c, idx]
// ------------------------------------------------------------------------------------------------
__global__ void gather_pivots_kernel(const int64_t *__restrict__ idx, int m, const double *__restrict__ X,
                                     const double *__restrict__ xnorm, int d, int64_t ldx, double *__restrict__ Xpiv,
                                     double *__restrict__ pnorm, int64_t ldp, const double *__restrict__ G, int64_t ldg, int c,
                                     double *__restrict__ P, int64_t ldP, int64_t lo, int64_t hi) {
    int blk = blockIdx.x;
    if (blk < c) {
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const int64_t g = idx[j];
            P[(int64_t)blk * ldP + j] = (g >= lo && g < hi) ? G[(int64_t)blk * ldg + (g - lo)] : 0.0;
        }
    } else {
        int j = blk - c;
        const int64_t g = idx[j];
        const bool mine = g >= lo && g < hi;
        const int64_t src = g - lo;
        for (int f = threadIdx.x; f < ldp; f += blockDim.x) Xpiv[(int64_t)j * ldp + f] = (mine && f < d) ? X[src * ldx + f] : 0.0;
        if (threadIdx.x == 0 && xnorm && pnorm) pnorm[j] = mine ? xnorm[src] : 0.0;
    }
}

extern "C" int rpc_gather_pivots(rpc_ctx *ctx, void *stream, const int64_t *idx, int m, const double *X,
                                 const double *xnorm, int d, int64_t ldx, double *Xpiv, double *pnorm, int64_t ldp,
                                 const double *G, int64_t ldg, int c, double *P, int64_t ldP, int64_t lo, int64_t hi) {
    RPC_CHECK_ARG(ctx && idx && m >= 1 && c >= 0, "null pointer or bad sizes");
    RPC_CHECK_ARG(c == 0 || (G && P), "G and P required when c > 0");
    RPC_CHECK_ARG(!X || Xpiv, "Xpiv required when X is given");
    RPC_CHECK_ARG(hi >= lo, "empty ownership range");
    int blocks = c + (X ? m : 0);
    if (blocks == 0) return 0;
    gather_pivots_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(idx, m, X, xnorm, d, ldx, Xpiv, pnorm, ldp, G, ldg, c, P,
                                                                  ldP, lo, hi);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// residual Gram: H -= P^T P.  32x32 output tile per CTA, split over K in chunks of GR_KC rows with
// partial sums in the workspace, then a fixed-order reduction (deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
constexpr int GR_T = 32;
constexpr int GR_KC = 128;

__global__ void __launch_bounds__(256) gram_partial_kernel(const double *__restrict__ P, int c, int m, int64_t ldP,
                                                           double *__restrict__ part) {
    __shared__ double Pi[GR_T][GR_T + 1];
    __shared__ double Pj[GR_T][GR_T + 1];
    const int i0 = blockIdx.y * GR_T, j0 = blockIdx.x * GR_T;
    const int k0 = blockIdx.z * GR_KC;
    const int k1 = min(c, k0 + GR_KC);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 2 x 2 outputs each
    double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
    for (int kb = k0; kb < k1; kb += GR_T) {
        __syncthreads();
        for (int q = threadIdx.x; q < GR_T * GR_T; q += 256) {
            int kk = q / GR_T, col = q - kk * GR_T;
            int k = kb + kk;
            Pi[kk][col] = (k < k1 && i0 + col < m) ? P[(int64_t)k * ldP + i0 + col] : 0.0;
            Pj[kk][col] = (k < k1 && j0 + col < m) ? P[(int64_t)k * ldP + j0 + col] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < GR_T; ++kk) {
            double x0 = Pi[kk][ty * 2], x1 = Pi[kk][ty * 2 + 1];
            double y0 = Pj[kk][tx * 2], y1 = Pj[kk][tx * 2 + 1];
            a00 = fma(x0, y0, a00); a01 = fma(x0, y1, a01);
            a10 = fma(x1, y0, a10); a11 = fma(x1, y1, a11);
        }
    }
    double *dst = part + (size_t)blockIdx.z * m * m;
    int i = i0 + ty * 2, j = j0 + tx * 2;
    if (i < m && j < m) dst[(size_t)i * m + j] = a00;
    if (i < m && j + 1 < m) dst[(size_t)i * m + j + 1] = a01;
    if (i + 1 < m && j < m) dst[(size_t)(i + 1) * m + j] = a10;
    if (i + 1 < m && j + 1 < m) dst[(size_t)(i + 1) * m + j + 1] = a11;
}

__global__ void gram_finish_kernel(const double *__restrict__ part, int nsplit, int m, double *__restrict__ H, int64_t ldh) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m * m) return;
    int i = q / m, j = q - i * m;
    double s = 0.0;
    for (int z = 0; z < nsplit; ++z) s += part[(size_t)z * m * m + q];
    H[(int64_t)i * ldh + j] -= s;
}

extern "C" int rpc_gram_residual(rpc_ctx *ctx, void *stream, const double *P, int c, int m, int64_t ldP, double *H,
                                 int64_t ldh) {
    RPC_CHECK_ARG(ctx && H && m >= 1 && c >= 0, "null pointer or bad sizes");
    if (c == 0) return 0;
    RPC_CHECK_ARG(P != nullptr, "P required when c > 0");
    cudaStream_t st = (cudaStream_t)stream;
    int nsplit = (c + GR_KC - 1) / GR_KC;
    int rc = rpc_ws_reserve(ctx, (size_t)nsplit * m * m * sizeof(double));
    if (rc) return rc;
    int t = (m + GR_T - 1) / GR_T;
    gram_partial_kernel<<<dim3(t, t, nsplit), 256, 0, st>>>(P, c, m, ldP, ctx->ws);
    RPC_LAUNCHED(ctx);
    gram_finish_kernel<<<(m * m + 255) / 256, 256, 0, st>>>(ctx->ws, nsplit, m, H, ldh);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// rejection_cholesky (rpcholesky.py:140-157) / shifted Cholesky (rpcholesky.py:106): ONE CTA, the
// lower triangle of H packed in shared memory when it fits (m <= 232), right-looking.
// ------------------------------------------------------------------------------------------------
constexpr int RC_THREADS = 512;
constexpr int RC_SMEM_MAX_M = 224;

template <bool PACKED>
__device__ __forceinline__ size_t tri_at(int i, int l, int64_t ld) {
    return PACKED ? (size_t)i * (i + 1) / 2 + l : (size_t)i * ld + l;
}

// Blocked right-looking formulation, RB columns at a time:
//   step 1 (warp 0): the RB x RB diagonal block is eliminated in registers, lane r holding row r: the b sequential
//           accept tests u_j H0[j,j] < H[j,j] (rpcholesky.py:151), sqrt and column scaling never touch a barrier;
//   step 2 (all threads): rows below the block get their entries of the accepted columns by a tiny forward
//           substitution against the block;
//   step 3 (all warps): ONE rank-(#accepted) update of the trailing lower triangle.
// Three barriers per RB columns instead of two per column, and RB FMAs per trailing-matrix access.
constexpr int RB = 8;

template <bool PACKED>
__global__ void __launch_bounds__(RC_THREADS, 1) rejection_cholesky_kernel(double *__restrict__ H, int m, int64_t ldh,
                                                                          const double *__restrict__ u, int max_accept,
                                                                          int mode, double shift, double *__restrict__ Lfull,
                                                                          double *__restrict__ L, int64_t ldl,
                                                                          int32_t *__restrict__ acc, int32_t *__restrict__ info) {
    extern __shared__ double sh[];
    double *thr = sh;                                  // [m] u[j] * original diagonal: acceptance threshold of step j
    double *Lb = sh + m;                               // [RB][m] entries of the block's accepted columns, all rows
    int32_t *accpos = (int32_t *)(sh + m + RB * m);    // [m] accepted positions so far
    int32_t *blk = accpos + m;                         // [0] #accepted in the block, [1] stop flag, [2..2+RB) their columns
    double *rpiv = sh + m + RB * m + (m + 16) / 2 + 8; // [RB] 1/sqrt(pivot) of the block's accepted columns
    double *T = PACKED ? rpiv + RB : H;
    const int64_t ld = ldh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = RC_THREADS / 32;
    const unsigned full = 0xffffffffu;

    if (PACKED) {
        // lower triangle only, a warp per row (coalesced, no index division), four rows in flight per warp
        for (int i0 = warp * 4; i0 < m; i0 += NW * 4) {
            for (int l = lane; l <= i0 + 3; l += 32) {
                double v[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) v[w] = (i0 + w < m && l <= i0 + w) ? H[(int64_t)(i0 + w) * ldh + l] : 0.0;
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    if (i0 + w < m && l <= i0 + w) T[tri_at<true>(i0 + w, l, ld)] = v[w];
            }
        }
    }
    // trace check (np.trace(H) <= 0 -> RuntimeError): only the sign matters, so the summation order does not
    if (warp == 0) {
        double tr = 0.0;
        for (int j = lane; j < m; j += 32) tr += H[(int64_t)j * ldh + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(full, tr, o);
        if (lane == 0) {
            info[2] = (mode == 0 && !(tr > 0.0)) ? 1 : 0;
            info[1] = 0;
        }
    }
    for (int j = tid; j < m; j += RC_THREADS) thr[j] = (mode == 0 ? u[j] : 0.0) * H[(int64_t)j * ldh + j];
    __syncthreads();
    if (mode == 1 && shift != 0.0) {
        for (int j = tid; j < m; j += RC_THREADS) T[tri_at<PACKED>(j, j, ld)] += shift;
    }
    __syncthreads();

    int nacc = 0;
    for (int j0 = 0; j0 < m; j0 += RB) {
        const int nb = min(RB, m - j0);
        // ---- step 1: eliminate the diagonal block in registers (warp 0; lane r < nb holds row j0 + r)
        if (warp == 0) {
            const int r = lane;
            double drow[RB], lrow[RB];
#pragma unroll
            for (int cc = 0; cc < RB; ++cc) {
                drow[cc] = (r < nb && cc <= r) ? T[tri_at<PACKED>(j0 + r, j0 + cc, ld)] : 0.0;
                lrow[cc] = 0.0;
            }
            int na = 0, stop = 0, cols[RB];
            double rinvs[RB], thr8[RB];
#pragma unroll
            for (int jj = 0; jj < RB; ++jj) thr8[jj] = jj < nb ? thr[j0 + jj] : 0.0;    // off the serial chain
#pragma unroll
            for (int jj = 0; jj < RB; ++jj) {
                if (jj < nb && !stop) {                                   // warp-uniform
                    const double hjj = __shfl_sync(full, drow[jj], jj);
                    bool accept;
                    if (mode == 1) {
                        accept = true;
                        if (!(hjj > 0.0)) {                               // numpy.linalg.cholesky would raise LinAlgError
                            if (lane == 0) info[1] = j0 + jj + 1;
                            stop = 1;
                            accept = false;
                        }
                    } else {
                        accept = thr8[jj] < hjj;                          // rpcholesky.py:151
                    }
                    if (accept && nacc + na == max_accept) { stop = 1; accept = false; }   // truncation, :193-196
                    if (accept) {
                        // L[:,j] = H[:,j] / sqrt(H[j,j]) as a multiplication by rsqrt (1-2 ulp from the division;
                        // sqrt + divide are ~300 dependent cycles on the one serial chain of the whole factorisation)
                        const double rinv = rsqrt(hjj);
                        const double lv = (r >= jj && r < nb) ? (r == jj ? hjj * rinv : drow[jj] * rinv) : 0.0;
#pragma unroll
                        for (int cc = jj + 1; cc < RB; ++cc) {
                            const double lcc = __shfl_sync(full, lv, cc);
                            if (cc <= r) drow[cc] = drow[cc] - lv * lcc;
                        }
#pragma unroll
                        for (int a = 0; a < RB; ++a) if (a == na) { lrow[a] = lv; cols[a] = j0 + jj; rinvs[a] = rinv; }
                        ++na;
                    }
                }
            }
            if (r < nb) {
#pragma unroll
                for (int a = 0; a < RB; ++a)
                    if (a < na) {
                        Lb[a * m + j0 + r] = lrow[a];
                        Lfull[(size_t)(j0 + r) * m + nacc + a] = lrow[a];
                    }
            }
            if (lane == 0) {
                blk[0] = na;
                blk[1] = stop;
#pragma unroll
                for (int a = 0; a < RB; ++a) if (a < na) { blk[2 + a] = cols[a]; accpos[nacc + a] = cols[a]; rpiv[a] = rinvs[a]; }
            }
        }
        __syncthreads();
        const int na = blk[0];
        const int stop = blk[1];
        if (na > 0 && !stop) {
            // ---- step 2: rows below the block, one thread per row: l[a] = (T[i, col_a] - sum_{a'<a} l[a'] L[col_a, a']) / L[col_a, a]
            for (int i = j0 + nb + tid; i < m; i += RC_THREADS) {
                double l[RB];
#pragma unroll
                for (int a = 0; a < RB; ++a) {
                    l[a] = 0.0;
                    if (a < na) {
                        const int ca = blk[2 + a];
                        double v = T[tri_at<PACKED>(i, ca, ld)];
#pragma unroll
                        for (int a2 = 0; a2 < RB; ++a2)
                            if (a2 < a) v = v - l[a2] * Lb[a2 * m + ca];
                        l[a] = v * rpiv[a];
                        Lb[a * m + i] = l[a];
                        Lfull[(size_t)i * m + nacc + a] = l[a];
                    }
                }
            }
            __syncthreads();
            // ---- step 3: trailing lower triangle T[i,l] -= sum_a L[i,a] L[l,a], j0+nb <= l <= i.
            if (PACKED && nb == RB) {
                // FP64 tensor cores: the rank-8 update of an 8 x 8 tile is two DMMA.8x8x4 (A = -L[i-block, 0:8],
                // B = L[l-block, 0:8]^T), i.e. 14 instructions per 64 entries instead of ~70.  A warp owns whole
                // block rows ("strips", A fragments loaded once per strip); strips are dealt longest-first in snake
                // order so every warp gets about the same number of tiles.
                const int t0 = (j0 + RB) >> 3, mt = (m + 7) >> 3, nst = mt - t0;
                const int r = lane >> 2, kq = lane & 3;
                for (int pass = 0; pass < 2; ++pass) {
                    const int si = pass == 0 ? warp : 2 * NW - 1 - warp;          // strip rank by length (0 = longest)
                    if (si >= nst) continue;
                    const int I = mt - 1 - si, i0 = I << 3, i = i0 + r;
                    const bool iok = i < m;
                    const double a0 = (iok && kq < na) ? -Lb[kq * m + i] : 0.0;
                    const double a1 = (iok && 4 + kq < na) ? -Lb[(4 + kq) * m + i] : 0.0;
                    const size_t rowbase = (size_t)i * (i + 1) / 2;
#pragma unroll 4
                    for (int J = t0; J <= I; ++J) {
                        const int l0 = J << 3, lb = l0 + r, l = l0 + 2 * kq;
                        const bool lok = lb < m;
                        const double b0 = (lok && kq < na) ? Lb[kq * m + lb] : 0.0;
                        const double b1 = (lok && 4 + kq < na) ? Lb[(4 + kq) * m + lb] : 0.0;
                        const bool v0 = iok && l <= i, v1 = iok && l + 1 <= i;
                        double cfrag[2];
                        cfrag[0] = v0 ? T[rowbase + l] : 0.0;
                        cfrag[1] = v1 ? T[rowbase + l + 1] : 0.0;
                        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                            : "+d"(cfrag[0]), "+d"(cfrag[1]) : "d"(a0), "d"(b0));
                        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                            : "+d"(cfrag[0]), "+d"(cfrag[1]) : "d"(a1), "d"(b1));
                        if (v0) T[rowbase + l] = cfrag[0];
                        if (v1) T[rowbase + l + 1] = cfrag[1];
                    }
                }
            } else
            // a warp takes two rows at a time (two independent FMA chains per L[l,a] load)
            for (int i = j0 + nb + 2 * warp; i < m; i += 2 * NW) {
                const bool two = i + 1 < m;
                double li0[RB], li1[RB];
#pragma unroll
                for (int a = 0; a < RB; ++a) {
                    li0[a] = a < na ? Lb[a * m + i] : 0.0;
                    li1[a] = (a < na && two) ? Lb[a * m + i + 1] : 0.0;
                }
                const int lmax = two ? i + 1 : i;
                for (int l = j0 + nb + lane; l <= lmax; l += 32) {
                    const bool in0 = l <= i, in1 = two;            // row i+1 holds columns up to i+1
                    const size_t at0 = tri_at<PACKED>(i, in0 ? l : i, ld);
                    const size_t at1 = tri_at<PACKED>(two ? i + 1 : i, l <= i + 1 ? l : i, ld);
                    double t0 = in0 ? T[at0] : 0.0, t1 = in1 ? T[at1] : 0.0;
#pragma unroll
                    for (int a = 0; a < RB; ++a)
                        if (a < na) {
                            const double lb = Lb[a * m + l];
                            t0 = t0 - li0[a] * lb;
                            t1 = t1 - li1[a] * lb;
                        }
                    if (in0) T[at0] = t0;
                    if (in1) T[at1] = t1;
                }
            }
        }
        nacc += na;
        __syncthreads();
        if (stop) break;
    }
    __syncthreads();
    // L = Lfull[acc, acc] (rpcholesky.py:155-156)
    // a warp per four rows (four independent loads in flight): no index division, no loads above the diagonal
    for (int r0 = warp * 4; r0 < nacc; r0 += NW * 4) {
        const double *src[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) src[w] = Lfull + (size_t)accpos[r0 + w < nacc ? r0 + w : r0] * m;
        for (int a = lane; a < nacc; a += 32) {
            double v[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) v[w] = (r0 + w < nacc && a <= r0 + w) ? src[w][a] : 0.0;
#pragma unroll
            for (int w = 0; w < 4; ++w)
                if (r0 + w < nacc) L[(int64_t)(r0 + w) * ldl + a] = v[w];
        }
    }
    for (int q = tid; q < nacc; q += RC_THREADS) acc[q] = accpos[q];
    if (tid == 0) info[0] = nacc;
}

extern "C" int rpc_rejection_cholesky(rpc_ctx *ctx, void *stream, double *H, int m, int64_t ldh, const double *u,
                                      int max_accept, int mode, double shift, double *L, int64_t ldl, int32_t *acc,
                                      int32_t *info) {
    RPC_CHECK_ARG(ctx && H && L && acc && info, "null pointer");
    RPC_CHECK_ARG(m >= 1 && ldh >= m && ldl >= m, "bad sizes");
    RPC_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (rejection) or 1 (shifted Cholesky)");
    RPC_CHECK_ARG(mode == 1 || u != nullptr, "uniforms required for rejection sampling");
    if (max_accept > m || max_accept < 0) max_accept = m;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = rpc_ws_reserve(ctx, (size_t)m * m * sizeof(double));
    if (rc) return rc;
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(rejection_cholesky_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      220 * 1024));
        RPC_CUDA(cudaFuncSetAttribute(rejection_cholesky_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      220 * 1024));
        configured |= 1ull << (ctx->device & 63);
    }
    // thr[m] | Lb[RB][m] | accpos[m] + blk ints | (packed lower triangle)
    const size_t head = ((size_t)m + RB * m + (m + 16) / 2 + 8 + RB) * sizeof(double);
    RPC_CHECK_ARG(m <= RPC_MAX_BLOCK, "block larger than RPC_MAX_BLOCK");
    if (m <= RC_SMEM_MAX_M) {
        size_t shm = head + (size_t)m * (m + 1) / 2 * sizeof(double);
        rejection_cholesky_kernel<true><<<1, RC_THREADS, shm, st>>>(H, m, ldh, u, max_accept, mode, shift, ctx->ws, L, ldl,
                                                                  acc, info);
    } else {
        size_t shm = head;
        rejection_cholesky_kernel<false><<<1, RC_THREADS, shm, st>>>(H, m, ldh, u, max_accept, mode, shift, ctx->ws, L, ldl,
                                                                   acc, info);
    }
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// panel assembly.  mode 0: At = [-P_acc; I; 0] L^{-T}, one triangular solve per panel row (warp per
// TS_RW rows, the running solution in shared memory, L streamed from L2 with coalesced row reads).
// mode 1 (eigh fallback): At = [-P_acc; I; 0] T^T with a dense T.
// ------------------------------------------------------------------------------------------------
constexpr int TS_RW = 4;        // panel rows solved together by one warp (re-uses every L load 4 times)
constexpr int TS_WARPS = 4;
constexpr int TS_NJ = 4;        // columns solved per dependent step

// Right-looking (column-oriented) forward substitution: once z_j is known every lane updates ITS OWN right-hand
// side entries l > j, so no warp reduction sits on the j -> j+1 dependency chain.  L is staged transposed
// (Lt[j][l] = L[l][j], packed rows j of length s-j) so the column reads are contiguous, plus 1/L_jj.
template <bool LSMEM>
__global__ void __launch_bounds__(TS_WARPS * 32) panel_trsm_kernel(const double *__restrict__ P, int64_t ldP, int c,
                                                                   const int32_t *__restrict__ acc, int s,
                                                                   const double *__restrict__ L, int64_t ldl,
                                                                   double *__restrict__ At, int64_t ldat, int k_pad,
                                                                   const int32_t *__restrict__ s_dev,
                                                                   const int32_t *__restrict__ ld_table) {
    extern __shared__ double zs[];                  // [TS_WARPS][TS_RW][s] | inv[s] | Lt packed (when LSMEM)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (s_dev) {
        // the row count is still on the device (the host has not read the round back yet): derive the panel shape here
        s = *s_dev;
        ldat = s <= 256 ? ld_table[s] : (s + 255) / 256 * 256;
        k_pad = (c + s + RPC_KT - 1) / RPC_KT * RPC_KT;
    }
    const int r0 = (blockIdx.x * TS_WARPS + warp) * TS_RW;
    double *inv = zs + (size_t)TS_WARPS * TS_RW * s;
    double *Lt = inv + s;
    // row j of Lt starts at j*s - j*(j-1)/2 and holds L[j..s-1][j]
    auto lt_row = [&](int j) -> size_t { return (size_t)j * s - (size_t)j * (j - 1) / 2; };
    for (int j = threadIdx.x; j < s; j += blockDim.x) inv[j] = 1.0 / L[(int64_t)j * ldl + j];
    if (LSMEM) {
        for (int i = warp; i < s; i += TS_WARPS)     // coalesced read of row i of L, scattered write into the columns
            for (int l = lane; l <= i; l += 32) Lt[lt_row(l) + (i - l)] = L[(int64_t)i * ldl + l];
    }
    __syncthreads();
    if (r0 >= k_pad) return;
    double *z = zs + (size_t)warp * TS_RW * s;
    if (r0 < c + s) {
        // rows r < c start from -P[r, acc]; row c+i is e_i, whose solution is zero before column i
        const int jstart = r0 >= c ? r0 - c : 0;
        for (int j = lane; j < s; j += 32) {
#pragma unroll
            for (int w = 0; w < TS_RW; ++w) {
                int r = r0 + w;
                double v = 0.0;
                if (r < c) v = -P[(int64_t)r * ldP + (acc ? acc[j] : j)];
                else if (r - c == j) v = 1.0;
                z[w * s + j] = v;                  // right-hand side, overwritten by the solution column by column
            }
        }
        __syncwarp();
        // TS_NJ columns per dependent step: the small TS_NJ x TS_NJ triangle is solved redundantly by every lane
        // from broadcast reads, then each lane applies the TS_NJ columns to its own entries l >= j + TS_NJ
        for (int j = jstart; j < s; j += TS_NJ) {
            const int nj = min(TS_NJ, s - j);
            double zj[TS_RW][TS_NJ];
#pragma unroll
            for (int q = 0; q < TS_NJ; ++q) {
                if (q < nj) {
#pragma unroll
                    for (int w = 0; w < TS_RW; ++w) {
                        double v = z[w * s + j + q];
#pragma unroll
                        for (int q2 = 0; q2 < TS_NJ; ++q2)
                            if (q2 < q) {
                                const double lv = LSMEM ? Lt[lt_row(j + q2) + (q - q2)] : L[(int64_t)(j + q) * ldl + j + q2];
                                v = fma(-lv, zj[w][q2], v);
                            }
                        zj[w][q] = v * inv[j + q];
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < TS_RW; ++w) zj[w][q] = 0.0;
                }
            }
            __syncwarp();
            if (lane < TS_RW * TS_NJ) {                       // lane (w, q) stores z_{j+q} of row w
                const int w = lane / TS_NJ, q = lane % TS_NJ;
                double v = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < TS_RW; ++w2)
#pragma unroll
                    for (int q2 = 0; q2 < TS_NJ; ++q2)
                        if (w2 == w && q2 == q) v = zj[w2][q2];
                if (q < nj) z[w * s + j + q] = v;
            }
            for (int l = j + nj + lane; l < s; l += 32) {
                double lv[TS_NJ];
#pragma unroll
                for (int q = 0; q < TS_NJ; ++q)
                    lv[q] = q < nj ? (LSMEM ? Lt[lt_row(j + q) + (l - j - q)] : L[(int64_t)l * ldl + j + q]) : 0.0;
#pragma unroll
                for (int w = 0; w < TS_RW; ++w) {
                    double v = z[w * s + l];
#pragma unroll
                    for (int q = 0; q < TS_NJ; ++q) v = fma(-lv[q], zj[w][q], v);
                    z[w * s + l] = v;
                }
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int w = 0; w < TS_RW; ++w) {
        int r = r0 + w;
        if (r >= k_pad) break;
        const bool live = r < c + s;
        for (int j = lane; j < ldat; j += 32) At[(int64_t)r * ldat + j] = (live && j < s) ? z[w * s + j] : 0.0;
    }
}

// MiT given (lds), At rows: blockIdx.x covers PB rows
constexpr int PB = 8;
__global__ void __launch_bounds__(256) panel_kernel(const double *__restrict__ P, int64_t ldP, int c, const int32_t *__restrict__ acc,
                                                    int s, const double *__restrict__ MiT, int lds, int tri,
                                                    double *__restrict__ At, int64_t ldat, int k_pad) {
    extern __shared__ double Ps[];   // [PB][s]
    const int r0 = blockIdx.x * PB;
    if (r0 < c) {
        for (int q = threadIdx.x; q < PB * s; q += blockDim.x) {
            int rr = q / s, i = q - rr * s;
            int r = r0 + rr;
            Ps[q] = (r < c) ? P[(int64_t)r * ldP + (acc ? acc[i] : i)] : 0.0;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < ldat; j += blockDim.x) {
        double out[PB];
#pragma unroll
        for (int rr = 0; rr < PB; ++rr) out[rr] = 0.0;
        if (j < s) {
            // rows of P: -(sum_i P[r][acc_i] * Minv[j][i]), Minv[j][i] = MiT[i][j]; i <= j when triangular
            const int iend = tri ? j + 1 : s;
            bool anyP = r0 < c;
            if (anyP) {
                for (int i = 0; i < iend; ++i) {
                    const double mv = MiT[(size_t)i * lds + j];
#pragma unroll
                    for (int rr = 0; rr < PB; ++rr) out[rr] = fma(Ps[rr * s + i], mv, out[rr]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < PB; ++rr) {
                int r = r0 + rr;
                if (r < c) out[rr] = -out[rr];
                else if (r < c + s) out[rr] = MiT[(size_t)(r - c) * lds + j];
                else out[rr] = 0.0;
            }
        }
#pragma unroll
        for (int rr = 0; rr < PB; ++rr) {
            int r = r0 + rr;
            if (r < k_pad) At[(int64_t)r * ldat + j] = out[rr];
        }
    }
}

__global__ void compact_selected_kernel(const int32_t *__restrict__ acc, int s, const int64_t *__restrict__ idx,
                                        const double *__restrict__ Xpiv, const double *__restrict__ pnorm, int64_t ldp,
                                        int64_t *__restrict__ sel_idx, double *__restrict__ Xsel, double *__restrict__ nsel,
                                        const int32_t *__restrict__ s_dev);

static int launch_panel_trsm(rpc_ctx *ctx, cudaStream_t st, const double *P, int64_t ldP, int c, const int32_t *acc, int s,
                             const double *L, int64_t ldl, double *At, int64_t ldat, int k_pad, const int32_t *s_dev,
                             const int32_t *ld_table) {
    // s is the row count, or its upper bound when the count is read on the device (s_dev)
    size_t shm = ((size_t)TS_WARPS * TS_RW * s + s) * sizeof(double);
    size_t shm_l = shm + (size_t)s * (s + 1) / 2 * sizeof(double);
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(panel_trsm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        RPC_CUDA(cudaFuncSetAttribute(panel_trsm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured |= 1ull << (ctx->device & 63);
    }
    int rows_per_cta = TS_WARPS * TS_RW;
    int blocks = (k_pad + rows_per_cta - 1) / rows_per_cta;
    if (shm_l <= 227 * 1024)
        panel_trsm_kernel<true><<<blocks, TS_WARPS * 32, shm_l, st>>>(P, ldP, c, acc, s, L, ldl, At, ldat, k_pad, s_dev, ld_table);
    else
        panel_trsm_kernel<false><<<blocks, TS_WARPS * 32, shm, st>>>(P, ldP, c, acc, s, L, ldl, At, ldat, k_pad, s_dev, ld_table);
    RPC_LAUNCHED(ctx);
    return 0;
}

// rpc_build_panel (mode 0) + rpc_compact_selected with the accepted count read ON THE DEVICE from *s_dev (<= s_max): both
// can be enqueued right behind the rejection kernel, before the host has read the round back, so the host's wake-up and
// launch latency hides behind them.  The panel uses ldat = rpc_panel_ld(*s_dev) and k_pad = (c + *s_dev) rounded up to 16.
extern "C" int rpc_build_panel_dev(rpc_ctx *ctx, void *stream, const double *P, int64_t ldP, int c, const int32_t *acc,
                                   const int32_t *s_dev, int s_max, const double *L, int64_t ldl, double *At,
                                   const int64_t *idx, const double *Xpiv, const double *pnorm, int64_t ldp, int64_t *sel_idx,
                                   double *Xsel, double *nsel) {
    RPC_CHECK_ARG(ctx && L && At && acc && s_dev && idx && sel_idx, "null pointer");
    RPC_CHECK_ARG(s_max >= 1 && s_max <= RPC_MAX_BLOCK && c >= 0 && (c == 0 || P), "bad sizes (s_max <= RPC_MAX_BLOCK)");
    cudaStream_t st = (cudaStream_t)stream;
    if (!ctx->panel_ld_table) {
        int32_t host[257];
        for (int q = 0; q <= 256; ++q) host[q] = q == 0 ? rpc_panel_ld(1) : rpc_panel_ld(q);
        RPC_CUDA(cudaMalloc((void **)&ctx->panel_ld_table, sizeof(host)));
        RPC_CUDA(cudaMemcpy(ctx->panel_ld_table, host, sizeof(host), cudaMemcpyHostToDevice));
    }
    const int k_pad_max = (c + s_max + RPC_KT - 1) / RPC_KT * RPC_KT;
    int rc = launch_panel_trsm(ctx, st, P, ldP, c, acc, s_max, L, ldl, At, 0, k_pad_max, s_dev, ctx->panel_ld_table);
    if (rc) return rc;
    compact_selected_kernel<<<s_max, 64, 0, st>>>(acc, s_max, idx, Xpiv, pnorm, ldp, sel_idx, Xsel, nsel, s_dev);
    RPC_LAUNCHED(ctx);
    return 0;
}

extern "C" int rpc_build_panel(rpc_ctx *ctx, void *stream, const double *P, int64_t ldP, int c, const int32_t *acc, int s,
                               const double *L, int64_t ldl, int mode, double *At, int64_t ldat, int k_pad) {
    RPC_CHECK_ARG(ctx && L && At, "null pointer");
    RPC_CHECK_ARG(s >= 1 && s <= RPC_MAX_BLOCK && c >= 0 && (c == 0 || P), "bad sizes (s <= RPC_MAX_BLOCK)");
    RPC_CHECK_ARG(k_pad >= c + s && ldat >= s, "k_pad >= c+s and ldat >= s required");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) return launch_panel_trsm(ctx, st, P, ldP, c, acc, s, L, ldl, At, ldat, k_pad, nullptr, nullptr);
    // Minv given in L (dense s x s): MiT[i][j] = L[j][i]
    const int lds = (s + 7) & ~7;
    int rc = rpc_ws_reserve(ctx, (size_t)s * lds * sizeof(double));
    if (rc) return rc;
    double *MiT = ctx->ws;
    rc = rpc_transpose_pad(ctx, st, L, ldl, s, s, MiT, lds, s);
    if (rc) return rc;
    size_t shm = (size_t)PB * s * sizeof(double);
    int blocks = (k_pad + PB - 1) / PB;
    panel_kernel<<<blocks, 256, shm, st>>>(P, ldP, c, acc, s, MiT, lds, 0, At, ldat, k_pad);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// selected pivots S = idx[acc]: compact the payload rows of the accepted proposals (rpcholesky.py:198 `idx = idx[accepted]`)
// ------------------------------------------------------------------------------------------------
__global__ void compact_selected_kernel(const int32_t *__restrict__ acc, int s, const int64_t *__restrict__ idx,
                                        const double *__restrict__ Xpiv, const double *__restrict__ pnorm, int64_t ldp,
                                        int64_t *__restrict__ sel_idx, double *__restrict__ Xsel, double *__restrict__ nsel,
                                        const int32_t *__restrict__ s_dev) {
    const int j = blockIdx.x;
    if (s_dev && j >= *s_dev) return;
    const int src = acc ? acc[j] : j;
    if (threadIdx.x == 0) {
        sel_idx[j] = idx[src];
        if (pnorm && nsel) nsel[j] = pnorm[src];
    }
    if (Xpiv && Xsel)
        for (int f = threadIdx.x; f < ldp; f += blockDim.x) Xsel[(int64_t)j * ldp + f] = Xpiv[(int64_t)src * ldp + f];
}

extern "C" int rpc_compact_selected(rpc_ctx *ctx, void *stream, const int32_t *acc, int s, const int64_t *idx,
                                    const double *Xpiv, const double *pnorm, int64_t ldp, int64_t *sel_idx, double *Xsel,
                                    double *nsel) {
    RPC_CHECK_ARG(ctx && idx && sel_idx && s >= 1, "null pointer or bad sizes");
    compact_selected_kernel<<<s, 64, 0, (cudaStream_t)stream>>>(acc, s, idx, Xpiv, pnorm, ldp, sel_idx, Xsel, nsel, nullptr);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// greedy pivoting: np.argmax(diags) (first maximum, rpcholesky.py:35) and the number of entries tied at the
// maximum (rgreedy draws among them on the host, rpcholesky.py:33)
// ------------------------------------------------------------------------------------------------
struct MaxRec { double v; long long i; long long cnt; };

__device__ __forceinline__ MaxRec max_combine(MaxRec a, MaxRec b) {     // a precedes b in index order
    if (b.v > a.v) return b;
    if (b.v == a.v) { a.cnt += b.cnt; if (a.cnt == b.cnt) a.i = b.i; return a; }
    return a;
}

__global__ void __launch_bounds__(256) argmax_partial_kernel(const double *__restrict__ d, int64_t n, MaxRec *__restrict__ part) {
    __shared__ MaxRec sm[256];
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
    // thread t scans a contiguous sub-range so that combining in thread order keeps "first maximum"
    const int64_t sub = (hi - lo + 255) / 256;
    const int64_t a = lo + (int64_t)threadIdx.x * sub, b = a + sub < hi ? a + sub : hi;
    MaxRec r{-INFINITY, -1, 0};
    for (int64_t j = a; j < b; ++j) {
        const double v = d[j];
        if (v > r.v) { r.v = v; r.i = j; r.cnt = 1; }
        else if (v == r.v) r.cnt++;
    }
    sm[threadIdx.x] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
        MaxRec t = sm[0];
        for (int q = 1; q < 256; ++q) if (sm[q].cnt > 0) t = t.cnt > 0 ? max_combine(t, sm[q]) : sm[q];
        part[blockIdx.x] = t;
    }
}

__global__ void argmax_final_kernel(const MaxRec *__restrict__ part, int nparts, int64_t *__restrict__ idx, double *__restrict__ stats) {
    MaxRec t = part[0];
    for (int q = 1; q < nparts; ++q) if (part[q].cnt > 0) t = t.cnt > 0 ? max_combine(t, part[q]) : part[q];
    idx[0] = t.i;
    stats[0] = t.v;
    stats[1] = (double)t.cnt;
}

extern "C" int rpc_argmax(rpc_ctx *ctx, void *stream, const double *diag, int64_t n, int64_t *idx, double *stats) {
    RPC_CHECK_ARG(ctx && diag && idx && stats && n > 0, "null pointer or empty");
    cudaStream_t st = (cudaStream_t)stream;
    int parts = (int)((n + 8191) / 8192);
    if (parts > 2 * ctx->sm_count) parts = 2 * ctx->sm_count;
    int rc = rpc_ws_reserve(ctx, (size_t)parts * sizeof(MaxRec));
    if (rc) return rc;
    MaxRec *part = reinterpret_cast<MaxRec *>(ctx->ws);
    argmax_partial_kernel<<<parts, 256, 0, st>>>(diag, n, part);
    RPC_LAUNCHED(ctx);
    argmax_final_kernel<<<1, 1, 0, st>>>(part, parts, idx, stats);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// the round's host read-back: [seq | stats(4) | info(4) | acc(m) | idx(m)] written as doubles into PINNED HOST
// memory (UVA: the device writes it directly), the sequence number last, so the host can spin on one word
// instead of paying a stream synchronisation.
// ------------------------------------------------------------------------------------------------
__global__ void publish_round_kernel(const double *__restrict__ stats, const int32_t *__restrict__ info,
                                     const int32_t *__restrict__ acc, const int64_t *__restrict__ idx, int m,
                                     volatile double *host, double seq) {
    const int t = threadIdx.x;
    if (t < 4) host[1 + t] = stats[t];
    if (t < 4) host[5 + t] = info ? (double)info[t] : 0.0;
    for (int j = t; j < m; j += blockDim.x) {
        host[9 + j] = acc ? (double)acc[j] : 0.0;
        host[9 + m + j] = (double)idx[j];
    }
    __threadfence_system();
    __syncthreads();
    if (t == 0) {
        __threadfence_system();
        host[0] = seq;
    }
}

extern "C" int rpc_publish_round(rpc_ctx *ctx, void *stream, const double *stats, const int32_t *info, const int32_t *acc,
                                 const int64_t *idx, int m, double *host_mapped, double seq) {
    RPC_CHECK_ARG(ctx && stats && idx && host_mapped && m >= 0, "null pointer or bad sizes");
    publish_round_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(stats, info, acc, idx, m, host_mapped, seq);
    RPC_LAUNCHED(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// RBRP: diagonally pivoted Cholesky of the b x b residual block (`_greedy_cholesky`, rpcholesky.py:52-63 = LAPACK
// dpstrf(lower) + the trailing-sum rank rule).  ONE CTA, left-looking and swap free: step j takes the first maximum
// of the residual diagonal a_ii - sum_a l_ia^2 (dpstf2's WORK recurrence, sequential in a), stops like dpstrf when it
// falls to m * eps * max_i a_ii, and forms column j from row p of H and the columns computed so far.  The factor is kept
// column-major (LfT[a][i]) so a column update reads coalesced rows.
// ------------------------------------------------------------------------------------------------
constexpr int PC_THREADS = 512;

__global__ void __launch_bounds__(PC_THREADS, 1) pivoted_cholesky_kernel(const double *__restrict__ H, int m, int64_t ldh,
                                                                        int use_tol, double rtol, double atol,
                                                                        double *__restrict__ LfT, double *__restrict__ L,
                                                                        int64_t ldl, int32_t *__restrict__ acc,
                                                                        int32_t *__restrict__ info) {
    extern __shared__ double psh[];
    double *work = psh;                     // [m] sum_a l_ia^2
    double *hd = work + m;                  // [m] a_ii
    double *Lp = hd + m;                    // [m] row p of the factor (first j entries)
    double *cn = Lp + m;                    // [m] squared column norms
    int32_t *piv = (int32_t *)(cn + m);     // [m] pivot order
    int32_t *chosen = piv + m;              // [m]
    __shared__ double red_v[PC_THREADS / 32];
    __shared__ int red_i[PC_THREADS / 32];
    __shared__ double s_ajj, s_stop;
    __shared__ int s_p, s_rank;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    for (int i = tid; i < m; i += PC_THREADS) {
        work[i] = 0.0;
        hd[i] = H[(int64_t)i * ldh + i];
        chosen[i] = 0;
    }
    if (tid == 0) s_rank = m;
    __syncthreads();
    int rank = m;
    for (int j = 0; j < m; ++j) {
        // first maximum of the residual diagonal over the rows not chosen yet (Fortran MAXLOC; NaN never wins a '>')
        double bv = -1.0 / 0.0;
        int bi = 0x7fffffff;
        bool nan_seen = false;
        for (int i = tid; i < m; i += PC_THREADS) {
            if (chosen[i]) continue;
            const double v = hd[i] - work[i];
            if (v != v) nan_seen = true;
            if (v > bv || bi == 0x7fffffff) { if (v > bv || bi == 0x7fffffff) { bv = v; bi = i; } }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(full, bv, o);
            const int oi = __shfl_xor_sync(full, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        nan_seen = __any_sync(full, nan_seen);
        if (lane == 0) { red_v[warp] = nan_seen ? 0.0 / 0.0 : bv; red_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double v = -1.0 / 0.0;
            int p = 0x7fffffff;
            bool bad = false;
            for (int w = 0; w < PC_THREADS / 32; ++w) {
                if (red_i[w] == 0x7fffffff) continue;
                if (red_v[w] != red_v[w]) { bad = true; continue; }
                if (p == 0x7fffffff || red_v[w] > v || (red_v[w] == v && red_i[w] < p)) { v = red_v[w]; p = red_i[w]; }
            }
            if (j == 0) s_stop = (double)m * 2.220446049250313e-16 * v;     // dpstrf: DSTOP = N * eps * max diag (tol < 0)
            // dpstrf tests the first pivot for <= 0 and the later ones against DSTOP
            const bool stop = bad || p == 0x7fffffff || (j == 0 ? !(v > 0.0) : !(v > s_stop));
            s_ajj = v;
            s_p = stop ? -1 : p;
            if (stop) s_rank = j;
        }
        __syncthreads();
        const int p = s_p;
        if (p < 0) { rank = j; break; }
        const double ajj = sqrt(s_ajj);
        for (int a = tid; a < j; a += PC_THREADS) Lp[a] = LfT[(size_t)a * m + p];
        __syncthreads();
        for (int i = tid; i < m; i += PC_THREADS) {
            if (i == p) {
                LfT[(size_t)j * m + i] = ajj;
            } else if (!chosen[i]) {
                // dgemv then dscal: column j = (H[:, p] - sum_a L[:, a] L[p, a]) * (1 / ajj)
                double v = H[(int64_t)p * ldh + i];
                for (int a = 0; a < j; ++a) v = fma(-LfT[(size_t)a * m + i], Lp[a], v);
                const double l = v * (1.0 / ajj);
                LfT[(size_t)j * m + i] = l;
                work[i] += l * l;
            } else {
                LfT[(size_t)j * m + i] = 0.0;
            }
        }
        if (tid == 0) { piv[j] = p; chosen[p] = 1; }
        __syncthreads();
    }
    __syncthreads();
    rank = s_rank;
    if (use_tol && rank > 0) {
        // rank = argmax(trace - cumsum(norm(L, axis=0)**2) < atol + rtol * trace) + 1   (rpcholesky.py:59-62); the column
        // sums run over the rows in pivot order, like numpy's reduction along axis 0
        for (int a = tid; a < rank; a += PC_THREADS) {
            double s = 0.0;
            for (int r = a; r < m; ++r) {
                int row = -1;
                if (r < rank) row = piv[r];
                if (row >= 0) { const double x = LfT[(size_t)a * m + row]; s += x * x; }
            }
            // rows never chosen (rank < m) still carry their entries of the computed columns
            if (rank < m)
                for (int i = 0; i < m; ++i)
                    if (!chosen[i]) { const double x = LfT[(size_t)a * m + i]; s += x * x; }
            const double nrm = sqrt(s);
            cn[a] = nrm * nrm;
        }
        __syncthreads();
        if (tid == 0) {
            double tr = 0.0;
            for (int i = 0; i < m; ++i) tr += hd[i];
            const double thr = atol + rtol * tr;
            double cs = 0.0;
            int newrank = 1;                     // np.argmax of an all-False array is 0
            for (int a = 0; a < rank; ++a) {
                cs += cn[a];
                if (tr - cs < thr) { newrank = a + 1; break; }
            }
            s_rank = newrank;
        }
        __syncthreads();
        rank = s_rank;
    }
    for (int q = tid; q < rank * rank; q += PC_THREADS) {
        const int r = q / rank, a = q - r * rank;
        L[(int64_t)r * ldl + a] = a <= r ? LfT[(size_t)a * m + piv[r]] : 0.0;
    }
    for (int q = tid; q < rank; q += PC_THREADS) acc[q] = piv[q];
    if (tid == 0) { info[0] = rank; info[1] = 0; info[2] = 0; }
}

extern "C" int rpc_pivoted_cholesky(rpc_ctx *ctx, void *stream, const double *H, int m, int64_t ldh, int use_tol, double rtol,
                                    double atol, double *L, int64_t ldl, int32_t *acc, int32_t *info) {
    RPC_CHECK_ARG(ctx && H && L && acc && info, "null pointer");
    RPC_CHECK_ARG(m >= 1 && m <= 4096 && ldh >= m && ldl >= m, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = rpc_ws_reserve(ctx, (size_t)m * m * sizeof(double));
    if (rc) return rc;
    const size_t shm = (size_t)m * (4 * sizeof(double) + 2 * sizeof(int32_t));
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(pivoted_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured |= 1ull << (ctx->device & 63);
    }
    pivoted_cholesky_kernel<<<1, PC_THREADS, shm, st>>>(H, m, ldh, use_tol, rtol, atol, ctx->ws, L, ldl, acc, info);
    RPC_LAUNCHED(ctx);
    return 0;
}


// ------------------------------------------------------------------------------------------------
// dst[i, :] = src[rowmap[i], :] for i < *count: compaction of the speculatively evaluated kernel rows of the accepted
// proposals into rows[c:c+s] (= A[S,:]).  The count is read on the device, so the copy can be enqueued before the host has
// read the round back; it is HBM bound and runs beside the compute-bound Schur update.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const double *__restrict__ src, int64_t lds, const int32_t *__restrict__ rowmap,
                                                          const int32_t *__restrict__ count, double *__restrict__ dst, int64_t ldd,
                                                          int64_t n_pad) {
    // A FEW persistent CTAs (the launch leaves most SM slots free: the copy runs beside the next round's sampler, whose many small
    // launches must not queue behind it), eight 16-byte loads in flight per thread.
    const int rows = *count;
    const int64_t per_row = n_pad / 2;                                   // double2 elements per row
    const int64_t step = (int64_t)gridDim.x * 256;
    for (int i = 0; i < rows; ++i) {
        const double2 *s2 = reinterpret_cast<const double2 *>(src + (int64_t)rowmap[i] * lds);
        double2 *d2 = reinterpret_cast<double2 *>(dst + (int64_t)i * ldd);
        int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
        for (; q + 7 * step < per_row; q += 8 * step) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = s2[q + u * step];
#pragma unroll
            for (int u = 0; u < 8; ++u) d2[q + u * step] = v[u];
        }
        for (; q < per_row; q += step) d2[q] = s2[q];
    }
}

extern "C" int rpc_gather_rows(rpc_ctx *ctx, void *stream, const double *src, int64_t lds, const int32_t *rowmap,
                               const int32_t *count_dev, int max_rows, double *dst, int64_t ldd, int64_t n_pad) {
    RPC_CHECK_ARG(ctx && src && rowmap && count_dev && dst, "null pointer");
    RPC_CHECK_ARG(max_rows >= 1 && n_pad > 0 && n_pad % 2 == 0 && lds % 2 == 0 && ldd % 2 == 0, "bad sizes");
    int gx = (int)((n_pad / 2 + 256 * 4 - 1) / (256 * 4));
    if (gx < 1) gx = 1;
    if (gx > ctx->sm_count / 2) gx = ctx->sm_count / 2;                 // one CTA on every other SM at most
    gather_rows_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(src, lds, rowmap, count_dev, dst, ldd, n_pad);
    RPC_LAUNCHED(ctx);
    return 0;
}
