// Kernel rows for the kernels that are NOT a contraction: Laplace (L1 distance, kernels.py:16-21) and the
// direct-difference Euclidean form (extra_stability, kernels.py:35-36,74-75).  FP64 vector pipe bound:
// 2 DADD (or DADD + DFMA) per (pivot, point, feature).
//
// Persistent CTAs (one per SM): the transposed pivot panel Pt[f][m] is copied into shared memory ONCE per CTA and
// the point tiles (64 points x ldx, one contiguous block because X is point-major) stream through a two-stage
// TMA bulk-copy + mbarrier ring, so the FP64 pipe never waits for a load.  Thread (tx, ty) owns pivots
// ty + 16 i and points tx + 16 j; X is stored with an ODD point stride for these kernels so the 16 points a
// half warp reads per feature fall in 16 different bank pairs, and pivots are broadcast reads.  The L1 distance is
// accumulated strictly in feature order, which makes it bit-equal to scipy's cdist(..., 'cityblock').
#include "dmma_gemm.cuh"

using namespace dg;

constexpr int DE_TN = 64;          // points per tile
constexpr int DE_THREADS = 256;    // 16 x 16

// Laplace epilogue: exp(-l1 / bw) as exp_tab_batch(l1 * (-1 / bw)) -- 13 FP64-pipe instructions per entry instead of a division
// and libm's exp (~40), eight interleaved chains per out-of-line call (two pivot rows of the thread).  The L1 distance itself
// stays bit-equal to scipy's cdist; the kernel value is within ~2.5 ulp of the reference's (argument: one rounding of the
// reciprocal multiply; exponential: <= 1.2 ulp).  Stores go to the thread's four points tx + 16 j of each row.
static __device__ __noinline__ void l1_tab_store8(double *dstA, int okA, double *dstB, int okB, double tab_lane, double x0, double x1,
                                                  double x2, double x3, double x4, double x5, double x6, double x7) {
    double x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
    exp_tab_batch<8>(x, tab_lane);
    if (okA) { dstA[0] = x[0]; dstA[16] = x[1]; dstA[32] = x[2]; dstA[48] = x[3]; }
    if (okB) { dstB[0] = x[4]; dstB[16] = x[5]; dstB[32] = x[6]; dstB[48] = x[7]; }
}
static __device__ __noinline__ double4 l1_tab_wsum8(double tab_lane, double wA, double wB, double4 cs, double x0, double x1, double x2,
                                                    double x3, double x4, double x5, double x6, double x7) {
    double x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
    exp_tab_batch<8>(x, tab_lane);
    cs.x = fma(wB, x[4], fma(wA, x[0], cs.x));
    cs.y = fma(wB, x[5], fma(wA, x[1], cs.y));
    cs.z = fma(wB, x[6], fma(wA, x[2], cs.z));
    cs.w = fma(wB, x[7], fma(wA, x[3], cs.w));
    return cs;
}

// REDUCE (fused kernel-times-vector): out is a vector, out[col] = (accumulate ? out[col] : 0) + sum_row w[row] k(pivot row, X col)
template <bool L1, int PI, bool REDUCE>
__global__ void __launch_bounds__(DE_THREADS, 1) direct_rows_tma(const KernParams kp, const double *__restrict__ X, int64_t ldx,
                                                                 int d, const double *__restrict__ Pt, int ldpt, int s,
                                                                 double *__restrict__ out, int64_t ldo, int ntiles,
                                                                 const double *__restrict__ w, int accumulate) {
    extern __shared__ __align__(16) double sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm);          // [0] panel, [1..2] X tiles
    double *Ps = sm + 4;                                        // [d][ldpt]
    double *Xs = Ps + (size_t)d * ldpt;                         // [2][DE_TN * ldx]
    double *red = Xs + 2 * (size_t)DE_TN * ldx;                 // REDUCE: [16 thread rows][DE_TN] partial column sums
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tile_doubles = DE_TN * (int)ldx;
    const double tab_lane = c_exp2_tab[tid & 31];
    const double nib = -1.0 / kp.bw;                            // Laplace: exp(-l1 / bw) = exp(l1 * nib)

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int first = blockIdx.x;
    if (tid == 0) {
        const unsigned pb = (unsigned)((size_t)d * ldpt * 8);
        mbar_expect_tx(&bars[0], pb);
        bulk_g2s(Ps, Pt, pb, &bars[0]);
        if (first < ntiles) {
            mbar_expect_tx(&bars[1], (unsigned)tile_doubles * 8);
            bulk_g2s(Xs, X + (int64_t)first * DE_TN * ldx, (unsigned)tile_doubles * 8, &bars[1]);
        }
    }
    mbar_wait(&bars[0], 0);

    int it = 0;
    for (int tile = first; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int nxt = tile + gridDim.x;
        if (tid == 0 && nxt < ntiles) {     // buffer buf^1 was released by the barrier that ended the previous iteration
            mbar_expect_tx(&bars[1 + (buf ^ 1)], (unsigned)tile_doubles * 8);
            bulk_g2s(Xs + (buf ^ 1) * tile_doubles, X + (int64_t)nxt * DE_TN * ldx, (unsigned)tile_doubles * 8, &bars[1 + (buf ^ 1)]);
        }
        mbar_wait(&bars[1 + buf], (it >> 1) & 1);
        const double *xs = Xs + buf * tile_doubles;

        double acc[PI][4];
#pragma unroll
        for (int i = 0; i < PI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 2
        for (int f = 0; f < d; ++f) {
            double xp[PI], yv[4];
#pragma unroll
            for (int i = 0; i < PI; ++i) xp[i] = Ps[f * ldpt + ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) yv[j] = xs[(tx + 16 * j) * ldx + f];
#pragma unroll
            for (int i = 0; i < PI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double df = xp[i] - yv[j];
                    if (L1) acc[i][j] += fabs(df);                // sequential in f: bit-equal to cdist 'cityblock'
                    else acc[i][j] = fma(df, df, acc[i][j]);
                }
        }
        const int64_t col0 = (int64_t)tile * DE_TN;
        if (L1) {
            // every lane takes part (the table lives in the lanes' registers): rows past s are computed and not stored
            double4 cs4 = make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
            for (int i = 0; i < PI; i += 2) {
                constexpr int LAST = PI - 1;
                const int i1 = i + 1 < PI ? i + 1 : LAST;
                const int rowA = ty + 16 * i, rowB = ty + 16 * i1;
                const bool pair = i + 1 < PI;
                if (REDUCE) {
                    const double wA = rowA < s ? w[rowA] : 0.0, wB = (pair && rowB < s) ? w[rowB] : 0.0;
                    cs4 = l1_tab_wsum8(tab_lane, wA, wB, cs4, acc[i][0] * nib, acc[i][1] * nib, acc[i][2] * nib, acc[i][3] * nib,
                                       acc[i1][0] * nib, acc[i1][1] * nib, acc[i1][2] * nib, acc[i1][3] * nib);
                } else {
                    l1_tab_store8(out + (int64_t)rowA * ldo + col0 + tx, rowA < s, out + (int64_t)rowB * ldo + col0 + tx,
                                  pair && rowB < s, tab_lane, acc[i][0] * nib, acc[i][1] * nib, acc[i][2] * nib, acc[i][3] * nib,
                                  acc[i1][0] * nib, acc[i1][1] * nib, acc[i1][2] * nib, acc[i1][3] * nib);
                }
            }
            if (REDUCE) {
                red[ty * DE_TN + tx] = cs4.x; red[ty * DE_TN + tx + 16] = cs4.y;
                red[ty * DE_TN + tx + 32] = cs4.z; red[ty * DE_TN + tx + 48] = cs4.w;
                __syncthreads();
                if (tid < DE_TN) {
                    double t = 0.0;
#pragma unroll
                    for (int q = 0; q < 16; ++q) t += red[q * DE_TN + tid];          // fixed order: deterministic
                    out[col0 + tid] = (accumulate ? out[col0 + tid] : 0.0) + t;
                }
            }
            __syncthreads();      // every thread is done with xs (and red) before they are refilled
            continue;
        }
        if (REDUCE) {
            double cs[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int i = 0; i < PI; ++i) {
                const int row = ty + 16 * i;
                if (row < s) {
                    const double wr = w[row];
#pragma unroll
                    for (int j = 0; j < 4; j += 2) {
                        const double2 v = L1 ? kern_from_l1_pair(kp, acc[i][j], acc[i][j + 1])
                                             : kern_from_d2_pair(kp, acc[i][j], acc[i][j + 1]);
                        cs[j] = fma(wr, v.x, cs[j]);
                        cs[j + 1] = fma(wr, v.y, cs[j + 1]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) red[ty * DE_TN + tx + 16 * j] = cs[j];
            __syncthreads();
            if (tid < DE_TN) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) t += red[q * DE_TN + tid];          // fixed order: deterministic
                out[col0 + tid] = (accumulate ? out[col0 + tid] : 0.0) + t;
            }
            __syncthreads();
            continue;
        }
#pragma unroll
        for (int i = 0; i < PI; ++i) {
            const int row = ty + 16 * i;
            if (row < s) {
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const double2 v = L1 ? kern_from_l1_pair(kp, acc[i][j], acc[i][j + 1])
                                         : kern_from_d2_pair(kp, acc[i][j], acc[i][j + 1]);
                    out[(int64_t)row * ldo + col0 + tx + 16 * j] = v.x;
                    out[(int64_t)row * ldo + col0 + tx + 16 * (j + 1)] = v.y;
                }
            }
        }
        __syncthreads();      // every thread is done with xs before it is refilled two iterations later
    }
}

template <bool L1, int PI>
static int launch_direct(rpc_ctx *ctx, cudaStream_t st, const KernParams &kp, const double *X, int64_t ldx, int d, const double *Pt,
                         int ldpt, int s, double *out, int64_t ldo, int64_t n_pad, size_t shm, const double *w, int accumulate) {
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(direct_rows_tma<L1, PI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        RPC_CUDA(cudaFuncSetAttribute(direct_rows_tma<L1, PI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured |= 1ull << (ctx->device & 63);
    }
    const int ntiles = (int)(n_pad / DE_TN);
    const int ctas = rpc_persistent_ctas(ctx);
    const int grid = ntiles < ctas ? ntiles : ctas;
    if (w) direct_rows_tma<L1, PI, true><<<grid, DE_THREADS, shm, st>>>(kp, X, ldx, d, Pt, ldpt, s, out, ldo, ntiles, w, accumulate);
    else direct_rows_tma<L1, PI, false><<<grid, DE_THREADS, shm, st>>>(kp, X, ldx, d, Pt, ldpt, s, out, ldo, ntiles, nullptr, 0);
    RPC_LAUNCHED(ctx);
    return 0;
}

// s <= 256 pivots, panel Pt (d x ldpt, ldpt = 16 * PI) already built.  Returns -100 if the shape does not fit.
// w != NULL: the fused kernel-times-vector variant (out = vector of n_pad weighted column sums).
int rpc_direct_rows_tma(rpc_ctx *ctx, cudaStream_t st, const KernParams &kp, bool l1, const double *X, int64_t ldx, int d,
                        const double *Pt, int ldpt, int s, double *out, int64_t ldo, int64_t n_pad, const double *w, int accumulate) {
    const int pi = ldpt / 16;
    const size_t shm = (4 + (size_t)d * ldpt + 2 * (size_t)DE_TN * ldx + 16 * DE_TN) * sizeof(double);
    if (shm > 227 * 1024 || (ldx * 8 * DE_TN) % 16 != 0) return -100;
#define DE_CASE(L, P) \
    case P: return launch_direct<L, P>(ctx, st, kp, X, ldx, d, Pt, ldpt, s, out, ldo, n_pad, shm, w, accumulate);
    if (l1) {
        switch (pi) {
            DE_CASE(true, 1) DE_CASE(true, 2) DE_CASE(true, 4) DE_CASE(true, 6) DE_CASE(true, 8) DE_CASE(true, 10)
            DE_CASE(true, 11) DE_CASE(true, 12) DE_CASE(true, 13) DE_CASE(true, 14) DE_CASE(true, 16)
            default: return -100;
        }
    }
    switch (pi) {
        DE_CASE(false, 2) DE_CASE(false, 4) DE_CASE(false, 8) DE_CASE(false, 12) DE_CASE(false, 16)
        default: return -100;
    }
#undef DE_CASE
}

// pivots-per-thread variant that covers s rows (multiples instantiated above)
int rpc_direct_pi_for(int s, bool l1) {
    const int need = (s + 15) / 16;
    const int l1_set[] = {1, 2, 4, 6, 8, 10, 11, 12, 13, 14, 16};
    const int d2_set[] = {2, 4, 8, 12, 16};
    if (l1) { for (int p : l1_set) if (p >= need) return p; }
    else { for (int p : d2_set) if (p >= need) return p; }
    return 16;
}
