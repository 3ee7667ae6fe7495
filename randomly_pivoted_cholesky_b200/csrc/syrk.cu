// Normal equations of restricted-Nystrom kernel ridge regression (KRR_Nystrom.py:52-54) on the factor rows that the hot path
// leaves in HBM:   M = K_Mn K_nM  (k x k, contraction over the N_loc columns of rows = A[S,:]),   rhs = K_Mn y,
// and the two triangular solves of scipy.linalg.solve(..., assume_a='pos') once the k x k Cholesky factor is known.
//
// rpc_syrk_rows: FP64 DMMA, 128 x 128 output tiles of the LOWER triangle (the product is symmetric: 36 instead of 64 tiles at
// k = 1000), each split over the columns so that tiles x splits fill the SMs; the partial tiles go to the workspace and a second
// kernel adds them in a fixed order (deterministic, no atomics) and mirrors the result.  Both operands of a tile are row blocks
// of the SAME row-major matrix with the contraction index contiguous, so a k-chunk of a tile is a 2-D box (128 rows x 20
// columns): it is fetched with ONE tensor-map TMA copy (cp.async.bulk.tensor.2d -> SASS UTMALDG) per operand and stage; row by
// row bulk copies of 160 bytes would be bound by the TMA issue rate (~40 ns per copy).  The box is 20 doubles wide on
// purpose: the shared-memory rows are packed (pitch 20 = 4 mod 8 doubles), which makes the DMMA fragment loads
// (lane -> row lane>>2, k lane&3) bank-conflict free without a swizzle.  Rows >= k and columns >= n_cols of a box are
// zero-filled by the TMA unit (out-of-bounds fill), so there is no edge handling anywhere.
#include <cuda.h>
#include "dmma_gemm.cuh"

using namespace dg;

namespace sy {

constexpr int TM = 128, KC = 20, SY_STAGES = 4;
constexpr int WARPS_M = 2, WARPS_N = 4, SMI = TM / (8 * WARPS_M), SNI = TM / (8 * WARPS_N);     // 8 x 4 fragments per warp
constexpr int CONSUMERS = WARPS_M * WARPS_N;
constexpr int THREADS = (CONSUMERS + 4) * 32;
constexpr int BOX_DOUBLES = TM * KC;                              // one operand of one stage: 20480 bytes (128-byte aligned)
constexpr size_t SMEM_BYTES = (size_t)SY_STAGES * 2 * BOX_DOUBLES * 8 + 2 * SY_STAGES * 8 + 128;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the driver entry point, resolved through the runtime (the library does not link libcuda)
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

struct SyrkParams {
    int nb;            // row blocks of 128
    int ntile;         // nb (nb + 1) / 2 lower-triangular tiles
    int nsplit;        // column splits
    int chunks_per;    // k-chunks (of KC columns) per split
    int nchunks;       // k-chunks in total
    double *part;      // [nsplit][ntile][128][128]
};

__global__ void __launch_bounds__(THREADS, 1) syrk_kernel(const __grid_constant__ CUtensorMap tmap, const SyrkParams p) {
    extern __shared__ __align__(128) double smem_raw[];
    // TMA tile destinations need 128-byte alignment: align by hand (SMEM_BYTES carries 128 bytes of slack)
    double *smem = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    double *As = smem;                                            // [stage][128][KC]
    double *Bs = smem + SY_STAGES * BOX_DOUBLES;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + 2 * SY_STAGES * BOX_DOUBLES);
    uint64_t *empty = full + SY_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // tile (I, J), J <= I, from the linear index t = I (I + 1) / 2 + J
    const int t = blockIdx.x % p.ntile, split = blockIdx.x / p.ntile;
    int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const bool diag = I == J;
    const int ch0 = split * p.chunks_per, ch1 = min(p.nchunks, ch0 + p.chunks_per);

    if (tid == 0) {
        for (int s = 0; s < SY_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMERS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp >= CONSUMERS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp != CONSUMERS || lane != 0) return;
        // producer: one thread, one (diagonal tile) or two tensor-map copies per stage
        const unsigned bytes = (unsigned)(BOX_DOUBLES * 8) * (diag ? 1u : 2u);
        int it = 0;
        for (int ch = ch0; ch < ch1; ++ch, ++it) {
            const int s = it % SY_STAGES;
            mbar_wait(&empty[s], ((it / SY_STAGES) & 1) ^ 1);
            mbar_expect_tx(&full[s], bytes);
            tma_load_2d(As + s * BOX_DOUBLES, &tmap, ch * KC, I * TM, &full[s]);
            if (!diag) tma_load_2d(Bs + s * BOX_DOUBLES, &tmap, ch * KC, J * TM, &full[s]);
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int m0 = wm * SMI * 8, n0 = wn * SNI * 8;
    const int lr = lane >> 2, lk = lane & 3;
    double acc[SMI][SNI][2];
#pragma unroll
    for (int i = 0; i < SMI; ++i)
#pragma unroll
        for (int j = 0; j < SNI; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    int it = 0;
    for (int ch = ch0; ch < ch1; ++ch, ++it) {
        const int s = it % SY_STAGES;
        mbar_wait(&full[s], (it / SY_STAGES) & 1);
        const double *Ap = As + s * BOX_DOUBLES + (m0 + lr) * KC + lk;
        const double *Bp = (diag ? As : Bs) + s * BOX_DOUBLES + (n0 + lr) * KC + lk;
        double a[2][SMI], b[2][SNI];
#pragma unroll
        for (int i = 0; i < SMI; ++i) a[0][i] = Ap[i * 8 * KC];
#pragma unroll
        for (int j = 0; j < SNI; ++j) b[0][j] = Bp[j * 8 * KC];
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            const int cur = ks & 1, nxt = cur ^ 1;
            if (ks + 1 < KC / 4) {
#pragma unroll
                for (int i = 0; i < SMI; ++i) a[nxt][i] = Ap[i * 8 * KC + (ks + 1) * 4];
#pragma unroll
                for (int j = 0; j < SNI; ++j) b[nxt][j] = Bp[j * 8 * KC + (ks + 1) * 4];
            }
#pragma unroll
            for (int i = 0; i < SMI; ++i)
#pragma unroll
                for (int j = 0; j < SNI; ++j) dmma884(acc[i][j], a[cur][i], b[cur][j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    // partial tile -> workspace (row = m0 + i*8 + lr, cols = n0 + j*8 + 2*lk + {0,1})
    double *dst = p.part + ((size_t)split * p.ntile + t) * TM * TM;
#pragma unroll
    for (int i = 0; i < SMI; ++i)
#pragma unroll
        for (int j = 0; j < SNI; ++j)
            *reinterpret_cast<double2 *>(dst + (size_t)(m0 + i * 8 + lr) * TM + n0 + j * 8 + 2 * lk) = make_double2(acc[i][j][0], acc[i][j][1]);
}

// M[I*128 + r, J*128 + c] = sum over the splits (fixed order) of the partial tiles, mirrored into the upper triangle
__global__ void __launch_bounds__(256) syrk_reduce_kernel(const double *__restrict__ part, int ntile, int nsplit, int k,
                                                          double *__restrict__ M, int64_t ldm) {
    const int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    while (I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    for (int q = threadIdx.x; q < TM * TM; q += 256) {
        const int r = q / TM, c = q - r * TM;
        const int row = I * TM + r, col = J * TM + c;
        if (row >= k || col >= k) continue;
        double s = 0.0;
        for (int z = 0; z < nsplit; ++z) s += part[((size_t)z * ntile + t) * TM * TM + q];
        if (I == J && c > r) continue;                    // the diagonal tile's upper half is written by its mirror below
        M[(int64_t)row * ldm + col] = s;
        if (row != col) M[(int64_t)col * ldm + row] = s;
    }
}

// out[i] = sum_col R[i, col] * y[col]: one CTA per row, fixed-order tree (deterministic); HBM bound
__global__ void __launch_bounds__(256) rows_dot_kernel(const double *__restrict__ R, int64_t ldr, int64_t n, const double *__restrict__ y,
                                                       double *__restrict__ out) {
    __shared__ double red[256];
    const double *row = R + (int64_t)blockIdx.x * ldr;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int64_t c = threadIdx.x;
    for (; c + 768 < n; c += 1024) {
        a0 = fma(row[c], y[c], a0);
        a1 = fma(row[c + 256], y[c + 256], a1);
        a2 = fma(row[c + 512], y[c + 512], a2);
        a3 = fma(row[c + 768], y[c + 768], a3);
    }
    for (; c < n; c += 256) a0 = fma(row[c], y[c], a0);
    red[threadIdx.x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// out[col] = sum_i w[i] * R[i, col]: each thread owns two adjacent columns and keeps 8 row loads in flight; one pass over R
__global__ void __launch_bounds__(256) rows_tdot_kernel(const double *__restrict__ R, int64_t ldr, int k, int64_t n2,
                                                        const double *__restrict__ w, double *__restrict__ out) {
    extern __shared__ double ws[];
    for (int i = threadIdx.x; i < k; i += 256) ws[i] = w[i];
    __syncthreads();
    for (int64_t c2 = (int64_t)blockIdx.x * 256 + threadIdx.x; c2 < n2; c2 += (int64_t)gridDim.x * 256) {
        const double *Rc = R + 2 * c2;
        double2 acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = make_double2(0.0, 0.0);
        int i = 0;
        for (; i + 8 <= k; i += 8) {
            double2 g[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) g[q] = *reinterpret_cast<const double2 *>(Rc + (int64_t)(i + q) * ldr);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc[q].x = fma(ws[i + q], g[q].x, acc[q].x);
                acc[q].y = fma(ws[i + q], g[q].y, acc[q].y);
            }
        }
        for (; i < k; ++i) {
            const double2 g = *reinterpret_cast<const double2 *>(Rc + (int64_t)i * ldr);
            acc[0].x = fma(ws[i], g.x, acc[0].x);
            acc[0].y = fma(ws[i], g.y, acc[0].y);
        }
        const double sx = ((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x)) + ((acc[4].x + acc[5].x) + (acc[6].x + acc[7].x));
        const double sy = ((acc[0].y + acc[1].y) + (acc[2].y + acc[3].y)) + ((acc[4].y + acc[5].y) + (acc[6].y + acc[7].y));
        *reinterpret_cast<double2 *>(out + 2 * c2) = make_double2(sx, sy);
    }
}

// x <- (L L^T)^{-1} x for a lower-triangular L (k x k, row-major) and nrhs right-hand sides stored as columns of x (k x nrhs,
// row-major): ONE CTA, 32-wide blocks; forward substitution, then backward substitution with L^T.
constexpr int TS_B = 32;
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double *__restrict__ L, int64_t ldl, int k, double *__restrict__ x,
                                                          int64_t ldx, int nrhs) {
    __shared__ double xb[TS_B];
    const int tid = threadIdx.x;
    for (int c = 0; c < nrhs; ++c) {
        double *xc = x + c;
        // ---- forward: L u = b
        for (int j0 = 0; j0 < k; j0 += TS_B) {
            const int nb = min(TS_B, k - j0);
            if (tid < 32) {                                   // the diagonal block, one warp: lane r owns row j0 + r
                double v = tid < nb ? xc[(int64_t)(j0 + tid) * ldx] : 0.0;
                for (int jj = 0; jj < nb; ++jj) {
                    const double ljj = L[(int64_t)(j0 + jj) * ldl + j0 + jj];
                    const double xj = __shfl_sync(0xffffffffu, v, jj) / ljj;
                    if (tid == jj) v = xj;
                    else if (tid > jj && tid < nb) v -= L[(int64_t)(j0 + tid) * ldl + j0 + jj] * xj;
                }
                if (tid < nb) { xb[tid] = v; xc[(int64_t)(j0 + tid) * ldx] = v; }
            }
            __syncthreads();
            for (int i = j0 + nb + tid; i < k; i += blockDim.x) {       // rows below: b_i -= L[i, block] . x_block
                const double *Li = L + (int64_t)i * ldl + j0;
                double s = 0.0;
                for (int jj = 0; jj < nb; ++jj) s = fma(Li[jj], xb[jj], s);
                xc[(int64_t)i * ldx] -= s;
            }
            __syncthreads();
        }
        // ---- backward: L^T x = u
        for (int j1 = k; j1 > 0; j1 -= TS_B) {
            const int j0 = max(0, j1 - TS_B), nb = j1 - j0;
            if (tid < 32) {
                double v = tid < nb ? xc[(int64_t)(j0 + tid) * ldx] : 0.0;
                for (int jj = nb - 1; jj >= 0; --jj) {
                    const double ljj = L[(int64_t)(j0 + jj) * ldl + j0 + jj];
                    const double xj = __shfl_sync(0xffffffffu, v, jj) / ljj;
                    if (tid == jj) v = xj;
                    else if (tid < jj) v -= L[(int64_t)(j0 + jj) * ldl + j0 + tid] * xj;      // (L^T)[tid, jj] = L[jj, tid]
                }
                if (tid < nb) { xb[tid] = v; xc[(int64_t)(j0 + tid) * ldx] = v; }
            }
            __syncthreads();
            for (int j = tid; j < j0; j += blockDim.x) {                 // earlier entries: u_j -= sum_i L[i, j] x_i, i in the block
                double s = 0.0;
                for (int ii = 0; ii < nb; ++ii) s = fma(L[(int64_t)(j0 + ii) * ldl + j], xb[ii], s);
                xc[(int64_t)j * ldx] -= s;
            }
            __syncthreads();
        }
    }
}

}  // namespace sy

extern "C" int rpc_syrk_rows(rpc_ctx *ctx, void *stream, const double *R, int64_t ldr, int k, int64_t n_cols, double *M, int64_t ldm) {
    using namespace sy;
    RPC_CHECK_ARG(ctx && R && M, "null pointer");
    RPC_CHECK_ARG(k >= 1 && n_cols >= 1 && ldr >= n_cols && ldm >= k, "bad sizes");
    RPC_CHECK_ARG(ldr % 2 == 0 && ((uintptr_t)R & 15) == 0, "R must be 16-byte aligned with an even leading dimension (TMA)");
    cudaStream_t st = (cudaStream_t)stream;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) {
        rpc_set_error("rpc_syrk_rows: cuTensorMapEncodeTiled is not available from this driver");
        return -5;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)n_cols, (cuuint64_t)k};
    const cuuint64_t gstr[1] = {(cuuint64_t)ldr * 8};
    const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)TM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(R), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        rpc_set_error("rpc_syrk_rows: cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);
        return -5;
    }
    SyrkParams p{};
    p.nb = (k + TM - 1) / TM;
    p.ntile = p.nb * (p.nb + 1) / 2;
    p.nchunks = (int)((n_cols + KC - 1) / KC);
    const int ctas = rpc_persistent_ctas(ctx);
    int nsplit = ctas / p.ntile;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > p.nchunks) nsplit = p.nchunks;
    p.chunks_per = (p.nchunks + nsplit - 1) / nsplit;
    p.nsplit = (p.nchunks + p.chunks_per - 1) / p.chunks_per;
    int rc = rpc_ws_reserve(ctx, (size_t)p.nsplit * p.ntile * TM * TM * sizeof(double));
    if (rc) return rc;
    p.part = ctx->ws;
    static unsigned long long configured = 0;
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        configured |= 1ull << (ctx->device & 63);
    }
    syrk_kernel<<<p.ntile * p.nsplit, THREADS, SMEM_BYTES, st>>>(tmap, p);
    RPC_LAUNCHED(ctx);
    syrk_reduce_kernel<<<p.ntile, 256, 0, st>>>(p.part, p.ntile, p.nsplit, k, M, ldm);
    RPC_LAUNCHED(ctx);
    return 0;
}

extern "C" int rpc_rows_dot(rpc_ctx *ctx, void *stream, const double *R, int64_t ldr, int k, int64_t n_cols, const double *y, double *out) {
    RPC_CHECK_ARG(ctx && R && y && out && k >= 1 && n_cols >= 1 && ldr >= n_cols, "null pointer or bad sizes");
    sy::rows_dot_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(R, ldr, n_cols, y, out);
    RPC_LAUNCHED(ctx);
    return 0;
}

extern "C" int rpc_chol_solve(rpc_ctx *ctx, void *stream, const double *L, int64_t ldl, int k, double *x, int64_t ldx, int nrhs) {
    RPC_CHECK_ARG(ctx && L && x && k >= 1 && nrhs >= 1 && ldl >= k && ldx >= nrhs, "null pointer or bad sizes");
    sy::chol_solve_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(L, ldl, k, x, ldx, nrhs);
    RPC_LAUNCHED(ctx);
    return 0;
}

extern "C" int rpc_rows_tdot(rpc_ctx *ctx, void *stream, const double *R, int64_t ldr, int k, int64_t n_cols, const double *w, double *out) {
    RPC_CHECK_ARG(ctx && R && w && out && k >= 1 && n_cols >= 2, "null pointer or bad sizes");
    RPC_CHECK_ARG(n_cols % 2 == 0 && ldr % 2 == 0 && ldr >= n_cols, "n_cols and ldr must be even (16-byte column pairs)");
    RPC_CHECK_ARG((size_t)k * sizeof(double) <= 200 * 1024, "k too large for the shared-memory weight vector");
    const size_t shm = (size_t)k * sizeof(double);
    static unsigned long long configured = 0;
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(sy::rows_tdot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured |= 1ull << (ctx->device & 63);
    }
    int grid = (int)((n_cols / 2 + 255) / 256);
    const int cap = ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    sy::rows_tdot_kernel<<<grid, 256, shm, (cudaStream_t)stream>>>(R, ldr, k, n_cols / 2, w, out);
    RPC_LAUNCHED(ctx);
    return 0;
}
