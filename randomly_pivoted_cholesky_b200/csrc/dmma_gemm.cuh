// FP64 tensor-core (DMMA.8x8x4 via mma.sync) skinny GEMM core for sm_100a.
//
//   C[M x n_pad] = At^T[M x K] * B[K x n_pad]        M <= 256 (all rows of the update in one CTA)
//
// One CTA owns ALL M rows of a column tile, so the big operand B (the factor G, or the data X) is read
// from HBM exactly once per launch and the epilogue (residual-diagonal update / kernel function) is
// fused.  The small operand At (the pivot panel, K x M, zero padded) is re-read by every CTA from L2.
// Operands are staged in shared memory by a 4-stage cp.async pipeline; smem leading dimensions are
// == 4 (mod 16) doubles so that the DMMA fragment loads (lane -> (k = lane&3, row = lane>>2)) hit 16
// distinct 8-byte bank pairs per half warp.
//
// tcgen05 has no FP64 kind, so on Blackwell the FP64 tensor path is the warp-level mma.sync DMMA;
// measured peak on B200: 36.9 TFLOP/s (profiles/fp64_peaks_r01.json), same pipe as DFMA.
#pragma once
#include "rpc_common.cuh"

namespace dg {

constexpr int KT = RPC_KT;      // k-depth per stage
constexpr int STAGES = 4;
constexpr int LDBK = KT + 4;    // point-major B stage: [n][LDBK]

enum { MODE_SCHUR = 0, MODE_EVAL = 1 };

template <int WM_, int WN_, int MI_, int NI_>
struct Cfg {
    static constexpr int WM = WM_, WN = WN_, MI = MI_, NI = NI_;
    static constexpr int THREADS = WM * WN * 32;
    static constexpr int M_TILE = WM * MI * 8;
    static constexpr int N_TILE = WN * NI * 8;
    static constexpr int LDA = M_TILE + 4;
    static constexpr int LDB = N_TILE + 4;
    static constexpr int A_STAGE = KT * LDA;                                  // doubles
    static constexpr int B_STAGE_ROW = KT * LDB;                              // k-major B stage
    static constexpr int B_STAGE_PT = N_TILE * LDBK;                          // point-major B stage
    static constexpr int B_STAGE = B_STAGE_ROW > B_STAGE_PT ? B_STAGE_ROW : B_STAGE_PT;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);
    static_assert(M_TILE % 16 == 0 || M_TILE == 8, "M_TILE");
    static_assert(N_TILE % 16 == 0, "N_TILE");
};

struct Params {
    const double *At;   // k_pad x ldat, zero padded
    int64_t ldat;
    int k_pad;          // multiple of KT
    int m_valid;        // rows of C that exist (<= M_TILE)
    // MODE_SCHUR: B rows kk < c come from G, c <= kk < K from R; C -> Cout (ldc); diag updated
    const double *G;
    int64_t ldg;
    int c;
    const double *R;
    int64_t ldr;
    int K;
    double *Cout;
    int64_t ldc;
    double *diag;
    // MODE_EVAL: B[k][n] = X[n][k] (point-major, d features); C -> kernel values
    const double *X;
    int64_t ldx;
    int d;
    const double *xnorm;   // per point
    const double *pnorm;   // per pivot (row of C)
    KernParams kp;
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
// src_bytes = 0 -> 16 zero bytes are written
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <class C, int MODE>
__device__ __forceinline__ void load_stage(const Params &p, double *As, double *Bs, int k0, int64_t col0, int tid) {
    // A tile: KT rows x M_TILE columns of At
    constexpr int ACH = C::M_TILE / 2;   // 16-byte chunks per row
#pragma unroll 2
    for (int q = tid; q < KT * ACH; q += C::THREADS) {
        int r = q / ACH, c2 = q - r * ACH;
        cp_async16(As + r * C::LDA + c2 * 2, p.At + (int64_t)(k0 + r) * p.ldat + c2 * 2);
    }
    if (MODE == MODE_SCHUR) {
        constexpr int BCH = C::N_TILE / 2;
#pragma unroll 2
        for (int q = tid; q < KT * BCH; q += C::THREADS) {
            int r = q / BCH, c2 = q - r * BCH;
            int kk = k0 + r;
            const double *src;
            if (kk < p.c) src = p.G + (int64_t)kk * p.ldg;
            else if (kk < p.K) src = p.R + (int64_t)(kk - p.c) * p.ldr;
            else src = p.R;                       // padded k: At is zero there, any finite row will do
            cp_async16(Bs + r * C::LDB + c2 * 2, src + col0 + c2 * 2);
        }
    } else {
        constexpr int KCH = KT / 2;
#pragma unroll 2
        for (int q = tid; q < C::N_TILE * KCH; q += C::THREADS) {
            int n = q / KCH, c2 = q - n * KCH;
            int kk = k0 + c2 * 2;
            int ok = kk < p.d;
            const double *src = p.X + (col0 + n) * p.ldx + (ok ? kk : 0);
            cp_async16_zfill(Bs + n * LDBK + c2 * 2, src, ok ? 16 : 0);
        }
    }
}

template <class C, int MODE>
__global__ void __launch_bounds__(C::THREADS, 1) gemm_kernel(const Params p) {
    extern __shared__ __align__(16) double smem[];
    double *As_base = smem;
    double *Bs_base = smem + STAGES * C::A_STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp % C::WM, wn = warp / C::WM;
    const int m0 = wm * C::MI * 8, n0 = wn * C::NI * 8;
    const int lr = lane >> 2, lk = lane & 3;
    const int64_t col0 = (int64_t)blockIdx.x * C::N_TILE;

    double acc[C::MI][C::NI][2];
#pragma unroll
    for (int i = 0; i < C::MI; ++i)
#pragma unroll
        for (int j = 0; j < C::NI; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    const int nk = p.k_pad / KT;
    // prologue: STAGES-1 stages in flight
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage<C, MODE>(p, As_base + s * C::A_STAGE, Bs_base + s * C::B_STAGE, s * KT, col0, tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {   // refill the stage consumed in the previous iteration
            int nxt = kt + STAGES - 1;
            if (nxt < nk) {
                int s = nxt % STAGES;
                load_stage<C, MODE>(p, As_base + s * C::A_STAGE, Bs_base + s * C::B_STAGE, nxt * KT, col0, tid);
            }
            cp_async_commit();
        }
        const double *As = As_base + (kt % STAGES) * C::A_STAGE;
        const double *Bs = Bs_base + (kt % STAGES) * C::B_STAGE;
#pragma unroll
        for (int ks = 0; ks < KT / 4; ++ks) {
            const int kk = ks * 4 + lk;
            double a[C::MI], b[C::NI];
#pragma unroll
            for (int i = 0; i < C::MI; ++i) a[i] = As[kk * C::LDA + m0 + i * 8 + lr];
#pragma unroll
            for (int j = 0; j < C::NI; ++j) {
                if (MODE == MODE_SCHUR) b[j] = Bs[kk * C::LDB + n0 + j * 8 + lr];
                else b[j] = Bs[(n0 + j * 8 + lr) * LDBK + kk];
            }
#pragma unroll
            for (int i = 0; i < C::MI; ++i)
#pragma unroll
                for (int j = 0; j < C::NI; ++j) dmma884(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // all warps done with the stages: smem can be reused by the epilogue

    // C fragment: row = m0 + i*8 + lr, cols = n0 + j*8 + 2*lk + {0,1}
    if (MODE == MODE_SCHUR) {
        double csum[C::NI][2];
#pragma unroll
        for (int j = 0; j < C::NI; ++j) { csum[j][0] = 0.0; csum[j][1] = 0.0; }
#pragma unroll
        for (int i = 0; i < C::MI; ++i) {
            const int row = m0 + i * 8 + lr;
            if (row < p.m_valid) {
                double *dst = p.Cout + (int64_t)row * p.ldc + col0 + n0 + 2 * lk;
#pragma unroll
                for (int j = 0; j < C::NI; ++j) {
                    *reinterpret_cast<double2 *>(dst + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
            }
            // rows >= m_valid are exactly zero (At zero padded), so they add nothing below
#pragma unroll
            for (int j = 0; j < C::NI; ++j) {
                csum[j][0] = fma(acc[i][j][0], acc[i][j][0], csum[j][0]);
                csum[j][1] = fma(acc[i][j][1], acc[i][j][1], csum[j][1]);
            }
        }
        // reduce over the 8 row groups of the warp (lanes with equal lk)
#pragma unroll
        for (int j = 0; j < C::NI; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = csum[j][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                csum[j][e] = v;
            }
        double *red = smem;   // [WM][N_TILE]
        if (lr == 0) {
#pragma unroll
            for (int j = 0; j < C::NI; ++j) {
                red[wm * C::N_TILE + n0 + j * 8 + 2 * lk] = csum[j][0];
                red[wm * C::N_TILE + n0 + j * 8 + 2 * lk + 1] = csum[j][1];
            }
        }
        __syncthreads();
        for (int n = tid; n < C::N_TILE; n += C::THREADS) {
            double sres = 0.0;
#pragma unroll
            for (int w = 0; w < C::WM; ++w) sres += red[w * C::N_TILE + n];
            double dv = p.diag[col0 + n] - sres;
            p.diag[col0 + n] = dv > 0.0 ? dv : 0.0;      // .clip(min=0); NaN -> 0 never arises from finite data
        }
    } else {
        double xn[C::NI][2];
#pragma unroll
        for (int j = 0; j < C::NI; ++j) {
            xn[j][0] = p.xnorm[col0 + n0 + j * 8 + 2 * lk];
            xn[j][1] = p.xnorm[col0 + n0 + j * 8 + 2 * lk + 1];
        }
#pragma unroll
        for (int i = 0; i < C::MI; ++i) {
            const int row = m0 + i * 8 + lr;
            if (row < p.m_valid) {
                const double pn = p.pnorm[row];
                double *dst = p.Cout + (int64_t)row * p.ldc + col0 + n0 + 2 * lk;
#pragma unroll
                for (int j = 0; j < C::NI; ++j) {
                    // sklearn: D2 = ((-2 * X.Y^T) + XX) + YY, clamped at 0
                    double d0 = (-2.0 * acc[i][j][0] + pn) + xn[j][0];
                    double d1 = (-2.0 * acc[i][j][1] + pn) + xn[j][1];
                    d0 = d0 > 0.0 ? d0 : 0.0;
                    d1 = d1 > 0.0 ? d1 : 0.0;
                    *reinterpret_cast<double2 *>(dst + j * 8) =
                        make_double2(kern_from_d2(p.kp, d0), kern_from_d2(p.kp, d1));
                }
            }
        }
    }
}

}  // namespace dg
