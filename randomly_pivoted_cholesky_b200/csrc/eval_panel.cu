// Kernel-row evaluator with the PIVOT PANEL RESIDENT in shared memory (Gaussian / Matern in sklearn's expanded form,
// kernels.py:32-39, 71-85):   out[s x n_pad] = k(Xpiv, X)   with   D2 = ((-2 Xpiv.X^T) + |xpiv|^2) + |x|^2.
//
// dg::gemm_kernel<MODE_EVALX> keeps the X tile resident and streams the panel through the stage ring once per 64-column
// tile: at d = 50 that is 80 KB of panel copies and 4 barrier round trips per tile for 13 k4-steps of DMMAs (ncu, round 1:
// 13% of the stall samples sat on the stage barriers, 19% on the global loads of the row norms in the epilogue).  Here the
// roles are swapped, which is the natural order for a SHORT contraction (K = d):
//   * the panel Xpiv^T (d x s, leading dimension = the tile's LDA) and the pivot norms are loaded ONCE per persistent CTA;
//   * the point tiles X[col0 : col0 + N_TILE, :] (one contiguous block: X is point-major) and their norms stream through a
//     2-3 deep ring of TMA bulk copies issued by a producer warp (full/empty mbarriers per buffer);
//   * a consumer warp runs its whole k loop without any barrier, releases the buffer, and only then does the epilogue.
//     (Tried: starting the second warp of every SM sub-partition one k loop late so that one warp's DMMAs overlap the other's
//     epilogue on the shared FP64 pipe -- no gain, 0.834 vs 0.817 ms at s = 167: the eight-chain epilogue is pipe bound itself.)
// Epilogue (Gaussian): exp(g D2), g = -1/(2 bw^2), evaluated as exp(min(fma(x.y, -2g, g|xpiv|^2 + g|x|^2), 0)) with the
// 32-entry-table exponential below: 13 FP64-pipe instructions per entry instead of 25 (the DMMAs cost 2d/16 = 6.5 pipe
// cycles per entry at d = 52, the old epilogue 3.1).  The scaled form differs from the reference's rounding sequence by a
// few ulp of g(|xpiv|^2 + |x|^2), i.e. ~1e-16 relative in the kernel value (parity bar: 1e-10).
#include "dmma_gemm.cuh"

using namespace dg;

namespace ep {

struct EvalParams {
    const double *Pan;      // kp x LDA: Xpiv^T zero padded to the row tile
    int kp;                 // contraction length (d_pad, a multiple of 4)
    const double *pnorm;    // [m_valid] squared norms of the pivots
    int m_valid;
    const double *X;        // n_pad x ldx points
    int64_t ldx;
    const double *xnorm;    // [n_pad]
    KernParams kparm;
    double *out;
    int64_t ldo;
    int ntiles;             // n_pad / N_TILE
    int xb;                 // X buffers in the ring (2 or 3)
    // REDUCE variant (fused kernel-times-vector): out[col] = (accumulate ? out[col] : 0) + sum_row w[row] k(pivot row, X col);
    // `out` is then a vector of n_pad entries and no kernel value is written to memory
    const double *w;        // [m_valid]
    int accumulate;
};

// Eight scaled arguments g * D2 (not yet clamped) -> kernel values -> two groups of two double2 stores: either two fragment
// rows of an NI = 2 tile (dstA, dstB) or the two halves of one row of an NI = 4 tile (dstB = dstA + 16).  Eight interleaved
// chains per call: with four the dependent FP64 chain (about 12 instructions of ~8 cycles) was latency bound (ncu: 440 cycles
// per call for 112 cycles of FP64 pipe).  Out of line: inlined at every fragment row the epilogue outgrows the instruction cache.
static __device__ __noinline__ void gauss_tab_store8(double *dstA, int okA, double *dstB, int okB, double tab_lane, double x0, double x1,
                                                     double x2, double x3, double x4, double x5, double x6, double x7) {
    double x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = x[i] < 0.0 ? x[i] : 0.0;          // D2 clamped at 0 (sklearn), in the scaled domain
    exp_tab_batch<8>(x, tab_lane);
    if (okA) {
        *reinterpret_cast<double2 *>(dstA) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2 *>(dstA + 8) = make_double2(x[2], x[3]);
    }
    if (okB) {
        *reinterpret_cast<double2 *>(dstB) = make_double2(x[4], x[5]);
        *reinterpret_cast<double2 *>(dstB + 8) = make_double2(x[6], x[7]);
    }
}
static __device__ __noinline__ void gauss_tab_store4(double *dst, int ok, double tab_lane, double x0, double x1, double x2, double x3) {
    double x[4] = {x0, x1, x2, x3};
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = x[i] < 0.0 ? x[i] : 0.0;
    exp_tab_batch<4>(x, tab_lane);
    if (ok) {
        *reinterpret_cast<double2 *>(dst) = make_double2(x[0], x[1]);
        *reinterpret_cast<double2 *>(dst + 8) = make_double2(x[2], x[3]);
    }
}

// REDUCE variant: eight scaled arguments (two fragment rows A, B of an NI = 2 tile, columns c0..c3 each) -> kernel values ->
// the thread's four running weighted column sums, returned in registers
static __device__ __noinline__ double4 gauss_tab_wsum8(double tab_lane, double wA, double wB, double4 cs, double x0, double x1, double x2,
                                                       double x3, double x4, double x5, double x6, double x7) {
    double x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = x[i] < 0.0 ? x[i] : 0.0;
    exp_tab_batch<8>(x, tab_lane);
    cs.x = fma(wB, x[4], fma(wA, x[0], cs.x));
    cs.y = fma(wB, x[5], fma(wA, x[1], cs.y));
    cs.z = fma(wB, x[6], fma(wA, x[2], cs.z));
    cs.w = fma(wB, x[7], fma(wA, x[3], cs.w));
    return cs;
}

template <class C>
struct Smem {
    // doubles: panel | pivot norms | X ring | norm ring | barriers
    static size_t doubles(int kp, int64_t ldx, int xb) {
        return (size_t)kp * C::LDA + 2 * C::M_TILE + C::WM * C::N_TILE + (size_t)xb * (C::N_TILE * ldx + C::N_TILE) + 16;
    }
};

// acc += Pan[:, rows of the warp]^T X_tile[cols of the warp, :]^T over the whole contraction; fragments double buffered, two
// k4-steps per trip, no barrier inside
template <class C, int CMI>
__device__ __forceinline__ void k_loop(double (&acc)[C::MI][C::NI][2], const double *Ap, const double *Bp, int ldx, int lk, int nks) {
    double a0[C::MI], b0[C::NI], a1[C::MI], b1[C::NI];
    auto load = [&](double (&a)[C::MI], double (&bb)[C::NI], int ks) {
        const int kk = ks * 4 + lk;
#pragma unroll
        for (int i = 0; i < CMI; ++i) a[i] = Ap[kk * C::LDA + i * C::RSTEP];
#pragma unroll
        for (int j = 0; j < C::NI; ++j) bb[j] = Bp[j * 8 * ldx + kk];
    };
    auto mma = [&](const double (&a)[C::MI], const double (&bb)[C::NI]) {
#pragma unroll
        for (int i = 0; i < CMI; ++i)
#pragma unroll
            for (int j = 0; j < C::NI; ++j) dmma884(acc[i][j], a[i], bb[j]);
    };
    load(a0, b0, 0);
    int ks = 0;
    for (; ks + 1 < nks; ks += 2) {
        load(a1, b1, ks + 1);
        mma(a0, b0);
        if (ks + 2 < nks) load(a0, b0, ks + 2);
        mma(a1, b1);
    }
    if (nks & 1) mma(a0, b0);
}

template <class C, bool GAUSS, bool REDUCE>
__global__ void __launch_bounds__(C::THREADS, 1) eval_panel_kernel(const EvalParams p) {
    static_assert(!REDUCE || C::NI == 2, "the fused kernel-times-vector variant is written for the NI = 2 tiles");
    extern __shared__ __align__(16) double smem[];
    const int ldx = (int)p.ldx;
    const int xtile = C::N_TILE * ldx;
    double *Pan = smem;
    double *pns = Pan + (size_t)p.kp * C::LDA;
    double *wts = pns + C::M_TILE;                                              // REDUCE: weights of the rows (0 past m_valid)
    double *red = wts + C::M_TILE;                                              // REDUCE: [WM][N_TILE] column sums of the warp rows
    double *Xs = red + C::WM * C::N_TILE;
    double *xns = Xs + (size_t)p.xb * xtile;
    uint64_t *bars = reinterpret_cast<uint64_t *>(xns + p.xb * C::N_TILE);      // [0] panel, [1 + b] full, [1 + xb + b] empty
    uint64_t *xfull = bars + 1, *xempty = bars + 1 + p.xb;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double g = GAUSS ? p.kparm.gscale : 1.0;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        for (int b = 0; b < p.xb; ++b) {
            mbar_init(&xfull[b], 1);
            mbar_init(&xempty[b], C::CONSUMERS);
        }
        fence_barrier_init();
    }
    for (int r = tid; r < C::M_TILE; r += C::THREADS) {
        pns[r] = r < p.m_valid ? g * p.pnorm[r] : 0.0;
        wts[r] = (REDUCE && r < p.m_valid) ? p.w[r] : 0.0;
    }
    __syncthreads();

    if (warp >= C::CONSUMERS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp != C::CONSUMERS || lane != 0) return;
        // ------------------------------- producer: one thread, one or two bulk copies per tile -------------------
        const unsigned pan_bytes = (unsigned)((size_t)p.kp * C::LDA * 8);
        mbar_expect_tx(&bars[0], pan_bytes);
        bulk_g2s(Pan, p.Pan, pan_bytes, &bars[0]);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
            const int b = it % p.xb;
            mbar_wait(&xempty[b], ((it / p.xb) & 1) ^ 1);             // first lap: returns immediately
            const int64_t col0 = (int64_t)tile * C::N_TILE;
            mbar_expect_tx(&xfull[b], (unsigned)((xtile + C::N_TILE) * 8));
            bulk_g2s(Xs + (size_t)b * xtile, p.X + col0 * p.ldx, (unsigned)(xtile * 8), &xfull[b]);
            bulk_g2s(xns + b * C::N_TILE, p.xnorm + col0, (unsigned)(C::N_TILE * 8), &xfull[b]);
        }
        return;
    }

    // ----------------------------------- consumer warps ------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    int wm, wn;
    if (C::ODD) {
        const int half = warp >> 2, sp = warp & 3;
        wm = (sp + half) & 1;
        wn = (sp >> 1) + 2 * half;
    } else {
        wm = warp % C::WM;
        wn = warp / C::WM;
    }
    const int n0 = wn * C::NI * 8;
    const int lr = lane >> 2, lk = lane & 3;
    constexpr int CMI_FULL = C::MI;
    const int mi_cnt = (C::ODD && wm == 1) ? C::MI - 1 : C::MI;
    const int m0 = wm * C::WARP_ROW0;
    const int nks = p.kp >> 2;
    const double tab_lane = c_exp2_tab[lane];
    const double m2g = -2.0 * g;
    const double *Ap = Pan + m0 + lr;

    mbar_wait(&bars[0], 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int b = it % p.xb;
        const int64_t col0 = (int64_t)tile * C::N_TILE;
        mbar_wait(&xfull[b], (it / p.xb) & 1);
        const double *Bp = Xs + (size_t)b * xtile + (size_t)(n0 + lr) * ldx;

        double acc[C::MI][C::NI][2];
#pragma unroll
        for (int i = 0; i < C::MI; ++i)
#pragma unroll
            for (int j = 0; j < C::NI; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        // k loop over the whole contraction (the fragment count is a compile-time constant of each branch, as in the Schur kernel)
        if (C::ODD && mi_cnt != C::MI) k_loop<C, (C::ODD ? C::MI - 1 : C::MI)>(acc, Ap, Bp, ldx, lk, nks);
        else k_loop<C, C::MI>(acc, Ap, Bp, ldx, lk, nks);

        // the tile's (scaled) point norms, then the buffer goes back to the producer
        double xn[C::NI][2];
#pragma unroll
        for (int j = 0; j < C::NI; ++j) {
            const double2 v = *reinterpret_cast<const double2 *>(xns + b * C::N_TILE + n0 + j * 8 + 2 * lk);
            xn[j][0] = g * v.x;
            xn[j][1] = g * v.y;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&xempty[b]);

        if (REDUCE) {
            // weighted column sums over the rows of the tile: in the thread, over the 8 row groups of the warp (shuffles), over
            // the warp rows (shared memory), in a fixed order
            double4 cs = make_double4(0.0, 0.0, 0.0, 0.0);          // columns n0 + {2lk, 2lk+1, 8+2lk, 8+2lk+1}
            constexpr int J1 = C::NI > 1 ? 1 : 0;
#pragma unroll
            for (int i = 0; i < C::MI; i += 2) {
                const int rA = m0 + i * C::RSTEP;
                if (i < mi_cnt && rA < p.m_valid) {                         // warp uniform
                    const int rowA = rA + lr;
                    const int i1 = (i + 1 < C::MI) ? i + 1 : i;
                    const bool pairB = i + 1 < C::MI && i + 1 < mi_cnt;
                    const int rowB = pairB ? rowA + C::RSTEP : rowA;
                    const double pnA = pns[rowA], pnB = pns[rowB];
                    const double wA = wts[rowA], wB = pairB ? wts[rowB] : 0.0;
                    if (GAUSS) {
                        cs = gauss_tab_wsum8(tab_lane, wA, wB, cs, fma(acc[i][0][0], m2g, pnA + xn[0][0]), fma(acc[i][0][1], m2g, pnA + xn[0][1]),
                                             fma(acc[i][J1][0], m2g, pnA + xn[J1][0]), fma(acc[i][J1][1], m2g, pnA + xn[J1][1]),
                                             fma(acc[i1][0][0], m2g, pnB + xn[0][0]), fma(acc[i1][0][1], m2g, pnB + xn[0][1]),
                                             fma(acc[i1][J1][0], m2g, pnB + xn[J1][0]), fma(acc[i1][J1][1], m2g, pnB + xn[J1][1]));
                    } else {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int ih = h ? i1 : i;
                            const double pn = h ? pnB : pnA, wv = h ? wB : wA;
                            double d00 = (-2.0 * acc[ih][0][0] + pn) + xn[0][0], d01 = (-2.0 * acc[ih][0][1] + pn) + xn[0][1];
                            double d10 = (-2.0 * acc[ih][J1][0] + pn) + xn[J1][0], d11 = (-2.0 * acc[ih][J1][1] + pn) + xn[J1][1];
                            const double2 v0 = kern_from_d2_pair(p.kparm, d00 > 0.0 ? d00 : 0.0, d01 > 0.0 ? d01 : 0.0);
                            const double2 v1 = kern_from_d2_pair(p.kparm, d10 > 0.0 ? d10 : 0.0, d11 > 0.0 ? d11 : 0.0);
                            cs.x = fma(wv, v0.x, cs.x); cs.y = fma(wv, v0.y, cs.y);
                            cs.z = fma(wv, v1.x, cs.z); cs.w = fma(wv, v1.y, cs.w);
                        }
                    }
                }
            }
            double c4[4] = {cs.x, cs.y, cs.z, cs.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double v = c4[q];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                c4[q] = v;
            }
            consumer_bar_sync(C::CONSUMERS * 32);                  // the previous tile's readers of `red` are done
            if (lr == 0) {
                red[wm * C::N_TILE + n0 + 2 * lk] = c4[0];
                red[wm * C::N_TILE + n0 + 2 * lk + 1] = c4[1];
                red[wm * C::N_TILE + n0 + 8 * J1 + 2 * lk] = c4[2];
                red[wm * C::N_TILE + n0 + 8 * J1 + 2 * lk + 1] = c4[3];
            }
            consumer_bar_sync(C::CONSUMERS * 32);
            for (int n = tid; n < C::N_TILE; n += C::CONSUMERS * 32) {
                double sres = 0.0;
#pragma unroll
                for (int wq = 0; wq < C::WM; ++wq) sres += red[wq * C::N_TILE + n];
                p.out[col0 + n] = (p.accumulate ? p.out[col0 + n] : 0.0) + sres;
            }
            continue;
        }
        double *out0 = p.out + col0 + n0 + 2 * lk;
        if (GAUSS && C::NI == 2) {
            // two fragment rows (8 values) per call
#pragma unroll
            for (int i = 0; i < C::MI; i += 2) {
                const int rA = m0 + i * C::RSTEP, rB = rA + C::RSTEP;
                if (i < mi_cnt && rA < p.m_valid) {                         // warp uniform: all lanes take part in the shuffles
                    const int rowA = rA + lr, rowB = rB + lr;
                    const double pnA = pns[rowA];
                    if (i + 1 < C::MI && i + 1 < mi_cnt) {
                        const double pnB = pns[rowB];
                        const int i1 = i + 1 < C::MI ? i + 1 : i;
                        gauss_tab_store8(out0 + (int64_t)rowA * p.ldo, rowA < p.m_valid, out0 + (int64_t)rowB * p.ldo, rowB < p.m_valid,
                                         tab_lane, fma(acc[i][0][0], m2g, pnA + xn[0][0]), fma(acc[i][0][1], m2g, pnA + xn[0][1]),
                                         fma(acc[i][1][0], m2g, pnA + xn[1][0]), fma(acc[i][1][1], m2g, pnA + xn[1][1]),
                                         fma(acc[i1][0][0], m2g, pnB + xn[0][0]), fma(acc[i1][0][1], m2g, pnB + xn[0][1]),
                                         fma(acc[i1][1][0], m2g, pnB + xn[1][0]), fma(acc[i1][1][1], m2g, pnB + xn[1][1]));
                    } else {
                        gauss_tab_store4(out0 + (int64_t)rowA * p.ldo, rowA < p.m_valid, tab_lane, fma(acc[i][0][0], m2g, pnA + xn[0][0]),
                                         fma(acc[i][0][1], m2g, pnA + xn[0][1]), fma(acc[i][1][0], m2g, pnA + xn[1][0]),
                                         fma(acc[i][1][1], m2g, pnA + xn[1][1]));
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < C::MI; ++i) {
                const int rbase = m0 + i * C::RSTEP;
                if (i < mi_cnt && rbase < p.m_valid) {
                    const int row = rbase + lr;
                    const int ok = row < p.m_valid;
                    const double pn = pns[row];
                    double *dst = out0 + (int64_t)row * p.ldo;
                    if (GAUSS) {                                            // NI == 4: one row = 8 values
                        gauss_tab_store8(dst, ok, dst + 16, ok, tab_lane, fma(acc[i][0][0], m2g, pn + xn[0][0]),
                                         fma(acc[i][0][1], m2g, pn + xn[0][1]), fma(acc[i][1 % C::NI][0], m2g, pn + xn[1 % C::NI][0]),
                                         fma(acc[i][1 % C::NI][1], m2g, pn + xn[1 % C::NI][1]), fma(acc[i][2 % C::NI][0], m2g, pn + xn[2 % C::NI][0]),
                                         fma(acc[i][2 % C::NI][1], m2g, pn + xn[2 % C::NI][1]), fma(acc[i][3 % C::NI][0], m2g, pn + xn[3 % C::NI][0]),
                                         fma(acc[i][3 % C::NI][1], m2g, pn + xn[3 % C::NI][1]));
                    } else if (ok) {
#pragma unroll
                        for (int j = 0; j < C::NI; ++j) {
                            // sklearn: D2 = ((-2 * X.Y^T) + XX) + YY, clamped at 0
                            double d0 = (-2.0 * acc[i][j][0] + pn) + xn[j][0], d1 = (-2.0 * acc[i][j][1] + pn) + xn[j][1];
                            d0 = d0 > 0.0 ? d0 : 0.0;
                            d1 = d1 > 0.0 ? d1 : 0.0;
                            *reinterpret_cast<double2 *>(dst + j * 8) = kern_from_d2_pair(p.kparm, d0, d1);
                        }
                    }
                }
            }
        }
    }
}

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <class C, bool REDUCE = false>
static int launch(rpc_ctx *ctx, cudaStream_t st, EvalParams p, int64_t n_pad) {
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(eval_panel_kernel<C, true, REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        RPC_CUDA(cudaFuncSetAttribute(eval_panel_kernel<C, false, REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        configured |= 1ull << (ctx->device & 63);
    }
    p.ntiles = (int)(n_pad / C::N_TILE);
    p.xb = Smem<C>::doubles(p.kp, p.ldx, 3) * 8 <= SMEM_LIMIT ? 3 : 2;
    const size_t shm = Smem<C>::doubles(p.kp, p.ldx, p.xb) * 8;
    const int ctas = rpc_persistent_ctas(ctx);
    const unsigned grid = (unsigned)(p.ntiles < ctas ? p.ntiles : ctas);
    if (p.kparm.kind == RPC_KERN_GAUSSIAN) eval_panel_kernel<C, true, REDUCE><<<grid, C::THREADS, shm, st>>>(p);
    else eval_panel_kernel<C, false, REDUCE><<<grid, C::THREADS, shm, st>>>(p);
    RPC_LAUNCHED(ctx);
    return 0;
}

template <class C>
static bool fits(int kp, int64_t ldx) { return Smem<C>::doubles(kp, ldx, 2) * 8 <= SMEM_LIMIT; }

// the row tiles of the Schur update (csrc/big_kernels.cu): 8-row steps above 128 rows
#define RPC_EP_FINE(X) X(136, 9, true) X(144, 9, false) X(152, 10, true) X(160, 10, false) X(168, 11, true) X(176, 11, false) \
    X(184, 12, true) X(192, 12, false) X(200, 13, true) X(208, 13, false) X(216, 14, true) X(224, 14, false) X(232, 15, true)   \
    X(240, 15, false) X(248, 16, true)
typedef Cfg<4, 2, 8, 4> E256;
typedef Cfg<2, 4, 8, 4> E128;
typedef Cfg<2, 4, 6, 4> E96;
typedef Cfg<2, 4, 4, 4> E64;
typedef Cfg<1, 8, 4, 2> E32;
typedef Cfg<1, 8, 2, 2> E16;

}  // namespace ep

// rows tile for m <= 256 rows of the evaluator (same rule as the Schur update)
static int ep_tile_rows(int m) {
    if (m > 128) return m <= 248 ? (m + 7) / 8 * 8 : 256;
    const int tiles[] = {16, 32, 64, 96, 128};
    for (int t : tiles) if (m <= t) return t;
    return 128;
}

// leading dimension the panel must be built with for m rows, or 0 when the shape does not fit in shared memory
int rpc_eval_panel_ld(int m, int kp, int64_t ldx) {
    using namespace ep;
    switch (ep_tile_rows(m)) {
#define RPC_EP_LD(M, MI, ODD) case M: return fits<Cfg<2, 4, MI, 2, ODD>>(kp, ldx) ? Cfg<2, 4, MI, 2, ODD>::LDA : 0;
        RPC_EP_FINE(RPC_EP_LD)
#undef RPC_EP_LD
        case 16: return fits<E16>(kp, ldx) ? E16::LDA : 0;
        case 32: return fits<E32>(kp, ldx) ? E32::LDA : 0;
        case 64: return fits<E64>(kp, ldx) ? E64::LDA : 0;
        case 96: return fits<E96>(kp, ldx) ? E96::LDA : 0;
        case 128: return fits<E128>(kp, ldx) ? E128::LDA : 0;
        default: return fits<E256>(kp, ldx) ? E256::LDA : 0;
    }
}

// out[m x n_pad] = k(Xpiv, X) for m <= 256 pivots whose transposed, zero-padded panel Pan (kp x rpc_eval_panel_ld) is given
int rpc_eval_panel_rows(rpc_ctx *ctx, cudaStream_t st, const KernParams &kparm, const double *X, const double *xnorm, int64_t n_pad,
                        int kp, int64_t ldx, const double *Pan, const double *pnorm, int m, double *out, int64_t ldo) {
    using namespace ep;
    EvalParams p{};
    p.Pan = Pan; p.kp = kp; p.pnorm = pnorm; p.m_valid = m; p.X = X; p.ldx = ldx; p.xnorm = xnorm; p.kparm = kparm;
    p.out = out; p.ldo = ldo;
    switch (ep_tile_rows(m)) {
#define RPC_EP_CASE(M, MI, ODD) case M: return launch<Cfg<2, 4, MI, 2, ODD>>(ctx, st, p, n_pad);
        RPC_EP_FINE(RPC_EP_CASE)
#undef RPC_EP_CASE
        case 16: return launch<E16>(ctx, st, p, n_pad);
        case 32: return launch<E32>(ctx, st, p, n_pad);
        case 64: return launch<E64>(ctx, st, p, n_pad);
        case 96: return launch<E96>(ctx, st, p, n_pad);
        case 128: return launch<E128>(ctx, st, p, n_pad);
        default: return launch<E256>(ctx, st, p, n_pad);
    }
}

// ---- fused kernel-times-vector: out[col] (+)= sum_{i < m} w[i] k(Xpiv_i, X_col), no kernel value written to memory ----------
// row tiles of the REDUCE variant (all NI = 2): 16, 32, 64, 128, 192, 248
namespace ep {
typedef Cfg<2, 4, 4, 2> R64;
typedef Cfg<2, 4, 8, 2> R128;
typedef Cfg<2, 4, 12, 2> R192;
typedef Cfg<2, 4, 16, 2, true> R248;
}  // namespace ep

static int ep_wsum_tile_rows(int m) {
    const int tiles[] = {16, 32, 64, 128, 192, 248};
    for (int t : tiles) if (m <= t) return t;
    return 248;
}

int rpc_eval_panel_wsum_ld(int m, int kp, int64_t ldx) {
    using namespace ep;
    switch (ep_wsum_tile_rows(m)) {
        case 16: return fits<E16>(kp, ldx) ? E16::LDA : 0;
        case 32: return fits<E32>(kp, ldx) ? E32::LDA : 0;
        case 64: return fits<R64>(kp, ldx) ? R64::LDA : 0;
        case 128: return fits<R128>(kp, ldx) ? R128::LDA : 0;
        case 192: return fits<R192>(kp, ldx) ? R192::LDA : 0;
        default: return fits<R248>(kp, ldx) ? R248::LDA : 0;
    }
}

int rpc_eval_panel_wsum(rpc_ctx *ctx, cudaStream_t st, const KernParams &kparm, const double *X, const double *xnorm, int64_t n_pad,
                        int kp, int64_t ldx, const double *Pan, const double *pnorm, int m, const double *w, double *out, int accumulate) {
    using namespace ep;
    EvalParams p{};
    p.Pan = Pan; p.kp = kp; p.pnorm = pnorm; p.m_valid = m; p.X = X; p.ldx = ldx; p.xnorm = xnorm; p.kparm = kparm;
    p.out = out; p.ldo = 0; p.w = w; p.accumulate = accumulate;
    switch (ep_wsum_tile_rows(m)) {
        case 16: return launch<E16, true>(ctx, st, p, n_pad);
        case 32: return launch<E32, true>(ctx, st, p, n_pad);
        case 64: return launch<R64, true>(ctx, st, p, n_pad);
        case 128: return launch<R128, true>(ctx, st, p, n_pad);
        case 192: return launch<R192, true>(ctx, st, p, n_pad);
        default: return launch<R248, true>(ctx, st, p, n_pad);
    }
}
