// FP64 tensor-core (DMMA.8x8x4 via mma.sync) skinny GEMM core for sm_100a.
//
//   C[M x n_pad] = At^T[M x K] * B[K x n_pad]        M <= 256 (all rows of the update in one CTA)
//
// One CTA owns ALL M rows of a column tile, so the big operand B (the factor G, or the data X) is read
// from HBM exactly once per launch and the epilogue (residual-diagonal update / kernel function) is
// fused.  The small operand At (the pivot panel, K x M, zero padded) is re-read by every CTA from L2.
//
// Pipeline (Blackwell/Hopper style, no CTA-wide barrier in the main loop):
//   * one PRODUCER warp stages the operands with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx):
//     per k-stage of KT=16 rows, lane r issues one row of the pivot panel (M_TILE*8 bytes) and one row of
//     B (N_TILE*8 bytes) into a 4-deep ring of shared-memory stages guarded by full/empty mbarriers;
//   * WM x WN CONSUMER warps wait on the stage's full barrier, run MI x NI DMMA.8x8x4 per k4-step from
//     double-buffered register fragments, and release the stage with one arrive per warp.
// Shared-memory leading dimensions are == 4 (mod 16) doubles so that the DMMA fragment loads
// (lane -> (k = lane&3, row = lane>>2)) hit 16 distinct 8-byte bank pairs per half warp.
//
// tcgen05 has no FP64 kind, so on Blackwell the FP64 tensor path is the warp-level mma.sync DMMA;
// measured peak on B200: 36.9 TFLOP/s (profiles/fp64_peaks_r01.json), the same pipe as DFMA.
#pragma once
#include "rpc_common.cuh"

#ifndef RPC_ROW_INTERLEAVE
#define RPC_ROW_INTERLEAVE 1     // 0: contiguous row blocks per warp row (diagnostic builds only)
#endif

namespace dg {

constexpr int KT = RPC_KT;      // k-depth per stage
constexpr int STAGES = 4;
constexpr int LDBK = KT + 4;    // point-major B stage: [n][LDBK]

enum { MODE_SCHUR = 0, MODE_EVAL = 1, MODE_EVALX = 3 };   // EVALX: EVAL with the X tile resident in shared memory

// ODD (only with WM == 2): the second warp row owns MI-1 fragments, so tiles come in 8-row steps.  The two warps an
// SM sub-partition hosts (w and w+4) are then given different warp rows, which keeps the four DMMA pipes balanced.
template <int WM_, int WN_, int MI_, int NI_, bool ODD_ = false>
struct Cfg {
    static constexpr int WM = WM_, WN = WN_, MI = MI_, NI = NI_;
    static constexpr bool ODD = ODD_;
    static_assert(!ODD_ || (WM_ == 2 && WN_ == 4), "ODD tiles are 2 x 4 warps");
    static constexpr int CONSUMERS = WM * WN;                // consumer warps
    static constexpr int THREADS = (CONSUMERS + 4) * 32;     // + one producer warpgroup (one warp of it works)
    static_assert(CONSUMERS % 4 == 0, "consumer warps must form whole warpgroups (setmaxnreg, even SMSP load)");
    static constexpr int M_TILE = (WM * MI - (ODD ? 1 : 0)) * 8;
    static constexpr int N_TILE = WN * NI * 8;
    // Row fragments are dealt to the warp rows round robin: fragment i of warp row wm covers rows (i * WM + wm) * 8 .. + 7.
    // The panel of the Schur update ends in the transpose of a triangular inverse (zero above the diagonal); with
    // interleaved rows every warp row owns the same share of the non-zero triangle (contiguous row blocks would leave
    // the last warp row with three times the work of the first).
    static constexpr int RSTEP = RPC_ROW_INTERLEAVE ? WM * 8 : 8;
    static constexpr int WARP_ROW0 = RPC_ROW_INTERLEAVE ? 8 : MI * 8;          // first row of warp row wm = wm * WARP_ROW0
    static constexpr int LDA = (M_TILE % 16 == 0) ? M_TILE + 4 : M_TILE + 12;   // == 4 (mod 16)
    static constexpr int LDB = N_TILE + 4;
    static constexpr int A_STAGE = KT * LDA;                                  // doubles
    static constexpr int B_STAGE_ROW = KT * LDB;                              // k-major B stage
    static constexpr int B_STAGE_PT = N_TILE * LDBK;                          // point-major B stage
    static constexpr int B_STAGE = B_STAGE_ROW > B_STAGE_PT ? B_STAGE_ROW : B_STAGE_PT;
    static constexpr int STAGE_DOUBLES = STAGES * (A_STAGE + B_STAGE);
    static constexpr int NBAR = 2 * STAGES + 4;                               // full[], empty[], xfull, xempty, 2 spare
    static constexpr int RED_DOUBLES = WM * N_TILE;                           // column-sum scratch of the epilogue
    static constexpr size_t SMEM_BYTES = (size_t)(STAGE_DOUBLES + NBAR + RED_DOUBLES) * sizeof(double);
    // MODE_EVALX keeps the whole X tile resident: N_TILE x ldx doubles (+16 zero doubles of slack)
    static constexpr size_t fused_smem_bytes(int ldx) { return SMEM_BYTES + ((size_t)N_TILE * ldx + 16) * sizeof(double); }
    static_assert(M_TILE % 8 == 0 && N_TILE % 16 == 0, "tile shape");
    static_assert(LDA % 16 == 4 && LDB % 16 == 4, "bank-conflict-free leading dimensions");
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
    const int32_t *rowmap;   // optional: new row i is row rowmap[i] of R (rows evaluated speculatively for ALL proposals)
    int K;
    // tri != 0: panel rows kk >= c hold the transpose of a LOWER-triangular s x s matrix (L^{-T}): At[c + i, j] == 0 for
    // j < i, so output row j only needs the k range kk <= c + j and the DMMAs of the structurally-zero half are skipped.
    // row_base = panel column (= row of the whole update) of this launch's row 0 (row chunks of updates with s > 256).
    int tri, row_base;
    // rows_out != NULL (speculative row evaluation: R holds the rows of ALL proposals, the new rows are picked through rowmap):
    // the kernel also writes new row i to rows_out[i, :] -- the B stages it fetches for kk >= c ARE those rows, so the
    // compaction rows[c:c+s] = rows_all[acc] costs no extra pass over memory (it used to be a separate copy kernel that
    // collided with the panel solve or the next sampler)
    double *rows_out;
    int64_t ld_rows_out;
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
    // MODE_EVALX: rows = k(Xsel, X) with the panel At_e = Xsel^T (ke_pad x ldat_e) streamed and the X tile resident; output R
    const double *At_e;
    int64_t ldat_e;
    int ke_pad;
    int ntiles;            // column tiles (n_pad / N_TILE); the kernel is persistent over them
    // MODE_SCHUR, row-sharded runs: the updated residual diagonal of this rank's columns is ALSO stored straight into
    // every rank's copy of the full diagonal (peer memory over NVLink; entry r already points at this rank's block),
    // and the last CTA to finish publishes `epoch` in each rank's flag word for this source (csrc/peers.cu)
    // MODE_SCHUR, tail launch (chunk_rows > 0): the "tile" index runs over (row chunk, column tile) pairs of the last partial
    // wave of column tiles, so that wave is spread over all SMs.  Chunk q covers panel rows [rstart, rstart + M_TILE) with
    // rstart = min(q * chunk_rows, chunk_last_start); rows below q * chunk_rows are duplicates of the previous chunk and are
    // neither stored nor summed.  The column sums of squares go to sq_out[q * sq_ld + column] (the diagonal is updated by a
    // finishing kernel once all chunks are in), tile_col0 is the first column of the tail.
    int64_t tile_col0;
    int chunk_rows, chunk_last_start, ncol_tiles;
    int64_t chunk_panel_stride;   // > 0: At holds one re-packed panel per chunk (ld = LDA of the tail tile: one bulk copy per stage)
    double *sq_out;
    int64_t sq_ld;
    int publish;           // with npeers > 0: the last CTA publishes the epoch (0: a later kernel does)
    int npeers;
    double *peer_diag[RPC_MAX_PEERS];
    unsigned long long *peer_flag[RPC_MAX_PEERS];
    unsigned long long epoch;
    unsigned int *done_counter;
};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// ---- mbarrier / TMA bulk-copy primitives (PTX ISA 8.x, sm_90+) ----------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA engine, no tensor map), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (TMA store), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void consumer_bar_sync(int nthreads) {
    asm volatile("bar.sync 1, %0;\n" ::"r"(nthreads) : "memory");
}

// bytes of the panel part of a stage: one contiguous block (padding columns included) when the panel was
// built with leading dimension LDA, else KT separate rows
template <class C>
__device__ __forceinline__ unsigned panel_bytes(int64_t ldat) {
    return ldat == C::LDA ? KT * C::LDA * 8 : KT * C::M_TILE * 8;
}
template <class C>
__device__ __forceinline__ void produce_panel(double *As, const double *At, int64_t ldat, int k0, uint64_t *full, int lane) {
    if (ldat == C::LDA) {          // the TMA engine is rate limited per copy (~40 ns): one 33 KB copy, not 16 rows
        if (lane == 0) bulk_g2s(As, At + (int64_t)k0 * ldat, KT * C::LDA * 8, full);
    } else if (lane < KT) {
        bulk_g2s(As + lane * C::LDA, At + (int64_t)(k0 + lane) * ldat, C::M_TILE * 8, full);
    }
}

template <class C, int MODE>
__device__ __forceinline__ unsigned stage_bytes(const Params &p, int k0) {
    unsigned a = panel_bytes<C>(p.ldat);
    if (MODE == MODE_SCHUR) return a + KT * C::N_TILE * 8;
    int kb = p.d - k0;
    kb = kb > KT ? KT : kb;
    return a + (unsigned)C::N_TILE * kb * 8;
}

// the producer warp: the panel stage in one copy, lane r issues row r of the B tile
template <class C, int MODE>
__device__ __forceinline__ void produce_stage(const Params &p, const double *At, double *As, double *Bs, uint64_t *full, int k0,
                                              int64_t col0, int lane) {
    if (lane == 0) mbar_expect_tx(full, stage_bytes<C, MODE>(p, k0));
    __syncwarp();
    produce_panel<C>(As, At, p.ldat, k0, full, lane);
    if (MODE == MODE_SCHUR) {
        if (lane >= 32 - KT) {
            const int r = lane - (32 - KT);
            const int kk = k0 + r;
            const double *src;
            if (kk < p.c) src = p.G + (int64_t)kk * p.ldg;
            else if (kk < p.K) src = p.R + (int64_t)(p.rowmap ? p.rowmap[kk - p.c] : kk - p.c) * p.ldr;
            else src = p.R;                       // padded k: At is zero there, any finite row will do
            bulk_g2s(Bs + r * C::LDB, src + col0, C::N_TILE * 8, full);
        }
    } else {
        int kb = p.d - k0;
        kb = kb > KT ? KT : kb;                   // features left (multiple of 4); the rest of the stage row keeps
        for (int n = lane; n < C::N_TILE; n += 32)  // stale-but-finite data that meets zero rows of At
            bulk_g2s(Bs + n * LDBK, p.X + (col0 + n) * p.ldx + k0, (unsigned)kb * 8, full);
    }
}

// one k-stage of DMMAs with double-buffered fragments; `ksteps` (1..4) k4-steps hold non-zero panel rows.
// BKIND 0: B stage is k-major [KT][LDB]; BKIND 1: point-major [N_TILE][LDBK]; BKIND 2: resident X tile [N_TILE][ldx]
template <class C, int BKIND, int CMI>
__device__ __forceinline__ void consume_stage_n(double (&acc)[C::MI][C::NI][2], const double *As, const double *Bs, int ldx,
                                                int lk, int ksteps) {
    constexpr int mi_cnt = CMI;
    double a[2][C::MI], b[2][C::NI];
    auto ldb = [&](int j, int kk) -> double {
        if (BKIND == 0) return Bs[kk * C::LDB + j * 8];
        if (BKIND == 1) return Bs[j * 8 * LDBK + kk];
        return Bs[j * 8 * ldx + kk];
    };
#pragma unroll
    for (int i = 0; i < C::MI; ++i) a[0][i] = i < mi_cnt ? As[lk * C::LDA + i * C::RSTEP] : 0.0;
#pragma unroll
    for (int j = 0; j < C::NI; ++j) b[0][j] = ldb(j, lk);
#pragma unroll
    for (int ks = 0; ks < KT / 4; ++ks) {
        const int cur = ks & 1, nxt = cur ^ 1;
        if (ks + 1 < KT / 4 && ks + 1 < ksteps) {
            const int kk = (ks + 1) * 4 + lk;
#pragma unroll
            for (int i = 0; i < C::MI; ++i) a[nxt][i] = i < mi_cnt ? As[kk * C::LDA + i * C::RSTEP] : 0.0;
#pragma unroll
            for (int j = 0; j < C::NI; ++j) b[nxt][j] = ldb(j, kk);
        }
        if (ks < ksteps) {
#pragma unroll
            for (int i = 0; i < C::MI; ++i)
                if (i < mi_cnt) {                      // warp-uniform: this warp owns mi_cnt row fragments
#pragma unroll
                    for (int j = 0; j < C::NI; ++j) dmma884(acc[i][j], a[cur][i], b[cur][j]);
                }
        }
    }
}

// warp rows of an ODD tile own MI or MI-1 fragments: two instantiations, chosen by a warp-uniform branch
template <class C, int BKIND>
__device__ __forceinline__ void consume_stage(double (&acc)[C::MI][C::NI][2], const double *As, const double *Bs, int ldx,
                                              int lk, int ksteps, int mi_cnt) {
    if (C::ODD && mi_cnt != C::MI) consume_stage_n<C, BKIND, (C::ODD ? C::MI - 1 : C::MI)>(acc, As, Bs, ldx, lk, ksteps);
    else consume_stage_n<C, BKIND, C::MI>(acc, As, Bs, ldx, lk, ksteps);
}

// A k-stage inside the triangular block of the Schur panel (k-major B).  Fragment i of the warp (rows r_i .. r_i + 7 of the
// update, r_i = first row of the warp + i * RSTEP) meets non-zero panel entries only in stages that start at or below panel row
// c + r_i + 7, so the fragments that take part in a stage are a SUFFIX i >= first of the warp's fragments.  ptxas turns a
// predicated mma.sync into a branch around it, and a branch per DMMA pair stalls the pipe (measured: slower than the dense
// stage), so the stage is written fragment-major -- per fragment 4 k4-steps x NI DMMAs from the stage's B fragments, the next
// fragment's A values in flight -- and entered at fragment `first` through ONE switch with fall-through cases: one indirect
// branch per stage, no code duplication, and inside a fragment every DMMA depends on the one NI instructions earlier only.
template <class C, int CMI, int I>
__device__ __forceinline__ void tri_fragment(double (&acc)[C::MI][C::NI][2], double (&a)[2][KT / 4], const double (&b)[KT / 4][C::NI],
                                             const double *As, int lk) {
    if (I < CMI) {
        if (I + 1 < CMI) {
#pragma unroll
            for (int ks = 0; ks < KT / 4; ++ks) a[(I + 1) & 1][ks] = As[(ks * 4 + lk) * C::LDA + (I + 1) * C::RSTEP];
        }
#pragma unroll
        for (int ks = 0; ks < KT / 4; ++ks)
#pragma unroll
            for (int j = 0; j < C::NI; ++j) dmma884(acc[I < C::MI ? I : 0][j], a[I & 1][ks], b[ks][j]);
    }
}

template <class C, int CMI>
__device__ __forceinline__ void consume_stage_tri_n(double (&acc)[C::MI][C::NI][2], const double *As, const double *Bs, int lk,
                                                    int first) {
    static_assert(C::MI <= 16, "the fall-through switch below lists 16 fragments");
    if (first >= CMI) return;                       // this warp's rows all lie above the stage: nothing to add
    double a[2][KT / 4], b[KT / 4][C::NI];
#pragma unroll
    for (int ks = 0; ks < KT / 4; ++ks) {
#pragma unroll
        for (int j = 0; j < C::NI; ++j) b[ks][j] = Bs[(ks * 4 + lk) * C::LDB + j * 8];
        a[0][ks] = As[(ks * 4 + lk) * C::LDA + first * C::RSTEP];      // the entry fragment's values, in both register sets
        a[1][ks] = a[0][ks];
    }
#define RPC_TRI_CASE(I) case I: tri_fragment<C, CMI, I>(acc, a, b, As, lk); [[fallthrough]];
    switch (first) {
        RPC_TRI_CASE(0) RPC_TRI_CASE(1) RPC_TRI_CASE(2) RPC_TRI_CASE(3) RPC_TRI_CASE(4) RPC_TRI_CASE(5) RPC_TRI_CASE(6) RPC_TRI_CASE(7)
        RPC_TRI_CASE(8) RPC_TRI_CASE(9) RPC_TRI_CASE(10) RPC_TRI_CASE(11) RPC_TRI_CASE(12) RPC_TRI_CASE(13) RPC_TRI_CASE(14)
        RPC_TRI_CASE(15)
        default: break;
    }
#undef RPC_TRI_CASE
}

template <class C>
__device__ __forceinline__ void consume_stage_tri(double (&acc)[C::MI][C::NI][2], const double *As, const double *Bs, int lk,
                                                  int mi_cnt, int first) {
    if (C::ODD && mi_cnt != C::MI) consume_stage_tri_n<C, (C::ODD ? C::MI - 1 : C::MI)>(acc, As, Bs, lk, first);
    else consume_stage_tri_n<C, C::MI>(acc, As, Bs, lk, first);
}

// kernel values from the x.y accumulators -> Cout (MODE_EVAL, MODE_EVALX)
template <class C>
__device__ __forceinline__ void eval_epilogue(const Params &p, double (&acc)[C::MI][C::NI][2], double *out, int64_t ldo,
                                              int64_t col0, int m0, int n0, int lr, int lk, int mi_cnt) {
    double xn[C::NI][2];
#pragma unroll
    for (int j = 0; j < C::NI; ++j) {
        xn[j][0] = p.xnorm[col0 + n0 + j * 8 + 2 * lk];
        xn[j][1] = p.xnorm[col0 + n0 + j * 8 + 2 * lk + 1];
    }
#pragma unroll
    for (int i = 0; i < C::MI; ++i) {
        const int row = m0 + i * C::RSTEP + lr;
        if (i < mi_cnt && row < p.m_valid) {
            const double pn = p.pnorm[row];
            double *dst = out + (int64_t)row * ldo + col0 + n0 + 2 * lk;
            // sklearn: D2 = ((-2 * X.Y^T) + XX) + YY, clamped at 0
            double dd[C::NI][2];
#pragma unroll
            for (int j = 0; j < C::NI; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double v = (-2.0 * acc[i][j][e] + pn) + xn[j][e];
                    dd[j][e] = v > 0.0 ? v : 0.0;
                }
            if (p.kp.kind == RPC_KERN_GAUSSIAN && (C::NI == 4 || C::NI == 2)) {
                // the whole fragment row in one out-of-line call: 2*NI interleaved exp chains
                if (C::NI == 4)
                    gauss_store8(p.kp.gscale, dst, dd[0][0], dd[0][1], dd[1][0], dd[1][1], dd[2 % C::NI][0], dd[2 % C::NI][1],
                                 dd[3 % C::NI][0], dd[3 % C::NI][1]);
                else
                    gauss_store4(p.kp.gscale, dst, dd[0][0], dd[0][1], dd[1][0], dd[1][1]);
            } else {
#pragma unroll
                for (int j = 0; j < C::NI; ++j)
                    *reinterpret_cast<double2 *>(dst + j * 8) = kern_from_d2_pair(p.kp, dd[j][0], dd[j][1]);
            }
        }
    }
}

// k4-steps of stage kt that hold real (non-padded) panel rows
__device__ __forceinline__ int live_ksteps(int kt, int k_real) {
    int left = k_real - kt * KT;
    return left >= KT ? KT / 4 : (left + 3) / 4;
}

template <class C, int MODE>
__global__ void __launch_bounds__(C::THREADS, 1) gemm_kernel(const Params p) {
    extern __shared__ __align__(16) double smem[];
    double *As_base = smem;
    double *Bs_base = smem + STAGES * C::A_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::STAGE_DOUBLES);
    uint64_t *empty = full + STAGES;
    uint64_t *xfull = empty + STAGES;
    uint64_t *xempty = xfull + 1;
    double *red = smem + C::STAGE_DOUBLES + C::NBAR;         // [WM][N_TILE]
    double *Xs = red + C::RED_DOUBLES;                       // MODE_EVALX only: [N_TILE][ldx] + 16 zeros

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nk = p.k_pad / KT;
    constexpr bool XRES = (MODE == MODE_EVALX);              // evaluation with the X tile resident
    const int nke = XRES ? p.ke_pad / KT : 0;
    const int nk2 = (MODE == MODE_EVALX) ? 0 : nk;
    const int ldx = (int)p.ldx;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], C::CONSUMERS + ((MODE == MODE_SCHUR && p.rows_out) ? 1 : 0));     // + the row-store warp
        }
        mbar_init(xfull, 1);
        mbar_init(xempty, C::CONSUMERS);
        fence_barrier_init();
    }
    if (MODE == MODE_EVAL) {
        // partial-row copies of the last k tile leave the tail of a stage row untouched: make it finite
        for (int q = tid; q < STAGES * C::B_STAGE; q += C::THREADS) Bs_base[q] = 0.0;
        fence_proxy_async();
    }
    if (XRES) {
        // the k loop of phase 1 runs to ke_pad >= ldx: the last point's reads run into this slack, which
        // must be finite because it meets zero rows of the panel
        if (tid < 16) Xs[C::N_TILE * ldx + tid] = 0.0;
    }
    __syncthreads();

    if (warp >= C::CONSUMERS) {
        // the producer warpgroup hands its registers to the consumers (the kernel launches with 168/thread)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (MODE == MODE_SCHUR && p.rows_out && warp == C::CONSUMERS + 1) {
            // ------------------------- row-store warp (speculative rows only) ------------------------------------------
            // The B stages fetched for panel rows kk >= c ARE the accepted rows A[S, tile]: this warp waits for each stage like
            // a consumer and sends those rows on to rows_out with TMA bulk stores (lane r = stage row r), so the compaction
            // rows[c:c+s] = rows_all[acc] costs neither a pass over memory nor an instruction of the DMMA warps.
            int it = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                int ct = tile, chunk = 0;
                if (p.chunk_rows) {
                    chunk = tile / p.ncol_tiles;
                    ct = tile - chunk * p.ncol_tiles;
                }
                const int64_t col0 = p.tile_col0 + (int64_t)ct * C::N_TILE;
                for (int kt = 0; kt < nk; ++kt, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full[s], (it / STAGES) & 1);
                    if (chunk == 0 && kt * KT + KT > p.c) {            // every row chunk of a tail tile streams the same rows: one stores
                        const int kk = kt * KT + lane;
                        if (lane < KT && kk >= p.c && kk < p.K)
                            bulk_s2g(p.rows_out + (int64_t)(kk - p.c) * p.ld_rows_out + col0, Bs_base + s * C::B_STAGE + lane * C::LDB,
                                     C::N_TILE * 8);
                        bulk_commit();
                        bulk_wait_read0();                               // the stage may be refilled once the reads are done
                        __syncwarp();
                    }
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
            }
            bulk_wait0();
            return;
        }
        if (warp != C::CONSUMERS) return;
        // ------------------------------- producer warp: runs ahead over the tiles -------------------
        int it = 0;
        unsigned tpar = 0;                                   // parity of this CTA's tile counter
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, tpar ^= 1) {
            int ct = tile, rstart = 0;
            if (MODE == MODE_SCHUR && p.chunk_rows) {        // tail launch: (row chunk, column tile)
                const int chunk = tile / p.ncol_tiles;
                ct = tile - chunk * p.ncol_tiles;
                rstart = min(chunk * p.chunk_rows, p.chunk_last_start);
            }
            const int64_t col0 = p.tile_col0 + (int64_t)ct * C::N_TILE;
            const double *At_tile = p.At + rstart;
            if (MODE == MODE_SCHUR && p.chunk_rows && p.chunk_panel_stride) At_tile = p.At + (int64_t)(tile / p.ncol_tiles) * p.chunk_panel_stride;
            if (XRES) {
                mbar_wait(xempty, tpar ^ 1);                 // consumers are done with the previous X tile
                if (lane == 0) {
                    const unsigned bytes = (unsigned)(C::N_TILE * ldx * 8);
                    mbar_expect_tx(xfull, bytes);
                    bulk_g2s(Xs, p.X + col0 * p.ldx, bytes, xfull);   // the whole X tile in one bulk copy
                }
                for (int kt = 0; kt < nke; ++kt, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    if (lane == 0) mbar_expect_tx(&full[s], panel_bytes<C>(p.ldat_e));
                    __syncwarp();
                    produce_panel<C>(As_base + s * C::A_STAGE, p.At_e, p.ldat_e, kt * KT, &full[s], lane);
                }
            }
            for (int kt = 0; kt < nk2; ++kt, ++it) {
                const int s = it % STAGES;
                mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);   // first lap: returns immediately
                produce_stage<C, MODE>(p, At_tile, As_base + s * C::A_STAGE, Bs_base + s * C::B_STAGE, &full[s], kt * KT, col0, lane);
            }
        }
        return;
    }

    // ----------------------------------- consumer warps ------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    // Warp w runs on SM sub-partition w % 4.  ODD tiles pair an MI-fragment warp with an (MI-1)-fragment warp on every
    // sub-partition: wm = (w + w/4) % 2.  (Dealing fragments at RUN time with predicated DMMAs was measured 2% slower
    // than padding; here the fragment count is a compile-time constant of each branch.)
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
    const int mi_cnt = (C::ODD && wm == 1) ? C::MI - 1 : C::MI;
    const int m0 = wm * C::WARP_ROW0;                        // fragment i: rows m0 + i * RSTEP .. + 7
    const int k_real = (MODE == MODE_EVAL) ? p.d : p.K;

    int it = 0;
    unsigned tpar = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, tpar ^= 1) {
        int ct = tile, chunk = 0, rstart = 0, first_new = 0;
        if (MODE == MODE_SCHUR && p.chunk_rows) {            // tail launch: (row chunk, column tile)
            chunk = tile / p.ncol_tiles;
            ct = tile - chunk * p.ncol_tiles;
            rstart = min(chunk * p.chunk_rows, p.chunk_last_start);
            first_new = chunk * p.chunk_rows - rstart;
        }
        const int64_t col0 = p.tile_col0 + (int64_t)ct * C::N_TILE;
        const int m_valid = p.m_valid - rstart;              // rows of this tile that exist
        double acc[C::MI][C::NI][2];
#pragma unroll
        for (int i = 0; i < C::MI; ++i)
#pragma unroll
            for (int j = 0; j < C::NI; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        if (XRES) {
            // ---- rows tile = k(Xsel, X tile) with the X tile resident, written to R
            mbar_wait(xfull, tpar);
            for (int kt = 0; kt < nke; ++kt, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                consume_stage<C, 2>(acc, As_base + s * C::A_STAGE + m0 + lr, Xs + (n0 + lr) * ldx + kt * KT, ldx, lk,
                                    live_ksteps(kt, p.d), mi_cnt);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            if (lane == 0) mbar_arrive(xempty);            // the X tile may be overwritten by the next one
            eval_epilogue<C>(p, acc, const_cast<double *>(p.R), p.ldr, col0, m0, n0, lr, lk, mi_cnt);
            continue;
        }
        for (int kt = 0; kt < nk2; ++kt, ++it) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            if (MODE == MODE_EVAL)
                consume_stage<C, 1>(acc, As_base + s * C::A_STAGE + m0 + lr, Bs_base + s * C::B_STAGE + (n0 + lr) * LDBK, 0,
                                    lk, live_ksteps(kt, k_real), mi_cnt);
            else {
                // rel0 > 0: the stage starts below the diagonal of this warp's first fragment(s); fragment i is needed iff
                // i * RSTEP >= rel0 (a stage past K's last real row meets zero panel rows and finite B rows: harmless)
                const int rel0 = kt * KT - p.c - p.row_base - rstart - m0 - 7;
                if (MODE == MODE_SCHUR && p.tri && rel0 > 0)
                    consume_stage_tri<C>(acc, As_base + s * C::A_STAGE + m0 + lr, Bs_base + s * C::B_STAGE + n0 + lr, lk, mi_cnt,
                                         (rel0 + C::RSTEP - 1) / C::RSTEP);
                else
                    consume_stage<C, 0>(acc, As_base + s * C::A_STAGE + m0 + lr, Bs_base + s * C::B_STAGE + n0 + lr, 0, lk,
                                        live_ksteps(kt, k_real), mi_cnt);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);         // this warp is done reading the stage
        }

        // C fragment: row = m0 + i*RSTEP + lr, cols = n0 + j*8 + 2*lk + {0,1}
        if (MODE != MODE_EVAL) {
            double csum[C::NI][2];
#pragma unroll
            for (int j = 0; j < C::NI; ++j) { csum[j][0] = 0.0; csum[j][1] = 0.0; }
#pragma unroll
            for (int i = 0; i < C::MI; ++i) {
                const int row = m0 + i * C::RSTEP + lr;
                const bool live = i < mi_cnt && row < m_valid && row >= first_new;   // fragments >= mi_cnt do not exist
                if (live) {
                    double *dst = p.Cout + (int64_t)(rstart + row) * p.ldc + col0 + n0 + 2 * lk;
#pragma unroll
                    for (int j = 0; j < C::NI; ++j)
                        *reinterpret_cast<double2 *>(dst + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
                // rows >= m_valid are exactly zero (At zero padded), so they add nothing below; a tail chunk also holds
                // duplicate rows (< first_new) and, past the panel's last column, rows fed by the next panel row
                if (!(MODE == MODE_SCHUR && p.chunk_rows) || live) {
#pragma unroll
                    for (int j = 0; j < C::NI; ++j) {
                        csum[j][0] = fma(acc[i][j][0], acc[i][j][0], csum[j][0]);
                        csum[j][1] = fma(acc[i][j][1], acc[i][j][1], csum[j][1]);
                    }
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
            consumer_bar_sync(C::CONSUMERS * 32);          // previous tile's readers of `red` are done
            if (lr == 0) {
#pragma unroll
                for (int j = 0; j < C::NI; ++j) {
                    red[wm * C::N_TILE + n0 + j * 8 + 2 * lk] = csum[j][0];
                    red[wm * C::N_TILE + n0 + j * 8 + 2 * lk + 1] = csum[j][1];
                }
            }
            consumer_bar_sync(C::CONSUMERS * 32);
            for (int n = tid; n < C::N_TILE; n += C::CONSUMERS * 32) {
                double sres = 0.0;
#pragma unroll
                for (int w = 0; w < C::WM; ++w) sres += red[w * C::N_TILE + n];
                if (MODE == MODE_SCHUR && p.sq_out) {      // tail chunk: the finishing kernel updates the diagonal
                    p.sq_out[chunk * p.sq_ld + (col0 - p.tile_col0) + n] = sres;
                    continue;
                }
                double dv = p.diag[col0 + n] - sres;
                dv = dv > 0.0 ? dv : 0.0;                  // .clip(min=0)
                p.diag[col0 + n] = dv;
                if (MODE == MODE_SCHUR)
                    for (int r = 0; r < p.npeers; ++r) p.peer_diag[r][col0 + n] = dv;   // the diag "all-gather", tile by tile
            }
        } else {
            eval_epilogue<C>(p, acc, p.Cout, p.ldc, col0, m0, n0, lr, lk, mi_cnt);
        }
    }
    if (MODE == MODE_SCHUR && p.npeers > 0 && p.publish) {
        // release pattern: this CTA's peer stores -> system fence -> CTA counter; the last CTA publishes the epoch
        consumer_bar_sync(C::CONSUMERS * 32);
        if (tid == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(p.done_counter, 1u);
            if (prev == gridDim.x - 1) {
                __threadfence_system();
                *p.done_counter = 0;                                   // every CTA has arrived: re-arm for the next launch
                for (int r = 0; r < p.npeers; ++r)
                    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p.peer_flag[r]), "l"(p.epoch) : "memory");
            }
        }
    }
}

}  // namespace dg
