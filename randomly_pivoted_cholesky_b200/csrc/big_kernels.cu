// N-proportional kernels: kernel-row evaluation (K1/K2), fused Schur update (K5), fused simple step (K4).
#include <stdlib.h>
#include "dmma_gemm.cuh"

using namespace dg;

// ------------------------------------------------------------------------------------------------
// DMMA GEMM dispatch: pick the narrowest M tile that holds the rows of the update
// ------------------------------------------------------------------------------------------------
constexpr size_t SMEM_LIMIT = 227 * 1024;

template <class C, int MODE>
static int launch_gemm(rpc_ctx *ctx, cudaStream_t st, const Params &p, int64_t n_pad) {
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    size_t shm = (MODE == MODE_EVALX) ? C::fused_smem_bytes((int)p.ldx) : C::SMEM_BYTES;
    if (shm > SMEM_LIMIT) {
        rpc_set_error("launch_gemm: %zu bytes of shared memory needed, %zu available", shm, SMEM_LIMIT);
        return -4;
    }
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(gemm_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (MODE == MODE_EVALX) ? (int)SMEM_LIMIT : (int)C::SMEM_BYTES));
        configured |= 1ull << (ctx->device & 63);
    }
    // persistent: one CTA per SM walks the column tiles; the producer warp prefetches across tile boundaries
    Params q = p;
    q.ntiles = (int)(n_pad / C::N_TILE);
    const int ctas = rpc_persistent_ctas(ctx);
    unsigned grid = (unsigned)(q.ntiles < ctas ? q.ntiles : ctas);
    gemm_kernel<C, MODE><<<grid, C::THREADS, shm, st>>>(q);
    RPC_LAUNCHED(ctx);
    return 0;
}

// row tiles (M_TILE = 8*WM*MI).  Every tile has 8 consumer warps (two per SM sub-partition: 5- and 3-row-warp
// shapes were measured 15-20% slower because the DMMA work no longer divides evenly over the 4 SMSPs).
typedef Cfg<4, 2, 8, 4> C256;   // 256 x 64
typedef Cfg<4, 2, 7, 4> C224;   // 224 x 64
typedef Cfg<4, 2, 6, 4> C192;   // 192 x 64
typedef Cfg<4, 2, 5, 4> C160;   // 160 x 64
// above 128 rows: 2 x 4 warps of (MI x 2) fragments, N tile 64; ODD tiles give 8-row steps (Schur update only)
#define RPC_FINE_TILES(X) X(136, 9, true) X(144, 9, false) X(152, 10, true) X(160, 10, false) X(168, 11, true) \
    X(176, 11, false) X(184, 12, true) X(192, 12, false) X(200, 13, true) X(208, 13, false) X(216, 14, true) \
    X(224, 14, false) X(232, 15, true) X(240, 15, false) X(248, 16, true)                                       \
    X(72, 5, true) X(80, 5, false) X(88, 6, true) X(104, 7, true) X(112, 7, false) X(120, 8, true)
#define RPC_DECL_TILE(M, MI, ODD) typedef Cfg<2, 4, MI, 2, ODD> F##M;
RPC_FINE_TILES(RPC_DECL_TILE)
#undef RPC_DECL_TILE
// short updates (the truncated last round of a factorisation, small blocks): 1 x 8 warps of (MI x 2) fragments, N tile 128
#define RPC_SHORT_TILES(X) X(24, 3) X(40, 5) X(48, 6) X(56, 7)
#define RPC_DECL_SHORT(M, MI) typedef Cfg<1, 8, MI, 2> F##M;
RPC_SHORT_TILES(RPC_DECL_SHORT)
#undef RPC_DECL_SHORT
typedef Cfg<2, 4, 8, 4> C128;   // 128 x 128
typedef Cfg<2, 4, 6, 4> C96;    //  96 x 128
typedef Cfg<2, 4, 4, 4> C64;    //  64 x 128
typedef Cfg<1, 8, 4, 2> C32;    //  32 x 128
typedef Cfg<1, 8, 2, 2> C16;    //  16 x 128

static int tile_rows_for(int m, bool fine = false) {
    if (fine && m > 16) return m <= 256 ? (m + 7) / 8 * 8 : 256;          // 24, 32, ..., 256 (Schur update: 8-row steps)
    const int tiles[] = {16, 32, 64, 96, 128, 160, 192, 224, 256};
    for (int t : tiles) if (m <= t) return t;
    return 256;
}

template <int MODE>
static int dispatch_gemm(rpc_ctx *ctx, cudaStream_t st, const Params &p, int m, int64_t n_pad) {
    if (MODE == MODE_SCHUR) {
        switch (tile_rows_for(m, true)) {
#define RPC_CASE_TILE(M, MI, ODD) case M: return launch_gemm<F##M, MODE_SCHUR>(ctx, st, p, n_pad);
            RPC_FINE_TILES(RPC_CASE_TILE)
#undef RPC_CASE_TILE
#define RPC_CASE_SHORT(M, MI) case M: return launch_gemm<F##M, MODE_SCHUR>(ctx, st, p, n_pad);
            RPC_SHORT_TILES(RPC_CASE_SHORT)
#undef RPC_CASE_SHORT
            default: break;
        }
    }
    switch (tile_rows_for(m)) {
        case 16: return launch_gemm<C16, MODE>(ctx, st, p, n_pad);
        case 32: return launch_gemm<C32, MODE>(ctx, st, p, n_pad);
        case 64: return launch_gemm<C64, MODE>(ctx, st, p, n_pad);
        case 96: return launch_gemm<C96, MODE>(ctx, st, p, n_pad);
        case 128: return launch_gemm<C128, MODE>(ctx, st, p, n_pad);
        case 160: return launch_gemm<C160, MODE>(ctx, st, p, n_pad);
        case 192: return launch_gemm<C192, MODE>(ctx, st, p, n_pad);
        case 224: return launch_gemm<C224, MODE>(ctx, st, p, n_pad);
        default: return launch_gemm<C256, MODE>(ctx, st, p, n_pad);
    }
}

// leading dimension of a panel whose k-stages can be fetched with ONE bulk copy by the tile chosen for m rows
static int panel_ld_for(int m, bool fine = false) {
    if (fine) {
        switch (tile_rows_for(m, true)) {
#define RPC_LD_TILE(M, MI, ODD) case M: return F##M::LDA;
            RPC_FINE_TILES(RPC_LD_TILE)
#undef RPC_LD_TILE
#define RPC_LD_SHORT(M, MI) case M: return F##M::LDA;
            RPC_SHORT_TILES(RPC_LD_SHORT)
#undef RPC_LD_SHORT
            default: break;
        }
    }
    switch (tile_rows_for(m)) {
        case 16: return C16::LDA;
        case 32: return C32::LDA;
        case 64: return C64::LDA;
        case 96: return C96::LDA;
        case 128: return C128::LDA;
        case 160: return C160::LDA;
        case 192: return C192::LDA;
        case 224: return C224::LDA;
        default: return C256::LDA;
    }
}
extern "C" int rpc_panel_ld(int s) { return s <= 256 ? panel_ld_for(s, true) : (s + 255) / 256 * 256; }

// does the X-resident evaluator (MODE_EVALX) fit in shared memory for this row count / point stride?
static bool fused_fits(int m, int64_t ldx) {
    int t = tile_rows_for(m);
    size_t need;
    if (t <= 32) need = (t == 16 ? C16::fused_smem_bytes((int)ldx) : C32::fused_smem_bytes((int)ldx));
    else if (t == 64) need = C64::fused_smem_bytes((int)ldx);
    else if (t == 96) need = C96::fused_smem_bytes((int)ldx);
    else if (t == 128) need = C128::fused_smem_bytes((int)ldx);
    else if (t == 160) need = C160::fused_smem_bytes((int)ldx);
    else if (t == 192) need = C192::fused_smem_bytes((int)ldx);
    else if (t == 224) need = C224::fused_smem_bytes((int)ldx);
    else need = C256::fused_smem_bytes((int)ldx);
    return need <= SMEM_LIMIT;
}

// ---- tail of the column-tile waves (see schur_update_impl) -------------------------------------------------------------
struct TailFinish {
    const double *sq;          // [nchunk][sq_ld] column sums of squares of the row chunks
    int64_t sq_ld;
    int nchunk, ncols;
    double *diag;              // first tail column
    int npeers;
    double *peer_diag[RPC_MAX_PEERS];
    unsigned long long *peer_flag[RPC_MAX_PEERS];
    unsigned long long epoch;
};

__global__ void __launch_bounds__(1024) schur_tail_finish_kernel(TailFinish f) {
    for (int col = threadIdx.x; col < f.ncols; col += blockDim.x) {
        double tot = 0.0;
        for (int q = 0; q < f.nchunk; ++q) tot += f.sq[q * f.sq_ld + col];       // fixed order: deterministic
        double dv = f.diag[col] - tot;
        dv = dv > 0.0 ? dv : 0.0;                                               // .clip(min=0)
        f.diag[col] = dv;
        for (int r = 0; r < f.npeers; ++r) f.peer_diag[r][col] = dv;
    }
    if (f.npeers > 0) {
        // the main launch stored its columns into the peers' copies before it completed; this single CTA publishes
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            for (int r = 0; r < f.npeers; ++r)
                asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(f.peer_flag[r]), "l"(f.epoch) : "memory");
        }
    }
}

// one panel per row chunk with the leading dimension of the tail tile, so that a 16-row stage of it is ONE bulk copy (with
// the main panel's stride the producer warp needs 16 row copies per stage and the TMA issue rate bounds the tail CTAs)
__global__ void __launch_bounds__(256) tail_panel_repack_kernel(const double *__restrict__ At, int64_t ldat, int k_pad, int nchunk,
                                                                int chunk_rows, int chunk_last_start, int ld_out,
                                                                double *__restrict__ out) {
    const int64_t total = (int64_t)nchunk * k_pad * ld_out;
    for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < total; q += (int64_t)gridDim.x * 256) {
        const int j = (int)(q % ld_out);
        const int64_t t = q / ld_out;
        const int r = (int)(t % k_pad), chunk = (int)(t / k_pad);
        const int rstart = min(chunk * chunk_rows, chunk_last_start);
        out[q] = (j < chunk_rows && rstart + j < ldat) ? At[(int64_t)r * ldat + rstart + j] : 0.0;
    }
}

static bool g_no_tail = getenv("RPC_B200_NO_TAIL") != nullptr;       // diagnostics: keep every column tile in the main launch
static bool g_no_evalp = getenv("RPC_B200_NO_EVALP") != nullptr;     // diagnostics: evaluator with the panel streamed (MODE_EVALX)
static bool g_no_tri = getenv("RPC_B200_NO_TRI") != nullptr;         // diagnostics: run the zero half of the triangular block too

static int schur_update_impl(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At, int64_t ldat,
                             int k_pad, const double *rows_new, int64_t ldr, const int32_t *rowmap, double *diag, int64_t n_pad,
                             const rpc_peers *peers, unsigned long long epoch, int panel_lower, double *rows_out = nullptr,
                             int64_t ld_rows_out = 0) {
    RPC_CHECK_ARG(ctx && G && At && rows_new && diag, "null pointer");
    RPC_CHECK_ARG(!rows_out || (ld_rows_out >= n_pad && ld_rows_out % 2 == 0), "rows_out needs an even leading dimension >= n_pad");
    const int tri = (panel_lower && !g_no_tri) ? 1 : 0;
    RPC_CHECK_ARG(s >= 1 && c >= 0, "s >= 1, c >= 0");
    RPC_CHECK_ARG(n_pad > 0 && n_pad % RPC_COL_ALIGN == 0, "n_pad must be a positive multiple of RPC_COL_ALIGN");
    RPC_CHECK_ARG(k_pad % RPC_KT == 0 && k_pad >= c + s, "k_pad must be a multiple of RPC_KT and >= c+s");
    RPC_CHECK_ARG(ldg % 2 == 0 && ldr % 2 == 0 && ldat % 2 == 0, "leading dimensions must be even (16-byte rows)");
    if (peers) {
        RPC_CHECK_ARG(peers->world >= 1 && peers->world <= RPC_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
                      "bad peer table");
        RPC_CHECK_ARG(n_pad <= peers->shard, "the local shard is wider than the peers' slot for it");
        int rc = rpc_sync_reserve(ctx);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // Tail of the column-tile waves.  The persistent grid walks n_pad/64 column tiles; when they do not fill the last wave
    // (8-GPU shard of T*: 1954 tiles on 148 SMs = 13.2 waves) the leftover tiles are taken out of the main launch and done
    // by a second launch of 32-row x 128-column CTAs over (row chunk, column tile) pairs, which spreads them over all SMs
    // (about half a tile-time instead of a full one).  The chunks leave their column sums of squares in the workspace and
    // schur_tail_finish_kernel applies them to the diagonal in a fixed order (deterministic).
    if (s > 128 && s <= 256 && !g_no_tail) {
        const int ctas = rpc_persistent_ctas(ctx);
        const int ntiles = (int)(n_pad / 64);
        const int left = ntiles % ctas;
        const int nchunk = (s + 31) / 32;
        const int ncol = left / 2;                                  // n_pad is a multiple of 128, so `left` is even
        if (left > 0 && left % 2 == 0 && ntiles / ctas >= 4 && ncol * nchunk <= ctas && ldat >= 32 && tile_rows_for(s, true) <= ldat) {
            const int64_t tail_cols = (int64_t)left * 64, col_lo = n_pad - tail_cols;
            const size_t sq_doubles = ((size_t)nchunk * tail_cols + 15) & ~(size_t)15;      // keeps the panels 128-byte aligned
            const size_t panel_doubles = (size_t)nchunk * k_pad * C32::LDA;
            int rc = rpc_ws_reserve(ctx, (sq_doubles + panel_doubles) * sizeof(double));
            if (rc) return rc;
            double *tail_panels = ctx->ws + sq_doubles;
            Params p{};
            p.At = At; p.ldat = ldat; p.k_pad = k_pad; p.m_valid = s;
            p.G = G; p.ldg = ldg; p.c = c; p.R = rows_new; p.ldr = ldr; p.K = c + s; p.rowmap = rowmap;
            p.Cout = G + (int64_t)c * ldg; p.ldc = ldg; p.diag = diag;
            p.tri = tri; p.row_base = 0;
            p.rows_out = rows_out; p.ld_rows_out = ld_rows_out;
            if (peers) {
                p.npeers = peers->world;
                for (int r = 0; r < peers->world; ++r) {
                    RPC_CHECK_ARG(peers->diag_full[r] && peers->flags[r], "peer table has null entries");
                    p.peer_diag[r] = peers->diag_full[r] + (int64_t)peers->rank * peers->shard;
                }
                p.publish = 0;                                      // the finishing kernel publishes
            }
            // main launch: the full waves
            rc = dispatch_gemm<MODE_SCHUR>(ctx, st, p, s, col_lo);
            if (rc) return rc;
            // tail launch
            Params t = p;
            t.npeers = 0;
            t.tile_col0 = col_lo; t.chunk_rows = 32; t.ncol_tiles = ncol;
            t.chunk_last_start = (int)((ldat - 32) & ~(int64_t)1);
            t.sq_out = ctx->ws; t.sq_ld = tail_cols;
            {
                const int64_t total = (int64_t)panel_doubles;
                int grid = (int)((total + 256 * 4 - 1) / (256 * 4));
                if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
                tail_panel_repack_kernel<<<grid, 256, 0, st>>>(At, ldat, k_pad, nchunk, 32, t.chunk_last_start, C32::LDA, tail_panels);
                RPC_LAUNCHED(ctx);
            }
            t.At = tail_panels; t.ldat = C32::LDA; t.chunk_panel_stride = (int64_t)k_pad * C32::LDA;
            rc = launch_gemm<C32, MODE_SCHUR>(ctx, st, t, (int64_t)ncol * nchunk * C32::N_TILE);
            if (rc) return rc;
            TailFinish f{};
            f.sq = ctx->ws; f.sq_ld = tail_cols; f.nchunk = nchunk; f.diag = diag + col_lo; f.ncols = (int)tail_cols;
            f.npeers = p.npeers;
            for (int r = 0; r < p.npeers; ++r) {
                f.peer_diag[r] = p.peer_diag[r] + col_lo;
                f.peer_flag[r] = peers->flags[r] + peers->rank;
            }
            f.epoch = epoch;
            schur_tail_finish_kernel<<<1, 1024, 0, st>>>(f);
            RPC_LAUNCHED(ctx);
            return 0;
        }
    }
    for (int m0 = 0; m0 < s; m0 += 256) {
        int m = s - m0 < 256 ? s - m0 : 256;
        RPC_CHECK_ARG(ldat >= m0 + tile_rows_for(m, true), "ldat too small for the zero-padded row tile");
        Params p{};
        p.At = At + m0; p.ldat = ldat; p.k_pad = k_pad; p.m_valid = m;
        p.G = G; p.ldg = ldg; p.c = c; p.R = rows_new; p.ldr = ldr; p.K = c + s; p.rowmap = rowmap;
        p.Cout = G + (int64_t)(c + m0) * ldg; p.ldc = ldg; p.diag = diag;
        p.tri = tri; p.row_base = m0;
        if (m0 == 0) { p.rows_out = rows_out; p.ld_rows_out = ld_rows_out; }      // every row chunk streams all K rows: one copies
        if (peers && m0 + 256 >= s) {
            // the last row chunk leaves the final diagonal: it stores it into every rank's copy and publishes the epoch
            p.npeers = peers->world;
            for (int r = 0; r < peers->world; ++r) {
                RPC_CHECK_ARG(peers->diag_full[r] && peers->flags[r], "peer table has null entries");
                p.peer_diag[r] = peers->diag_full[r] + (int64_t)peers->rank * peers->shard;
                p.peer_flag[r] = peers->flags[r] + peers->rank;
            }
            p.epoch = epoch;
            p.publish = 1;
            p.done_counter = ctx->sync_counter;
        }
        int rc = dispatch_gemm<MODE_SCHUR>(ctx, st, p, m, n_pad);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int rpc_schur_update(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At,
                                int64_t ldat, int k_pad, const double *rows_new, int64_t ldr, double *diag,
                                int64_t n_pad) {
    return schur_update_impl(ctx, stream, G, ldg, c, s, At, ldat, k_pad, rows_new, ldr, nullptr, diag, n_pad, nullptr, 0, 0);
}

extern "C" int rpc_schur_update_ex(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At,
                                   int64_t ldat, int k_pad, const double *rows_src, int64_t ldr, const int32_t *rowmap,
                                   double *diag, int64_t n_pad, const rpc_peers *peers, unsigned long long epoch,
                                   int panel_lower) {
    return schur_update_impl(ctx, stream, G, ldg, c, s, At, ldat, k_pad, rows_src, ldr, rowmap, diag, n_pad, peers, epoch,
                             panel_lower);
}

extern "C" int rpc_schur_update_rows(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At,
                                     int64_t ldat, int k_pad, const double *rows_src, int64_t ldr, const int32_t *rowmap,
                                     double *rows_out, int64_t ld_rows_out, double *diag, int64_t n_pad, const rpc_peers *peers,
                                     unsigned long long epoch, int panel_lower) {
    RPC_CHECK_ARG(rows_out != nullptr, "rows_out required (use rpc_schur_update_ex otherwise)");
    return schur_update_impl(ctx, stream, G, ldg, c, s, At, ldat, k_pad, rows_src, ldr, rowmap, diag, n_pad, peers, epoch,
                             panel_lower, rows_out, ld_rows_out);
}

extern "C" int rpc_schur_update_peers(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At,
                                      int64_t ldat, int k_pad, const double *rows_new, int64_t ldr, double *diag, int64_t n_pad,
                                      const rpc_peers *peers, unsigned long long epoch) {
    RPC_CHECK_ARG(peers != nullptr, "peer table required");
    return schur_update_impl(ctx, stream, G, ldg, c, s, At, ldat, k_pad, rows_new, ldr, nullptr, diag, n_pad, peers, epoch, 0);
}

// ------------------------------------------------------------------------------------------------
// CUDA-core evaluator: Laplace (L1, strictly sequential over features like scipy cdist 'cityblock')
// and the direct-difference Euclidean form (extra_stability).  4 pivots x 4 points per thread.
// ------------------------------------------------------------------------------------------------
constexpr int DK_TP = 64;    // pivots per CTA pass (16 thread rows x 4)
constexpr int DK_TN = 64;    // points per CTA (16 thread cols x 4)
constexpr int DK_FT = 32;    // features per smem chunk

// REDUCE (fused kernel-times-vector): out is a vector, out[col] = (accumulate ? out[col] : 0) + sum_row w[row] k(pivot row, X col)
template <bool L1, bool REDUCE>
__global__ void __launch_bounds__(256) direct_rows_kernel(KernParams kp, const double *__restrict__ X, int64_t ldx, int d,
                                                          const double *__restrict__ Xpiv, int64_t ldp, int s,
                                                          double *__restrict__ out, int64_t ldo, const double *__restrict__ w,
                                                          int accumulate) {
    __shared__ double Ps[DK_FT][DK_TP + 4];
    __shared__ double Ys[DK_FT][DK_TN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t col0 = (int64_t)blockIdx.x * DK_TN;
    double cs[4] = {0.0, 0.0, 0.0, 0.0};
    for (int p0 = 0; p0 < s; p0 += DK_TP) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        for (int f0 = 0; f0 < d; f0 += DK_FT) {
            __syncthreads();
            for (int q = tid; q < DK_FT * DK_TP; q += 256) {
                int pi = q / DK_FT, f = q - pi * DK_FT;
                double v = 0.0;
                if (p0 + pi < s && f0 + f < d) v = Xpiv[(int64_t)(p0 + pi) * ldp + f0 + f];
                Ps[f][pi] = v;
            }
            for (int q = tid; q < DK_FT * DK_TN; q += 256) {
                int ni = q / DK_FT, f = q - ni * DK_FT;
                double v = 0.0;
                if (f0 + f < d) v = X[(col0 + ni) * ldx + f0 + f];
                Ys[f][ni] = v;
            }
            __syncthreads();
            const int fmax = d - f0 < DK_FT ? d - f0 : DK_FT;
            for (int f = 0; f < fmax; ++f) {
                double xp[4], yv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xp[i] = Ps[f][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) yv[j] = Ys[f][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double df = xp[i] - yv[j];
                        if (L1) acc[i][j] += fabs(df);            // sequential in f: bit-equal to cdist
                        else acc[i][j] = fma(df, df, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int row = p0 + ty * 4 + i;
            if (row < s) {
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = L1 ? kern_from_l1(kp, acc[i][j]) : kern_from_d2(kp, acc[i][j]);
                if (REDUCE) {
                    const double wr = w[row];
#pragma unroll
                    for (int j = 0; j < 4; ++j) cs[j] = fma(wr, v[j], cs[j]);
                } else {
                    double *dst = out + (int64_t)row * ldo + col0 + tx * 4;
                    *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]);
                    *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]);
                }
            }
        }
    }
    if (REDUCE) {
        // sum over the 16 thread rows in a fixed order (Ps is free: every thread is past its last read of it)
        __syncthreads();
        double *red = &Ps[0][0];                                   // 16 x 64 doubles fit in the 32 x 68 panel buffer
#pragma unroll
        for (int j = 0; j < 4; ++j) red[ty * DK_TN + tx * 4 + j] = cs[j];
        __syncthreads();
        if (tid < DK_TN) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q) t += red[q * DK_TN + tid];
            out[col0 + tid] = (accumulate ? out[col0 + tid] : 0.0) + t;
        }
    }
}

extern "C" int rpc_kernel_rows(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *X,
                               const double *xnorm, int64_t n_pad, int d, int64_t ldx, const double *Xpiv,
                               const double *pnorm, int s, int64_t ldp, double *out, int64_t ldo) {
    RPC_CHECK_ARG(ctx && spec && X && Xpiv && out, "null pointer");
    RPC_CHECK_ARG(s >= 1, "s >= 1");
    RPC_CHECK_ARG(n_pad > 0 && n_pad % RPC_COL_ALIGN == 0, "n_pad must be a positive multiple of RPC_COL_ALIGN");
    RPC_CHECK_ARG(d >= 1 && d % RPC_FEAT_ALIGN == 0 && ldx >= d, "d must be a multiple of RPC_FEAT_ALIGN, ldx >= d");
    RPC_CHECK_ARG(ldo % 2 == 0, "ldo must be even");
    RPC_CHECK_ARG(spec->kind >= 0 && spec->kind <= 2, "unknown kernel kind");
    RPC_CHECK_ARG(spec->kind != RPC_KERN_MATERN || spec->nu2 == 1 || spec->nu2 == 3 || spec->nu2 == 5,
                  "Matern nu must be 0.5, 1.5 or 2.5");
    cudaStream_t st = (cudaStream_t)stream;
    KernParams kp = make_kern_params(spec);
    if (spec->kind == RPC_KERN_LAPLACE || spec->extra_stability) {
        const bool l1 = spec->kind == RPC_KERN_LAPLACE;
        // fast path: pivots resident in shared memory, point tiles streamed by TMA (csrc/direct_eval.cu)
        bool done = true;
        for (int m0 = 0; m0 < s && done; m0 += 256) {
            const int m = s - m0 < 256 ? s - m0 : 256;
            const int ldpt = 16 * rpc_direct_pi_for(m, l1);
            int rc = rpc_ws_reserve(ctx, (size_t)d * ldpt * sizeof(double));
            if (rc) return rc;
            rc = rpc_transpose_pad(ctx, st, Xpiv + (int64_t)m0 * ldp, ldp, m, d, ctx->ws, ldpt, d);
            if (rc) return rc;
            rc = rpc_direct_rows_tma(ctx, st, kp, l1, X, ldx, d, ctx->ws, ldpt, m, out + (int64_t)m0 * ldo, ldo, n_pad);
            if (rc == -100) { done = false; break; }
            if (rc) return rc;
        }
        if (done) return 0;
        dim3 grid((unsigned)(n_pad / DK_TN));
        if (spec->kind == RPC_KERN_LAPLACE)
            direct_rows_kernel<true, false><<<grid, 256, 0, st>>>(kp, X, ldx, d, Xpiv, ldp, s, out, ldo, nullptr, 0);
        else
            direct_rows_kernel<false, false><<<grid, 256, 0, st>>>(kp, X, ldx, d, Xpiv, ldp, s, out, ldo, nullptr, 0);
        RPC_LAUNCHED(ctx);
        return 0;
    }
    RPC_CHECK_ARG(xnorm && pnorm, "row norms required for the expanded Euclidean form");
    RPC_CHECK_ARG(ldx % 2 == 0, "the DMMA evaluator needs an even point stride (16-byte rows)");
    // A operand = Xpiv^T (d x s) zero padded to (k_pad x ld) in the context workspace
    const int k_pad = (d + RPC_KT - 1) / RPC_KT * RPC_KT;
    for (int m0 = 0; m0 < s; m0 += 256) {
        int m = s - m0 < 256 ? s - m0 : 256;
        // short contractions: the pivot panel stays resident in shared memory and the point tiles stream (csrc/eval_panel.cu)
        const int ld_res = (ldx % 2 == 0 && !g_no_evalp) ? rpc_eval_panel_ld(m, d, ldx) : 0;
        if (ld_res > 0) {
            int rc = rpc_ws_reserve(ctx, (size_t)d * ld_res * sizeof(double));
            if (rc) return rc;
            rc = rpc_transpose_pad(ctx, st, Xpiv + (int64_t)m0 * ldp, ldp, m, d, ctx->ws, ld_res, d);
            if (rc) return rc;
            rc = rpc_eval_panel_rows(ctx, st, kp, X, xnorm, n_pad, d, ldx, ctx->ws, pnorm + m0, m, out + (int64_t)m0 * ldo, ldo);
            if (rc) return rc;
            continue;
        }
        const int ld = panel_ld_for(m);
        int rc = rpc_ws_reserve(ctx, (size_t)k_pad * ld * sizeof(double));
        if (rc) return rc;
        rc = rpc_transpose_pad(ctx, st, Xpiv + (int64_t)m0 * ldp, ldp, m, d, ctx->ws, ld, k_pad);
        if (rc) return rc;
        Params p{};
        p.m_valid = m;
        p.X = X; p.ldx = ldx; p.d = d; p.xnorm = xnorm; p.pnorm = pnorm + m0; p.kp = kp;
        if (fused_fits(m, ldx) && ldx % RPC_FEAT_ALIGN == 0) {
            // X tile resident in shared memory (one bulk copy per tile), panel staged through the ring
            p.At_e = ctx->ws; p.ldat_e = ld; p.ke_pad = k_pad;
            p.At = ctx->ws; p.ldat = ld; p.k_pad = RPC_KT;
            p.R = out + (int64_t)m0 * ldo; p.ldr = ldo;
            rc = dispatch_gemm<MODE_EVALX>(ctx, st, p, m, n_pad);
        } else {
            p.At = ctx->ws; p.ldat = ld; p.k_pad = k_pad;
            p.Cout = out + (int64_t)m0 * ldo; p.ldc = ldo;
            rc = dispatch_gemm<MODE_EVAL>(ctx, st, p, m, n_pad);
        }
        if (rc) return rc;
    }
    return 0;
}

// Fused kernel-times-vector: out[col] = (accumulate ? out[col] : 0) + sum_{i < s} w[i] k(Xpiv_i, X_col) for all n_pad points, without
// writing a kernel value to memory -- the blocked kernel mat-vec of Nystrom-PCG (block_experiments/pes_code.py:500-506) and the
// prediction K(Xts, Xtr[S]) @ sol of KRR (KRR_Nystrom.py:60-66) are this operation with the roles of rows and columns chosen so
// that the reduction runs over the (at most 256 at a time) resident pivots.  Row blocks are applied one after the other.
extern "C" int rpc_kernel_rows_wsum(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *X, const double *xnorm,
                                    int64_t n_pad, int d, int64_t ldx, const double *Xpiv, const double *pnorm, int s, int64_t ldp,
                                    const double *w, double *out, int accumulate) {
    RPC_CHECK_ARG(ctx && spec && X && Xpiv && w && out, "null pointer");
    RPC_CHECK_ARG(s >= 1, "s >= 1");
    RPC_CHECK_ARG(n_pad > 0 && n_pad % RPC_COL_ALIGN == 0, "n_pad must be a positive multiple of RPC_COL_ALIGN");
    RPC_CHECK_ARG(d >= 1 && d % RPC_FEAT_ALIGN == 0 && ldx >= d, "d must be a multiple of RPC_FEAT_ALIGN, ldx >= d");
    RPC_CHECK_ARG(spec->kind >= 0 && spec->kind <= 2, "unknown kernel kind");
    RPC_CHECK_ARG(spec->kind != RPC_KERN_MATERN || spec->nu2 == 1 || spec->nu2 == 3 || spec->nu2 == 5,
                  "Matern nu must be 0.5, 1.5 or 2.5");
    cudaStream_t st = (cudaStream_t)stream;
    KernParams kp = make_kern_params(spec);
    if (spec->kind == RPC_KERN_LAPLACE || spec->extra_stability) {
        const bool l1 = spec->kind == RPC_KERN_LAPLACE;
        bool done = true;
        int acc_flag = accumulate;
        for (int m0 = 0; m0 < s && done; m0 += 256) {
            const int m = s - m0 < 256 ? s - m0 : 256;
            const int ldpt = 16 * rpc_direct_pi_for(m, l1);
            int rc = rpc_ws_reserve(ctx, (size_t)d * ldpt * sizeof(double));
            if (rc) return rc;
            rc = rpc_transpose_pad(ctx, st, Xpiv + (int64_t)m0 * ldp, ldp, m, d, ctx->ws, ldpt, d);
            if (rc) return rc;
            rc = rpc_direct_rows_tma(ctx, st, kp, l1, X, ldx, d, ctx->ws, ldpt, m, out, 0, n_pad, w + m0, acc_flag);
            if (rc == -100) { done = false; break; }
            if (rc) return rc;
            acc_flag = 1;
        }
        if (done) return 0;
        // the pivot panel does not fit in shared memory (large d): chunked kernel, all pivots in one launch.  A first row block
        // may already have been applied above only if it fitted, and then every block fits: nothing was applied here.
        dim3 grid((unsigned)(n_pad / DK_TN));
        if (l1) direct_rows_kernel<true, true><<<grid, 256, 0, st>>>(kp, X, ldx, d, Xpiv, ldp, s, out, 0, w, accumulate);
        else direct_rows_kernel<false, true><<<grid, 256, 0, st>>>(kp, X, ldx, d, Xpiv, ldp, s, out, 0, w, accumulate);
        RPC_LAUNCHED(ctx);
        return 0;
    }
    RPC_CHECK_ARG(xnorm && pnorm, "row norms required for the expanded Euclidean form");
    RPC_CHECK_ARG(ldx % 2 == 0, "the DMMA evaluator needs an even point stride (16-byte rows)");
    int acc_flag = accumulate;
    for (int m0 = 0; m0 < s; m0 += 248) {
        const int m = s - m0 < 248 ? s - m0 : 248;
        const int ld_res = rpc_eval_panel_wsum_ld(m, d, ldx);
        if (ld_res <= 0) {
            rpc_set_error("rpc_kernel_rows_wsum: a %d x %d pivot panel does not fit in shared memory (d too large for the fused DMMA path)", d, m);
            return -4;
        }
        int rc = rpc_ws_reserve(ctx, (size_t)d * ld_res * sizeof(double));
        if (rc) return rc;
        rc = rpc_transpose_pad(ctx, st, Xpiv + (int64_t)m0 * ldp, ldp, m, d, ctx->ws, ld_res, d);
        if (rc) return rc;
        rc = rpc_eval_panel_wsum(ctx, st, kp, X, xnorm, n_pad, d, ldx, ctx->ws, pnorm + m0, m, w + m0, out, acc_flag);
        if (rc) return rc;
        acc_flag = 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// simple RPCholesky step, fully fused (HBM-bound GEMV over the i x n prefix of G)
// ------------------------------------------------------------------------------------------------
constexpr int SS_THREADS = 256;
constexpr int SS_MAXD = 1024;   // features cached in smem

__device__ __forceinline__ double simple_row_entry(int has_spec, const KernParams &kp, const double *__restrict__ X,
                                                   const double *__restrict__ xnorm, int d, int64_t ldx,
                                                   const double *__restrict__ A, int64_t lda, int64_t n, int64_t piv,
                                                   const double *xp, double pn, int64_t col) {
    if (!has_spec) return col < n ? A[piv * lda + col] : 0.0;
    if (kp.kind == RPC_KERN_LAPLACE) {
        double l1 = 0.0;
        for (int f = 0; f < d; ++f) l1 += fabs(xp[f] - X[col * ldx + f]);
        return kern_from_l1(kp, l1);
    }
    if (kp.stab) {
        double d2 = 0.0;
        for (int f = 0; f < d; ++f) { double df = xp[f] - X[col * ldx + f]; d2 = fma(df, df, d2); }
        return kern_from_d2(kp, d2);
    }
    double dot = 0.0;
    for (int f = 0; f < d; ++f) dot = fma(xp[f], X[col * ldx + f], dot);
    double d2 = (-2.0 * dot + pn) + xnorm[col];
    return kern_from_d2(kp, d2 > 0.0 ? d2 : 0.0);
}

// Each thread owns two adjacent columns (16-byte loads) and keeps 8 independent row loads in flight: the kernel
// streams G[:i] once, 8 N i bytes, which is all the HBM traffic of a simple-RPCholesky step.
__global__ void __launch_bounds__(SS_THREADS) simple_step_kernel(int has_spec, KernParams kp, const double *__restrict__ X,
                                                                 const double *__restrict__ xnorm, int d, int64_t ldx,
                                                                 const double *__restrict__ A, int64_t lda, int64_t n,
                                                                 const int64_t *__restrict__ idx_dev, int i,
                                                                 double *__restrict__ G, int64_t ldg,
                                                                 double *__restrict__ rows, int64_t ldr,
                                                                 double *__restrict__ diag_in, double *__restrict__ diag_out,
                                                                 int64_t n_pad) {
    extern __shared__ double sm[];          // [i] pivot column of G, then [d] pivot point
    double *pcol = sm;
    double *xp = sm + i;
    const int64_t piv = idx_dev[i];
    for (int r = threadIdx.x; r < i; r += SS_THREADS) pcol[r] = G[(int64_t)r * ldg + piv];
    if (has_spec)
        for (int f = threadIdx.x; f < d; f += SS_THREADS) xp[f] = X[piv * ldx + f];
    __syncthreads();
    const double dpiv = diag_in[piv];
    const double scale = sqrt(dpiv);
    const double pn = has_spec && xnorm ? xnorm[piv] : 0.0;
    for (int64_t col = 2 * ((int64_t)blockIdx.x * SS_THREADS + threadIdx.x); col < n_pad;
         col += 2 * (int64_t)gridDim.x * SS_THREADS) {
        const double row0 = simple_row_entry(has_spec, kp, X, xnorm, d, ldx, A, lda, n, piv, xp, pn, col);
        const double row1 = simple_row_entry(has_spec, kp, X, xnorm, d, ldx, A, lda, n, piv, xp, pn, col + 1);
        // g = (row - G[:i,piv]^T G[:i,col]) / sqrt(diag[piv])   (rpcholesky.py:41)
        double2 acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = make_double2(0.0, 0.0);
        const double *Gc = G + col;
        int r = 0;
        for (; r + 8 <= i; r += 8) {
            double2 g[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) g[q] = *reinterpret_cast<const double2 *>(Gc + (int64_t)(r + q) * ldg);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double pc = pcol[r + q];
                acc[q].x = fma(pc, g[q].x, acc[q].x);
                acc[q].y = fma(pc, g[q].y, acc[q].y);
            }
        }
        for (; r < i; ++r) {
            const double2 g = *reinterpret_cast<const double2 *>(Gc + (int64_t)r * ldg);
            acc[0].x = fma(pcol[r], g.x, acc[0].x);
            acc[0].y = fma(pcol[r], g.y, acc[0].y);
        }
        const double s0 = ((acc[0].x + acc[1].x) + (acc[2].x + acc[3].x)) + ((acc[4].x + acc[5].x) + (acc[6].x + acc[7].x));
        const double s1 = ((acc[0].y + acc[1].y) + (acc[2].y + acc[3].y)) + ((acc[4].y + acc[5].y) + (acc[6].y + acc[7].y));
        const double g0 = (row0 - s0) / scale, g1 = (row1 - s1) / scale;
        *reinterpret_cast<double2 *>(rows + (int64_t)i * ldr + col) = make_double2(row0, row1);
        *reinterpret_cast<double2 *>(G + (int64_t)i * ldg + col) = make_double2(g0, g1);
        const double2 dv = *reinterpret_cast<const double2 *>(diag_in + col);
        const double d0 = dv.x - g0 * g0, d1 = dv.y - g1 * g1;
        *reinterpret_cast<double2 *>(diag_out + col) = make_double2(d0 > 0.0 ? d0 : 0.0, d1 > 0.0 ? d1 : 0.0);
    }
}

// one step reading the residual diagonal from diag_in and writing the updated one to diag_out (the step reads
// diag_in[piv] while other CTAs write the new diagonal, so the two must not alias)
static int simple_step_launch(rpc_ctx *ctx, cudaStream_t st, const rpc_kernel_spec *spec, const double *X, const double *xnorm,
                              int d, int64_t ldx, const double *A, int64_t lda, int64_t n, const int64_t *idx_dev, int i,
                              double *G, int64_t ldg, double *rows, int64_t ldr, double *diag_in, double *diag_out,
                              int64_t n_pad) {
    RPC_CHECK_ARG(ctx && idx_dev && G && rows && diag_in && diag_out, "null pointer");
    RPC_CHECK_ARG((spec && X) || A, "either a kernel spec with X or a dense A is required");
    RPC_CHECK_ARG(i >= 0 && n_pad > 0 && n_pad % 2 == 0 && ldg % 2 == 0 && ldr % 2 == 0, "i >= 0, even n_pad / ldg / ldr");
    RPC_CHECK_ARG(!spec || d <= SS_MAXD, "d too large for the fused simple step");
    KernParams kp{};
    if (spec) kp = make_kern_params(spec);
    int grid = (int)((n_pad / 2 + SS_THREADS - 1) / SS_THREADS);
    int cap = ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    size_t shm = (size_t)(i + (spec ? d : 0) + 1) * sizeof(double);
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(simple_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured |= 1ull << (ctx->device & 63);
    }
    RPC_CHECK_ARG(shm <= 200 * 1024, "rank too large for the fused simple step");
    simple_step_kernel<<<grid, SS_THREADS, shm, st>>>(spec ? 1 : 0, kp, X, xnorm, d, ldx, A, lda, n, idx_dev, i, G, ldg,
                                                      rows, ldr, diag_in, diag_out, n_pad);
    RPC_LAUNCHED(ctx);
    return 0;
}

extern "C" int rpc_simple_step(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *X,
                               const double *xnorm, int d, int64_t ldx, const double *A, int64_t lda, int64_t n,
                               const int64_t *idx_dev, int i, double *G, int64_t ldg, double *rows, int64_t ldr,
                               double *diag, int64_t n_pad) {
    RPC_CHECK_ARG(ctx && diag && n_pad > 0, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // in-place semantics for the caller: double-buffer through the workspace
    int rc = rpc_ws_reserve(ctx, (size_t)n_pad * sizeof(double));
    if (rc) return rc;
    rc = simple_step_launch(ctx, st, spec, X, xnorm, d, ldx, A, lda, n, idx_dev, i, G, ldg, rows, ldr, diag, ctx->ws, n_pad);
    if (rc) return rc;
    RPC_CUDA(cudaMemcpyAsync(diag, ctx->ws, (size_t)n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// The whole simple-RPCholesky / greedy loop (rpcholesky.py:29-43) enqueued by ONE host call: steps [i0, i1) of
// sample-or-argmax -> fused step.  The uniforms were uploaded in bulk (the proposal stream is state independent) and
// nothing is read back, so the k steps run back to back on the stream without any host involvement; stats row i
// (4 doubles) keeps the sampler status of step i for the caller to inspect afterwards.  The residual diagonal
// ping-pongs between `diag` and a second buffer at the end of the workspace (no per-step copy).
extern "C" int rpc_simple_run(rpc_ctx *ctx, void *stream, int greedy, const rpc_kernel_spec *spec, const double *X,
                              const double *xnorm, int d, int64_t ldx, const double *A, int64_t lda, int64_t n,
                              const double *u_all, int64_t *idx_all, double *stats_all, int i0, int i1, double *G,
                              int64_t ldg, double *rows, int64_t ldr, double *diag, int64_t n_pad) {
    RPC_CHECK_ARG(ctx && idx_all && stats_all && G && rows && diag, "null pointer");
    RPC_CHECK_ARG(greedy || u_all, "uniforms required for randomly pivoted steps");
    RPC_CHECK_ARG(i0 >= 0 && i1 >= i0 && n_pad > 0, "bad step range");
    cudaStream_t st = (cudaStream_t)stream;
    // the sampler / argmax scratch lives at the front of the workspace (never more than n_pad/8 + 4096 doubles); reserving
    // the whole range first keeps ctx->ws fixed for the duration of the loop
    const size_t front = ((size_t)n_pad / 8 + 4096 + 255) & ~(size_t)255;
    int rc = rpc_ws_reserve(ctx, (front + (size_t)n_pad) * sizeof(double));
    if (rc) return rc;
    double *other = ctx->ws + front;
    double *cur = diag, *nxt = other;
    for (int i = i0; i < i1; ++i) {
        rc = greedy ? rpc_argmax(ctx, stream, cur, n, idx_all + i, stats_all + 4 * (int64_t)i)
                    : rpc_sample_pivots(ctx, stream, cur, n_pad, u_all + i, 1, idx_all + i, stats_all + 4 * (int64_t)i);
        if (rc) return rc;
        rc = simple_step_launch(ctx, st, spec, X, xnorm, d, ldx, A, lda, n, idx_all, i, G, ldg, rows, ldr, cur, nxt, n_pad);
        if (rc) return rc;
        double *t = cur; cur = nxt; nxt = t;
    }
    if (cur != diag) RPC_CUDA(cudaMemcpyAsync(diag, cur, (size_t)n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return 0;
}
