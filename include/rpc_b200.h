/*
 * rpc_b200.h -- C ABI of the B200 (sm_100a) RPCholesky hot path.
 *
 * The reference (eepperly/Randomly-Pivoted-Cholesky) is pure Python and has no FFI layer; its
 * boundary is the Python operator protocol (matrix.py) plus the three factorisation loops
 * (rpcholesky.py).  Every entry point below replaces ONE numpy/scipy/sklearn call site of that
 * path; the site is cited next to the prototype as `file:line` relative to the reference root.
 * The Python host in randomly_pivoted_cholesky_b200/ binds these with ctypes.
 *
 * Conventions
 *   - all matrices are float64, row-major, with an explicit leading dimension in ELEMENTS;
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless the prototype says so;
 *   - return value: 0 = ok, <0 = bad argument, >0 = CUDA error code; rpc_last_error() returns a
 *     thread-local description.  Nothing throws across the boundary;
 *   - the caller owns every buffer; the library only owns the scratch inside rpc_ctx;
 *   - one rpc_ctx per device, not thread-safe.
 *   - "padded" operands: X, G, rows and diag are allocated by the host with n_pad = a multiple of
 *     RPC_COL_ALIGN columns/points; pad points are all-zero, pad diagonal entries are 0 so they are
 *     never sampled.  Kernels run over n_pad columns without edge checks.
 */
#ifndef RPC_B200_H
#define RPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPC_ABI_VERSION 1
#define RPC_COL_ALIGN 128   /* n_pad granularity (widest column tile of the DMMA kernels) */
#define RPC_FEAT_ALIGN 4    /* d_pad granularity (one DMMA k-step) */
#define RPC_KT 16           /* k-depth of one shared-memory stage; panel K is padded to this */
#define RPC_MAX_BLOCK 1536  /* largest proposal block b / accepted block s of one round (shared memory of the single-CTA
                               rejection Cholesky and of the panel solve); the reference accepts any b, the drivers raise
                               ValueError above this and cap the auto-tuned b at it */

/* kernel functions, reference kernels.py */
#define RPC_KERN_GAUSSIAN 0 /* kernels.py:32-39 */
#define RPC_KERN_LAPLACE 1  /* kernels.py:16-21 */
#define RPC_KERN_MATERN 2   /* kernels.py:71-85, nu2 in {1,3,5} = 2*nu */

typedef struct rpc_ctx rpc_ctx;

typedef struct rpc_kernel_spec {
    int32_t kind;            /* RPC_KERN_* */
    int32_t nu2;             /* Matern: 2*nu in {1,3,5}; ignored otherwise */
    int32_t extra_stability; /* kernels.py:35-36,74-75: direct-difference distance */
    int32_t reserved;
    double bandwidth;
} rpc_kernel_spec;

int rpc_abi_version(void);
/* sizeof of the structs that cross the ABI, for bindings to check their layout: 0 = rpc_kernel_spec, 1 = rpc_peers. */
int rpc_abi_sizeof(int which);
const char *rpc_last_error(void);

int rpc_ctx_create(int device, rpc_ctx **out);
int rpc_ctx_destroy(rpc_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t rpc_ctx_launch_count(const rpc_ctx *ctx);

/* ---- kernel evaluator: KernelMatrix.__getitem__ / diag (matrix.py:122-138,185-189) ---------- */

/* xnorm[i] = sum_f X[i,f]^2 : sklearn row_norms(X, squared=True) inside euclidean_distances
 * (called from kernels.py:38,77). */
int rpc_row_norms(rpc_ctx *ctx, void *stream, const double *X, int64_t n, int d, int64_t ldx, double *xnorm);

/* A.diag() for the built-in kernels: distance exactly 0 -> exactly 1.0 (kernels.py:11-14,27-30,56-66).
 * Entries [n, n_pad) are set to 0. */
int rpc_kernel_diag(rpc_ctx *ctx, void *stream, double *diag, int64_t n, int64_t n_pad);

/* out[j, i] = k(Xpiv[j,:], X[i,:]) for j<s, i<n_pad : A[idx,:] -> kernel_mtx(X[idx], X)
 * (matrix.py:188-189 -> kernels.py:16-21 / 32-39 / 71-85).  Gaussian/Matern without
 * extra_stability use D2 = ((-2 x.y) + |x|^2) + |y|^2 clamped at 0 (sklearn euclidean_distances)
 * with the x.y contraction on FP64 tensor cores (DMMA); Laplace sums |x_f - y_f| in feature order
 * (scipy cdist 'cityblock').  d must be a multiple of RPC_FEAT_ALIGN (zero-padded features). */
int rpc_kernel_rows(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *X, const double *xnorm,
                    int64_t n_pad, int d, int64_t ldx, const double *Xpiv, const double *pnorm, int s, int64_t ldp,
                    double *out, int64_t ldo);

/* H[i, j] = k(Xpiv[i,:], Xpiv[j,:]) : A[idx,idx] outer block (matrix.py:99,135; rpcholesky.py:189). */
int rpc_kernel_block(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *Xpiv, const double *pnorm,
                     int b, int d, int64_t ldp, double *H, int64_t ldh);

/* Dense PSDMatrix operator (matrix.py:86-102): diag gather, row gather A[idx,:], block A[idx,idx]. */
int rpc_psd_diag(rpc_ctx *ctx, void *stream, const double *A, int64_t n, int64_t lda, double *diag, int64_t n_pad);
int rpc_psd_rows(rpc_ctx *ctx, void *stream, const double *A, int64_t n, int64_t lda, const int64_t *idx, int s,
                 double *out, int64_t ldo, int64_t n_pad);
int rpc_psd_block(rpc_ctx *ctx, void *stream, const double *A, int64_t lda, const int64_t *idx, int b, double *H,
                  int64_t ldh);

/* ---- pivot sampler: rng.choice(range(n), size=m, p=diags/sum(diags)) (rpcholesky.py:31,91,184) -- */

/* idx[j] = #{ i : cdf[i] <= u[j] } with cdf = cumsum(diag/sum(diag)) / cdf[-1] over diag[0:n_pad], every sum
 * rounded exactly as numpy's sequential left-to-right fp64 loop rounds it (the device reproduces the sequentially
 * rounded prefix sums in parallel, csrc/sampler.cu), so the pivots are bit-identical to the reference's for the
 * same uniforms and stats[0] equals python's sum(diags) bit for bit.
 * stats[0] = sum(diag), stats[1] = 1.0 if the probabilities are invalid (sum <= 0, NaN or negative entries:
 * numpy raises ValueError), stats[2] = 1.0 on an internal consistency failure (never observed).  m may be 0. */
int rpc_sample_pivots(rpc_ctx *ctx, void *stream, const double *diag, int64_t n_pad, const double *u, int m,
                      int64_t *idx, double *stats);

/* Greedy pivoting (rpcholesky.py:33,35): idx[0] = np.argmax(diag[0:n]) (first maximum), stats[0] = the maximum,
 * stats[1] = number of entries equal to it (the randomized tie-break of 'rgreedy' draws among them on the host). */
int rpc_argmax(rpc_ctx *ctx, void *stream, const double *diag, int64_t n, int64_t *idx, double *stats);

/* ---- per-round small dense work (rpcholesky.py:102-107, 140-157, 189-190, 206-207) -------------- */

/* Xpiv[j,:] = X[idx[j],:], pnorm[j] = xnorm[idx[j]] (matrix.py:189 `self.data[vec_i,:]`) and
 * P[r, j] = G[r, idx[j]] for r<c (rpcholesky.py:102,189,206 `G[0:counter, idx]`).
 * X/xnorm may be NULL (dense operator), G may be NULL when c == 0.
 * Row-sharded operation: idx holds GLOBAL indices and this shard owns [lo, hi) (local row = idx - lo);
 * entries of pivots owned by another shard are written as 0, so that a sum all-reduce of the payload over
 * the shards assembles it exactly (one non-zero contributor per slot).  Single GPU: lo = 0, hi = N. */
int rpc_gather_pivots(rpc_ctx *ctx, void *stream, const int64_t *idx, int m, const double *X, const double *xnorm,
                      int d, int64_t ldx, double *Xpiv, double *pnorm, int64_t ldp, const double *G, int64_t ldg,
                      int c, double *P, int64_t ldP, int64_t lo, int64_t hi);

/* Selected pivots S = idx[acc] (rpcholesky.py:198): sel_idx[j] = idx[acc[j]], Xsel[j,:] = Xpiv[acc[j],:],
 * nsel[j] = pnorm[acc[j]] for j < s.  acc NULL = identity; Xpiv/pnorm NULL for the dense operator. */
int rpc_compact_selected(rpc_ctx *ctx, void *stream, const int32_t *acc, int s, const int64_t *idx, const double *Xpiv,
                         const double *pnorm, int64_t ldp, int64_t *sel_idx, double *Xsel, double *nsel);

/* H -= P^T P (H is m x m, P is c x m): residual Gram of the proposals (rpcholesky.py:189). */
int rpc_gram_residual(rpc_ctx *ctx, void *stream, const double *P, int c, int m, int64_t ldP, double *H, int64_t ldh);

/* rejection_cholesky(H) (rpcholesky.py:140-157) with u[j] replacing the j-th np.random.rand().
 * mode 0: rejection sampling; accepted positions -> acc[0:s] (ascending), stops after max_accept
 *         (the truncation of rpcholesky.py:193-196).
 * mode 1: plain Cholesky of H + shift*I (block RPCholesky, rpcholesky.py:106); info[1] = index+1 of
 *         the first non-positive pivot (numpy would raise LinAlgError), else 0.
 * L (ldl >= m) receives the s x s lower-triangular factor L[acc,acc] (zeros above the diagonal).
 * info[0] = s, info[2] = 1 if trace(H) <= 0 (RuntimeError at rpcholesky.py:145). */
int rpc_rejection_cholesky(rpc_ctx *ctx, void *stream, double *H, int m, int64_t ldh, const double *u, int max_accept,
                           int mode, double shift, double *L, int64_t ldl, int32_t *acc, int32_t *info);

/* RBRP block pivoting: `_greedy_cholesky(H, rtol, atol)` (rpcholesky.py:52-63), i.e. LAPACK dpstrf(lower) -- diagonal
 * pivoting on the first maximum of the residual diagonal, stop when it falls to m*eps*max(diag H) -- followed, when
 * use_tol != 0, by rank = argmax(trace - cumsum(norm(L,axis=0)**2) < atol + rtol*trace) + 1.  H (m x m, symmetric) is not
 * modified.  acc[0:rank] = the pivot positions piv[0:rank]-1 (rpcholesky.py:118), L (ldl >= m) = the rank x rank leading
 * block of the factor in pivoted order (zeros above the diagonal), info[0] = rank. */
int rpc_pivoted_cholesky(rpc_ctx *ctx, void *stream, const double *H, int m, int64_t ldh, int use_tol, double rtol,
                         double atol, double *L, int64_t ldl, int32_t *acc, int32_t *info);

/* Block-greedy pivots: idx[0:b] = the positions of the b largest entries of diag[0:n], i.e. the SET returned by
 * np.argpartition(diags, -b)[-b:] (rpcholesky.py:94-95), in ascending index order; among entries tied at the threshold
 * the lowest indices are kept (numpy's choice there is an artefact of its introselect). */
int rpc_select_largest(rpc_ctx *ctx, void *stream, const double *diag, int64_t n, int b, int64_t *idx);

/* Panel for the fused Schur update.  With Minv = L^{-1} (mode 0) or Minv = T given (mode 1, the
 * eigh fallback of rpcholesky.py:109-114):
 *    At[r, j]     = -sum_i P[r, acc[i]] * Minv[j, i]   r < c       (= -(P_acc Minv^T))
 *    At[c + i, j] = Minv[j, i]                          i < s
 *    zero elsewhere (rows up to k_pad, columns up to m_pad).
 * so that G_new = At^T [G[:c]; rows_new] = L^{-1} (rows_new - P_acc^T G[:c])
 * (rpcholesky.py:102+107, 206+207; np.linalg.solve(L, .) is folded into the panel by one forward substitution
 * per panel row -- no explicit inverse is formed; mode 1 takes a dense Minv for the eigh fallback).
 * acc may be NULL (= identity selection). */
int rpc_build_panel(rpc_ctx *ctx, void *stream, const double *P, int64_t ldP, int c, const int32_t *acc, int s,
                    const double *L, int64_t ldl, int mode, double *At, int64_t ldat, int k_pad);

/* Leading dimension to build the panel At with so that rpc_schur_update fetches each k-stage of it
 * with ONE TMA bulk copy (the shared-memory stride of the row tile chosen for s rows).  Any even ldat >= the tile
 * height works; this one is the fast one. */
int rpc_panel_ld(int s);

/* The round's host read-back without a stream synchronisation: writes [seq | stats(4) | info(4) | acc(m) | idx(m)] as
 * doubles into `host_mapped`, a PINNED host buffer of at least 9 + 2m doubles addressed through UVA, the sequence number
 * last (after a system-scope fence); the host spins until host_mapped[0] == seq.  info / acc may be NULL. */
int rpc_publish_round(rpc_ctx *ctx, void *stream, const double *stats, const int32_t *info, const int32_t *acc,
                      const int64_t *idx, int m, double *host_mapped, double seq);

/* ---- the N-proportional update (rpcholesky.py:41-43, 102-107+128-129, 206-209) ---------------- */

/* G[c:c+s, :] = At^T [G[0:c, :]; rows_new] ; diag = max(diag - colsum(G[c:c+s,:]^2), 0).
 * FP64 DMMA GEMM (M = s, K = c + s, N = n_pad) with the diagonal update fused in the epilogue.
 * G rows [c, c+s) are written; rows_new is s x n_pad (normally rows + c*ldr). */
int rpc_schur_update(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At, int64_t ldat,
                     int k_pad, const double *rows_new, int64_t ldr, double *diag, int64_t n_pad);

/* One step of simple RPCholesky (rpcholesky.py:40-43), fused: row = A[idx,:]; g = (row - G[:i,idx]^T
 * G[:i,:]) / sqrt(diag[idx]); rows[i,:] = row; G[i,:] = g; diag = max(diag - g^2, 0).
 * idx_dev points at the pivot index on the device (no host round trip).  For the dense operator pass
 * spec = NULL and A/lda. */
int rpc_simple_step(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *X, const double *xnorm, int d,
                    int64_t ldx, const double *A, int64_t lda, int64_t n, const int64_t *idx_dev, int i, double *G,
                    int64_t ldg, double *rows, int64_t ldr, double *diag, int64_t n_pad);

/* Steps [i0, i1) of the simple loop (rpcholesky.py:29-43) enqueued by one call: per step rpc_sample_pivots with the
 * pre-uploaded uniform u_all[i] (greedy != 0: rpc_argmax, rpcholesky.py:35) into idx_all[i], then rpc_simple_step.
 * stats_all holds 4 doubles per step (the sampler status of that step; for greedy [max, #ties]). */
int rpc_simple_run(rpc_ctx *ctx, void *stream, int greedy, const rpc_kernel_spec *spec, const double *X, const double *xnorm,
                   int d, int64_t ldx, const double *A, int64_t lda, int64_t n, const double *u_all, int64_t *idx_all,
                   double *stats_all, int i0, int i1, double *G, int64_t ldg, double *rows, int64_t ldr, double *diag,
                   int64_t n_pad);

/* rpc_build_panel (mode 0) followed by rpc_compact_selected, with the row count s read on the device from *s_dev
 * (0 <= *s_dev <= s_max): both can be enqueued before the host knows s (accelerated RPCholesky: right behind
 * rpc_rejection_cholesky, whose info[0] is s_dev).  The panel is laid out with ldat = rpc_panel_ld(*s_dev) and
 * k_pad = (c + *s_dev) rounded up to 16, exactly what the host then passes to rpc_schur_update. */
int rpc_build_panel_dev(rpc_ctx *ctx, void *stream, const double *P, int64_t ldP, int c, const int32_t *acc,
                        const int32_t *s_dev, int s_max, const double *L, int64_t ldl, double *At, const int64_t *idx,
                        const double *Xpiv, const double *pnorm, int64_t ldp, int64_t *sel_idx, double *Xsel, double *nsel);

struct rpc_peers;
/* Persistent kernels launched after this call use (SM count - count) CTAs: room for a single-CTA kernel (the rejection
 * Cholesky) to run beside them on another stream.  count = 0 restores the default. */
int rpc_ctx_reserve_sms(rpc_ctx *ctx, int count);

/* dst[i, :] = src[rowmap[i], :] for i < *count_dev (<= max_rows), n_pad columns: the accepted rows of a speculatively
 * evaluated proposal block, compacted into rows[c:c+s] = A[S,:] (rpcholesky.py:199).  The count is read on the device. */
int rpc_gather_rows(rpc_ctx *ctx, void *stream, const double *src, int64_t lds, const int32_t *rowmap,
                    const int32_t *count_dev, int max_rows, double *dst, int64_t ldd, int64_t n_pad);

/* rpc_schur_update with the new rows taken from rows_src through a row map (new row i = rows_src[rowmap[i], :]; rowmap
 * may be NULL) and, when peers != NULL, the fused diagonal exchange of rpc_schur_update_peers.
 * panel_lower != 0 declares that panel rows c..c+s-1 hold the transpose of a LOWER-triangular matrix, At[c + i, j] == 0
 * for j < i -- what rpc_build_panel mode 0 / rpc_build_panel_dev produce (L^{-T}; np.linalg.solve(L, .) of
 * rpcholesky.py:107,207 is a triangular solve) -- so the kernel skips the structurally-zero half of that block: the
 * executed flops are then the algorithmic s (2c + s) n_pad instead of 2 s (c + s) n_pad.  Pass 0 for a dense panel
 * (rpc_build_panel mode 1, the eigh fallback of rpcholesky.py:109-114). */
int rpc_schur_update_ex(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At, int64_t ldat,
                        int k_pad, const double *rows_src, int64_t ldr, const int32_t *rowmap, double *diag, int64_t n_pad,
                        const struct rpc_peers *peers, unsigned long long epoch, int panel_lower);

/* ---- "next" row 1 of the scope table: KRR_Nystrom on the factor rows the hot path leaves in HBM (KRR_Nystrom.py:34-58) ---- */

/* M = R R^T for R = rows[0:k, 0:n_cols] (row-major, ldr >= n_cols): `KMn @ KnM` of KRR_Nystrom.py:54.  FP64 DMMA over the
 * 128 x 128 tiles of the lower triangle, operands fetched by tensor-map TMA copies, split over the columns with a
 * deterministic fixed-order reduction; both triangles of M (k x k, ldm >= k) are written. */
int rpc_syrk_rows(rpc_ctx *ctx, void *stream, const double *R, int64_t ldr, int k, int64_t n_cols, double *M, int64_t ldm);

/* out[i] = sum_c R[i, c] * y[c], i < k: `KMn @ Ytr` of KRR_Nystrom.py:54 (one pass over R, deterministic). */
int rpc_rows_dot(rpc_ctx *ctx, void *stream, const double *R, int64_t ldr, int k, int64_t n_cols, const double *y, double *out);

/* out[c] = sum_{i < k} w[i] * R[i, c], c < n_cols (n_cols even): the transposed product G^T w of the Nystrom preconditioner
 * (block_experiments/pes_code.py:471-473), one pass over R. */
int rpc_rows_tdot(rpc_ctx *ctx, void *stream, const double *R, int64_t ldr, int k, int64_t n_cols, const double *w, double *out);

/* x <- (L L^T)^{-1} x for the lower-triangular Cholesky factor L (k x k, row-major) and nrhs right-hand sides stored as the
 * columns of x (k x nrhs, row-major, ldx >= nrhs): the two triangular solves of scipy.linalg.solve(..., assume_a='pos')
 * (KRR_Nystrom.py:54) after the k x k factorisation. */
int rpc_chol_solve(rpc_ctx *ctx, void *stream, const double *L, int64_t ldl, int k, double *x, int64_t ldx, int nrhs);

/* Fused kernel-times-vector: out[col] = (accumulate ? out[col] : 0) + sum_{i < s} w[i] * k(Xpiv[i], X[col]) for the n_pad points of
 * X, no kernel value written to memory (same operands as rpc_kernel_rows; w has s entries, out n_pad).  Replaces
 * `K_pred[:, :] @ sol` of KRR_Nystrom.predict_Nystrom (KRR_Nystrom.py:60-66, intent) and the blocked mat-vec
 * `np.dot(A[blocks[i]:blocks[i+1], :], p)` of Nystrom-PCG (block_experiments/pes_code.py:500-506), using k(x, y) = k(y, x) to
 * put the reduction over the resident pivots.  Deterministic (fixed-order sums). */
int rpc_kernel_rows_wsum(rpc_ctx *ctx, void *stream, const rpc_kernel_spec *spec, const double *X, const double *xnorm, int64_t n_pad,
                         int d, int64_t ldx, const double *Xpiv, const double *pnorm, int s, int64_t ldp, const double *w,
                         double *out, int accumulate);

/* Debug aid for bindings: counts the violations of the padding contract stated at the top of this file -- non-zero (or NaN)
 * entries among the pad features [d, ldx) of the n real points, the pad points [n, n_pad) and, when diag != NULL, the pad
 * entries of the residual diagonal -- and returns the count through the HOST pointer.  SYNCHRONISES the stream.  The Python host
 * runs it on every operator it builds when RPC_B200_VALIDATE=1 is set. */
int rpc_check_padding(rpc_ctx *ctx, void *stream, const double *X, int64_t n, int64_t n_pad, int d, int64_t ldx, const double *diag,
                      int32_t *violations_host);

/* Measurement aid (bench.py's roofline denominator): enqueues one launch of a pure DMMA.8x8x4 stream (one 256-thread CTA
 * per SM, 16 independent accumulator pairs per warp, `iters` rounds) and returns its flop count through the HOST pointer
 * flops_per_launch; the caller times it with CUDA events on `stream`.  Replaces no reference call site. */
int rpc_probe_fp64_peak(rpc_ctx *ctx, void *stream, int iters, double *flops_per_launch_host);

/* rpc_schur_update_ex that ALSO writes the new rows it consumes to rows_out[i, :] = rows_src[rowmap[i], :], i < s: with
 * speculatively evaluated proposal rows (rows_src = the rows of all b proposals, rowmap = the accepted positions) this is
 * `rows[c:c+s] = A[S,:]` of rpcholesky.py:205 at no extra pass over memory -- the kernel copies each B stage of new rows out of
 * shared memory while it multiplies it. */
int rpc_schur_update_rows(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At, int64_t ldat,
                          int k_pad, const double *rows_src, int64_t ldr, const int32_t *rowmap, double *rows_out,
                          int64_t ld_rows_out, double *diag, int64_t n_pad, const struct rpc_peers *peers,
                          unsigned long long epoch, int panel_lower);

/* ---- peer-memory exchange for row-sharded runs (one process per GPU of one NVSwitch box) --------------------------
 * The two per-round exchanges of the sharded factorisation (the residual-diagonal all-gather and
 * the pivot-payload exchange) are done by the producing kernels themselves: they store their results straight into every
 * rank's buffers through NVLink peer mappings and publish a monotone epoch in a flag word; the consumer side waits on
 * its local flag words.  No NCCL call is left inside a round. */
#define RPC_MAX_PEERS 8
typedef struct rpc_peers {
    int32_t world, rank;
    int64_t shard;                                  /* points per rank (dist.ShardInfo.shard) */
    double *diag_full[RPC_MAX_PEERS];               /* rank r's copy of the full residual diagonal (world * shard) */
    double *payload[RPC_MAX_PEERS];                 /* rank r's pivot payload [X[idx] | |x|^2 | G[:c, idx]] */
    unsigned long long *flags[RPC_MAX_PEERS];       /* rank r's flag words: [src] diag epochs, [RPC_MAX_PEERS + src] payload */
} rpc_peers;

/* A zero-initialised device buffer that other processes of the box can map: handle = its 64-byte cudaIpcMemHandle_t. */
int rpc_peer_alloc(rpc_ctx *ctx, int64_t bytes, void **ptr, unsigned char handle[64]);
int rpc_peer_free(rpc_ctx *ctx, void *ptr);
/* Map / unmap a buffer exported by another process (cudaIpcOpenMemHandle with lazy peer access). */
int rpc_peer_open(rpc_ctx *ctx, const unsigned char handle[64], void **ptr);
int rpc_peer_close(rpc_ctx *ctx, void *ptr);
/* Enqueue a wait until flags[i] >= epoch for all i < count (a one-warp kernel spinning on local memory).  If that does not
 * happen within timeout_ms the kernel gives up and sets *status (device int32) to 1. */
int rpc_peer_wait(rpc_ctx *ctx, void *stream, const unsigned long long *flags, int count, unsigned long long epoch,
                  int timeout_ms, int32_t *status);
/* rpc_schur_update that also stores the updated residual diagonal into every rank's diag_full and publishes `epoch` in
 * flags[r][rank] of every rank r (the fused diag all-gather). */
int rpc_schur_update_peers(rpc_ctx *ctx, void *stream, double *G, int64_t ldg, int c, int s, const double *At, int64_t ldat,
                           int k_pad, const double *rows_new, int64_t ldr, double *diag, int64_t n_pad, const rpc_peers *peers,
                           unsigned long long epoch);
/* rpc_gather_pivots whose owners store their entries into EVERY rank's payload (offsets in doubles: Xpiv at 0 with row
 * stride ldp, |x|^2 at off_norm, P at off_P with row stride ldP) and publish `epoch` in flags[r][RPC_MAX_PEERS + rank]. */
int rpc_gather_pivots_peers(rpc_ctx *ctx, void *stream, const int64_t *idx, int m, const double *X, const double *xnorm, int d,
                            int64_t ldx, int64_t ldp, int64_t off_norm, const double *G, int64_t ldg, int c, int64_t off_P,
                            int64_t ldP, int64_t lo, int64_t hi, const rpc_peers *peers, unsigned long long epoch);

#ifdef __cplusplus
}
#endif
#endif /* RPC_B200_H */
