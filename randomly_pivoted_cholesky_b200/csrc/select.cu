// Block-greedy pivot selection: the b largest entries of the residual diagonal
//     idx = np.argpartition(diags, -b)[-b:]                                          (rpcholesky.py:94-95)
// as a radix select on the order-preserving integer image of the doubles (6 passes of 11 bits: privatised shared-memory
// histograms, one small CTA picks the digit that holds the b-th largest key), followed by an index-ordered compaction.
// Which of several entries TIED at the threshold numpy's introselect keeps is an implementation detail of numpy; the
// device keeps the lowest indices, and returns the selection in ascending index order.
#include "rpc_common.cuh"

constexpr int SEL_BITS = 11;
constexpr int SEL_BINS = 1 << SEL_BITS;
constexpr int SEL_PASSES = 6;                 // 6 * 11 = 66 >= 64
constexpr int SEL_THREADS = 256;
constexpr int SEL_CHUNK = 4096;               // elements per CTA in the compaction passes

struct SelState {
    unsigned long long prefix;                // the digits fixed so far (high bits of the threshold key)
    unsigned long long mask;                  // which bits of `prefix` are fixed
    long long need;                           // how many keys are still to be taken from the candidates
};

__device__ __forceinline__ unsigned long long sel_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);      // larger double <=> larger key (NaN-free input)
}

__device__ __forceinline__ int sel_shift(int pass) {           // digits from the top; the last one is 9 bits wide
    const int s = 64 - SEL_BITS * (pass + 1);
    return s < 0 ? 0 : s;
}

__global__ void __launch_bounds__(SEL_THREADS) sel_hist_kernel(const double *__restrict__ d, int64_t n, const SelState *st,
                                                               int pass, unsigned int *__restrict__ hist) {
    __shared__ unsigned int sh[SEL_BINS];
    for (int q = threadIdx.x; q < SEL_BINS; q += SEL_THREADS) sh[q] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix, mask = st->mask;
    const int shift = sel_shift(pass);
    const unsigned int dmask = pass == SEL_PASSES - 1 ? (1u << (64 - SEL_BITS * (SEL_PASSES - 1))) - 1 : SEL_BINS - 1;
    for (int64_t j = (int64_t)blockIdx.x * SEL_THREADS + threadIdx.x; j < n; j += (int64_t)gridDim.x * SEL_THREADS) {
        const unsigned long long k = sel_key(d[j]);
        if ((k & mask) == prefix) atomicAdd(&sh[(unsigned int)(k >> shift) & dmask], 1u);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < SEL_BINS; q += SEL_THREADS)
        if (sh[q]) atomicAdd(&hist[q], sh[q]);
}

// one CTA: the highest digit D with  #(digit > D) < need <= #(digit >= D); fixes it and clears the histogram
__global__ void __launch_bounds__(SEL_THREADS) sel_pick_kernel(SelState *st, int pass, unsigned int *__restrict__ hist, int b) {
    __shared__ unsigned int sh[SEL_BINS];
    if (pass == 0 && threadIdx.x == 0) { st->prefix = 0; st->mask = 0; st->need = b; }
    for (int q = threadIdx.x; q < SEL_BINS; q += SEL_THREADS) { sh[q] = hist[q]; hist[q] = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long need = st->need, above = 0;
        int D = 0;
        for (int q = SEL_BINS - 1; q >= 0; --q) {
            if (above + sh[q] >= need) { D = q; break; }
            above += sh[q];
        }
        const int shift = sel_shift(pass);
        const unsigned long long dm = pass == SEL_PASSES - 1 ? (1ull << (64 - SEL_BITS * (SEL_PASSES - 1))) - 1 : SEL_BINS - 1;
        st->prefix |= (unsigned long long)D << shift;
        st->mask |= dm << shift;
        st->need = need - above;
    }
}

// per-chunk counts of keys above / equal to the threshold
__global__ void __launch_bounds__(SEL_THREADS) sel_count_kernel(const double *__restrict__ d, int64_t n, const SelState *st,
                                                                int *__restrict__ cnt) {
    __shared__ int sgt, seq;
    if (threadIdx.x == 0) { sgt = 0; seq = 0; }
    __syncthreads();
    const unsigned long long T = st->prefix;
    const int64_t lo = (int64_t)blockIdx.x * SEL_CHUNK;
    int gt = 0, eq = 0;
    for (int q = threadIdx.x; q < SEL_CHUNK; q += SEL_THREADS) {
        const int64_t j = lo + q;
        if (j < n) { const unsigned long long k = sel_key(d[j]); gt += k > T; eq += k == T; }
    }
    atomicAdd(&sgt, gt);
    atomicAdd(&seq, eq);
    __syncthreads();
    if (threadIdx.x == 0) { cnt[2 * blockIdx.x] = sgt; cnt[2 * blockIdx.x + 1] = seq; }
}

// exclusive scan of the chunk counts (one CTA, sequential over a few hundred chunks per thread at most)
__global__ void __launch_bounds__(32) sel_scan_kernel(int *__restrict__ cnt, int nchunk) {
    if (threadIdx.x == 0) {
        int g = 0, e = 0;
        for (int q = 0; q < nchunk; ++q) {
            const int cg = cnt[2 * q], ce = cnt[2 * q + 1];
            cnt[2 * q] = g; cnt[2 * q + 1] = e;
            g += cg; e += ce;
        }
    }
}

__device__ __forceinline__ int block_excl_scan(int v, int *warp_tot, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();                                 // warp_tot may still be read from the previous call
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < SEL_THREADS / 32; ++w) { if (w < warp) base += warp_tot[w]; total += warp_tot[w]; }
    return base + inc - v;
}

// selected = key > T, or key == T among the first `need` such entries in index order; written in ascending index order
__global__ void __launch_bounds__(SEL_THREADS) sel_write_kernel(const double *__restrict__ d, int64_t n, const SelState *st,
                                                                const int *__restrict__ cnt, int b, int64_t *__restrict__ out) {
    __shared__ int wt[SEL_THREADS / 32];
    const unsigned long long T = st->prefix;
    const int need = (int)st->need;
    int eq_base = cnt[2 * blockIdx.x + 1];
    int out_base = cnt[2 * blockIdx.x] + (eq_base < need ? eq_base : need);
    const int64_t lo = (int64_t)blockIdx.x * SEL_CHUNK;
    for (int q0 = 0; q0 < SEL_CHUNK; q0 += SEL_THREADS) {
        const int64_t j = lo + q0 + threadIdx.x;
        unsigned long long k = 0;
        const bool in = j < n;
        if (in) k = sel_key(d[j]);
        const int eq = in && k == T;
        int eq_tot, sel_tot;
        const int eq_rank = block_excl_scan(eq, wt, eq_tot);
        const int sel = in && (k > T || (eq && eq_base + eq_rank < need));
        const int pos = block_excl_scan(sel, wt, sel_tot);
        if (sel && out_base + pos < b) out[out_base + pos] = j;
        eq_base += eq_tot;
        out_base += sel_tot;
    }
}

extern "C" int rpc_select_largest(rpc_ctx *ctx, void *stream, const double *diag, int64_t n, int b, int64_t *idx) {
    RPC_CHECK_ARG(ctx && diag && idx, "null pointer");
    RPC_CHECK_ARG(n >= 1 && b >= 1 && b <= n, "need 1 <= b <= n");
    cudaStream_t st = (cudaStream_t)stream;
    const int nchunk = (int)((n + SEL_CHUNK - 1) / SEL_CHUNK);
    // int workspace: hist[SEL_BINS] | cnt[2 * nchunk] | SelState (8-byte aligned)
    const size_t ints = (size_t)SEL_BINS + 2 * (size_t)nchunk + 2;
    int rc = rpc_flags_reserve(ctx, (ints + 8) * sizeof(int32_t));
    if (rc) return rc;
    unsigned int *hist = (unsigned int *)ctx->flags;
    int *cnt = ctx->flags + SEL_BINS;
    SelState *state = (SelState *)(ctx->flags + ((SEL_BINS + 2 * (size_t)nchunk + 1) & ~(size_t)1));
    RPC_CUDA(cudaMemsetAsync(hist, 0, SEL_BINS * sizeof(unsigned int), st));
    int grid = (int)((n + SEL_THREADS - 1) / SEL_THREADS);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    for (int pass = 0; pass < SEL_PASSES; ++pass) {
        if (pass == 0) {
            // the state is initialised by the first pick; the first histogram must see mask = 0
            RPC_CUDA(cudaMemsetAsync(state, 0, sizeof(SelState), st));
        }
        sel_hist_kernel<<<grid, SEL_THREADS, 0, st>>>(diag, n, state, pass, hist);
        RPC_LAUNCHED(ctx);
        sel_pick_kernel<<<1, SEL_THREADS, 0, st>>>(state, pass, hist, b);
        RPC_LAUNCHED(ctx);
    }
    sel_count_kernel<<<nchunk, SEL_THREADS, 0, st>>>(diag, n, state, cnt);
    RPC_LAUNCHED(ctx);
    sel_scan_kernel<<<1, 32, 0, st>>>(cnt, nchunk);
    RPC_LAUNCHED(ctx);
    sel_write_kernel<<<nchunk, SEL_THREADS, 0, st>>>(diag, n, state, cnt, b, idx);
    RPC_LAUNCHED(ctx);
    return 0;
}
