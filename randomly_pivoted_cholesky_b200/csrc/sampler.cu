// Residual-proportional pivot sampler, bit-identical to numpy's
//     rng.choice(range(n), size=m, p=diags/sum(diags), replace=True)          (rpcholesky.py:31,91,184)
// i.e.  S = sum(diags) [python builtin: left-to-right fp64];  p = diags / S;  cdf = cumsum(p) [left-to-right];
//       cdf /= cdf[-1];  idx = searchsorted(cdf, u, side='right').
//
// A parallel scan rounds differently from numpy's sequential one, so the device computes the
// SEQUENTIALLY ROUNDED prefix sums exactly, in parallel, with this observation: while the running
// sum c stays inside one binade [2^e, 2^(e+1)) with ulp q = 2^(e-52), fl(c + x) = c + q*rne(x/q)
// (ties aside), and all quantities are integer multiples of q below 2^53 q, so they add EXACTLY in
// any order.  Per chunk of SC_CH elements:
//   1. plain parallel sums give an approximate prefix P with a rigorous bound |c - P| <= delta*P;
//   2. a chunk whose running sums provably stay in one binade and that holds no rounding tie is
//      "regular": its exact contribution D = q * sum_j rne(x_j/q) is computed in parallel;
//   3. one CTA walks the chunks: runs of regular chunks of one binade are folded with exact prefix sums, the
//      few others (binade crossings, ties: ~20 per scan) are resolved with real fp64 adds at 8-element grain;
//   4. each draw binary-searches the exact chunk carries with numpy's own predicate fl(c/c_last) <= u
//      (as a division-free threshold test) and replays one chunk sequentially to land on the element.
// The same machinery run on diag itself yields python's sum(diags) bit-exactly (stoptol tests,
// rpcholesky.py:45,133,224, therefore agree with the reference too).
#include "rpc_common.cuh"

constexpr int SC_CH = 256;            // elements per chunk (one warp, 8 per lane)
constexpr int SC_WARPS = 8;           // chunks per CTA in the parallel passes


__device__ __forceinline__ double pow2d(int e) {   // 2^e for e in [-1022, 1023]
    return __longlong_as_double((long long)(e + 1023) << 52);
}

// exponent of a positive NORMAL double from its bit pattern (the serial walk cannot afford ilogb's special-case handling);
// zero / subnormal give -1023 and inf / NaN give 1024, both outside the range the callers accept
__device__ __forceinline__ int fast_ilogb(double c) { return (int)((__double_as_longlong(c) >> 52) & 0x7ff) - 1023; }

template <bool DIV>
__device__ __forceinline__ double elem(const double *__restrict__ d, int64_t j, double S) {
    return DIV ? d[j] / S : d[j];
}

// 1. approximate chunk sums (+ validity: NaN or negative entries make numpy raise ValueError)
template <bool DIV>
__device__ __forceinline__ void chunk_sums_body(const double *__restrict__ d, int64_t n, const double *Sptr,
                                                double *__restrict__ csum, int32_t *__restrict__ bad, int64_t k, int lane) {
    const int64_t base = k * SC_CH;
    if (base >= n) return;
    const double S = DIV ? *Sptr : 1.0;
    double s = 0.0;
    bool isbad = false;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        int64_t j = base + lane * 8 + t;
        double x = j < n ? elem<DIV>(d, j, S) : 0.0;
        isbad |= !(x >= 0.0);
        s += x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (__any_sync(0xffffffffu, isbad) && lane == 0) atomicExch(bad, 1);
    if (lane == 0) csum[k] = s;
}

template <bool DIV>
__global__ void __launch_bounds__(SC_WARPS * 32) sc_chunk_sums(const double *__restrict__ d, int64_t n, const double *Sptr,
                                                                double *__restrict__ csum, int32_t *__restrict__ bad) {
    chunk_sums_body<DIV>(d, n, Sptr, csum, bad, (int64_t)blockIdx.x * SC_WARPS + (threadIdx.x >> 5), threadIdx.x & 31);
}

// 2. exclusive scan of the chunk sums, one CTA
__device__ __forceinline__ void scan_chunks_body(double *__restrict__ csum, int nch) {
    __shared__ double part[1024];
    const int tid = threadIdx.x;
    const int per = (nch + 1023) / 1024;
    const int b0 = tid * per;
    double s = 0.0;
    for (int i = 0; i < per; ++i) if (b0 + i < nch) s += csum[b0 + i];
    part[tid] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int o = 1; o < 1024; o <<= 1) {
        double v = tid >= o ? part[tid - o] : 0.0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    double run = tid > 0 ? part[tid - 1] : 0.0;
    for (int i = 0; i < per; ++i) {
        if (b0 + i < nch) { double v = csum[b0 + i]; csum[b0 + i] = run; run += v; }
    }
    if (tid == 1023) csum[nch] = part[1023];
}

__global__ void __launch_bounds__(1024) sc_scan_chunks(double *__restrict__ csum, int nch) { scan_chunks_body(csum, nch); }

// 3. classify each chunk and compute the exact contribution of the regular ones
template <bool DIV>
__device__ __forceinline__ void classify_body(const double *__restrict__ d, int64_t n, const double *Sptr,
                                              const double *__restrict__ P, double delta, double *__restrict__ D,
                                              int32_t *__restrict__ cls, int64_t k, int lane) {
    const int64_t base = k * SC_CH;
    if (base >= n) return;
    const double S = DIV ? *Sptr : 1.0;
    // pass B (x = d/S): the exact pass-A carries divided by S enclose the running sums of the quotients
    // (each quotient and each sequential add is off by at most one rounding: delta covers both passes)
    const double lo = (DIV ? P[k] / S : P[k]) * (1.0 - delta), hi = (DIV ? P[k + 1] / S : P[k + 1]) * (1.0 + delta);
    bool regular = lo > 0.0 && isfinite(hi);
    int e = 0;
    if (regular) {
        e = ilogb(lo);                               // lo in [2^e, 2^(e+1))
        regular = e >= -960 && e <= 1000 && hi < pow2d(e + 1);
    }
    double A = 0.0;
    bool tie = false;
    if (regular) {                                   // warp-uniform
        const double scale = pow2d(52 - e);          // exact power of two
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            int64_t j = base + lane * 8 + t;
            double x = j < n ? elem<DIV>(d, j, S) : 0.0;
            double v = x * scale;                    // exact (or a harmless underflow towards 0)
            double kf = floor(v);
            double f = v - kf;                       // exact
            tie |= (f == 0.5);
            A += kf + (f > 0.5 ? 1.0 : 0.0);         // integers < 2^53: exact
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) A += __shfl_xor_sync(0xffffffffu, A, o);
        if (__any_sync(0xffffffffu, tie)) regular = false;
    }
    if (lane == 0) {
        cls[k] = regular ? e : INT_MIN;
        D[k] = regular ? A * pow2d(e - 52) : 0.0;    // exact: integer times a power of two
    }
}

template <bool DIV>
__global__ void __launch_bounds__(SC_WARPS * 32) sc_classify(const double *__restrict__ d, int64_t n, const double *Sptr,
                                                             const double *__restrict__ P, double delta,
                                                             double *__restrict__ D, int32_t *__restrict__ cls) {
    classify_body<DIV>(d, n, Sptr, P, delta, D, cls, (int64_t)blockIdx.x * SC_WARPS + (threadIdx.x >> 5), threadIdx.x & 31);
}

// 4. the walk: exact running sum before every chunk.  One CTA.
// A run of regular chunks of ONE binade is folded with a prefix sum: all its D are multiples of the same ulp q and
// their total is below 2^52 q, so the additions are exact in any order and carry + prefix reproduces the sequential
// running sum bit for bit.
//   phase A (all warps): per tile of 32 chunks, is it "clean" (32 regular chunks of one binade)?  If so its
//            within-tile exclusive prefix and total are computed with warp shuffles.
//   phase B (warp 0): carries across tiles -- one exact add per clean tile; dirty tiles (mixed binades, chunks
//            that need a replay) are resolved run by run, the chunks marked INT_MIN with real sequential adds at
//            8-element granularity.
//   phase C (all warps): cstart = tile carry + within-tile prefix for the clean tiles.
constexpr int SW_THREADS = 1024;

template <bool DIV>
__device__ __forceinline__ double walk_dirty_tile(const double *__restrict__ d, int64_t n, double S, int k0, int nch, int e,
                                                  double Dk, double carry, double *__restrict__ cstart, int lane,
                                                  int &violated, const double *xs) {
    const unsigned full = 0xffffffffu;
    const int k = k0 + lane;
    int pos = 0;
    const int kend = min(32, nch - k0);
    while (pos < kend) {
        const int e0 = __shfl_sync(full, e, pos);
        if (e0 == INT_MIN) {
            // Chunk k0+pos could not be classified a priori (binade crossing, rounding tie, start of the array).
            // Now the EXACT carry is known, so the same argument is applied at a finer grain: lane l owns 8
            // consecutive elements; all lanes up to the first one that crosses the binade of the carry (or holds a
            // tie) are folded in bulk, that lane's 8 elements are added one by one, repeat.
            const int64_t base = (int64_t)(k0 + pos) * SC_CH + lane * 8;
            // phase A staged the chunk's elements in shared memory (slot id parked in the unused D entry)
            const int slot = (int)(-__shfl_sync(full, Dk, pos)) - 1;
            double x[8];
            if (slot >= 0) {
#pragma unroll
                for (int t = 0; t < 8; ++t) x[t] = xs[slot * SC_CH + t * 32 + lane];
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) x[t] = (base + t) < n ? elem<DIV>(d, base + t, S) : 0.0;
            }
            if (lane == 0) cstart[k0 + pos] = carry;
            int lpos = 0;
            while (lpos < 32) {
                const int eb = carry > 0.0 ? fast_ilogb(carry) : INT_MIN;
                int f = lpos;                                  // first lane that needs the sequential treatment
                if (eb >= -960 && eb <= 1000) {
                    const double scale = pow2d(52 - eb), q = pow2d(eb - 52), top = pow2d(eb + 1);
                    double A = 0.0;
                    bool tie = false;
                    if (lane >= lpos) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            double v = x[t] * scale;
                            double kf = floor(v);
                            double fr = v - kf;
                            tie |= (fr == 0.5);
                            A += kf + (fr > 0.5 ? 1.0 : 0.0);
                        }
                    }
                    double pre = A;                            // inclusive prefix over lanes >= lpos (exact integers)
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        double t2 = __shfl_up_sync(full, pre, o);
                        if (lane >= o) pre += t2;
                    }
                    const bool bad = lane >= lpos && (tie || !isfinite(pre) || !(carry + pre * q < top));
                    const unsigned badmask = __ballot_sync(full, bad);
                    f = badmask ? __ffs(badmask) - 1 : 32;
                    if (f > lpos) {
                        const double upto = __shfl_sync(full, pre, f - 1);
                        carry += upto * q;                     // exact: all terms are multiples of q below 2^53 q
                    }
                }
                if (f < 32) {
                    // the real sequential adds: every lane adds ITS eight elements to the (identical) carry in its own registers and
                    // lane f's result is broadcast -- the same eight fp64 additions in the same order as a chain of eight
                    // shuffle + add pairs, with one shuffle latency on the serial path instead of eight
                    double cf = carry;
#pragma unroll
                    for (int t = 0; t < 8; ++t) cf += x[t];
                    carry = __shfl_sync(full, cf, f);
                    lpos = f + 1;
                } else {
                    lpos = 32;
                }
            }
            ++pos;
            continue;
        }
        // maximal run [pos, end) of chunks with binade e0
        const unsigned same = __ballot_sync(full, e == e0);
        const unsigned from = same >> pos;                        // bit i <-> lane pos+i
        const int len = (~from == 0u) ? 32 - pos : __ffs(~from) - 1;
        const int end = pos + len;
        const bool in_run = lane >= pos && lane < end;
        double v = in_run ? Dk : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(full, v, o);
            if (lane >= o) v += t;
        }
        const double total = __shfl_sync(full, v, 31);
        if (in_run) cstart[k] = carry + (v - Dk);                  // exclusive prefix, exact
        if (fast_ilogb(carry) != e0 || !(carry + total <= pow2d(e0 + 1))) violated = 1;
        carry += total;                                            // exact
        pos = end;
    }
    return carry;
}

constexpr int SW_SLOTS = 64;      // replay chunks staged in shared memory (64 x 2 KB); more fall back to global loads

template <bool DIV>
__device__ __forceinline__ void walk_body(const double *__restrict__ d, int64_t n, const double *Sptr,
                                          double *__restrict__ D, const int32_t *__restrict__ cls, int nch,
                                          double *__restrict__ cstart, double *__restrict__ tile_tot,
                                          double *__restrict__ tile_carry, int32_t *__restrict__ tile_e,
                                          int32_t *__restrict__ bad, double *__restrict__ stats, double *xs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const double S = DIV ? *Sptr : 1.0;
    const int ntile = (nch + 31) / 32;
    constexpr int NW = SW_THREADS / 32;
    // xs: [SW_SLOTS][SC_CH], element (lane*8+t) at [t*32+lane]
    __shared__ int nslots;
    if (threadIdx.x == 0) nslots = 0;
    __syncthreads();

    // ---- phase A
    for (int t = warp; t < ntile; t += NW) {
        const int k = t * 32 + lane;
        const bool live = k < nch;
        const int e = live ? cls[k] : INT_MIN + 1;
        double Dk = live ? D[k] : 0.0;
        // stage the elements of the chunks that phase B has to replay (keeps global latency off the serial path)
        unsigned dirty = __ballot_sync(full, e == INT_MIN);
        while (dirty) {
            const int src = __ffs(dirty) - 1;
            dirty &= dirty - 1;
            int slot = 0;
            if (lane == 0) slot = atomicAdd(&nslots, 1);
            slot = __shfl_sync(full, slot, 0);
            if (slot < SW_SLOTS) {
                const int64_t base = (int64_t)(t * 32 + src) * SC_CH + lane * 8;
#pragma unroll
                for (int q = 0; q < 8; ++q) xs[slot * SC_CH + q * 32 + lane] = (base + q) < n ? elem<DIV>(d, base + q, S) : 0.0;
                if (lane == src) D[k] = -(double)(slot + 1);
            }
        }
        const int e0 = __shfl_sync(full, e, 0);
        const bool clean = __all_sync(full, e == e0) && e0 != INT_MIN && e0 != INT_MIN + 1;   // a full tile, one binade
        double v = Dk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t2 = __shfl_up_sync(full, v, o);
            if (lane >= o) v += t2;
        }
        if (clean) cstart[k] = v - Dk;                           // within-tile exclusive prefix (exact)
        if (lane == 31) { tile_tot[t] = v; tile_e[t] = clean ? e0 : INT_MIN; }
    }
    __syncthreads();

    // ---- phase B
    if (warp == 0) {
        double carry = 0.0;                                      // identical in all lanes
        int violated = 0;
        for (int t0 = 0; t0 < ntile; t0 += 32) {
            const int tl = t0 + lane;
            const int te = tl < ntile ? tile_e[tl] : INT_MIN;
            const double tt = tl < ntile ? tile_tot[tl] : 0.0;
            const int tend = min(32, ntile - t0);
            int q = 0;
            while (q < tend) {
                const int e0 = __shfl_sync(full, te, q);
                if (e0 == INT_MIN) {
                    if (lane == 0) tile_carry[t0 + q] = carry;
                    const int k0 = (t0 + q) * 32, k = k0 + lane;
                    const bool live = k < nch;
                    carry = walk_dirty_tile<DIV>(d, n, S, k0, nch, live ? cls[k] : INT_MIN + 1, live ? D[k] : 0.0, carry, cstart,
                                                 lane, violated, xs);
                    ++q;
                    continue;
                }
                // maximal run [q, end) of clean tiles of binade e0: their totals are multiples of one ulp and add up
                // exactly in any order, so the carries of the whole run come from one warp scan (lanes past ntile
                // hold INT_MIN and never match)
                const unsigned same = __ballot_sync(full, te == e0);
                const unsigned from = same >> q;
                const int len = (~from == 0u) ? 32 - q : __ffs(~from) - 1;
                const int end = q + len;
                const bool in_run = lane >= q && lane < end;
                double v = in_run ? tt : 0.0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t2 = __shfl_up_sync(full, v, o);
                    if (lane >= o) v += t2;
                }
                const double total = __shfl_sync(full, v, 31);
                if (in_run) tile_carry[t0 + lane] = carry + (v - tt);          // exact
                // non-negative terms: the running sum stays in the binade iff it starts and ends in it
                if (!(carry > 0.0) || fast_ilogb(carry) != e0 || !(carry + total <= pow2d(e0 + 1))) violated = 1;
                carry += total;                                                // exact
                q = end;
            }
        }
        if (lane == 0) {
            cstart[nch] = carry;
            if (violated) atomicExch(bad, 2);
            if (DIV && stats) {
                // numpy raises ValueError when p holds NaN / negatives or does not sum to 1 (S == 0, inf, NaN)
                const int b = violated ? 2 : *bad;
                const bool invalid = (b == 1) || !(S > 0.0) || !isfinite(S) || !(fabs(carry - 1.0) <= 1.4901161193847656e-08);
                stats[0] = S;
                stats[1] = invalid ? 1.0 : 0.0;
                stats[2] = (b == 2) ? 1.0 : 0.0;
            }
        }
    }
    __syncthreads();

    // ---- phase C
    for (int t = warp; t < ntile; t += NW) {
        if (tile_e[t] == INT_MIN) continue;
        const int k = t * 32 + lane;
        cstart[k] = tile_carry[t] + cstart[k];                   // exact (carry and prefix are multiples of q)
    }
}

template <bool DIV>
__global__ void __launch_bounds__(SW_THREADS) sc_walk(const double *__restrict__ d, int64_t n, const double *Sptr,
                                                      double *__restrict__ D, const int32_t *__restrict__ cls, int nch,
                                                      double *__restrict__ cstart, double *__restrict__ tile_tot,
                                                      double *__restrict__ tile_carry, int32_t *__restrict__ tile_e,
                                                      int32_t *__restrict__ bad, double *__restrict__ stats) {
    extern __shared__ double xs_dyn[];
    walk_body<DIV>(d, n, Sptr, D, cls, nch, cstart, tile_tot, tile_carry, tile_e, bad, stats, xs_dyn);
}

// 5. locate the draws: idx = #{ j : fl(c_j / c_last) <= u }.  Division by c_last > 0 is monotone, so the predicate
// equals c_j <= t with t the LARGEST double whose quotient rounds to <= u; t is found once per draw and the
// per-element division disappears from the sequential replay.
__device__ __forceinline__ double quotient_threshold(double u, double clast) {
    double t = u * clast;
    if (!(t == t) || !(clast > 0.0) || !isfinite(clast)) return t;
    for (int it = 0; it < 8 && !(t / clast <= u); ++it) t = __longlong_as_double(__double_as_longlong(t) - 1);   // t > 0 here
    for (int it = 0; it < 8; ++it) {
        const double up = t == 0.0 ? __longlong_as_double(1) : __longlong_as_double(__double_as_longlong(t) + 1);
        if (up / clast <= u) t = up; else break;
    }
    return t;
}

__device__ __forceinline__ void locate_body(const double *__restrict__ d, int64_t n, const double *Sptr,
                                            const double *__restrict__ cstart, int nch, double uq, int64_t *__restrict__ out,
                                            int lane, double *xs) {
    const double S = *Sptr;
    const double clast = cstart[nch];
    const double thr = quotient_threshold(uq, clast);
    // first chunk whose END value exceeds the threshold: cdf is non-decreasing, cstart[k+1] is the value at chunk k's end
    int lo = 0, hi = nch - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cstart[mid + 1] > thr) hi = mid; else lo = mid + 1;
    }
    const int64_t base = (int64_t)lo * SC_CH;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        int64_t j = base + t * 32 + lane;
        xs[t * 32 + lane] = j < n ? d[j] / S : 0.0;
    }
    __syncwarp();
    if (lane == 0) {
        double c = cstart[lo];
        int cnt = 0;
        for (int j = 0; j < SC_CH; ++j) {
            c += xs[j];
            if (c <= thr) ++cnt; else break;
        }
        int64_t r = base + cnt;
        *out = r < n ? r : n - 1;
    }
}

__global__ void __launch_bounds__(32) sc_locate(const double *__restrict__ d, int64_t n, const double *Sptr,
                                                const double *__restrict__ cstart, int nch, const double *__restrict__ u,
                                                int64_t *__restrict__ idx) {
    __shared__ double xs[SC_CH];
    locate_body(d, n, Sptr, cstart, nch, u[blockIdx.x], idx + blockIdx.x, threadIdx.x, xs);
}

// Small problems (n_pad <= SC_FUSED_MAX): all seven phases in ONE CTA separated by CTA barriers -- the multi-kernel
// pipeline above is a chain of latency-bound launches there (simple RPCholesky samples once per step).
constexpr int64_t SC_FUSED_MAX = 65536;

__global__ void __launch_bounds__(SW_THREADS) sc_fused_small(const double *__restrict__ d, int64_t n, const double *__restrict__ u,
                                                             int m, int64_t *__restrict__ idx, double *__restrict__ stats,
                                                             double *__restrict__ csum, double *__restrict__ D,
                                                             double *__restrict__ cA, double *__restrict__ cB,
                                                             double *__restrict__ tile_tot, double *__restrict__ tile_carry,
                                                             int32_t *__restrict__ cls, int32_t *__restrict__ bad,
                                                             int32_t *__restrict__ tile_e, int nch, double delta, double delta_b) {
    extern __shared__ double xs_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = SW_THREADS / 32;
    const double *Sptr = cA + nch;
    if (threadIdx.x < 2) bad[threadIdx.x] = 0;
    __syncthreads();
    for (int k = warp; k < nch; k += NW) chunk_sums_body<false>(d, n, nullptr, csum, bad, k, lane);
    __syncthreads();
    scan_chunks_body(csum, nch);
    __syncthreads();
    for (int k = warp; k < nch; k += NW) classify_body<false>(d, n, nullptr, csum, delta, D, cls, k, lane);
    __syncthreads();
    walk_body<false>(d, n, nullptr, D, cls, nch, cA, tile_tot, tile_carry, tile_e, bad, nullptr, xs_dyn);
    __syncthreads();
    for (int k = warp; k < nch; k += NW) classify_body<true>(d, n, Sptr, cA, delta_b, D, cls, k, lane);
    __syncthreads();
    walk_body<true>(d, n, Sptr, D, cls, nch, cB, tile_tot, tile_carry, tile_e, bad, stats, xs_dyn);
    __syncthreads();
    for (int q = warp; q < m; q += NW) locate_body(d, n, Sptr, cB, nch, u[q], idx + q, lane, xs_dyn + warp * SC_CH);
}

extern "C" int rpc_sample_pivots(rpc_ctx *ctx, void *stream, const double *diag, int64_t n_pad, const double *u, int m,
                                 int64_t *idx, double *stats) {
    RPC_CHECK_ARG(ctx && diag && stats, "null pointer");
    RPC_CHECK_ARG(n_pad > 0 && m >= 0, "bad sizes");
    RPC_CHECK_ARG(m == 0 || (u && idx), "uniforms and idx required when m > 0");
    cudaStream_t st = (cudaStream_t)stream;
    const int nch = (int)((n_pad + SC_CH - 1) / SC_CH);
    // workspace layout (doubles): csum[nch+1] | D[nch] | cstartA[nch+1] | cstartB[nch+1] | tile_tot | tile_carry ;
    // ints: cls[nch] | bad[2] | tile_e
    const int ntile = (nch + 31) / 32;
    size_t nd = (size_t)4 * (nch + 2) + 2 * (size_t)(ntile + 2);
    int rc = rpc_ws_reserve(ctx, nd * sizeof(double));
    if (rc) return rc;
    rc = rpc_flags_reserve(ctx, (size_t)(nch + ntile + 16) * sizeof(int32_t));
    if (rc) return rc;
    double *csum = ctx->ws, *D = csum + (nch + 2), *cA = D + (nch + 2), *cB = cA + (nch + 2);
    double *tile_tot = cB + (nch + 2), *tile_carry = tile_tot + (ntile + 2);
    int32_t *cls = ctx->flags, *bad = ctx->flags + nch, *tile_e = ctx->flags + nch + 8;
    static unsigned long long configured = 0;   // one bit per device: function attributes are per device
    if (!rpc_dev_bit(configured, ctx)) {
        RPC_CUDA(cudaFuncSetAttribute(sc_walk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SLOTS * SC_CH * 8));
        RPC_CUDA(cudaFuncSetAttribute(sc_walk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SLOTS * SC_CH * 8));
        RPC_CUDA(cudaFuncSetAttribute(sc_fused_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_SLOTS * SC_CH * 8));
        configured |= 1ull << (ctx->device & 63);
    }
    const double delta = ((double)n_pad + 256.0) * 1.1102230246251565e-16 * 1.01;
    if (n_pad <= SC_FUSED_MAX) {
        const double delta_b = (2.0 * (double)n_pad + 512.0) * 1.1102230246251565e-16 * 1.01;
        sc_fused_small<<<1, SW_THREADS, SW_SLOTS * SC_CH * sizeof(double), st>>>(diag, n_pad, u, m, idx, stats, csum, D, cA, cB,
                                                                                 tile_tot, tile_carry, cls, bad, tile_e, nch,
                                                                                 delta, delta_b);
        RPC_LAUNCHED(ctx);
        return 0;
    }
    RPC_CUDA(cudaMemsetAsync(bad, 0, 2 * sizeof(int32_t), st));
    const unsigned grid = (unsigned)((nch + SC_WARPS - 1) / SC_WARPS);
    const double *Sptr = cA + nch;     // exact python sum(diags)

    sc_chunk_sums<false><<<grid, SC_WARPS * 32, 0, st>>>(diag, n_pad, nullptr, csum, bad);
    RPC_LAUNCHED(ctx);
    sc_scan_chunks<<<1, 1024, 0, st>>>(csum, nch);
    RPC_LAUNCHED(ctx);
    sc_classify<false><<<grid, SC_WARPS * 32, 0, st>>>(diag, n_pad, nullptr, csum, delta, D, cls);
    RPC_LAUNCHED(ctx);
    sc_walk<false><<<1, SW_THREADS, SW_SLOTS * SC_CH * sizeof(double), st>>>(diag, n_pad, nullptr, D, cls, nch, cA, tile_tot, tile_carry,
                                                                             tile_e, bad, nullptr);
    RPC_LAUNCHED(ctx);

    // pass B: x_j = d_j / S.  Its approximate prefix comes from the exact pass-A carries (no second reduction).
    const double delta_b = (2.0 * (double)n_pad + 512.0) * 1.1102230246251565e-16 * 1.01;
    sc_classify<true><<<grid, SC_WARPS * 32, 0, st>>>(diag, n_pad, Sptr, cA, delta_b, D, cls);
    RPC_LAUNCHED(ctx);
    sc_walk<true><<<1, SW_THREADS, SW_SLOTS * SC_CH * sizeof(double), st>>>(diag, n_pad, Sptr, D, cls, nch, cB, tile_tot, tile_carry,
                                                                            tile_e, bad, stats);
    RPC_LAUNCHED(ctx);
    if (m > 0) {
        sc_locate<<<m, 32, 0, st>>>(diag, n_pad, Sptr, cB, nch, u, idx);
        RPC_LAUNCHED(ctx);
    }
    return 0;
}
