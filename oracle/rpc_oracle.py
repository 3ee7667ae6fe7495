"""CPU oracle for the RPCholesky hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference algorithm (eepperly/Randomly-Pivoted-Cholesky) for the
path named by BASELINE.json.north_star.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product
package ``randomly_pivoted_cholesky_b200`` never does.

Parity status: PINNED AGAINST THE LIVE REFERENCE.  The reference ships no golden vectors for this
path (its tests/test_lra.py covers lra.py only), so the oracle is pinned by
``tests/golden/make_golden.py``, which imports the unmodified reference from /root/reference under
the RNG replay patch and stores its outputs (pivots, G, rows, kernel blocks) as fixtures under
``tests/golden/``; ``tests/test_oracle.py`` checks this module against every fixture.

Third-party arithmetic restated here (not vendored in /root/reference, no version pinned upstream;
versions installed in the survey/build container: scikit-learn 1.9.0, scipy 1.18.1, numpy 2.3.5):
  * sklearn.metrics.pairwise.euclidean_distances (float64 branch, sklearn/metrics/pairwise.py:376-426):
        D = sqrt(max(((-2 * X @ Y.T) + ||x||^2) + ||y||^2, 0)),  norms by einsum('ij,ij->i').
  * sklearn.metrics.pairwise.manhattan_distances -> scipy cdist(X, Y, 'cityblock'): sum_f |x_f - y_f|
    accumulated strictly in feature order.
  * numpy Generator.choice(n, size, p=p): cdf = cumsum(p); cdf /= cdf[-1];
        idx = searchsorted(cdf, Generator.random(size), side='right').

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
import warnings

import numpy as np

EPS = float(np.finfo(float).eps)


# ----------------------------------------------------------------------------------------------
# kernel evaluators (reference kernels.py)
# ----------------------------------------------------------------------------------------------
def euclidean_distances(X, Y):
    """sklearn `_euclidean_distances`, float64 branch (called from kernels.py:38,77)."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    XX = np.einsum("ij,ij->i", X, X)[:, None]
    YY = np.einsum("ij,ij->i", Y, Y)[None, :]
    D = -2 * (X @ Y.T)
    D += XX
    D += YY
    np.maximum(D, 0, out=D)
    np.sqrt(D, out=D)
    return D


def manhattan_distances(X, Y):
    """sklearn manhattan_distances as reached from kernels.py:20 = scipy cdist(X, Y, 'cityblock') (the call the
    reference itself makes); without scipy, the restated loop below (bit-equal: tests/test_oracle.py)."""
    try:
        from scipy.spatial.distance import cdist
    except ImportError:
        return manhattan_distances_loop(X, Y)
    return cdist(np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64), "cityblock")


def manhattan_distances_loop(X, Y):
    """cdist(..., 'cityblock') restated: sum_f |x_f - y_f| accumulated strictly in feature order."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    D = np.zeros((X.shape[0], Y.shape[0]))
    tmp = np.empty_like(D)
    for f in range(X.shape[1]):
        np.subtract(X[:, f][:, None], Y[:, f][None, :], out=tmp)
        np.abs(tmp, out=tmp)
        D += tmp
    return D


def direct_distances(X, Y):
    """`np.linalg.norm(xx[:,None,:]-yy[None,:,:], axis=-1)` (kernels.py:36,75), chunked over X."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    out = np.empty((X.shape[0], Y.shape[0]))
    for i in range(X.shape[0]):
        out[i] = np.linalg.norm(X[i][None, :] - Y, axis=-1)
    return out


def gaussian_mtx(xx, yy, bandwidth=1.0, extra_stability=False):
    """GaussianKernel_mtx, kernels.py:32-39."""
    dsts = direct_distances(xx, yy) if extra_stability else euclidean_distances(xx, yy)
    return np.exp(-0.5 * dsts**2 / bandwidth**2)


def laplace_mtx(xx, yy, bandwidth=1.0, **_):
    """LaplaceKernel_mtx, kernels.py:16-21."""
    return np.exp(-manhattan_distances(xx, yy) / bandwidth)


def matern_mtx(xx, yy, bandwidth=1.0, nu=0.5, extra_stability=False):
    """MaternKernel_mtx for nu in {0.5, 1.5, 2.5}, kernels.py:71-85.

    The reference omits ``/bandwidth`` on the extra_stability branch (kernels.py:74-75); that quirk
    is reproduced because parity is defined against the reference as it is.
    """
    if extra_stability:
        d = direct_distances(xx, yy)
    else:
        d = euclidean_distances(xx, yy) / bandwidth
    if nu == 0.5:
        return np.exp(-d)
    if nu == 1.5:
        s3 = np.sqrt(3.0)
        return (1 + s3 * d) * np.exp(-s3 * d)
    if nu == 2.5:
        s5 = np.sqrt(5.0)
        return (1 + s5 * d + 5.0 / 3.0 * d * d) * np.exp(-s5 * d)
    raise NotImplementedError("generic-nu Matern is out of scope (its diagonal is NaN in the reference)")


_MTX = {"gaussian": gaussian_mtx, "laplace": laplace_mtx, "matern": matern_mtx}


class OracleKernelMatrix:
    """KernelMatrix restated (matrix.py:140-192) with the operator protocol of matrix.py:25-77.

    Only what the hot path calls: shape, diag(), A[idx,:], A[idx,idx] (outer block), trace(),
    num_queries(), reset().  Query counting follows matrix.py:25-39 (result size per call).
    """

    def __init__(self, X, kernel="gaussian", bandwidth=1.0, **kwargs):
        self.data = np.ascontiguousarray(X, dtype=np.float64)
        self.shape = (self.data.shape[0], self.data.shape[0])
        self.kernel_name = kernel
        self.bandwidth = float(bandwidth)
        self.kwargs = {k: v for k, v in kwargs.items() if k in ("nu", "extra_stability")}
        if kernel not in _MTX:
            raise RuntimeError("Kernel name {} not recognized".format(kernel))
        self.queries = 0

    def _mtx(self, xi, xj):
        return _MTX[self.kernel_name](xi, xj, bandwidth=self.bandwidth, **self.kwargs)

    def diag(self):
        # *_vec on identical points: distance exactly 0 -> exactly 1.0 (kernels.py:11-14,27-30,56-66)
        d = np.ones(self.shape[0])
        self.queries += d.size
        return d

    def rows(self, idx):
        """A[idx,:] (matrix.py:129-138,188-189: both operands are fancy-index copies of X)."""
        idx = np.atleast_1d(np.asarray(idx, dtype=np.int64))
        out = self._mtx(self.data[idx, :], self.data[list(range(self.shape[0])), :])
        self.queries += out.size
        return out

    def block(self, idx):
        """A[idx,idx] -> outer (m x m) block (matrix.py:99,135)."""
        idx = np.atleast_1d(np.asarray(idx, dtype=np.int64))
        out = self._mtx(self.data[idx, :], self.data[idx, :])
        self.queries += out.size
        return out

    def trace(self):
        return seq_sum(self.diag())

    def num_queries(self):
        return self.queries

    def reset(self):
        self.queries = 0


class OracleDenseMatrix:
    """PSDMatrix restated (matrix.py:79-102)."""

    def __init__(self, A):
        self.matrix = np.asarray(A, dtype=np.float64)
        self.shape = self.matrix.shape
        self.queries = 0

    def diag(self):
        d = self.matrix[range(self.shape[0]), range(self.shape[0])]
        self.queries += d.size
        return d

    def rows(self, idx):
        idx = np.atleast_1d(np.asarray(idx, dtype=np.int64))
        out = self.matrix[np.ix_(idx, range(self.shape[0]))]
        self.queries += out.size
        return out

    def block(self, idx):
        idx = np.atleast_1d(np.asarray(idx, dtype=np.int64))
        out = self.matrix[np.ix_(idx, idx)]
        self.queries += out.size
        return out

    def trace(self):
        return seq_sum(self.diag())

    def num_queries(self):
        return self.queries

    def reset(self):
        self.queries = 0


# ----------------------------------------------------------------------------------------------
# pivot sampler (numpy Generator.choice as called at rpcholesky.py:31,91,184)
# ----------------------------------------------------------------------------------------------
def seq_sum(v):
    """python builtin ``sum`` over float64 scalars = left-to-right fp64 accumulation; numpy's
    ``cumsum`` performs the same sequence of additions (checked in tests/test_oracle.py)."""
    v = np.asarray(v, dtype=np.float64)
    if v.size == 0:
        return 0.0
    return float(np.cumsum(v)[-1])


def sample_pivots(diags, u):
    """idx = rng.choice(range(n), size=len(u), p=diags/sum(diags), replace=True) given the uniforms.

    Raises ValueError like Generator.choice when the probabilities are not a distribution
    (all-zero or NaN residual diagonal; callers rely on it: experiments/comparison.py:51).
    """
    diags = np.asarray(diags, dtype=np.float64)
    S = seq_sum(diags)
    with np.errstate(invalid="ignore", divide="ignore"):
        p = diags / S
    if np.isnan(p).any():
        raise ValueError("probabilities contain NaN")
    if (p < 0).any():
        raise ValueError("probabilities are not non-negative")
    cdf = np.cumsum(p)
    if not np.isfinite(cdf[-1]) or abs(cdf[-1] - 1.0) > math.sqrt(EPS):
        raise ValueError("probabilities do not sum to 1")
    cdf /= cdf[-1]
    return np.searchsorted(cdf, np.asarray(u, dtype=np.float64), side="right").astype(np.int64)


# ----------------------------------------------------------------------------------------------
# the three factorisation loops (reference rpcholesky.py)
# ----------------------------------------------------------------------------------------------
class OracleResult:
    """PSDLowRank fields (lra.py:87-116) + the trace-norm error of utils.py:11-15."""

    def __init__(self, G, idx, rows):
        self.G, self.idx, self.rows = G, idx, rows

    def trace(self):
        return np.linalg.norm(self.G, "fro") ** 2


def _as_operator(A):
    if isinstance(A, (OracleKernelMatrix, OracleDenseMatrix)):
        return A
    return OracleDenseMatrix(A)


def simple_rpcholesky(A, k, rng, stoptol=0, alg="rp"):
    """cholesky_helper(A, k, alg, stoptol) -- rpcholesky.py:12-50.  `rng` supplies the proposal
    stream exactly as `np.random.default_rng()` does in the reference (one double per step for 'rp';
    'rgreedy' draws through rng.choice over the tied maxima, 'greedy' draws nothing)."""
    A = _as_operator(A)
    n = A.shape[0]
    diags = A.diag().copy()
    orig_trace = seq_sum(diags)
    if stoptol is None:
        stoptol = 0
    G = np.zeros((k, n))
    rows = np.zeros((k, n))
    arr_idx = []
    for i in range(k):
        if alg == "rp":
            idx = int(sample_pivots(diags, np.array([rng.random()]))[0])      # :31
        elif alg == "rgreedy":
            idx = int(rng.choice(np.where(diags == np.max(diags))[0]))        # :33
        elif alg == "greedy":
            idx = int(np.argmax(diags))                                        # :35
        else:
            raise RuntimeError("Algorithm '{}' not recognized".format(alg))
        arr_idx.append(idx)
        rows[i, :] = A.rows(idx)[0]                                            # :40
        G[i, :] = (rows[i, :] - G[:i, idx].T @ G[:i, :]) / np.sqrt(diags[idx])  # :41
        diags -= G[i, :] ** 2                                                  # :42
        diags = diags.clip(min=0)                                              # :43
        if stoptol > 0 and seq_sum(diags) <= stoptol * orig_trace:            # :45-48
            G = G[:i, :]
            rows = rows[:i, :]
            break
    return OracleResult(G, arr_idx, rows)


def _greedy_cholesky(H, rtol=None, atol=None):
    """_greedy_cholesky -- rpcholesky.py:52-63 (LAPACK dpstrf through scipy, like the reference)."""
    from scipy.linalg import get_lapack_funcs
    pstrf = get_lapack_funcs("pstrf", dtype=np.float64)
    trace = np.trace(H)
    L, piv, rank, info = pstrf(H, lower=True)
    L = np.tril(L)
    if not (rtol is None) and not (atol is None):
        trailing_sums = trace - np.cumsum(np.linalg.norm(L, axis=0) ** 2)
        rank = np.argmax(trailing_sums < (atol + rtol * trace)) + 1
    return L, piv, rank


def block_rpcholesky(A, k, b, rng, stoptol=1e-14, strategy="regularize", rbrp_atol=0.0, rbrp_rtol="1/b", alg="rp"):
    """block_cholesky_helper(A, k, b, alg, stoptol, strategy, rbrp_atol, rbrp_rtol) -- rpcholesky.py:65-138.
    alg 'greedy' takes the block_size largest residuals (:94-95) in ascending index order: numpy's argpartition order
    (and its choice among ties) is an implementation detail that the device path does not reproduce."""
    A = _as_operator(A)
    diags = A.diag().copy()
    n = A.shape[0]
    orig_trace = seq_sum(diags)
    scale = 2 * max(diags)                                                     # :72
    if stoptol is None:
        stoptol = 1e-14
    if rbrp_rtol == "1/b":
        rbrp_rtol = 1.0 / b                                                    # :75-76
    G = np.zeros((k, n))
    rows = np.zeros((k, n))
    arr_idx = []
    counter = 0
    conds = []          # diagnostics only, see g_tolerance
    while counter < k:
        block_size = min(k - counter, b)
        if alg == "rp":
            idx = sample_pivots(diags, rng.random(2 * block_size))            # :91
            idx = np.unique(idx)[:block_size]                                  # :92
            block_size = len(idx)
        elif alg == "greedy":
            # the set of np.argpartition(diags, -block_size)[-block_size:], lowest indices among ties at the threshold
            order = np.lexsort((np.arange(n), -diags))
            idx = np.sort(order[:block_size])
        else:
            raise RuntimeError("Algorithm '{}' not recognized".format(alg))
        if strategy in ("regularize", "regularized"):
            arr_idx.extend(idx)
            rows[counter:counter + block_size, :] = A.rows(idx)                # :101
            G[counter:counter + block_size, :] = (
                rows[counter:counter + block_size, :] - G[0:counter, idx].T @ G[0:counter, :])  # :102
            C = G[counter:counter + block_size, idx]                           # :103
            try:
                L = np.linalg.cholesky(C + EPS * b * scale * np.identity(block_size))  # :106
                conds.append(float(np.linalg.cond(L)))
                G[counter:counter + block_size, :] = np.linalg.solve(L, G[counter:counter + block_size, :])  # :107
            except np.linalg.LinAlgError:                                      # :109-114
                warnings.warn("Cholesky failed in block partial Cholesky. Falling back to eigendecomposition")
                evals, evecs = np.linalg.eigh(C)
                evals[evals > 0] = evals[evals > 0] ** (-0.5)
                evals[evals < 0] = 0
                G[counter:counter + block_size, :] = evals[:, np.newaxis] * (evecs.T @ G[counter:counter + block_size, :])
        elif strategy in ("rbrp", "pivoting", "pivoted"):                      # :115-123
            H = A.block(idx) - G[0:counter, idx].T @ G[0:counter, idx]
            L, piv, rank = _greedy_cholesky(H, rtol=rbrp_rtol, atol=rbrp_atol)
            idx = idx[piv[0:rank] - 1]
            conds.append(float(np.linalg.cond(L[0:rank, 0:rank])))
            arr_idx.extend(idx)
            rows[counter:counter + len(idx), :] = A.rows(idx)
            G[counter:counter + len(idx), :] = rows[counter:counter + len(idx), :] - G[0:counter, idx].T @ G[0:counter, :]
            G[counter:counter + len(idx), :] = np.linalg.solve(L[0:rank, 0:rank], G[counter:counter + len(idx), :])
        else:
            raise ValueError("'{}' is not a valid strategy for block RPCholesky".format(strategy))
        diags -= np.sum(G[counter:counter + len(idx), :] ** 2, axis=0)         # :128
        diags = diags.clip(min=0)                                              # :129
        counter += len(idx)
        if stoptol > 0 and seq_sum(diags) <= stoptol * orig_trace:            # :133-136
            G = G[:counter, :]
            rows = rows[:counter, :]
            break
    res = OracleResult(G, arr_idx, rows)
    res.cond_L = conds
    return res


def rejection_cholesky(H, u):
    """rejection_cholesky(H) -- rpcholesky.py:140-157; `u` replaces the b `np.random.rand()` calls.
    H is modified in place, as in the reference."""
    b = H.shape[0]
    if H.shape[0] != H.shape[1]:
        raise RuntimeError("rejection_cholesky requires a square matrix")
    if np.trace(H) <= 0:
        raise RuntimeError("rejection_cholesky requires a strictly positive trace")
    u0 = np.array([H[j, j] for j in range(b)])
    idx = []
    L = np.zeros((b, b))
    for j in range(b):
        if u[j] * u0[j] < H[j, j]:                                             # :151
            idx.append(j)
            L[j:, j] = H[j:, j] / np.sqrt(H[j, j])
            H[(j + 1):, (j + 1):] -= np.outer(L[(j + 1):, j], L[(j + 1):, j])
    idx = np.array(idx, dtype=np.int64)
    L = L[np.ix_(idx, idx)]
    return L, idx


def accelerated_rpcholesky(A, k, b, rng, legacy, stoptol=1e-13):
    """accelerated_rpcholesky(A, k, b, stoptol) with a FIXED b -- rpcholesky.py:159-232.
    `rng` = proposal stream (Generator), `legacy` = np.random.RandomState standing in for the global
    legacy stream that `np.random.rand()` reads at rpcholesky.py:151 (b draws per round)."""
    A = _as_operator(A)
    diags = A.diag().copy()
    n = A.shape[0]
    orig_trace = seq_sum(diags)
    if stoptol is None:
        stoptol = 1e-13
    G = np.zeros((k, n))
    rows = np.zeros((k, n))
    arr_idx = np.zeros(k, dtype=int)
    counter = 0
    rounds = []
    conds = []          # diagnostics only: 2-norm condition number of each round's triangular factor (see g_tolerance)
    while counter < k:
        idx = sample_pivots(diags, rng.random(b))                              # :184
        H = A.block(idx) - G[0:counter, idx].T @ G[0:counter, idx]             # :189
        L, accepted = rejection_cholesky(H, legacy.random_sample(b))           # :190
        num_sel = len(accepted)
        if num_sel > k - counter:                                              # :193-196
            num_sel = k - counter
            accepted = accepted[:num_sel]
            L = L[:num_sel, :num_sel]
        idx = idx[accepted]
        conds.append(float(np.linalg.cond(L)) if num_sel > 0 else 1.0)
        arr_idx[counter:counter + num_sel] = idx
        rows[counter:counter + num_sel, :] = A.rows(idx)                       # :205
        G[counter:counter + num_sel, :] = (
            rows[counter:counter + num_sel, :] - G[0:counter, idx].T @ G[0:counter, :])  # :206
        G[counter:counter + num_sel, :] = np.linalg.solve(L, G[counter:counter + num_sel, :])  # :207
        diags -= np.sum(G[counter:counter + num_sel, :] ** 2, axis=0)          # :208
        diags = diags.clip(min=0)                                              # :209
        counter += num_sel
        rounds.append(num_sel)
        if stoptol > 0 and seq_sum(diags) <= stoptol * orig_trace:            # :224-227
            G = G[:counter, :]
            rows = rows[:counter, :]
            break
    res = OracleResult(G, arr_idx, rows)
    res.rounds = rounds
    res.cond_L = conds
    return res


def g_tolerance(res, floor=1e-10):
    """Bar for ||G - G_ref||_F / ||G_ref||_F.  north_star's 1e-10 wherever the rounds' triangular factors are well
    conditioned; a block whose factor L has cond_2(L) = kappa turns the round-off of ANY backward-stable solve
    (the reference's LU `np.linalg.solve(L, .)`, rpcholesky.py:107,207, as much as the device's forward substitution)
    into a relative forward error of order kappa * eps in G_new, so two correct implementations may differ by that much:
    the bar is max(1e-10, 8 * eps * max_round kappa).  rows = A[S,:] involve no solve and keep the flat 1e-10."""
    kappa = max(getattr(res, "cond_L", None) or [1.0])
    return max(floor, 8.0 * EPS * kappa)


def relative_trace_error(A, res):
    """utils.approximation_error(A, lra, relative=True) -- utils.py:11-15."""
    tr = A.trace()
    return (tr - res.trace()) / tr


def make_streams(seed):
    """The replay recipe of SURVEY.md section 8(c): proposal stream = Generator(PCG64(seed)) (what the
    patched `np.random.default_rng()` returns), rejection stream = legacy MT19937 seeded by
    `np.random.seed(seed)`."""
    return np.random.Generator(np.random.PCG64(seed)), np.random.RandomState(seed)
