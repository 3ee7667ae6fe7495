"""CPU oracle for the first "next" row of the hot path: KRR_Nystrom -- TEST INFRASTRUCTURE ONLY.

Numpy restatement of the reference's KRR_Nystrom.fit_Nystrom (KRR_Nystrom.py:34-58) on top of the RPCholesky oracle,
and of the evident intent of predict_Nystrom (KRR_Nystrom.py:60-66 is broken upstream: it builds
NonsymmetricKernelMatrix(Xts, Xtr[S]) and then multiplies an undefined `KtM`; the intent is K_pred[:,:] @ sol).
Pinned to the live reference by tests/golden/make_golden_krr.py (fixtures krr_*.npz).
"""
import numpy as np
import scipy.linalg

from . import rpc_oracle as orc


def fit_nystrom(Xtr, Ytr, lamb, k, kernel, bandwidth, alg, b, seed, **kw):
    """Returns (sol, sample_idx, reltrace_err, queries) like KRR_Nystrom.fit_Nystrom(..., solve_method='Direct')."""
    A = orc.OracleKernelMatrix(Xtr, kernel=kernel, bandwidth=bandwidth, **kw)
    rng, legacy = orc.make_streams(seed)
    if alg == "accel":
        lra = orc.accelerated_rpcholesky(A, k, b, rng, legacy)
    else:
        lra = orc.block_rpcholesky(A, k, b, rng)
    idx = np.asarray(lra.idx, dtype=np.int64)
    KMn = lra.rows                                                    # :40
    queries = A.num_queries()                                         # :44
    trK = A.trace()                                                   # :46 (counts N more queries, then reset)
    reltrace_err = (trK - lra.trace()) / trK                          # :47
    KMM = KMn[:, idx]                                                 # :52
    KnM = KMn.T
    M = KMn @ KnM + KnM.shape[0] * lamb * KMM + 100 * KMM.max() * np.finfo(float).eps * np.identity(k)   # :54
    sol = scipy.linalg.solve(M, KMn @ Ytr, assume_a="pos")
    return sol, idx, reltrace_err, queries


def predict_nystrom(Xts, Xtr, sample_idx, sol, kernel, bandwidth, **kw):
    """K(Xts, Xtr[S]) @ sol with the reference's kernel_mtx (NonsymmetricKernelMatrix._function_mtx, matrix.py:208-209)."""
    mtx = {"gaussian": orc.gaussian_mtx, "laplace": orc.laplace_mtx, "matern": orc.matern_mtx}[kernel]
    return mtx(Xts, Xtr[np.asarray(sample_idx), :], bandwidth=bandwidth, **kw) @ sol
