"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the Nystrom-preconditioned CG of the reference's experiment script
(block_experiments/pes_code.py:463-524; the script itself needs MD17 data and cannot run here, so this piece is pinned
through its building blocks: the preconditioner against the reference's own CompactEigenvalueDecomposition.from_G(G).krr
(lra.py:51-55, 78-80, fixture tests/golden/pcg_precond.npz) and the solve against numpy.linalg.solve).
Never imported by the product package."""
import numpy as np


def nystrom_preconditioner(G, mu):
    """pes_code.py:464-475 (numpy qr instead of scipy's: same LAPACK routine)."""
    Q, C = np.linalg.qr(G.T, mode="reduced")
    C, Sigma, _ = np.linalg.svd(C)
    U = Q @ C

    def precondition(v):
        diag = 1 / (Sigma ** 2 + mu) - 1 / mu
        return U @ (diag * (U.T @ v)) + v / mu
    return precondition


def pcg(Afull, y, mu, precondition=None, tol=1e-6, max_iters=100):
    """pes_code.py:479-511 on a dense matrix."""
    if precondition is None:
        precondition = lambda v: v            # noqa: E731
    x = np.zeros_like(y)
    r = np.copy(y)
    resid = [np.sum(r ** 2) ** .5 / np.sum(y ** 2) ** .5]
    z = precondition(r)
    p = np.copy(z)
    rz = r @ z
    for _ in range(1, max_iters + 1):
        Kp = Afull @ p
        Kp += mu * p
        alpha = rz / (p @ Kp)
        x = x + alpha * p
        r = r - alpha * Kp
        resid.append(np.sum(r ** 2) ** .5 / np.sum(y ** 2) ** .5)
        if resid[-1] < tol:
            break
        z = precondition(r)
        beta = (r @ z) / rz
        rz = r @ z
        p = z + beta * p
    return x, resid
