"""Nystrom-preconditioned conjugate gradient for kernel ridge regression -- SURVEY.md section 8(f) rank 2.

The reference has this only inline in an experiment script (block_experiments/pes_code.py:463-524); the pieces are kept
under the script's own names:

    precondition = NystromPreconditioner(lra, mu)     # :464-475  thin QR of G^T, SVD of R, P^{-1}v = U((S^2+mu)^{-1} - mu^{-1})U^T v + v/mu
    x, resid = pcg(A, y, mu, precondition, tol, max_iters)   # :479-511  solves (A + mu I) x = y

`A` is a device-backed KernelMatrix / PSDMatrix: the mat-vec A @ p is evaluated block of rows by block of rows with the
CUDA kernel-row evaluator (the script's `np.dot(A[blocks[i]:blocks[i+1], :], p)`, :502-505), the factor `lra.G` stays in
HBM, QR / SVD / GEMV are library calls (cuSOLVER / cuBLAS through torch).  `CompactEigenvalueDecomposition.from_G(G).krr`
(lra.py:51-55, 78-80) is the same operator in eigen form.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .matrix import AbstractPSDMatrix


class NystromPreconditioner:
    """pes_code.py:464-475 for a factor G (k x N, A ~ G^T G) and shift mu."""

    def __init__(self, lra, mu):
        G = lra.G_device if hasattr(lra, "G_device") else lra
        if not isinstance(G, torch.Tensor):
            raise RuntimeError("NystromPreconditioner needs a device-backed PSDLowRank (or a CUDA tensor G)")
        Qm, R = torch.linalg.qr(G.T, mode="reduced")          # :467
        Cm, Sigma, _ = torch.linalg.svd(R)                     # :468
        self.U = Qm @ Cm                                       # :469
        self.Sigma = Sigma
        self.mu = float(mu)
        self.diag = 1.0 / (Sigma ** 2 + self.mu) - 1.0 / self.mu

    def __call__(self, v):
        t = self.U.T @ v
        t = self.diag * t if t.ndim == 1 else self.diag[:, None] * t
        return self.U @ t + v / self.mu                        # :471-473


def kernel_matvec(A: AbstractPSDMatrix, p: torch.Tensor, block_rows: int = 256) -> torch.Tensor:
    """A @ p with the rows of A evaluated block by block on the device (pes_code.py:502-505); counts N^2 queries."""
    n = A.shape[0]
    dev = A.device
    out = torch.empty((n,) + tuple(p.shape[1:]), dtype=torch.float64, device=dev)
    buf = torch.empty((block_rows, A.n_pad), dtype=torch.float64, device=dev)
    piv = A._dev_new_payload(block_rows) if hasattr(A, "_dev_new_payload") else None
    for r0 in range(0, n, block_rows):
        m = min(block_rows, n - r0)
        idx = torch.arange(r0, r0 + m, dtype=torch.int64, device=dev)
        if piv is not None:
            A._dev_gather(idx, m, piv)
        A._dev_rows_into(idx, m, piv, C.c_void_p(buf.data_ptr()), A.n_pad)
        out[r0:r0 + m] = buf[:m, :n] @ p
    A._count(n * n)
    return out


def pcg(A, y, mu, precondition=None, tol=1e-6, max_iters=100, block_rows=256):
    """Preconditioned CG on (A + mu I) x = y, iterate for iterate as pes_code.py:479-511.
    Returns (x as numpy, list of relative residuals ||r|| / ||y||, one per iteration starting with iteration 0)."""
    dev = A.device
    yd = y.to(dev, torch.float64) if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(dev)
    if precondition is None:
        precondition = lambda v: v                              # noqa: E731  (:476-477)
    x = torch.zeros_like(yd)
    r = yd.clone()
    ynorm = torch.sum(yd ** 2) ** 0.5
    resid = [float(torch.sum(r ** 2) ** 0.5 / ynorm)]
    z = precondition(r)
    p = z.clone()
    rz = r @ z
    for _ in range(1, max_iters + 1):
        Kp = kernel_matvec(A, p, block_rows) + mu * p          # :500-506
        alpha = rz / (p @ Kp)
        x = x + alpha * p
        r = r - alpha * Kp
        resid.append(float(torch.sum(r ** 2) ** 0.5 / ynorm))
        if resid[-1] < tol:                                     # :511-512
            break
        z = precondition(r)
        rz_new = r @ z
        beta = rz_new / rz
        rz = rz_new
        p = z + beta * p
    return x.cpu().numpy(), resid
