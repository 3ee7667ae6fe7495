"""Nystrom-preconditioned conjugate gradient for kernel ridge regression -- SURVEY.md section 8(f) rank 2.

The reference has this only inline in an experiment script (block_experiments/pes_code.py:463-524); the pieces are kept
under the script's own names:

    precondition = NystromPreconditioner(lra, mu)     # :464-475  thin QR of G^T, SVD of R, P^{-1}v = U((S^2+mu)^{-1} - mu^{-1})U^T v + v/mu
    x, resid = pcg(A, y, mu, precondition, tol, max_iters)   # :479-511  solves (A + mu I) x = y

`A` is a device-backed KernelMatrix / PSDMatrix.  Every N-proportional operation runs in the library's own kernels: the
mat-vec A @ p is the fused kernel-times-vector kernel (`rpc_kernel_rows_wsum`: the script's blocked
`np.dot(A[blocks[i]:blocks[i+1], :], p)`, :502-505, without ever writing a block of A), the preconditioner works on the factor
`lra.G` in HBM through `rpc_syrk_rows` / `rpc_rows_dot` / `rpc_rows_tdot`; the only library call is one k x k `eigh`.
`CompactEigenvalueDecomposition.from_G(G).krr` (lra.py:51-55, 78-80) is the same operator in eigen form.  The loop is pinned to
the reference's own statements by tests/golden/pcg_loop_*.npz (tests/golden/make_golden_pcg.py executes them unchanged).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import ops
from .matrix import AbstractPSDMatrix


class NystromPreconditioner:
    """pes_code.py:464-475 for a factor G (k x N, A ~ G^T G) and shift mu:  P^{-1} v = U ((S^2 + mu)^{-1} - mu^{-1}) U^T v + v / mu
    with G^T = U S W^T.  The reference forms U explicitly from a thin QR of G^T and an SVD of R; here U = G^T W S^{-1} stays
    implicit: W, S^2 come from the k x k eigenproblem of G G^T (`rpc_syrk_rows`, then one small library eigh), and an application
    is two passes over G in HBM (`rpc_rows_dot` for G v, `rpc_rows_tdot` for G^T w) -- no N x k matrix besides G is ever formed.
    Directions with S^2 below eps * S_max^2 * k are dropped: their coefficient (S^2 + mu)^{-1} - mu^{-1} vanishes with S^2."""

    def __init__(self, lra, mu):
        G = lra.G_device if hasattr(lra, "G_device") else lra
        if not isinstance(G, torch.Tensor):
            raise RuntimeError("NystromPreconditioner needs a device-backed PSDLowRank (or a CUDA tensor G)")
        self.G = G
        self.n = lra.shape[0] if hasattr(lra, "shape") and not isinstance(lra, torch.Tensor) else G.shape[1]
        k = G.shape[0]
        S2, W = torch.linalg.eigh(ops.syrk_rows(G, self.n))       # G G^T = W S^2 W^T (k x k)
        keep = S2 > S2.max() * k * torch.finfo(torch.float64).eps
        self.Sigma = torch.sqrt(S2.clamp_min(0.0))
        inv = torch.where(keep, 1.0 / self.Sigma.clamp_min(1e-300), torch.zeros_like(S2))
        self.WSinv = W * inv[None, :]                              # k x k: U = G^T (W S^{-1})
        self.mu = float(mu)
        self.diag = torch.where(keep, 1.0 / (S2 + self.mu) - 1.0 / self.mu, torch.zeros_like(S2))

    def __call__(self, v):
        if v.ndim != 1:
            return torch.stack([self(v[:, c]) for c in range(v.shape[1])], dim=1)
        t = self.WSinv.T @ ops.rows_dot(self.G, v, self.n)         # U^T v       (k)
        w = self.WSinv @ (self.diag * t)                           # k x k products: tiny
        return ops.rows_tdot(self.G, w, self.n) + v / self.mu      # :471-473


def kernel_matvec(A: AbstractPSDMatrix, p: torch.Tensor, block_rows: int = 256) -> torch.Tensor:
    """A @ p (pes_code.py:500-506, `np.dot(A[blocks[i]:blocks[i+1], :], p)` block by block); counts N^2 queries.
    KernelMatrix: the fused kernel-times-vector kernel `rpc_kernel_rows_wsum` -- out[i] = sum_j p[j] k(x_j, x_i) with blocks of
    pivots j resident and all points i streamed, so no block of A is ever written to memory.  Dense PSDMatrix: `rpc_rows_dot`."""
    n = A.shape[0]
    dev = A.device
    if p.ndim != 1:
        return torch.stack([kernel_matvec(A, p[:, c].contiguous(), block_rows) for c in range(p.shape[1])], dim=1)
    if hasattr(A, "_spec"):
        out = ops.kernel_weighted_sum(A, A._Xd, A._xnorm, n, p.contiguous())[:n]
    else:
        out = ops.rows_dot(A._Ad, p.contiguous(), n)
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
