"""Result type of the hot path: PSDLowRank (reference lra.py:8-41, 87-116).

`PSDLowRank(G, idx=..., rows=...)` keeps the reference's fields and methods.  G / rows may be numpy
arrays (host container semantics, exactly the reference's) or CUDA tensors produced by the
factorisation drivers; device-backed results expose `.G` / `.rows` as numpy arrays materialised on
first access, and `.G_device` / `.rows_device` as (k, N) views of the padded device buffers so that
a 16 GB factor never has to leave HBM.
"""
from __future__ import annotations

import numpy as np
import torch


class AbstractPSDLowRank:
    """lra.py:8-41."""

    def __init__(self, **kwargs):
        self.idx = kwargs["idx"] if ("idx" in kwargs) else None
        self._rows = kwargs["rows"] if ("rows" in kwargs) else None
        self._rows_host = None
        # row-sharded results (dist.ShardInfo): the tensors hold this rank's columns; `.G`/`.rows` assemble the
        # full matrices with an all-gather, so every rank must touch them together
        self.shard_info = kwargs["shard_info"] if ("shard_info" in kwargs) else None

    def _sharded(self):
        return self.shard_info is not None and self.shard_info.world > 1

    def _materialise(self, t):
        if self._sharded():
            from . import dist as rdist
            return rdist.gather_columns(self.shard_info, t).cpu().numpy()
        return t.cpu().numpy()

    @property
    def rows(self):
        if isinstance(self._rows, torch.Tensor):
            if self._rows_host is None:
                self._rows_host = self._materialise(self._rows)
            return self._rows_host
        return self._rows

    @rows.setter
    def rows(self, value):
        self._rows = value
        self._rows_host = None

    @property
    def rows_device(self):
        return self._rows if isinstance(self._rows, torch.Tensor) else None

    def matrix(self):
        return self @ np.identity(self.shape[0])

    def get_rows(self):
        if not (self._rows is None):
            return self.rows
        raise RuntimeError("Rows are not defined for this low-rank approximation")

    def get_indices(self):
        if not (self.idx is None):
            return self.idx
        raise RuntimeError("Indices are not defined for this low-rank approximation")


class CompactEigenvalueDecomposition(AbstractPSDLowRank):
    """lra.py:43-85 (host numpy; off the hot path, kept so PSDLowRank.eigenvalue_decomposition works)."""

    def __init__(self, V, Lambda, **kwargs):
        super().__init__(**kwargs)
        self.V = V
        self.Lambda = Lambda
        self.shape = (self.V.shape[0], self.V.shape[0])   # the reference stores (V[0], V[0]) (lra.py:49, a bug)

    @staticmethod
    def from_G(G, **kwargs):
        Q, R = np.linalg.qr(np.asarray(G).T, "reduced")
        U, S, _ = np.linalg.svd(R)
        return CompactEigenvalueDecomposition(Q @ U, S ** 2, **kwargs)

    def evals(self):
        return self.Lambda

    def evecs(self):
        return self.V

    def trace(self):
        return sum(self.Lambda)

    def __matmul__(self, other):
        if other.ndim == 1:
            return self.V @ (self.Lambda * (self.V.T @ other))
        return self.V @ (self.Lambda[:, np.newaxis] * (self.V.T @ other))

    def rank(self):
        return len(self.Lambda)

    def eigenvalue_decomposition(self):
        return self

    def krr(self, b, lam):
        VTb = self.V.T @ b
        diag = self.Lambda + lam
        if b.ndim == 1:
            return b / lam + self.V @ (np.divide(VTb, diag) - VTb / lam)
        return b / lam + self.V @ (np.divide(VTb, diag[:, np.newaxis]) - VTb / lam)

    def krr_vec(self, b, lam):
        """lra.py:82-85: (A_hat + lam I)^{-1} b for a block of right-hand sides b of shape (N, m)."""
        VTb = self.V.T @ b
        diag = np.asarray(self.Lambda).reshape(VTb.shape[0]) + lam
        return b / lam + self.V @ (np.divide(VTb, diag[:, np.newaxis]) - VTb / lam)


class PSDLowRank(AbstractPSDLowRank):
    """A ~= G^T G with G of shape (k, N): lra.py:87-116."""

    def __init__(self, G, n=None, **kwargs):
        super().__init__(**kwargs)
        self._G = G
        self._G_host = None
        self._trace_hint = None
        ncols = G.shape[1] if n is None else n
        self.shape = (ncols, ncols)

    # numpy view of the factor (materialised from the device on first use)
    @property
    def G(self):
        if isinstance(self._G, torch.Tensor):
            if self._G_host is None:
                self._G_host = self._materialise(self._G)
            return self._G_host
        return self._G

    @G.setter
    def G(self, value):
        self._G = value
        self._G_host = None
        self._trace_hint = None

    @property
    def G_device(self):
        return self._G if isinstance(self._G, torch.Tensor) else None

    def _on_device(self):
        return isinstance(self._G, torch.Tensor) and self._G.is_cuda

    def trace(self):
        if self._trace_hint is not None:
            # the driver's own account of the factor's trace: trace(A) - sum(residual diagonal), both exact sequential sums
            # of the device sampler.  ||G||_F^2 (lra.py:94-95) is the same number up to the rounding of the diagonal updates
            # (~1e-15 relative), without a pass over the k x N factor (16 GB at the headline size).
            return self._trace_hint
        if self._on_device():
            if self._sharded():                                # local columns only (the shard is padded), then sum
                import torch.distributed as tdist
                t = torch.linalg.vector_norm(self._G[:, :self.shard_info.n_loc]) ** 2      # strided reduction, no copy
                tdist.all_reduce(t, group=self.shard_info.group)
                return float(t)
            return float(torch.linalg.vector_norm(self._G) ** 2)   # lra.py:94-95, ||G||_F^2 (no contiguous copy of the view)
        return np.linalg.norm(self.G, "fro") ** 2

    def __matmul__(self, other):
        if self._on_device() and not self._sharded():
            o = torch.as_tensor(np.asarray(other, dtype=np.float64), device=self._G.device)
            return (self._G.T @ (self._G @ o)).cpu().numpy()
        return self.G.T @ (self.G @ other)

    def rank(self):
        return self._G.shape[0]

    def matrix(self):
        if self._on_device() and not self._sharded():
            return (self._G.T @ self._G).cpu().numpy()
        return self.G.T @ self.G

    def eigenvalue_decomposition(self):
        return CompactEigenvalueDecomposition.from_G(self.G, idx=self.idx, rows=self.rows)

    def scale(self, scaling):
        if self._on_device() and not self._sharded():
            sc = torch.as_tensor(np.asarray(scaling, dtype=np.float64), device=self._G.device)
            return PSDLowRank(self._G * sc[None, :], idx=self.idx, rows=self._rows)
        return PSDLowRank(self.G * scaling[np.newaxis, :], idx=self.idx, rows=self.rows)

    def get_left_factor(self):
        return self.G.T

    def get_right_factor(self):
        return self.G


class NystromExtension(AbstractPSDLowRank):
    """A ~= rows^T core^{-1} rows, held as the k x k core and the k x N rows (lra.py:118-171).  Host numpy, off the
    hot path: the other samplers of the reference (uniform, leverage scores, DPP) return this type through
    utils.lra_from_sample.  Everything derived from the core is lazy: the Cholesky factor L of the symmetrised core shifted
    by trace(core) * eps (lra.py:131-133), the factor G = L^{-1} rows, and the interpolation matrix."""

    def __init__(self, core, factor=None, interpolation=None, **kwargs):
        if "rows" not in kwargs:
            raise RuntimeError("Need to specify rows for Nystrom extension")
        super().__init__(**kwargs)
        core = np.asarray(core)
        self.C = 0.5 * (core + core.T)
        n = self.rows.shape[1]
        self.shape = (n, n)
        self.L, self.G, self.W = None, factor, interpolation

    # -- lazily built pieces ------------------------------------------------------------------------------------------
    def get_core_factor(self):
        if self.L is None:
            shifted = self.C.copy()
            shifted[np.diag_indices_from(shifted)] += np.trace(self.C) * np.finfo(float).eps
            self.L = np.linalg.cholesky(shifted)
        return self.L

    def get_factor(self):
        if self.G is None:
            from scipy.linalg import solve_triangular
            self.G = solve_triangular(self.get_core_factor(), self.rows, lower=True)
        return self.G

    def get_interpolation(self):
        if self.W is None:
            from scipy.linalg import solve_triangular
            # the reference solves L^T against the LEFT (N x k) factor here (lra.py:140-143), which only has matching shapes
            # for N == k; the operand is kept as the reference has it
            self.W = solve_triangular(self.get_core_factor().T, self.get_left_factor(), lower=False)
        return self.W

    def get_right_factor(self):
        return self.get_factor()

    def get_left_factor(self):
        return self.get_right_factor().T

    # -- the low-rank operator ----------------------------------------------------------------------------------------------
    def rank(self):
        return self.C.shape[0]

    def trace(self):
        return np.linalg.norm(self.get_factor()) ** 2

    def __matmul__(self, other):
        G = self.get_factor()
        return G.T @ (G @ other)

    def matrix(self):
        G = self.get_factor()
        return G.T @ G

    def eigenvalue_decomposition(self):
        return CompactEigenvalueDecomposition.from_G(self.get_factor(), idx=self.idx, rows=self.rows)

    def scale(self, scaling):
        scaling = np.asarray(scaling)
        scaled_factor = None if self.G is None else self.G * np.sqrt(scaling)[np.newaxis, :]
        return NystromExtension(self.C, idx=self.idx, rows=self.rows * scaling[np.newaxis, :], factor=scaled_factor,
                                interpolation=self.W)
