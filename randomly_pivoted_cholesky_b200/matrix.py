"""Operator boundary of the hot path: AbstractPSDMatrix / PSDMatrix / KernelMatrix.

Mirrors the reference's matrix.py (class names, constructor arguments, `shape`, `diag`, `__getitem__`,
`trace`, `num_queries`, `reset`, `to_matrix`, index normalisation and query counting: matrix.py:10-77,
79-102, 140-192) with the data resident in B200 HBM and every entry evaluated by the CUDA kernels
behind include/rpc_b200.h.  No CPU evaluation path exists: without the CUDA library/device these
classes raise.

Device layout (DESIGN.md "data layout"): points X are stored row-major as (n_pad, d_pad) float64 with
n_pad = N rounded up to 128 and d_pad = d rounded up to 4; pad points/features are zero.  Squared row
norms (n_pad,) are precomputed once for the expanded Euclidean form.
"""
from __future__ import annotations

import ctypes as C
import numbers
import os

import numpy as np
import torch

from . import _lib
from . import dist as rdist
from ._lib import COL_ALIGN, FEAT_ALIGN, KernelSpec, check


def _round_up(x, m):
    return (x + m - 1) // m * m


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _normalise_indices(vec, n):
    """numpy fancy-index semantics for a list of point indices (the reference evaluates `self.data[vec, :]`, matrix.py:189):
    negative indices wrap, anything outside [-n, n) raises IndexError.  The device gather itself treats an index outside the
    owned range as "not mine" (it writes zeros), so the check has to happen here."""
    arr = np.asarray(vec, dtype=np.int64).ravel()
    bad = (arr < -n) | (arr >= n)
    if bad.any():
        raise IndexError("index {} is out of bounds for axis 0 with size {}".format(int(arr[bad][0]), n))
    return np.where(arr < 0, arr + n, arr)


def _cuda_device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("randomly_pivoted_cholesky_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


class AbstractPSDMatrix:
    """matrix.py:10-77 — query counter + index protocol; subclasses provide the device evaluation."""

    def __init__(self, **kwargs):
        self.queries = 0
        self.count_queries = kwargs["count_queries"] if ("count_queries" in kwargs) else True
        self.dpp_stuff = None

    # -- host-facing protocol (returns numpy like the reference) ---------------------------------
    def _diag_helper(self, *args):
        raise NotImplementedError

    def _getitem_helper(self, *args):
        raise NotImplementedError

    def __getitem__(self, *args):
        to_return = self._getitem_helper(*args)
        if isinstance(to_return, np.ndarray):
            self.queries += to_return.size
        else:
            self.queries += 1
        return to_return

    def diag(self, *args):
        to_return = self._diag_helper(*args)
        if isinstance(to_return, np.ndarray):
            self.queries += to_return.size
        else:
            self.queries += 1
        return to_return

    def __call__(self, *args):
        return self.__getitem__(*args)

    def trace(self):
        return sum(self.diag())                      # matrix.py:44-45: python sum, left to right

    def num_queries(self):
        return self.queries

    def reset(self):
        self.queries = 0
        self.dpp_stuff = None

    def to_matrix(self):
        return PSDMatrix(self[:, :])

    def _clean_index_input(self, args):              # matrix.py:57-77
        idx = args[0]
        if len(idx) == 1:
            idx = [idx, idx]
        else:
            idx = list(idx)
            assert len(idx) == 2
        for j in range(2):
            if isinstance(idx[j], np.ndarray):
                idx[j] = idx[j].ravel().tolist()
            elif isinstance(idx[j], slice):
                idx[j] = list(range(self.shape[0]))[idx[j]]
            elif isinstance(idx[j], list):
                pass
            elif isinstance(idx[j], numbers.Integral):
                idx[j] = [idx[j]]
            else:
                raise RuntimeError("Indexing not implemented with index of type {}".format(type(idx)))
        return idx

    # -- device-facing protocol used by the factorisation drivers ---------------------------------
    # n_pad, device, _ctx and the methods below are provided by the concrete classes:
    #   _dev_diag_into(diag)                     diag (n_pad,) <- A.diag(), pad entries 0
    #   _dev_gather(idx_dev, m, piv)             per-round pivot payload (points / nothing)
    #   _dev_block_into(idx_dev, m, piv, H, ldh) H <- A[idx, idx]
    #   _dev_rows_into(idx_dev, m, piv, out, ldo) out[m, n_pad] <- A[idx, :]
    def _count(self, nentries):
        self.queries += int(nentries)


class PSDMatrix(AbstractPSDMatrix):
    """Dense operator, matrix.py:79-102.  The matrix lives on the device; host indexing returns numpy."""

    def __init__(self, A, device=None, **kwargs):
        super().__init__(**kwargs)
        dev = _cuda_device(device if device is not None else (A.device if isinstance(A, torch.Tensor) and A.is_cuda else None))
        if isinstance(A, torch.Tensor):
            self.matrix = A.detach().cpu().numpy() if not A.is_cuda else None
            self._Ad = A.detach().to(dev, torch.float64).contiguous()
        else:
            self.matrix = np.asarray(A)
            self._Ad = torch.from_numpy(np.ascontiguousarray(self.matrix, dtype=np.float64)).to(dev)
        if self._Ad.ndim != 2 or self._Ad.shape[0] != self._Ad.shape[1]:
            raise RuntimeError("PSDMatrix requires a square matrix")
        self.shape = tuple(self._Ad.shape)
        self.device = dev
        self.n_pad = _round_up(self.shape[0], COL_ALIGN)
        self.shard_info = rdist.make_shard_info(self.shape[0], 1, 0)     # a dense N x N operator is not sharded
        self._ctx = _lib.Context.get(dev.index)

    def _host(self):
        if self.matrix is None:
            self.matrix = self._Ad.cpu().numpy()
        return self.matrix

    def _diag_helper(self, *args):
        M = self._host()
        X = list(args[0]) if len(args) > 0 else range(M.shape[0])
        return M[X, X]

    def _getitem_helper(self, *args):
        idx = self._clean_index_input(args)
        M = self._host()
        if len(idx[0]) == 1 and len(idx[1]) == 1:
            return M[idx[0][0], idx[1][0]]
        mtx = M[np.ix_(idx[0], idx[1])]
        if len(idx[0]) == 1:
            return mtx[0, :]
        if len(idx[1]) == 1:
            return mtx[:, 0]
        return mtx

    # device protocol
    def _dev_diag_into(self, diag):
        n = self.shape[0]
        check(self._ctx.lib.rpc_psd_diag(self._ctx.handle, _stream(), _ptr(self._Ad), n, self._Ad.stride(0), _ptr(diag),
                                         self.n_pad), "rpc_psd_diag")

    def _dev_new_payload(self, m_max):
        return None

    def _dev_gather(self, idx_dev, m, piv):
        return None

    def _dev_block_into(self, idx_dev, m, piv, H, ldh):
        check(self._ctx.lib.rpc_psd_block(self._ctx.handle, _stream(), _ptr(self._Ad), self._Ad.stride(0), _ptr(idx_dev), m,
                                          _ptr(H), ldh), "rpc_psd_block")

    def _dev_rows_into(self, idx_dev, m, piv, out_ptr, ldo):
        n = self.shape[0]
        check(self._ctx.lib.rpc_psd_rows(self._ctx.handle, _stream(), _ptr(self._Ad), n, self._Ad.stride(0), _ptr(idx_dev), m,
                                         out_ptr, ldo, self.n_pad), "rpc_psd_rows")

    def _simple_step_args(self):
        return None, None, None, 0, 0, self._Ad, self._Ad.stride(0)


class _PivotPayload:
    """Per-round device buffers holding the gathered pivot points (b x d_pad) and their norms."""

    def __init__(self, m_max, d_pad, device):
        self.X = torch.zeros((m_max, d_pad), dtype=torch.float64, device=device)
        self.norm = torch.zeros((m_max,), dtype=torch.float64, device=device)


class KernelMatrix(AbstractPSDMatrix):
    """KernelMatrix(X, kernel="gaussian", bandwidth=1.0, **kwargs) — matrix.py:140-192.

    kernel in {"gaussian", "laplace", "matern"} (+ nu in {0.5, 1.5, 2.5}, extra_stability).  A triple
    of Python callables (accepted by the reference, matrix.py:164-165) cannot run on the device and
    is rejected.  bandwidth may be a float, "median" or "approx_median" (matrix.py:168-176).
    """

    def __init__(self, X, kernel="gaussian", bandwidth=1.0, device=None, sharded=False, process_group=None, **kwargs):
        """`sharded=True` (under an initialised torch.distributed NCCL group, one process per GPU): every rank
        passes the SAME full X and keeps only its block of rows in HBM (dist.ShardInfo); the factorisation
        drivers then run row-sharded and return the local column shard of G / rows."""
        super().__init__(**kwargs)
        kwargs = {k: v for k, v in kwargs.items() if k != "count_queries"}
        if not hasattr(X, "ndim") or X.ndim != 2:
            raise RuntimeError("KernelMatrix requires an (N, d) array of points")
        n, d = int(X.shape[0]), int(X.shape[1])
        self.shape = (n, n)
        self.shard_info = rdist.current_shard_info(n, process_group) if sharded else rdist.make_shard_info(n, 1, 0)
        info = self.shard_info
        # only this rank's block of rows crosses to the device
        if isinstance(X, torch.Tensor):
            dev = _cuda_device(device if device is not None else (X.device if X.is_cuda else None))
            Xd = X.detach()[info.lo:info.hi].to(dev, torch.float64, non_blocking=True)
            self.data = X.detach().cpu().numpy() if not X.is_cuda else None
        else:
            dev = _cuda_device(device)
            self.data = np.asarray(X)
            Xd = torch.from_numpy(np.ascontiguousarray(self.data[info.lo:info.hi], dtype=np.float64)).to(dev)
        self.device = dev
        self.n_pad = info.shard                       # LOCAL padded point count (== round_up(N,128) when not sharded)
        self.d = d
        self.d_pad = _round_up(d, FEAT_ALIGN)
        # point stride == 4 (mod 8) doubles: the DMMA fragment reads of a resident X tile are then bank-conflict free
        # (the non-contraction kernels -- Laplace, extra_stability -- want an ODD stride instead, set below)
        self.ldx = self.d_pad if self.d_pad % 8 == 4 else self.d_pad + 4
        self._ctx = _lib.Context.get(dev.index)
        self.kernel_kwargs = dict(kwargs)

        if not isinstance(kernel, str):
            raise RuntimeError("user-supplied kernel callables cannot be evaluated on the device; "
                               "use kernel in {'gaussian','laplace','matern'}")
        # matrix.py:153-163 (the reference's `kernel in "laplace"` substring test is not reproduced)
        nu2 = 0
        if kernel == "gaussian":
            kind = _lib.KERN_GAUSSIAN
        elif kernel == "laplace":
            kind = _lib.KERN_LAPLACE
        elif kernel == "matern":
            kind = _lib.KERN_MATERN
            nu = kwargs["nu"] if "nu" in kwargs else 0.5     # MaternKernel_mtx requires nu; 0.5 is kernels.py's default
            nu2 = int(round(2 * float(nu)))
            if nu2 not in (1, 3, 5) or abs(2 * float(nu) - nu2) > 0:
                raise NotImplementedError("Matern kernel: only nu in {0.5, 1.5, 2.5} run on the device "
                                          "(the reference's generic-nu Bessel path has a NaN diagonal)")
        else:
            raise RuntimeError("Kernel name {} not recognized".format(kernel))
        self.kernel_name = kernel
        stab = bool(kwargs.get("extra_stability", False))
        if kind == _lib.KERN_LAPLACE:
            stab = False                                     # LaplaceKernel_mtx ignores it (kernels.py:16)
        if kind == _lib.KERN_LAPLACE or stab:
            self.ldx = self.d_pad + 1                        # 16 points x one feature -> 16 distinct bank pairs

        Xp = torch.zeros((self.n_pad, self.ldx), dtype=torch.float64, device=dev)
        Xp[:info.n_loc, :d] = Xd
        self._Xd = Xp
        self._xnorm = torch.zeros((self.n_pad,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(self._ctx.lib.rpc_row_norms(self._ctx.handle, _stream(), _ptr(Xp), self.n_pad, self.d_pad, self.ldx,
                                              _ptr(self._xnorm)), "rpc_row_norms")

        if os.environ.get("RPC_B200_VALIDATE", "0") == "1" and self.padding_violations() != 0:
            raise RuntimeError("internal error: the padded point buffer violates the padding contract of include/rpc_b200.h")

        if isinstance(bandwidth, str):
            if info.world > 1:
                raise NotImplementedError("'median' bandwidths need all pairwise distances: pass a float when sharded")
            bandwidth = self._median_bandwidth(bandwidth, kind)
        self.bandwidth = bandwidth
        self._spec = KernelSpec(kind, nu2, 1 if stab else 0, 0, float(bandwidth))

    def padding_violations(self, diag=None):
        """Number of entries that break the padding contract of include/rpc_b200.h (pad features / pad points / pad diagonal
        entries must be zero: the N-proportional kernels run over n_pad columns without edge checks).  Synchronises."""
        bad = C.c_int32(-1)
        with torch.cuda.device(self.device):
            check(self._ctx.lib.rpc_check_padding(self._ctx.handle, _stream(), _ptr(self._Xd), self.shard_info.n_loc, self.n_pad, self.d,
                                                  self.ldx, _ptr(diag) if diag is not None else None, C.byref(bad)), "rpc_check_padding")
        return int(bad.value)

    # matrix.py:168-176: median of ALL pairwise distances (zeros on the diagonal included)
    def _median_bandwidth(self, mode, kind):
        n = self.shape[0]
        X = self._Xd[:n, :self.d]
        if mode == "approx_median":
            sel = np.random.choice(n, size=min(n, 1000), replace=False)   # legacy global RNG, as matrix.py:172
            X = X[torch.as_tensor(sel, device=self.device)]
        elif mode != "median":
            raise RuntimeError("Bandwidth {} not recognized".format(mode))
        if X.shape[0] > 30000:
            raise RuntimeError("'median' bandwidth needs all N^2 distances; use 'approx_median' for N > 30000")
        dist = torch.cdist(X, X, p=1.0 if kind == _lib.KERN_LAPLACE else 2.0).flatten()
        vals, _ = torch.sort(dist)
        m = vals.numel()
        med = vals[m // 2] if m % 2 else 0.5 * (vals[m // 2 - 1] + vals[m // 2])
        return float(med)

    def trace(self):
        # matrix.py:44-45 is python's sum(self.diag()); the built-in kernels have an all-ones diagonal (distance
        # exactly 0), whose left-to-right float64 sum is exactly N -- same value, same query count, no O(N) host loop
        self.queries += self.shape[0]
        return np.float64(self.shape[0])

    # -- host-facing evaluation (numpy out) --------------------------------------------------------
    def _diag_helper(self, *args):
        if len(args) > 0:
            return np.ones(len(list(args[0])))
        return np.ones(self.shape[0])                # *_vec on identical points: exactly 1.0

    def _cross(self, vec_i, vec_j, other=None):
        """kernel_mtx(X[vec_i], Y[vec_j]) as a (len(vec_i), len(vec_j)) numpy array; Y = `other`'s points (default:
        this operator's own).  The rows X[vec_i] play the pivots of the row evaluator, Y the streamed points."""
        if self.shard_info.world > 1:
            raise NotImplementedError("host-side indexing A[i,j] of a sharded KernelMatrix is not supported; "
                                      "use the factorisation drivers (device protocol) or an unsharded operator")
        tgt = self if other is None else other
        n = tgt.shape[0]
        mi, mj = len(vec_i), len(vec_j)
        vec_i = _normalise_indices(vec_i, self.shape[0])
        ti = torch.as_tensor(vec_i, device=self.device)
        piv = _PivotPayload(mi, self.ldx, self.device)
        with torch.cuda.device(self.device):
            self._dev_gather(ti, mi, piv)
        vec_j = _normalise_indices(vec_j, n)
        full = (mj == n) and bool(np.array_equal(vec_j, np.arange(n)))
        if full:
            Y, ynorm, npad = tgt._Xd, tgt._xnorm, tgt.n_pad
        else:
            tj = torch.as_tensor(vec_j, device=self.device)
            npad = _round_up(mj, COL_ALIGN)
            Y = torch.zeros((npad, self.ldx), dtype=torch.float64, device=self.device)
            Y[:mj] = tgt._Xd[tj]
            ynorm = torch.zeros((npad,), dtype=torch.float64, device=self.device)
            ynorm[:mj] = tgt._xnorm[tj]
        out = torch.empty((mi, npad), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._ctx.lib.rpc_kernel_rows(self._ctx.handle, _stream(), C.byref(self._spec), _ptr(Y), _ptr(ynorm), npad,
                                                self.d_pad, self.ldx, _ptr(piv.X), _ptr(piv.norm), mi, self.ldx,
                                                _ptr(out), npad), "rpc_kernel_rows")
        return out[:, :mj].cpu().numpy()

    def _getitem_helper(self, *args):                # matrix.py:129-138
        idx = self._clean_index_input(args)
        if len(idx[0]) == 0 or len(idx[1]) == 0:
            return np.zeros((len(idx[0]), len(idx[1])))
        mtx = self._cross(idx[0], idx[1])
        if len(idx[0]) == 1 and len(idx[1]) == 1:
            return mtx[0, 0]
        if len(idx[0]) == 1:
            return mtx[0, :]
        if len(idx[1]) == 1:
            return mtx[:, 0]
        return mtx

    def _companion(self, X):
        """A KernelMatrix over other points with this operator's kernel, bandwidth and device."""
        return KernelMatrix(X, kernel=self.kernel_name, bandwidth=self.bandwidth, device=self.device, **self.kernel_kwargs)

    def out_of_sample(self, Xtest, vec):
        """kernel_mtx(Xtest, X[vec]) -> (len(Xtest), len(vec)) array (matrix.py:191-192)."""
        vec = np.atleast_1d(np.asarray(vec, dtype=np.int64)).ravel().tolist()
        Kt = self._companion(Xtest)
        return np.ascontiguousarray(self._cross(vec, list(range(Kt.shape[0])), other=Kt).T)

    # -- device protocol ---------------------------------------------------------------------------
    def _dev_diag_into(self, diag):
        check(self._ctx.lib.rpc_kernel_diag(self._ctx.handle, _stream(), _ptr(diag), self.shard_info.n_loc, self.n_pad),
              "rpc_kernel_diag")

    def _dev_new_payload(self, m_max):
        return _PivotPayload(m_max, self.ldx, self.device)

    def _dev_gather(self, idx_dev, m, piv):
        """piv <- X[idx] and their squared norms (global indices; assembled with a sum all-reduce when sharded)."""
        info = self.shard_info
        check(self._ctx.lib.rpc_gather_pivots(self._ctx.handle, _stream(), _ptr(idx_dev), m, _ptr(self._Xd), _ptr(self._xnorm),
                                              self.d_pad, self.ldx, _ptr(piv.X), _ptr(piv.norm), self.ldx, None, 0, 0,
                                              None, 0, info.lo, info.hi), "rpc_gather_pivots")
        if info.world > 1:
            rdist.exchange_payload(info, piv.X)
            rdist.exchange_payload(info, piv.norm)

    def _dev_block_into(self, idx_dev, m, piv, H, ldh):
        check(self._ctx.lib.rpc_kernel_block(self._ctx.handle, _stream(), C.byref(self._spec), _ptr(piv.X), _ptr(piv.norm), m,
                                             self.d_pad, self.ldx, _ptr(H), ldh), "rpc_kernel_block")

    def _dev_rows_into(self, idx_dev, m, piv, out_ptr, ldo):
        check(self._ctx.lib.rpc_kernel_rows(self._ctx.handle, _stream(), C.byref(self._spec), _ptr(self._Xd), _ptr(self._xnorm),
                                            self.n_pad, self.d_pad, self.ldx, _ptr(piv.X), _ptr(piv.norm), m, self.ldx,
                                            out_ptr, ldo), "rpc_kernel_rows")

    def _simple_step_args(self):
        return self._spec, self._Xd, self._xnorm, self.d_pad, self.ldx, None, 0


class NonsymmetricKernelMatrix(object):
    """K[i, j] = k(X[i], Y[j]) for two point sets (matrix.py:194-239): the cross-kernel operator the reference uses for
    prediction.  Indexing accepts ints, lists, arrays and slices per axis and returns numpy, like the reference's."""

    def __init__(self, X, Y, kernel="gaussian", bandwidth=1.0, device=None, **kwargs):
        self.X = X
        self.Y = Y
        self.shape = (X.shape[0], Y.shape[0])
        self._Kx = KernelMatrix(X, kernel=kernel, bandwidth=bandwidth, device=device, **kwargs)
        self._Ky = self._Kx._companion(Y)

    def _getitem_helper(self, *args):
        idx = args[0]
        if len(idx) == 1:
            idx = [idx, idx]
        else:
            idx = list(idx)
        for j in range(2):
            if isinstance(idx[j], np.ndarray):
                idx[j] = idx[j].ravel().tolist()
            elif isinstance(idx[j], slice):
                idx[j] = list(range(self.shape[j]))[idx[j]]
            elif isinstance(idx[j], list):
                pass
            elif isinstance(idx[j], numbers.Integral):
                idx[j] = [idx[j]]
            else:
                raise RuntimeError("Indexing not implemented with index of type {}".format(type(idx)))
        mtx = self._Kx._cross(idx[0], idx[1], other=self._Ky)
        if len(idx[0]) == 1 and len(idx[1]) == 1:
            return mtx[0, 0]
        if len(idx[0]) == 1:
            return mtx[0, :]
        if len(idx[1]) == 1:
            return mtx[:, 0]
        return mtx

    def __getitem__(self, *args):
        return self._getitem_helper(*args)
