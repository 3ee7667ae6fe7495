"""Accelerated randomly pivoted QR -- SURVEY.md section 8(f) rank 4, same entry points as the reference's rpqr.py (:7-67).

    Q, F, idx = rpqr(B, k)            # B ~ Q @ F.T, Q (m x k) orthonormal, F = B.T @ Q, idx = the k pivot columns

It reuses the hot path's exact pivot sampler (`rng.choice(range(n), size=b, p=u/sum(u))`, rpqr.py:27) and the device
rejection-sampling Cholesky (`rejection_cholesky(C.T @ C)`, rpqr.py:33); the dense m x n products and the thin QR of the
accepted block (rpqr.py:49-50) are plain library calls (cuBLAS / cuSOLVER through torch).  Random streams are consumed
like the reference: b proposal uniforms from `np.random.default_rng()` and b rejection uniforms from the legacy global
stream per round, so a seeded replay reproduces the reference's pivots.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from ._lib import COL_ALIGN, check
from .matrix import _cuda_device, _ptr, _round_up, _stream


def accelerated_rpqr(B, k, b="auto", accelerated=True, verbose=False, device=None, return_device=False):
    dev = _cuda_device(device if device is not None else (B.device if isinstance(B, torch.Tensor) and B.is_cuda else None))
    ctx = _lib.Context.get(dev.index)
    lib = ctx.lib
    if isinstance(B, torch.Tensor):
        Bd = B.detach().to(dev, torch.float64).contiguous()
    else:
        Bd = torch.from_numpy(np.ascontiguousarray(B, dtype=np.float64)).to(dev)
    m, n = Bd.shape
    k = int(k)
    f64 = dict(dtype=torch.float64, device=dev)
    Q = torch.zeros((m, k), **f64)
    F = torch.zeros((n, k), **f64)
    rng = np.random.default_rng()
    n_pad = _round_up(n, COL_ALIGN)
    u = torch.zeros((n_pad,), **f64)
    u[:n] = torch.linalg.vector_norm(Bd, dim=0) ** 2                      # rpqr.py:16
    arr_idx = np.zeros(k, dtype=int)
    auto_b = False
    if b == "auto":                                                        # rpqr.py:20-24
        b = int(np.ceil(k / 10))
        auto_b = True
    b = int(b)
    b_cap = max(b, int(np.ceil(k / 3)), 10) if auto_b else b
    u_dev = torch.empty((2 * b_cap,), **f64)
    idx_dev = torch.empty((b_cap,), dtype=torch.int64, device=dev)
    stats = torch.zeros((4,), **f64)
    L = torch.empty((b_cap, b_cap), **f64)
    acc = torch.zeros((b_cap,), dtype=torch.int32, device=dev)
    info = torch.zeros((4,), dtype=torch.int32, device=dev)
    counter = 0
    while counter < k:
        up = rng.random(b)                                                 # rpqr.py:27
        ur = np.random.random_sample(b)                                    # the b np.random.rand() calls of rejection_cholesky
        u_dev[:2 * b].copy_(torch.from_numpy(np.concatenate([up, ur])))
        check(lib.rpc_sample_pivots(ctx.handle, _stream(), _ptr(u), n_pad, _ptr(u_dev), b, _ptr(idx_dev), _ptr(stats)),
              "rpc_sample_pivots")
        start = time.perf_counter()
        idx_b = idx_dev[:b]
        Cm = Bd.index_select(1, idx_b) - Q[:, :counter] @ F.index_select(0, idx_b)[:, :counter].T      # rpqr.py:32
        H = (Cm.T @ Cm).contiguous()                                       # rpqr.py:33
        check(lib.rpc_rejection_cholesky(ctx.handle, _stream(), _ptr(H), b, b, C.c_void_p(u_dev.data_ptr() + 8 * b),
                                         k - counter, 0, 0.0, _ptr(L), b_cap, _ptr(acc), _ptr(info)),
              "rpc_rejection_cholesky")
        sh = stats.cpu().numpy()
        if sh[2] != 0.0:
            raise RuntimeError("internal error: exact-scan binade bound violated")
        if sh[1] != 0.0:
            raise ValueError("probabilities contain NaN" if not (sh[0] == sh[0]) or sh[0] == 0.0
                             else "probabilities are not non-negative")
        ih = info.cpu().numpy()
        if ih[2] != 0:
            raise RuntimeError("rejection_cholesky requires a strictly positive trace")
        num_sel = int(ih[0])                                               # already truncated to k - counter (rpqr.py:41-43)
        rejection_time = time.perf_counter() - start
        start = time.perf_counter()
        accepted = acc[:num_sel].to(torch.int64)
        new_counter = counter + num_sel
        arr_idx[counter:new_counter] = idx_b.index_select(0, accepted).cpu().numpy()
        Cm = Cm.index_select(1, accepted)
        Qc = Q[:, :counter]
        Qn, _ = torch.linalg.qr(Cm - Qc @ (Qc.T @ Cm), mode="reduced")     # rpqr.py:49
        Q[:, counter:new_counter] = Qn
        Fn = Bd.T @ Qn                                                     # rpqr.py:50
        F[:, counter:new_counter] = Fn
        if auto_b:                                                         # rpqr.py:52-56, on device-synchronised wall time
            torch.cuda.synchronize(dev)
            process_time = time.perf_counter() - start
            target = int(np.ceil(b * process_time / (4 * max(rejection_time, 1e-9))))
            b = max([min([target, int(np.ceil(1.5 * b)), int(np.ceil(k / 3))]), int(np.ceil(b / 3)), 10])
            b = min(b, b_cap)
        u[:n] -= torch.linalg.vector_norm(Fn, dim=1) ** 2                  # rpqr.py:59
        u.clamp_(min=0)                                                    # rpqr.py:60
        counter = new_counter
        if verbose:
            print("Accepted {} / {}".format(num_sel, b))
    if return_device:
        return Q, F, arr_idx
    return Q.cpu().numpy(), F.cpu().numpy(), arr_idx


def rpqr(B, k, **kwargs):
    return accelerated_rpqr(B, k, **kwargs)
