"""Tensor-level wrappers of single C-ABI entry points (include/rpc_b200.h), used by the tests, bench.py
and anyone who wants one kernel of the path without the drivers.  Inputs are numpy arrays or CUDA
tensors; outputs are CUDA tensors unless stated.  Every function runs the CUDA library — no fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import COL_ALIGN, KT, check
from .matrix import _cuda_device, _ptr, _round_up, _stream


def _dev_f64(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(dev, torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)


def _ctx(dev):
    return _lib.Context.get(dev.index)


def sample_pivots(diag, u, device=None):
    """numpy `Generator.choice(len(diag), size=len(u), p=diag/sum(diag))` given its uniforms, bit-exact.
    Returns (idx int64 tensor, python_sum_of_diag float).  Raises ValueError like numpy on bad p."""
    dev = _cuda_device(device)
    ctx = _ctx(dev)
    d = _dev_f64(diag, dev)
    n = d.numel()
    n_pad = _round_up(n, COL_ALIGN)
    dp = torch.zeros((n_pad,), dtype=torch.float64, device=dev)
    dp[:n] = d
    uu = _dev_f64(np.atleast_1d(u), dev)
    m = uu.numel()
    idx = torch.empty((max(m, 1),), dtype=torch.int64, device=dev)
    stats = torch.zeros((4,), dtype=torch.float64, device=dev)
    check(ctx.lib.rpc_sample_pivots(ctx.handle, _stream(), _ptr(dp), n_pad, _ptr(uu) if m else None, m,
                                    _ptr(idx) if m else None, _ptr(stats)), "rpc_sample_pivots")
    sh = stats.cpu().numpy()
    if sh[2] != 0:
        raise RuntimeError("internal error: exact-scan binade bound violated")
    if sh[1] != 0:
        raise ValueError("probabilities are not a distribution (sum <= 0, NaN or negative entries)")
    return idx[:m], float(sh[0])


def rejection_cholesky(H, u, max_accept=-1, device=None):
    """rejection_cholesky(H) of rpcholesky.py:140-157 with `u` replacing the np.random.rand() calls.
    Returns (L (s x s) tensor, accepted positions int64 numpy)."""
    dev = _cuda_device(device)
    ctx = _ctx(dev)
    Hd = _dev_f64(H, dev).clone()
    if Hd.ndim != 2 or Hd.shape[0] != Hd.shape[1]:
        raise RuntimeError("rejection_cholesky requires a square matrix")
    m = Hd.shape[0]
    ud = _dev_f64(u, dev)
    L = torch.zeros((m, m), dtype=torch.float64, device=dev)
    acc = torch.zeros((m,), dtype=torch.int32, device=dev)
    info = torch.zeros((4,), dtype=torch.int32, device=dev)
    check(ctx.lib.rpc_rejection_cholesky(ctx.handle, _stream(), _ptr(Hd), m, m, _ptr(ud), int(max_accept), 0, 0.0, _ptr(L), m,
                                         _ptr(acc), _ptr(info)), "rpc_rejection_cholesky")
    ih = info.cpu().numpy()
    if ih[2] != 0:
        raise RuntimeError("rejection_cholesky requires a strictly positive trace")
    s = int(ih[0])
    return L[:s, :s], acc[:s].cpu().numpy().astype(np.int64)


def shifted_cholesky(Cm, shift, device=None):
    """np.linalg.cholesky(C + shift*I) on the device; returns (L, info) with info = failing pivot + 1 or 0."""
    dev = _cuda_device(device)
    ctx = _ctx(dev)
    Hd = _dev_f64(Cm, dev).clone()
    m = Hd.shape[0]
    L = torch.zeros((m, m), dtype=torch.float64, device=dev)
    acc = torch.zeros((m,), dtype=torch.int32, device=dev)
    info = torch.zeros((4,), dtype=torch.int32, device=dev)
    check(ctx.lib.rpc_rejection_cholesky(ctx.handle, _stream(), _ptr(Hd), m, m, None, m, 1, float(shift), _ptr(L), m, _ptr(acc),
                                         _ptr(info)), "rpc_rejection_cholesky")
    ih = info.cpu().numpy()
    return L[:int(ih[0]), :int(ih[0])], int(ih[1])


def schur_update(G, c, rows_new, P, L, diag, device=None, panel_lower=True):
    """One block update on explicit operands (test / bench helper).

    G: (k, n_pad) tensor updated in place at rows [c, c+s); rows_new: (s, n_pad); P: (c, s) = G[:c, S];
    L: (s, s) lower triangular; diag: (n_pad,) updated in place.  Computes
    G[c:c+s] = L^{-1}(rows_new - P^T G[:c]); diag = max(diag - colsum(G[c:c+s]^2), 0).
    """
    dev = G.device
    ctx = _ctx(dev)
    s = rows_new.shape[0]
    n_pad = G.shape[1]
    k_pad = _round_up(c + s, KT)
    ldat = int(ctx.lib.rpc_panel_ld(s))
    At = torch.empty((k_pad, ldat), dtype=torch.float64, device=dev)
    Pd = P.contiguous() if c > 0 else None
    Ld = L.contiguous()
    check(ctx.lib.rpc_build_panel(ctx.handle, _stream(), _ptr(Pd), s, c, None, s, _ptr(Ld), s, 0, _ptr(At), ldat, k_pad),
          "rpc_build_panel")
    # the panel of rpc_build_panel mode 0 ends in L^{-T}: the kernel may skip the zero half of that block
    check(ctx.lib.rpc_schur_update_ex(ctx.handle, _stream(), _ptr(G), G.stride(0), c, s, _ptr(At), ldat, k_pad, _ptr(rows_new),
                                      rows_new.stride(0), None, _ptr(diag), n_pad, None, 0, 1 if panel_lower else 0),
          "rpc_schur_update_ex")
    return At


def pivoted_cholesky(H, rtol=None, atol=None, device=None):
    """`_greedy_cholesky(H, rtol, atol)` of rpcholesky.py:52-63 (dpstrf + trailing-sum rank rule).
    Returns (L (rank x rank, pivoted order) tensor, pivot positions int64 numpy (0-based), rank)."""
    dev = _cuda_device(device)
    ctx = _ctx(dev)
    Hd = _dev_f64(H, dev)
    if Hd.ndim != 2 or Hd.shape[0] != Hd.shape[1]:
        raise RuntimeError("pivoted_cholesky requires a square matrix")
    m = Hd.shape[0]
    L = torch.zeros((m, m), dtype=torch.float64, device=dev)
    acc = torch.zeros((m,), dtype=torch.int32, device=dev)
    info = torch.zeros((4,), dtype=torch.int32, device=dev)
    use_tol = rtol is not None and atol is not None
    check(ctx.lib.rpc_pivoted_cholesky(ctx.handle, _stream(), _ptr(Hd), m, m, 1 if use_tol else 0,
                                       float(rtol) if use_tol else 0.0, float(atol) if use_tol else 0.0, _ptr(L), m,
                                       _ptr(acc), _ptr(info)), "rpc_pivoted_cholesky")
    rank = int(info[0])
    return L[:rank, :rank], acc[:rank].cpu().numpy().astype(np.int64), rank


def select_largest(diag, b, device=None):
    """The set np.argpartition(diag, -b)[-b:] in ascending index order (lowest indices among ties at the threshold)."""
    dev = _cuda_device(device)
    ctx = _ctx(dev)
    d = _dev_f64(diag, dev)
    idx = torch.empty((b,), dtype=torch.int64, device=dev)
    check(ctx.lib.rpc_select_largest(ctx.handle, _stream(), _ptr(d), d.numel(), int(b), _ptr(idx)), "rpc_select_largest")
    return idx


def fp64_peak_tflops(device=None, iters=4000, reps=5):
    """FP64 tensor-pipe (DMMA) peak of this device in TFLOP/s, measured now: best of `reps` launches of the library's pure
    DMMA stream (rpc_probe_fp64_peak), timed with CUDA events.  ~10 ms per launch at iters = 4000."""
    dev = _cuda_device(device)
    ctx = _ctx(dev)
    fl = C.c_double(0.0)
    best = float("inf")
    with torch.cuda.device(dev):
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(ctx.lib.rpc_probe_fp64_peak(ctx.handle, _stream(), int(iters), C.byref(fl)), "rpc_probe_fp64_peak")
            e1.record()
            torch.cuda.synchronize(dev)
            if r > 0:
                best = min(best, e0.elapsed_time(e1))
    return fl.value / best / 1e9


# ---- N-proportional linear algebra on the factor rows (KRR_Nystrom / Nystrom-PCG, the "next" rows of the scope table) --------
def syrk_rows(R, n_cols=None):
    """M = R[:, :n_cols] @ R[:, :n_cols].T for a (k, ld) CUDA tensor with unit column stride: rpc_syrk_rows (DMMA + tensor-map TMA)."""
    if R.stride(1) != 1:
        raise ValueError("syrk_rows needs a row-major tensor")
    ctx = _ctx(R.device)
    k = R.shape[0]
    n_cols = R.shape[1] if n_cols is None else int(n_cols)
    if R.stride(0) % 2 or R.data_ptr() % 16:                 # tensor-map TMA: 16-byte aligned rows
        Rp = torch.zeros((k, n_cols + (n_cols & 1)), dtype=torch.float64, device=R.device)
        Rp[:, :n_cols] = R[:, :n_cols]
        R = Rp
    M = torch.empty((k, k), dtype=torch.float64, device=R.device)
    with torch.cuda.device(R.device):
        check(ctx.lib.rpc_syrk_rows(ctx.handle, _stream(), _ptr(R), R.stride(0), k, n_cols, _ptr(M), k), "rpc_syrk_rows")
    return M


def rows_dot(R, y, n_cols=None):
    """R[:, :n_cols] @ y for y of shape (n_cols,) or (n_cols, m): rpc_rows_dot, one pass over R per column of y."""
    ctx = _ctx(R.device)
    k = R.shape[0]
    n_cols = R.shape[1] if n_cols is None else int(n_cols)
    y2 = y.reshape(n_cols, -1)
    out = torch.empty((k, y2.shape[1]), dtype=torch.float64, device=R.device)
    with torch.cuda.device(R.device):
        for c in range(y2.shape[1]):
            yc = y2[:, c].contiguous()
            oc = torch.empty((k,), dtype=torch.float64, device=R.device)
            check(ctx.lib.rpc_rows_dot(ctx.handle, _stream(), _ptr(R), R.stride(0), k, n_cols, _ptr(yc), _ptr(oc)), "rpc_rows_dot")
            out[:, c] = oc
    return out.reshape((k,) + tuple(y.shape[1:]))


def rows_tdot(R, w, n_cols=None):
    """R[:, :n_cols].T @ w for w of shape (k,): rpc_rows_tdot (n_cols is rounded up to even inside the padded row)."""
    ctx = _ctx(R.device)
    k = R.shape[0]
    n_cols = R.shape[1] if n_cols is None else int(n_cols)
    n_even = n_cols + (n_cols & 1)
    if n_even > R.shape[1] and n_even > R.stride(0):
        raise ValueError("rows_tdot needs one column of padding for an odd column count")
    out = torch.empty((n_even,), dtype=torch.float64, device=R.device)
    wc = w.contiguous()
    with torch.cuda.device(R.device):
        check(ctx.lib.rpc_rows_tdot(ctx.handle, _stream(), _ptr(R), R.stride(0), k, n_even, _ptr(wc), _ptr(out)), "rpc_rows_tdot")
    return out[:n_cols]


def chol_solve(L, b):
    """(L L^T)^{-1} b for a lower-triangular (k, k) CUDA tensor and b of shape (k,) or (k, m): rpc_chol_solve."""
    ctx = _ctx(L.device)
    k = L.shape[0]
    x = b.reshape(k, -1).contiguous().clone()
    Lc = L.contiguous()
    with torch.cuda.device(L.device):
        check(ctx.lib.rpc_chol_solve(ctx.handle, _stream(), _ptr(Lc), Lc.stride(0), k, _ptr(x), x.stride(0), x.shape[1]), "rpc_chol_solve")
    return x.reshape(b.shape)


def kernel_weighted_sum(A, piv_X, piv_norm, m, w, out=None, accumulate=False):
    """out[col] (+)= sum_{i < m} w[i] k(piv_X[i], A's point col) over A's n_pad points: rpc_kernel_rows_wsum (no kernel value is
    written to memory).  piv_X: (>= m, A.ldx) points in A's padded layout, piv_norm their squared norms."""
    ctx = A._ctx
    if out is None:
        out = torch.empty((A.n_pad,), dtype=torch.float64, device=A.device)
        accumulate = False
    wc = w.contiguous()
    with torch.cuda.device(A.device):
        check(ctx.lib.rpc_kernel_rows_wsum(ctx.handle, _stream(), C.byref(A._spec), _ptr(A._Xd), _ptr(A._xnorm), A.n_pad, A.d_pad, A.ldx,
                                           _ptr(piv_X), _ptr(piv_norm), int(m), piv_X.stride(0), _ptr(wc), _ptr(out),
                                           1 if accumulate else 0), "rpc_kernel_rows_wsum")
    return out
