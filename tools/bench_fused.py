#!/usr/bin/env python3
"""Micro-benchmark: rpc_fused_update vs rpc_kernel_rows + rpc_schur_update on a Gaussian KernelMatrix.
usage: python tools/bench_fused.py [N] [d] [c,s ...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import randomly_pivoted_cholesky_b200 as rpc
from randomly_pivoted_cholesky_b200 import _lib
from randomly_pivoted_cholesky_b200.matrix import _ptr, _stream, _round_up
rpc.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 50
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or [(0, 200), (1000, 167), (1000, 192), (1000, 256), (1800, 200)]
dev = torch.device("cuda")
X = np.random.default_rng(0).standard_normal((n, d))
A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=float(np.sqrt(d)))
ctx = A._ctx; lib = ctx.lib
n_pad = A.n_pad
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for c, s in shapes:
    G = torch.randn((c + s, n_pad), dtype=torch.float64, device=dev) * 0.01
    rows = torch.empty((s, n_pad), dtype=torch.float64, device=dev)
    diag = torch.full((n_pad,), 1e9, dtype=torch.float64, device=dev)
    k_pad = _round_up(c + s, 16); ldat = int(lib.rpc_panel_ld(s))
    At = torch.randn((k_pad, ldat), dtype=torch.float64, device=dev) * 0.01
    At[:, s:] = 0; At[c + s:, :] = 0
    idx = torch.arange(0, s, dtype=torch.int64, device=dev) * 17
    piv = A._dev_new_payload(s); A._dev_gather(idx, s, piv)
    Gp = C.c_void_p(G.data_ptr()); rp = C.c_void_p(rows.data_ptr())
    def fused():
        _lib.check(lib.rpc_fused_update(ctx.handle, _stream(), C.byref(A._spec), _ptr(A._Xd), _ptr(A._xnorm), n_pad, A.d_pad, A.ldx,
                                        _ptr(piv.X), _ptr(piv.norm), A.ldx, Gp, n_pad, c, s, _ptr(At), ldat, k_pad, rp, n_pad, _ptr(diag)))
    def ev():
        A._dev_rows_into(idx, s, piv, rp, n_pad)
    def sch():
        _lib.check(lib.rpc_schur_update(ctx.handle, _stream(), Gp, n_pad, c, s, _ptr(At), ldat, k_pad, rp, n_pad, _ptr(diag), n_pad))
    tf, te, ts = timeit(fused), timeit(ev), timeit(sch)
    fl = 2.0 * s * (c + s) * n
    print(json.dumps({"c": c, "s": s, "fused_ms": tf, "eval_ms": te, "schur_ms": ts, "schur_tflops": fl / ts / 1e9,
                      "fused_tflops_incl_eval": (fl + 2.0 * s * d * n) / tf / 1e9}), flush=True)
    del G, rows, At
