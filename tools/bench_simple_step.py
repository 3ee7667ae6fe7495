#!/usr/bin/env python3
"""HBM roofline of the fused simple-RPCholesky step (GEMV over G[:i] + kernel row + diag update):
python tools/bench_simple_step.py [N] [d] [i ...]"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import randomly_pivoted_cholesky_b200 as rpc
from randomly_pivoted_cholesky_b200 import _lib
from randomly_pivoted_cholesky_b200.matrix import _ptr, _stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 10
steps = [int(v) for v in sys.argv[3:]] or [0, 100, 500, 1000]
X = np.random.default_rng(0).standard_normal((n, d))
A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=float(np.sqrt(d)))
dev = A.device; ctx = A._ctx; lib = ctx.lib
k = max(steps) + 1
G = torch.randn((k, A.n_pad), dtype=torch.float64, device=dev) * 0.01
rows = torch.empty((k, A.n_pad), dtype=torch.float64, device=dev)
diag = torch.ones((A.n_pad,), dtype=torch.float64, device=dev)
idx = torch.full((k,), 12345, dtype=torch.int64, device=dev)
spec, Xd, xn, dp, ldx, Ad, lda = A._simple_step_args()
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
for i in steps:
    def run():
        _lib.check(lib.rpc_simple_step(ctx.handle, _stream(), C.byref(spec), _ptr(Xd), _ptr(xn), dp, ldx, None, 0, n, _ptr(idx), i,
                                       _ptr(G), A.n_pad, _ptr(rows), A.n_pad, _ptr(diag), A.n_pad))
    run(); torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        diag.fill_(1.0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    nbytes = 8.0 * n * (i + 2 + 3 + d)      # read G[:i], write G[i] + rows[i], diag r/w + copy, X
    print(json.dumps({"i": i, "n": n, "d": d, "ms": best, "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / best / 1e6,
                      "frac_of_measured_hbm": nbytes / best / 1e6 / peak}), flush=True)
