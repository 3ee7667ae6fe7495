#!/usr/bin/env python3
"""Times the serial (per-round) phases in isolation on realistic operands:
sampler (N points, m draws), residual Gram, rejection Cholesky (b), panel solve (c, s)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import randomly_pivoted_cholesky_b200 as rpc
from randomly_pivoted_cholesky_b200 import _lib
from randomly_pivoted_cholesky_b200.matrix import _ptr, _stream, _round_up
rpc.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_064
b = int(sys.argv[2]) if len(sys.argv) > 2 else 200
c = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
reps = int(os.environ.get("REPS", "20"))
dev = torch.device("cuda"); ctx = _lib.Context.get(0); lib = ctx.lib
def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
rng = np.random.default_rng(0)
out = {}
# sampler on a residual-like diagonal (values in (0,1], varied)
diag = torch.from_numpy(rng.random(n) ** 2).to(dev); u = torch.from_numpy(rng.random(b)).to(dev)
idx = torch.empty(b, dtype=torch.int64, device=dev); stats = torch.zeros(4, dtype=torch.float64, device=dev)
out["sample_us"] = timeit(lambda: _lib.check(lib.rpc_sample_pivots(ctx.handle, _stream(), _ptr(diag), n, _ptr(u), b, _ptr(idx), _ptr(stats))))
ones = torch.ones(n, dtype=torch.float64, device=dev)
out["sample_ones_us"] = timeit(lambda: _lib.check(lib.rpc_sample_pivots(ctx.handle, _stream(), _ptr(ones), n, _ptr(u), b, _ptr(idx), _ptr(stats))))
# gram + rejection on a PSD residual block
P = torch.from_numpy(rng.standard_normal((c, b)) * 0.02).to(dev)
B = rng.standard_normal((b, b + 20)); H0 = torch.from_numpy(B @ B.T / (b + 20) + np.eye(b)).to(dev)
H = H0.clone(); L = torch.zeros((b, b), dtype=torch.float64, device=dev)
acc = torch.zeros(b, dtype=torch.int32, device=dev); info = torch.zeros(4, dtype=torch.int32, device=dev)
def gram():
    H.copy_(H0)
    _lib.check(lib.rpc_gram_residual(ctx.handle, _stream(), _ptr(P), c, b, b, _ptr(H), b))
out["gram_us(incl copy)"] = timeit(gram)
def rej():
    H.copy_(H0)
    _lib.check(lib.rpc_rejection_cholesky(ctx.handle, _stream(), _ptr(H), b, b, _ptr(u), b, 0, 0.0, _ptr(L), b, _ptr(acc), _ptr(info)))
out["reject_us(incl copy)"] = timeit(rej)
s = int(info.cpu()[0]); out["accepted"] = s
k_pad = _round_up(c + s, 16); ldat = int(lib.rpc_panel_ld(s))
At = torch.empty((k_pad, ldat), dtype=torch.float64, device=dev)
out["panel_us"] = timeit(lambda: _lib.check(lib.rpc_build_panel(ctx.handle, _stream(), _ptr(P), b, c, _ptr(acc), s, _ptr(L), b, 0, _ptr(At), ldat, k_pad)))
print(json.dumps(out))
