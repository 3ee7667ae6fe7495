#!/usr/bin/env python3
"""Times rpc_kernel_rows alone: python tools/bench_eval.py kernel N d s"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import randomly_pivoted_cholesky_b200 as rpc
kernel = sys.argv[1] if len(sys.argv) > 1 else "laplace"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 50
s = int(sys.argv[4]) if len(sys.argv) > 4 else 200
X = np.random.default_rng(0).standard_normal((n, d))
A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=float(d), **({"nu": 2.5} if kernel == "matern" else {}))
dev = A.device
idx = torch.arange(0, s, dtype=torch.int64, device=dev) * 17
piv = A._dev_new_payload(s); A._dev_gather(idx, s, piv)
out = torch.empty((s, A.n_pad), dtype=torch.float64, device=dev)
def run(): A._dev_rows_into(idx, s, piv, C.c_void_p(out.data_ptr()), A.n_pad)
run(); torch.cuda.synchronize()
best = 1e9
for _ in range(int(os.environ.get("REPS", "3"))):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print(json.dumps({"kernel": kernel, "n": n, "d": d, "s": s, "ms": best, "entries_per_s": s * n / best * 1e3,
                  "fp64_ops_per_s": 2.0 * s * n * d / best * 1e3}))
