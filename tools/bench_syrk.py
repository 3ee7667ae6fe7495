#!/usr/bin/env python3
"""Times rpc_syrk_rows (M = R R^T, DMMA + tensor-map TMA) alone: python tools/bench_syrk.py [k] [n]  -> one JSON line.
flops counts the lower-triangular tiles the kernel computes (36 of 64 at k = 1000) and, beside it, the full 2 k^2 n of a GEMM."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from randomly_pivoted_cholesky_b200 import ops, load
load()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
R = torch.randn((k, (n + 127) // 128 * 128), dtype=torch.float64, device="cuda")
M = ops.syrk_rows(R, n); torch.cuda.synchronize()
best = 1e9
for _ in range(int(os.environ.get("REPS", "5"))):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); M = ops.syrk_rows(R, n); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
Rn = R[:, :n].contiguous(); _ = Rn @ Rn.T; torch.cuda.synchronize()
e0.record(); Mc = Rn @ Rn.T; e1.record(); torch.cuda.synchronize()
nb = (k + 127) // 128
tri = nb * (nb + 1) // 2 * 128 * 128 * 2.0 * n
print(json.dumps({"k": k, "n": n, "ms": best, "executed_tflops": tri / best / 1e9, "gemm_equivalent_tflops": 2.0 * k * k * n / best / 1e9,
                  "cublas_gemm_ms": e0.elapsed_time(e1), "max_abs_diff_vs_cublas": float((M - Mc).abs().max())}))
