#!/usr/bin/env python3
"""Micro-benchmark of the fused Schur-update DMMA kernel (rpc_schur_update_ex) in isolation.
usage: [TRI=0|1] python tools/bench_schur.py [N] [c,s c,s ...]   -> one JSON line per shape
TRI=1 (default): the kernel skips the structurally-zero half of the panel's L^{-T} block (panel_lower); TRI=0: dense panel.
`tflops` counts the ALGORITHMIC flops s (2c + s) N in both cases; `dense_tflops` counts 2 s (c + s) N."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from randomly_pivoted_cholesky_b200 import ops, load
load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_064
n = (n + 127) // 128 * 128
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]] or [(1000, 256), (1000, 200), (1000, 192), (1800, 200), (200, 200), (1000, 128), (1000, 64)]
dev = torch.device("cuda")
reps = int(os.environ.get("REPS", "3"))
tri = int(os.environ.get("TRI", "1"))
for c, s in shapes:
    torch.manual_seed(0)
    G = torch.randn((c + s, n), dtype=torch.float64, device=dev)
    rows = torch.randn((s, n), dtype=torch.float64, device=dev)
    P = torch.randn((c, s), dtype=torch.float64, device=dev)
    L = torch.tril(torch.randn((s, s), dtype=torch.float64, device=dev)) + 30 * torch.eye(s, dtype=torch.float64, device=dev)
    diag = torch.full((n,), 1e9, dtype=torch.float64, device=dev)
    ops.schur_update(G, c, rows, P, L, diag)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        At = ops.schur_update(G, c, rows, P, L, diag)   # includes the tiny panel build; timed separately below
        torch.cuda.synchronize()
        from randomly_pivoted_cholesky_b200 import _lib
        from randomly_pivoted_cholesky_b200.matrix import _ptr, _stream, _round_up
        ctx = _lib.Context.get(0)
        k_pad = _round_up(c + s, 16)
        e0.record()
        _lib.check(ctx.lib.rpc_schur_update_ex(ctx.handle, _stream(), _ptr(G), G.stride(0), c, s, _ptr(At), At.stride(0), k_pad,
                                               _ptr(rows), rows.stride(0), None, _ptr(diag), n, None, 0, tri))
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = 1.0 * s * (2 * c + s) * n
    print(json.dumps({"c": c, "s": s, "n": n, "tri": tri, "ms": best, "tflops": fl / best / 1e9, "frac_of_36.9": fl / best / 1e9 / 36.9,
                      "dense_tflops": 2.0 * s * (c + s) * n / best / 1e9, "hbm_gbs": 8.0 * n * (c + 2 * s) / best / 1e6}), flush=True)
    del G, rows
