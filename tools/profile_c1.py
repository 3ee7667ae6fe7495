#!/usr/bin/env python3
"""Where does a small simple_rpcholesky (configs[0]) spend its time?  Host enqueue vs device execution."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import randomly_pivoted_cholesky_b200 as rpc

n, d, k = 10_000, 10, 100
X = np.random.default_rng(0).standard_normal((n, d))
A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=float(np.sqrt(d)))
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lra = rpc.simple_rpcholesky(A, k)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("call %.3f ms, +sync %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for it in range(5):
    lra = rpc.simple_rpcholesky(A, k)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
