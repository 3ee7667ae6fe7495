#!/usr/bin/env python3
"""Small end-to-end runs of every driver / kernel family, meant to be run under compute-sanitizer."""
import os, sys
from unittest import mock
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import randomly_pivoted_cholesky_b200 as rpc
from randomly_pivoted_cholesky_b200 import ops
def replay(seed):
    np.random.seed(seed)
    return mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed)))
rng = np.random.default_rng(0)
for kernel, kw, d in [("gaussian", {}, 7), ("laplace", {}, 5), ("matern", {"nu": 2.5}, 3), ("gaussian", {"extra_stability": True}, 2)]:
    X = rng.standard_normal((1500, d))
    A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=1.5, **kw)
    with replay(1): a = rpc.accelerated_rpcholesky(A, 90, b=40)
    with replay(2): b = rpc.block_rpcholesky(A, 70, b=24)
    with replay(3): c = rpc.simple_rpcholesky(A, 20)
    print(kernel, kw, a.G.shape, b.G.shape, c.G.shape, float(a.trace()), flush=True)
B = rng.standard_normal((40, 300)); D = B.T @ B
with replay(4): r = rpc.rpcholesky(D, 30, b=10)
print("dense", r.G.shape)
X = rng.standard_normal((3000, 4)); A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=2.0)
with replay(5): r = rpc.accelerated_rpcholesky(A, 600, b=300)
print("large b", r.G.shape, r.rounds)
idx, S = ops.sample_pivots(rng.random(100003), rng.random(64)); print("sampler", int(idx[0]), S)
print("sanitize run finished")
