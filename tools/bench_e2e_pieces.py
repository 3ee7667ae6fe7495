import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unittest import mock
import numpy as np, torch
import randomly_pivoted_cholesky_b200 as rpc
n, d, k, b = 1_000_000, 50, 2000, 200
X = np.random.default_rng(0).standard_normal((n, d)); Xpin = torch.from_numpy(X).pin_memory()
dev = torch.device("cuda", 0)
def sync(): torch.cuda.synchronize()
for it in range(4):
    sync(); t0 = time.perf_counter()
    A = rpc.KernelMatrix(Xpin, kernel="gaussian", bandwidth=float(np.sqrt(d)), device=dev); sync(); t1 = time.perf_counter()
    np.random.seed(1234)
    with mock.patch("numpy.random.default_rng", lambda *a, **kw: np.random.Generator(np.random.PCG64(1234))):
        lra = rpc.accelerated_rpcholesky(A, k, b=b)
    sync(); t2 = time.perf_counter()
    idx = np.asarray(lra.idx); t3 = time.perf_counter()
    tr = lra.trace(); sync(); t4 = time.perf_counter()
    del lra, A; sync(); t5 = time.perf_counter()
    print("iter %d: ctor %.1f ms, factorise %.1f ms, idx %.2f ms, trace %.1f ms, free %.1f ms" % (it, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4)), flush=True)
