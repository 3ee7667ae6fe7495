#!/usr/bin/env python3
"""Run the exact pivot sampler at N=1e6 on a mid-factorisation-like diagonal (for ncu / timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from randomly_pivoted_cholesky_b200 import ops

n = 1_000_000
rng = np.random.default_rng(1)
diag = np.clip(rng.random(n) ** 2, 0, 1)
u = rng.random(200)
d = torch.from_numpy(diag).cuda()
for it in range(3):
    idx, S = ops.sample_pivots(d, u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(20):
    idx, S = ops.sample_pivots(d, u)
e1.record(); torch.cuda.synchronize()
print("sampler N=1e6 m=200: %.1f us per call (incl. wrapper overhead)" % (e0.elapsed_time(e1) * 1000 / 20))
