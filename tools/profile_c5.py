import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import randomly_pivoted_cholesky_b200 as rpc
import bench
w = bench.WORKLOADS["c5"]
X = bench.make_points(w["N"], w["d"]); y = bench.krr_targets(X); Xts = np.random.default_rng(1).standard_normal((w["nt"], w["d"]))
dev = torch.device("cuda", 0)
for _ in range(2): bench.run_krr(rpc, X, y, Xts, w, dev, False)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); bench.run_krr(rpc, X, y, Xts, w, dev, False); torch.cuda.synchronize(); t1 = time.perf_counter()
pr.disable()
print("one step: %.1f ms" % (1e3 * (t1 - t0)))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
