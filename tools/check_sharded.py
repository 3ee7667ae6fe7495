#!/usr/bin/env python3
"""Run under torchrun on W GPUs: the row-sharded drivers must reproduce the oracle (and hence the 1-GPU path).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py"""
import os, sys
from unittest import mock
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import randomly_pivoted_cholesky_b200 as rpc
from oracle import rpc_oracle as orc

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def replay(seed):
    np.random.seed(seed)
    return mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed)))
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
ok = True
for (alg, kernel, n, d, k, b, bw, kw) in [("accel", "gaussian", 20000, 20, 600, 120, np.sqrt(20), {}),
                                           ("block", "laplace", 15000, 50, 400, 100, 50.0, {}),
                                           ("accel", "matern", 30011, 3, 500, 64, np.sqrt(3), {"nu": 2.5}),
                                           ("rbrp", "gaussian", 12000, 8, 300, 60, np.sqrt(8), {}),
                                           ("bgreedy", "laplace", 9000, 6, 200, 50, 6.0, {}),
                                           ("accel", "gaussian", 20000, 20, 600, 120, np.sqrt(20), {}),
                                           # shard of (4*148+30) column tiles and s > 128: the Schur update goes through the
                                           # row-chunked tail launch, whose finishing kernel publishes the diagonal epoch
                                           ("accel", "gaussian", world * (4 * 148 + 30) * 64, 12, 520, 200, np.sqrt(12), {}),
                                           ("block", "gaussian", world * (4 * 148 + 30) * 64, 12, 400, 180, np.sqrt(12), {})]:
    X = np.random.default_rng(3).standard_normal((n, d))
    A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=float(bw), sharded=True, **kw)
    with replay(77):
        if alg == "accel":
            lra = rpc.accelerated_rpcholesky(A, k, b=b)
        elif alg == "bgreedy":
            lra = rpc.greedy(A, k, b=b)
        else:
            lra = rpc.block_rpcholesky(A, k, b=b, **({"strategy": "rbrp"} if alg == "rbrp" else {}))
    G = lra.G; rows = lra.rows; tr = lra.trace()
    rng, legacy = orc.make_streams(77)
    A0 = orc.OracleKernelMatrix(X, kernel=kernel, bandwidth=float(bw), **kw)
    if alg == "accel":
        ref = orc.accelerated_rpcholesky(A0, k, b, rng, legacy)
    else:
        ref = orc.block_rpcholesky(A0, k, b, rng, alg="greedy" if alg == "bgreedy" else "rp",
                                   strategy="rbrp" if alg == "rbrp" else "regularize")
    same = [int(i) for i in lra.idx] == [int(i) for i in ref.idx]
    eg, er = rel(G, ref.G), rel(rows, ref.rows)
    good = same and eg <= 1e-10 and er <= 1e-10 and abs(tr - ref.trace()) <= 1e-9 * ref.trace() and A.num_queries() == A0.num_queries()
    ok &= good
    if rank == 0:
        print("%s %s N=%d world=%d: pivots %s, |dG|=%.2e, |drows|=%.2e, queries %d/%d -> %s" % (alg, kernel, n, world, same, eg, er, A.num_queries(), A0.num_queries(), "OK" if good else "FAIL"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
