#!/usr/bin/env python3
"""One rank of a row-sharded parity run (launched by tests/test_gpu_sharded.py, or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
        tests/sharded_worker.py --suite replay|unseeded [--jsonl gpurun_out/sharded.jsonl]

suite "replay":   the row-sharded drivers against the oracle on replayed draws (pivots bit-exact, G and rows to 1e-10,
                  trace, query counts) -- accelerated / block / RBRP / block greedy, shapes that go through the
                  row-chunked tail launch of the Schur update, s > 256 row chunks.
suite "unseeded": NO replay patch and a different legacy seed on every rank (what a user gets): the ranks must still
                  agree on the pivots (rank 0's draws are broadcast) and the result must be a valid partial Cholesky
                  factorisation (G[:,S]^T G = A[S,:], rows = A[S,:]); also b="auto", whose b is retuned from wall time.
The data plane (peer windows / NCCL) and the speculative row evaluation are selected by the environment
(RPC_B200_PEER, RPC_B200_SPEC) before the package is imported.  Rank 0 prints one JSON line per case.
"""
import argparse
import json
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

TOL = 1e-10      # north_star: "F within 1e-10 of the reference's"


def replay(seed):
    np.random.seed(seed)
    return mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed)))


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def replay_cases(world):
    tail_n = world * (4 * 148 + 30) * 64       # shard of (4*148+30) column tiles: with s > 128 the Schur update takes the tail launch
    return [
        ("accel", "gaussian", 20000, 20, 600, 120, np.sqrt(20), {}),
        ("block", "laplace", 15000, 50, 400, 100, 50.0, {}),
        ("accel", "matern", 30011, 3, 500, 64, np.sqrt(3), {"nu": 2.5}),
        ("rbrp", "gaussian", 12000, 8, 300, 60, np.sqrt(8), {}),
        ("bgreedy", "laplace", 9000, 6, 200, 50, 6.0, {}),
        ("accel", "gaussian", tail_n, 12, 520, 200, np.sqrt(12), {}),
        ("block", "gaussian", tail_n, 12, 400, 180, np.sqrt(12), {}),
        ("block", "gaussian", 16000, 10, 700, 300, np.sqrt(10), {}),          # s = 300 > 256: row-chunked update
    ]


def run_replay(rpc, orc, rank, world, emit):
    ok = True
    for (alg, kernel, n, d, k, b, bw, kw) in replay_cases(world):
        X = np.random.default_rng(3).standard_normal((n, d))
        A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=float(bw), sharded=True, **kw)
        with replay(77):
            if alg == "accel":
                lra = rpc.accelerated_rpcholesky(A, k, b=b)
            elif alg == "bgreedy":
                lra = rpc.greedy(A, k, b=b)
            else:
                lra = rpc.block_rpcholesky(A, k, b=b, **({"strategy": "rbrp"} if alg == "rbrp" else {}))
        G, rows, tr = lra.G, lra.rows, lra.trace()            # all-gathers: every rank takes part
        rng, legacy = orc.make_streams(77)
        A0 = orc.OracleKernelMatrix(X, kernel=kernel, bandwidth=float(bw), **kw)
        if alg == "accel":
            ref = orc.accelerated_rpcholesky(A0, k, b, rng, legacy)
        else:
            ref = orc.block_rpcholesky(A0, k, b, rng, alg="greedy" if alg == "bgreedy" else "rp",
                                       strategy="rbrp" if alg == "rbrp" else "regularize")
        same = [int(i) for i in lra.idx] == [int(i) for i in ref.idx]
        eg, er = rel(G, ref.G), rel(rows, ref.rows)
        good = bool(same and eg <= TOL and er <= TOL and abs(tr - ref.trace()) <= 1e-11 * ref.trace()
                    and A.num_queries() == A0.num_queries())
        flag = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)             # every rank checked its own copy of the gathered factor
        good = bool(flag.item())
        ok &= good
        emit({"suite": "replay", "alg": alg, "kernel": kernel, "N": n, "k": k, "b": b, "world": world, "pivots_equal": same,
              "rel_err_G": eg, "rel_err_rows": er, "queries": [A.num_queries(), A0.num_queries()], "ok": good})
    return ok


def run_unseeded(rpc, orc, rank, world, emit):
    """No replay patch, a different legacy seed per rank: only rank 0's draws may matter."""
    ok = True
    for (kernel, n, d, k, b, bw, kw) in [("gaussian", 24000, 10, 500, 100, np.sqrt(10), {}),
                                         ("matern", 20000, 4, 400, "auto", 2.0, {"nu": 1.5}),
                                         ("laplace", 16000, 12, 300, None, 12.0, {})]:       # None: block_rpcholesky
        np.random.seed(1000 + rank)                                # the ranks' legacy streams DIFFER
        X = np.random.default_rng(5).standard_normal((n, d))
        A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=float(bw), sharded=True, **kw)
        lra = rpc.block_rpcholesky(A, k, b=64) if b is None else rpc.accelerated_rpcholesky(A, k, b=b)
        idx = torch.as_tensor(np.asarray(lra.idx, dtype=np.int64), device="cuda")
        every = [torch.empty_like(idx) for _ in range(world)]
        dist.all_gather(every, idx)
        agree = all(bool(torch.equal(e, idx)) for e in every)
        G, rows = lra.G, lra.rows
        r = G.shape[0]
        S = np.asarray(lra.idx, dtype=np.int64)[:r]
        A0 = orc.OracleKernelMatrix(X, kernel=kernel, bandwidth=float(bw), **kw)
        want = A0.rows(S)
        e_rows = rel(rows, want)
        e_interp = rel(G[:, S].T @ G, want)                        # a partial Cholesky factorisation interpolates its pivot rows
        resid = 1.0 - np.einsum("ij,ij->j", G, G)                  # the residual diagonal stays non-negative
        good = bool(agree and e_rows <= TOL and e_interp <= (5e-6 if b is None else 1e-8) and resid.min() >= -1e-9
                    and len(set(S.tolist())) == r)
        flag = torch.tensor([1 if good else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        good = bool(flag.item())
        ok &= good
        emit({"suite": "unseeded", "kernel": kernel, "N": n, "k": k, "b": b if b is not None else "block b=64", "world": world,
              "ranks_agree_on_pivots": agree, "rel_err_rows": e_rows, "rel_err_interpolation": e_interp,
              "min_residual_diag": float(resid.min()), "rank": int(r), "ok": good})
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--suite", default="replay", choices=["replay", "unseeded"])
    ap.add_argument("--jsonl", default=None)
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import randomly_pivoted_cholesky_b200 as rpc
    from oracle import rpc_oracle as orc
    drv = sys.modules["randomly_pivoted_cholesky_b200.rpcholesky"]
    plane = {"peer_requested": drv.USE_PEER, "spec": drv.USE_SPEC}

    def emit(rec):
        if rank == 0:
            rec.update(plane)
            rec["data_plane"] = drv.LAST_DATA_PLANE
            line = json.dumps(rec)
            print(line, flush=True)
            if args.jsonl:
                os.makedirs(os.path.dirname(os.path.abspath(args.jsonl)), exist_ok=True)
                with open(args.jsonl, "a") as f:
                    f.write(line + "\n")

    ok = (run_replay if args.suite == "replay" else run_unseeded)(rpc, orc, rank, world, emit)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
