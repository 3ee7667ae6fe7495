"""Host-side logic of the row-sharded path on CPU: world_size-2 gloo processes exercise the partition, the
diagonal all-gather and the exact payload exchange (randomly_pivoted_cholesky_b200/dist.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from randomly_pivoted_cholesky_b200 import dist as rdist


def test_partition_properties():
    for n in (1000, 12345, 1_000_000, 4_000_000):
        for world in (1, 2, 4, 8):
            infos = [rdist.make_shard_info(n, world, r) for r in range(world)]
            assert all(i.shard % 128 == 0 for i in infos)
            assert infos[0].lo == 0 and infos[-1].hi == n
            for a, b in zip(infos, infos[1:]):
                assert a.hi == b.lo and a.n_loc == a.shard
            assert sum(i.n_loc for i in infos) == n
            assert infos[0].n_total_pad >= n
            idx = torch.tensor([0, n // 3, n - 1])
            own = rdist.owner_of(infos[0], idx)
            for j, o in zip(idx.tolist(), own.tolist()):
                assert infos[o].lo <= j < infos[o].hi
    with pytest.raises(ValueError):
        rdist.make_shard_info(200, 8, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, k, m):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        info = rdist.current_shard_info(n)
        assert info.world == world and info.rank == rank
        rng = np.random.default_rng(0)
        diag_full = rng.random(n)
        G_full = rng.standard_normal((k, n))
        X_full = rng.standard_normal((n, 5))
        # 1. diagonal all-gather reproduces the global order with zero padding
        loc = torch.zeros(info.shard, dtype=torch.float64)
        loc[:info.n_loc] = torch.from_numpy(diag_full[info.lo:info.hi])
        out = torch.empty(info.n_total_pad, dtype=torch.float64)
        full = rdist.gather_diag(info, loc, out)
        assert np.array_equal(full[:n].numpy(), diag_full)
        assert float(full[n:].abs().sum()) == 0.0
        # 2. payload exchange: owner writes, others write zeros, the sum is exact
        idx = torch.from_numpy(rng.integers(0, n, m))
        payload = torch.zeros(m * 5 + k * m, dtype=torch.float64)
        Xp = payload[:m * 5].view(m, 5)
        P = payload[m * 5:].view(k, m)
        for j, g in enumerate(idx.tolist()):
            if info.lo <= g < info.hi:
                Xp[j] = torch.from_numpy(X_full[g])
                P[:, j] = torch.from_numpy(G_full[:, g])
        rdist.exchange_payload(info, payload)
        assert np.array_equal(Xp.numpy(), X_full[idx.numpy()])
        assert np.array_equal(P.numpy(), G_full[:, idx.numpy()])
        # 2b. the round's uniforms: every rank draws from its OWN unseeded stream, rank 0's draws win everywhere
        mine = torch.from_numpy(np.random.default_rng(1000 + rank).random(2 * m))
        want = np.random.default_rng(1000).random(2 * m)
        rdist.broadcast_draws(info, mine)
        assert np.array_equal(mine.numpy(), want)
        assert rdist.agree_on_int(info, 100 + 7 * rank, "cpu") == 100
        # 3. column shards assemble back into the full factor
        Gl = torch.zeros((k, info.shard), dtype=torch.float64)
        Gl[:, :info.n_loc] = torch.from_numpy(G_full[:, info.lo:info.hi])
        got = rdist.gather_columns(info, Gl)
        assert np.array_equal(got.numpy(), G_full)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1000, 7, 12), nprocs=2, join=True)
