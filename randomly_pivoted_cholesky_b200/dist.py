"""Row sharding of the hot path over the GPUs of one box.

GPU r owns a contiguous block of points X[lo:hi] and the matching columns of G, rows and the residual
diagonal (= row sharding of F = G^T).  Every N-proportional kernel runs on the local shard with no
communication; per round there are exactly two small NCCL collectives over NVLink:

  1. all-gather of the residual-diagonal shards (8 bytes per point) so that EVERY rank runs the exact
     pivot sampler on the full diagonal and draws the same, reference-identical pivots;
  2. sum all-reduce of the pivot payload [X[idx] | |x|^2 | G[:c, idx]]: the owning shard writes the
     values, all others write zeros, so the sum has exactly one non-zero contributor per slot (exact).

The b x b work (residual Gram, rejection / shifted Cholesky, panel solve) is replicated: identical inputs
and identical kernels give identical results on every rank.  The functions below are plain torch /
torch.distributed and also run on CPU tensors with the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

COL_ALIGN = 128


def _round_up(x, m):
    return (x + m - 1) // m * m


@dataclass(frozen=True)
class ShardInfo:
    """Partition of N points over `world` ranks in equal blocks of `shard` (a multiple of 128) points."""
    n: int
    world: int
    rank: int
    shard: int          # points per rank (last ranks may own fewer real points)
    lo: int             # first global index owned
    hi: int             # one past the last global index owned (<= n)
    group: Optional[object] = None

    @property
    def n_loc(self):
        return self.hi - self.lo

    @property
    def n_total_pad(self):
        return self.shard * self.world


def make_shard_info(n: int, world: int, rank: int, group=None) -> ShardInfo:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    shard = _round_up(-(-n // world), COL_ALIGN)
    lo = min(rank * shard, n)
    hi = min(lo + shard, n)
    if (world - 1) * shard >= n and world > 1:
        raise ValueError("N=%d is too small to shard over %d ranks in blocks of %d points" % (n, world, shard))
    return ShardInfo(n=n, world=world, rank=rank, shard=shard, lo=lo, hi=hi, group=group)


def owner_of(info: ShardInfo, idx):
    """Rank that owns global index idx (tensor or int)."""
    return idx // info.shard


def gather_diag(info: ShardInfo, diag_loc: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[r*shard:(r+1)*shard] = rank r's diag shard: the full residual diagonal in global order, zero padded."""
    if info.world == 1:
        return diag_loc
    dist.all_gather_into_tensor(out, diag_loc, group=info.group)
    return out


def exchange_payload(info: ShardInfo, payload: torch.Tensor) -> torch.Tensor:
    """Sum all-reduce of a payload in which every slot was written by its owner and zeroed by the others."""
    if info.world > 1:
        dist.all_reduce(payload, op=dist.ReduceOp.SUM, group=info.group)
    return payload


def gather_columns(info: ShardInfo, local: torch.Tensor) -> torch.Tensor:
    """Assemble the (k, N) matrix from per-rank column shards (k, shard): host-side materialisation of G / rows."""
    if info.world == 1:
        return local[:, :info.n]
    k = local.shape[0]
    parts = [torch.empty((k, info.shard), dtype=local.dtype, device=local.device) for _ in range(info.world)]
    dist.all_gather(parts, local[:, :info.shard].contiguous(), group=info.group)
    return torch.cat(parts, dim=1)[:, :info.n]


def current_shard_info(n: int, group=None) -> ShardInfo:
    """ShardInfo from the initialised torch.distributed default (or given) group; world 1 if not initialised."""
    if dist.is_available() and dist.is_initialized():
        return make_shard_info(n, dist.get_world_size(group), dist.get_rank(group), group)
    return make_shard_info(n, 1, 0, None)
