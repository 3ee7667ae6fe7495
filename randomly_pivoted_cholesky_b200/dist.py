"""Row sharding of the hot path over the GPUs of one box.

GPU r owns a contiguous block of points X[lo:hi] and the matching columns of G, rows and the residual
diagonal (= row sharding of F = G^T).  Every N-proportional kernel runs on the local shard with no
communication; per round there are exactly two small exchanges over NVLink:

  1. the residual-diagonal shards (8 bytes per point), so that EVERY rank runs the exact pivot sampler on
     the full diagonal and draws the same, reference-identical pivots;
  2. the pivot payload [X[idx] | |x|^2 | G[:c, idx]], written by the shard that owns each pivot.

Data plane (recorded in bench.py's `config.data_plane`):
  * "peer"  (default on one box): `PeerWindow` below -- one buffer per rank mapped into every process over CUDA
    IPC; the Schur kernel's epilogue stores its block of the new diagonal into every rank's copy and the pivot
    gather stores the owners' entries into every rank's payload, each followed by a release/acquire epoch
    (csrc/peers.cu, csrc/dmma_gemm.cuh).  No NCCL call is left inside a round except the broadcast of the
    round's uniforms (control data, overlapped with the Schur update on a side stream);
  * "nccl"  (RPC_B200_PEER=0, or peer mapping unavailable): `gather_diag` = all-gather, `exchange_payload` =
    sum all-reduce in which every slot has exactly one non-zero contributor (exact).

The b x b work (residual Gram, rejection / shifted Cholesky, panel solve) is replicated: identical inputs
and identical kernels give identical results on every rank; `broadcast_draws` makes the inputs identical (rank
0's uniforms win).  The plain collectives below also run on CPU tensors with the gloo backend
(tests/test_dist_cpu.py).
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


def _src_rank(info: ShardInfo) -> int:
    """Global rank of the group's rank 0 (torch.distributed addresses broadcast sources by GLOBAL rank)."""
    if info.group is None:
        return 0
    return dist.get_global_rank(info.group, 0)


def broadcast_draws(info: ShardInfo, t: torch.Tensor) -> torch.Tensor:
    """Every rank must consume the SAME uniforms (replicated sampler / rejection Cholesky): rank 0's draws win.
    The reference draws from unseeded generators (rpcholesky.py:25,82,151,179), so nothing else ties the ranks'
    streams together.  In place; a no-op for world 1."""
    if info.world > 1:
        dist.broadcast(t, src=_src_rank(info), group=info.group)
    return t


def agree_on_int(info: ShardInfo, value: int, device) -> int:
    """Rank 0's value of a host-side integer decision (the auto-tuned proposal count b of accelerated RPCholesky is
    derived from rank-local wall-clock time, rpcholesky.py:211-220): one tiny broadcast + host read."""
    if info.world == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.broadcast(t, src=_src_rank(info), group=info.group)
    return int(t.item())


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


# ----------------------------------------------------------------------------------------------------------------
# Peer-memory exchange (CUDA only): the two per-round exchanges done by the producing kernels themselves
# ----------------------------------------------------------------------------------------------------------------
class PeerUnavailable(RuntimeError):
    """CUDA IPC / peer access is not available between the GPUs of this job."""


class _DevMem:
    """Wraps a raw device allocation so that torch can view it (torch.as_tensor via __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes, typestr, shape):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": shape, "typestr": typestr, "version": 2}
        self.nbytes = nbytes


class PeerWindow:
    """One buffer per rank, [diag_full (world*shard) | payload (payload_len) | flag words (2*8 u64)], mapped into every
    process of the box over CUDA IPC (include/rpc_b200.h, rpc_peers).  The Schur kernel stores its updated diagonal
    block into every rank's diag_full, the pivot gather stores the owners' entries into every rank's payload; each
    publishes a monotone epoch in the flag words, and `wait_*` enqueues the matching wait on the local stream.
    Windows are cached per (group, sizes): mapping a peer allocation costs milliseconds."""

    _cache = {}

    @classmethod
    def get(cls, info: ShardInfo, ctx, device, payload_len: int):
        """The cached window of this (group, shard, device), or None when peer mapping is unavailable on this box (decided
        collectively, so every rank takes the same path; the drivers then use the NCCL collectives)."""
        key = (id(info.group), info.world, info.rank, info.shard, device.index)
        if key in cls._cache:
            w = cls._cache[key]
            if w is None or w.payload_len >= payload_len:
                return w
            w.close()                                  # too small: unmap and free it before a larger one is made
        try:
            # the payload is sized generously so that one window serves every factorisation of this operator
            w = cls(info, ctx, device, max(payload_len, 1 << 20))
        except PeerUnavailable as exc:
            import warnings
            warnings.warn("peer-memory exchange unavailable (%s): using the NCCL collectives" % exc)
            w = None
        cls._cache[key] = w
        return w

    def __init__(self, info: ShardInfo, ctx, device, payload_len: int):
        import ctypes as C
        from . import _lib
        self.info, self.ctx, self.device = info, ctx, device
        self.payload_len = int(payload_len)
        self.n_total = info.n_total_pad
        ndbl = self.n_total + self.payload_len
        self.nbytes = 8 * ndbl + 8 * 2 * _lib.MAX_PEERS
        # every step below is executed by every rank even after a local failure, and the outcome is agreed on with one
        # MIN all-reduce: a rank must never leave the others waiting inside a collective
        ok, why = 1, ""
        base = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        if ctx.lib.rpc_peer_alloc(ctx.handle, self.nbytes, C.byref(base), handle) != 0:
            ok, why = 0, ctx.lib.rpc_last_error().decode()
        self.base = int(base.value or 0)
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
        allh = torch.empty((info.world, 64), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=info.group)
        allh = allh.cpu().numpy()
        self.bases = []
        self._opened = []
        for r in range(info.world):
            if r == info.rank:
                self.bases.append(self.base)
                continue
            p = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(bytes(allh[r].tobytes()))
            if ok and ctx.lib.rpc_peer_open(ctx.handle, hb, C.byref(p)) != 0:
                ok, why = 0, ctx.lib.rpc_last_error().decode()
            self.bases.append(int(p.value or 0))
            if p.value:
                self._opened.append(int(p.value))
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=info.group)
        if int(flag.item()) == 0:
            for q in self._opened:
                ctx.lib.rpc_peer_close(ctx.handle, C.c_void_p(q))
            if self.base:
                ctx.lib.rpc_peer_free(ctx.handle, C.c_void_p(self.base))
            raise PeerUnavailable(why or "another rank could not map the peer windows")
        peers = _lib.Peers()
        peers.world, peers.rank, peers.shard = info.world, info.rank, info.shard
        for r in range(info.world):
            peers.diag_full[r] = self.bases[r]
            peers.payload[r] = self.bases[r] + 8 * self.n_total
            peers.flags[r] = self.bases[r] + 8 * ndbl
        self.peers = peers
        self.flags_ptr = self.base + 8 * ndbl
        self.diag_full = torch.as_tensor(_DevMem(self.base, 8 * self.n_total, "<f8", (self.n_total,)), device=device)
        self.payload = torch.as_tensor(_DevMem(self.base + 8 * self.n_total, 8 * self.payload_len, "<f8", (self.payload_len,)),
                                       device=device)
        self.status = torch.zeros((1,), dtype=torch.int32, device=device)
        self.epoch_diag = 0
        self.epoch_payload = 0
        dist.barrier(group=info.group)            # every rank has mapped every window before anyone stores into one

    def close(self):
        """Unmap the peers' windows and free this rank's (collective: every rank closes its mappings before any window is
        freed).  The torch views of the window must not be used afterwards."""
        import ctypes as C
        torch.cuda.synchronize(self.device)
        for q in self._opened:
            self.ctx.lib.rpc_peer_close(self.ctx.handle, C.c_void_p(q))
        self._opened = []
        dist.barrier(group=self.info.group)
        if self.base:
            self.ctx.lib.rpc_peer_free(self.ctx.handle, C.c_void_p(self.base))
            self.base = 0
        self.diag_full = self.payload = None
