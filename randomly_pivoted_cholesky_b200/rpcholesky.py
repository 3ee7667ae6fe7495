"""The factorisation drivers: rpcholesky / simple_rpcholesky / block_rpcholesky / accelerated_rpcholesky.

Same entry points, keyword arguments, random-stream consumption and error behaviour as the reference's
rpcholesky.py (:12-50 simple, :65-138 block, :140-157 + :159-232 accelerated, :234-257 dispatch); every
N-proportional or b x b operation runs in the CUDA library (include/rpc_b200.h).  The host only draws
the uniforms (numpy, so that a seeded replay reproduces the reference's pivots bit for bit), reads back
the per-round pivot indices / accepted count, and keeps the pivot list.

Random streams:
  proposals  <- np.random.default_rng()      one double per step (simple), 2*min(k-c,b) per block, b per round
  rejections <- np.random.random_sample(b)   global legacy stream, b per round (== b calls of np.random.rand())
"""
from __future__ import annotations

import ctypes as C
import math
import time
from warnings import warn

import numpy as np
import torch

from . import _lib
from ._lib import KT, check
from .lra import PSDLowRank
from . import dist as rdist
from .matrix import AbstractPSDMatrix, PSDMatrix, _ptr, _round_up, _stream

EPS = float(np.finfo(float).eps)
AT_LD_ALIGN = 256          # panel leading dimension granularity (widest row tile of the DMMA kernel)

import os as _os
# Row-sharded runs exchange the residual diagonal and the pivot payload through NVLink peer memory, stored by the
# producing kernels themselves (dist.PeerWindow); RPC_B200_PEER=0 falls back to the two NCCL collectives per round
USE_PEER = _os.environ.get("RPC_B200_PEER", "1") != "0"
# Accelerated RPCholesky on a sharded operator evaluates the kernel rows of ALL b proposals on a second stream while the
# single-CTA rejection Cholesky runs (147 of 148 SMs would idle otherwise); the Schur update then reads the accepted rows
# through a row map.  Worth it once the evaluation is short compared with the serial chain, i.e. when sharded:
# RPC_B200_SPEC = "auto" (sharded only) | "1" (always) | "0" (never)
USE_SPEC = _os.environ.get("RPC_B200_SPEC", "auto")
# sharded accelerated runs draw + broadcast the NEXT round's uniforms at the start of a round (RPC_B200_PREFETCH=0: per round)
USE_PREFETCH = _os.environ.get("RPC_B200_PREFETCH", "1") != "0"
# how long a rank waits for another rank's epoch before the factorisation is aborted (0 = the library default, 20 s)
PEER_TIMEOUT_MS = int(_os.environ.get("RPC_B200_PEER_TIMEOUT_MS", "0"))

# bench.py sets this to a list to collect (start_event, end_event, flops, bytes, phase) per timed launch
KERNEL_TIMERS = None
# data plane the last factorisation actually ran on: "single" | "peer" (NVLink peer windows) | "nccl" (collectives)
LAST_DATA_PLANE = "single"


class _timed:
    """CUDA-event bracket around one launch on the current stream (only active under bench.py)."""

    def __init__(self, phase, flops, nbytes, shape=None):
        self.on = KERNEL_TIMERS is not None
        self.meta = (phase, flops, nbytes, shape)

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.on:
            self.e1.record()
            KERNEL_TIMERS.append((self.e0, self.e1) + self.meta)


def _on_operator_device(fn):
    """Run a driver with the operator's device current: streams, workspace allocations and launches of the C ABI all refer
    to the current device, which need not be the one the operator lives on."""
    import functools

    @functools.wraps(fn)
    def wrapped(A, *args, **kwargs):
        if not isinstance(A, AbstractPSDMatrix):           # a raw array is wrapped on the current device by the driver itself
            return fn(A, *args, **kwargs)
        with torch.cuda.device(A.device):
            return fn(A, *args, **kwargs)
    return wrapped


def _as_operator(A):
    if isinstance(A, AbstractPSDMatrix):
        return A
    if isinstance(A, (np.ndarray, torch.Tensor)):
        return PSDMatrix(A)                                   # rpcholesky.py:13-14, 66-67, 160-161
    raise RuntimeError("A must be an ndarray or an AbstractPSDMatrix")


class _Payload:
    """Views into the per-round pivot payload [X[idx] | |x|^2 | G[:c, idx]] (one flat buffer, one collective)."""

    def __init__(self, X, norm):
        self.X, self.norm = X, norm


class _State:
    """Device buffers of one factorisation (data layout in HBM).  Shard aware: every N-sized buffer
    holds this rank's block of columns; the small replicated buffers are identical on all ranks."""

    def __init__(self, A, k, m_max):
        self.A = A
        self.dev = A.device
        self.ctx = A._ctx
        self.lib = self.ctx.lib
        self.info = A.shard_info
        self.n = A.shape[0]                      # global N
        self.n_pad = A.n_pad                     # local padded column count
        self.n_loc = self.info.n_loc
        self.k = k
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.G = torch.empty((k, self.n_pad), **f64)
        self.rows = torch.empty((k, self.n_pad), **f64)
        self.diag = torch.empty((self.n_pad,), **f64)
        self.m_max = m_max
        self.u_dev = torch.empty((4 * m_max,), **f64)            # two slots of [proposal | rejection] uniforms (prefetch)
        self.idx_dev = torch.empty((m_max,), dtype=torch.int64, device=self.dev)
        self.sel_dev = torch.empty((m_max,), dtype=torch.int64, device=self.dev)
        self.stats = torch.zeros((4,), **f64)
        # payload: [Xpiv (m_max x ldx) | pnorm (m_max) | P (k x m_max)] -- the used prefix is one contiguous range
        self.ldx = getattr(A, "ldx", 0)
        nx = m_max * self.ldx
        payload_len = nx + (m_max if self.ldx else 0) + max(k, 1) * m_max
        self.peer = None
        self.diag_pending = False              # True: the last Schur update published the new diagonal into diag_full
        if (self.info.world > 1 and self.info.world <= _lib.MAX_PEERS and USE_PEER and self.dev.type == "cuda"
                and self.ldx):
            self.peer = rdist.PeerWindow.get(self.info, self.ctx, self.dev, payload_len)     # None: no peer access here
        global LAST_DATA_PLANE
        LAST_DATA_PLANE = "single" if self.info.world == 1 else ("peer" if self.peer is not None else "nccl")
        if self.peer is not None:
            self.payload = self.peer.payload[:payload_len]
            self.diag_full = self.peer.diag_full
        else:
            self.payload = torch.zeros((payload_len,), **f64)
            self.diag_full = torch.empty((self.info.n_total_pad,), **f64) if self.info.world > 1 else self.diag
        if self.ldx:
            self.piv = _Payload(self.payload[:nx].view(m_max, self.ldx), self.payload[nx:nx + m_max])
            self.sel = _Payload(torch.zeros((m_max, self.ldx), **f64), torch.zeros((m_max,), **f64))
            self.p_off = nx + m_max
        else:
            self.piv = self.sel = None
            self.p_off = 0
        self.P = self.payload[self.p_off:].view(max(k, 1), m_max)
        self.H = torch.empty((m_max, m_max), **f64)
        self.L = torch.empty((m_max, m_max), **f64)
        self.acc = torch.zeros((m_max,), dtype=torch.int32, device=self.dev)
        self.info_dev = torch.zeros((4,), dtype=torch.int32, device=self.dev)
        self.k_pad_max = _round_up(k + m_max, KT)
        self.At = torch.empty((self.k_pad_max * (_round_up(m_max, AT_LD_ALIGN) + 16),), **f64)   # panel, per-round ld
        # pinned staging
        self.u_host = torch.empty((4 * m_max,), dtype=torch.float64).pin_memory()
        self.sel_host = torch.empty((m_max,), dtype=torch.int64).pin_memory()
        self.rb_host = None
        self.spec = False
        self.draw_stream = None                # sharded: side stream of the uniform upload + broadcast
        self.ev_draw = [None, None]            # per slot: the upload (+ broadcast) of the slot's uniforms has completed
        A._dev_diag_into(self.diag)
        A._count(self.n)                                       # A.diag() counts N queries (matrix.py:33-39)

    def readback(self, m, with_chol=True):
        self.publish(m, with_chol)
        return self.collect(m)

    def publish(self, m, with_chol=True):
        """The round's host sync, device half.  A tiny kernel writes [seq | stats | info | acc[:m] | idx[:m]] straight into pinned
        host memory and the host spins on the sequence word: no stream synchronisation (~100 us of wake-up latency
        on this box), no staging copies.  Returns (stats[4], info[4], acc[m], idx[m]) as numpy arrays."""
        need = 9 + 2 * m
        if self.rb_host is None or self.rb_host.numel() < need:
            self.rb_host = torch.zeros((max(need, 4096),), dtype=torch.float64).pin_memory()
            self.rb_np = self.rb_host.numpy()
            self.rb_seq = 0
        self.rb_seq += 1
        seq = float(self.rb_seq)
        # a timed-out peer wait sets info_dev[3] (sticky): with peer windows the status word travels with every round
        check(self.lib.rpc_publish_round(self.ctx.handle, _stream(), _ptr(self.stats),
                                         _ptr(self.info_dev) if (with_chol or self.peer is not None) else None,
                                         _ptr(self.acc) if with_chol else None, _ptr(self.idx_dev), m,
                                         C.c_void_p(self.rb_host.data_ptr()), seq), "rpc_publish_round")

    def collect(self, m):
        """Host half of the round's sync: spin on the sequence word of the last publish()."""
        need = 9 + 2 * m
        seq = float(self.rb_seq)
        buf = self.rb_np
        spins = 0
        while buf[0] != seq:
            spins += 1
            if spins & 0xFFFF == 0 and torch.cuda.current_stream().query():     # the stream drained without publishing
                if buf[0] != seq:
                    torch.cuda.synchronize(self.dev)                            # surfaces the CUDA error, if any
                    if buf[0] != seq:
                        raise RuntimeError("read-back kernel did not publish its result")
        out = buf[1:need].copy()
        if self.peer is not None and out[7] != 0:
            raise RuntimeError("peer-memory exchange timed out waiting for another rank's epoch (this round's pivots are invalid)")
        return out[0:4], out[4:8].astype(np.int64), out[8:8 + m].astype(np.int64), out[8 + m:8 + 2 * m].astype(np.int64)

    # ---- thin wrappers over the C ABI ------------------------------------------------------------
    def upload_uniforms(self, up, ur=None, slot=0):
        """u_dev[slot] <- [proposal uniforms | rejection uniforms]; returns the slot's offset into u_dev.  Row-sharded runs use
        RANK 0's draws on every rank (the reference's generators are unseeded, so nothing else makes the ranks' streams agree).
        With two slots the accelerated driver uploads round t+1's draws at the start of round t."""
        m = len(up)
        off = slot * 2 * self.m_max
        self.u_host[off:off + m] = torch.from_numpy(up)
        tot = m
        if ur is not None:
            self.u_host[off + m:off + m + len(ur)] = torch.from_numpy(ur)
            tot = m + len(ur)
        # On the MAIN stream, on purpose.  The broadcast is an NCCL kernel that spins until every rank has joined; launched from
        # a side stream it lands beside the persistent Schur / evaluator kernels (the host runs one update ahead of the device),
        # takes an SM as soon as one of their CTAs exits and keeps it while the slowest rank arrives -- and the NEXT persistent
        # launch, whose 148 CTAs each need a whole SM and own a fixed share of the tiles, then finishes a whole tile-share late
        # (measured: +9 ms per factorisation on 4 GPUs, +10 ms on 8 without speculation).  In stream order it runs between two
        # updates, when the device has nothing else to do.
        self.u_dev[off:off + tot].copy_(self.u_host[off:off + tot], non_blocking=True)
        rdist.broadcast_draws(self.info, self.u_dev[off:off + tot])
        self.ev_draw[slot] = None
        return off

    def wait_uniforms(self, slot=0):
        if self.ev_draw[slot] is not None:
            torch.cuda.current_stream().wait_event(self.ev_draw[slot])
            self.ev_draw[slot] = None

    def sample(self, m, u_off=0):
        """Every rank samples on the FULL residual diagonal (all-gathered when sharded): identical pivots everywhere."""
        if self.peer is not None and self.diag_pending:
            # every rank's Schur kernel stored its block of the new diagonal here and published the epoch
            check(self.lib.rpc_peer_wait(self.ctx.handle, _stream(), C.c_void_p(self.peer.flags_ptr), self.info.world,
                                         self.peer.epoch_diag, PEER_TIMEOUT_MS, self._peer_status()), "rpc_peer_wait")
            self.diag_pending = False
            full = self.diag_full
        else:
            full = rdist.gather_diag(self.info, self.diag, self.diag_full)
        u = C.c_void_p(self.u_dev.data_ptr() + 8 * u_off)
        check(self.lib.rpc_sample_pivots(self.ctx.handle, _stream(), _ptr(full), full.numel(), u, m, _ptr(self.idx_dev),
                                         _ptr(self.stats)), "rpc_sample_pivots")

    def _peer_status(self):
        """Status word of the peer waits: info_dev[3], published to the host with every round's read-back."""
        return C.c_void_p(self.info_dev.data_ptr() + 12)

    def gather(self, idx_dev, m, c):
        """payload <- [X[idx] | |x|^2 | G[:c, idx]]; owners write, the others write zeros, then one sum all-reduce."""
        A = self.A
        X = getattr(A, "_Xd", None)
        xn = getattr(A, "_xnorm", None)
        d = getattr(A, "d_pad", 0)
        piv = self.piv
        if self.peer is not None:
            # owners store their entries into EVERY rank's payload through peer memory, then all wait for all epochs
            pw = self.peer
            pw.epoch_payload += 1
            check(self.lib.rpc_gather_pivots_peers(self.ctx.handle, _stream(), _ptr(idx_dev), m, _ptr(X), _ptr(xn), d, self.ldx,
                                                   self.ldx, self.m_max * self.ldx, _ptr(self.G) if c > 0 else None,
                                                   self.n_pad, c, self.p_off, self.m_max, self.info.lo, self.info.hi,
                                                   C.byref(pw.peers), pw.epoch_payload), "rpc_gather_pivots_peers")
            check(self.lib.rpc_peer_wait(self.ctx.handle, _stream(), C.c_void_p(pw.flags_ptr + 8 * _lib.MAX_PEERS),
                                         self.info.world, pw.epoch_payload, PEER_TIMEOUT_MS, self._peer_status()), "rpc_peer_wait")
            return
        check(self.lib.rpc_gather_pivots(self.ctx.handle, _stream(), _ptr(idx_dev), m, _ptr(X), _ptr(xn), d, self.ldx,
                                         _ptr(piv.X) if piv is not None else None,
                                         _ptr(piv.norm) if piv is not None else None, self.ldx,
                                         _ptr(self.G) if c > 0 else None, self.n_pad, c,
                                         _ptr(self.P) if c > 0 else None, self.m_max, self.info.lo, self.info.hi),
              "rpc_gather_pivots")
        if self.info.world > 1:
            rdist.exchange_payload(self.info, self.payload[:self.p_off + c * self.m_max])

    def residual_block(self, idx_dev, m, c, count=True):
        """H = A[idx,idx] - G[:c,idx]^T G[:c,idx]  (rpcholesky.py:189; also C of :103, which the reference
        slices out of G_new and therefore does not count as queries).  Replicated on every rank."""
        self.A._dev_block_into(idx_dev, m, self.piv, self.H, self.m_max)
        if count:
            self.A._count(m * m)
        check(self.lib.rpc_gram_residual(self.ctx.handle, _stream(), _ptr(self.P), c, m, self.m_max, _ptr(self.H),
                                         self.m_max), "rpc_gram_residual")

    def cholesky(self, m, mode, max_accept, shift, u_off):
        u = C.c_void_p(self.u_dev.data_ptr() + 8 * u_off)
        check(self.lib.rpc_rejection_cholesky(self.ctx.handle, _stream(), _ptr(self.H), m, self.m_max, u, max_accept, mode,
                                              float(shift), _ptr(self.L), self.m_max, _ptr(self.acc), _ptr(self.info_dev)),
              "rpc_rejection_cholesky")

    # ---- speculative evaluation of all proposals beside the rejection Cholesky (accelerated, sharded) ---------------------
    def enable_speculation(self):
        A = self.A
        # break-even is where evaluating b rows takes about as long as the rejection kernel: N/world <~ 1.5e5 points at d = 50
        # (round 2, faster evaluator: on 4 GPUs the plain order measured 41.6 ms against 43 with speculation)
        if not (hasattr(A, "_spec") and self.ldx) or USE_SPEC == "0" or (USE_SPEC == "auto" and self.info.world < 8):
            return
        self.spec = True
        self.rows_all = torch.empty((self.m_max, self.n_pad), dtype=torch.float64, device=self.dev)
        if getattr(self.ctx, "_side", None) is None:          # cached per device: a second rpc_ctx (own scratch: the side
            self.ctx._side = (_lib.Context.extra(self.dev.index), torch.cuda.Stream(device=self.dev))   # stream must not
        self.ctx2, self.side = self.ctx._side                 # share the main workspace) and the side stream
        self.ev_fork, self.ev_eval = torch.cuda.Event(), torch.cuda.Event()

    def launch_eval_all(self, m):
        """rows_all[:m] = A[idx, :] on the side stream, one SM left free for the rejection kernel."""
        A = self.A
        main = torch.cuda.current_stream()
        self.ev_fork.record(main)
        self.side.wait_event(self.ev_fork)
        with torch.cuda.stream(self.side):
            with _timed("eval", 2.0 * m * self.n_loc * A.d, 8.0 * self.n_loc * (m + A.d)):
                check(self.lib.rpc_ctx_reserve_sms(self.ctx2.handle, 1), "rpc_ctx_reserve_sms")
                check(self.lib.rpc_kernel_rows(self.ctx2.handle, _stream(), C.byref(A._spec), _ptr(A._Xd), _ptr(A._xnorm),
                                               self.n_pad, A.d_pad, A.ldx, _ptr(self.piv.X), _ptr(self.piv.norm), m, A.ldx,
                                               _ptr(self.rows_all), self.n_pad), "rpc_kernel_rows")
            self.ev_eval.record(self.side)

    def prelaunch_panel(self, c, s_max):
        """Panel solve + compaction of the accepted pivots with the accepted count read on the device (info_dev[0]):
        enqueued behind the rejection kernel BEFORE the host reads the round back, so the host's wake-up and launch
        latency hides behind them.  update(..., panel_done=True) then only evaluates the rows and runs the Schur update."""
        piv = self.piv
        check(self.lib.rpc_build_panel_dev(self.ctx.handle, _stream(), _ptr(self.P), self.m_max, c, _ptr(self.acc),
                                           _ptr(self.info_dev), s_max, _ptr(self.L), self.L.stride(0), _ptr(self.At),
                                           _ptr(self.idx_dev), _ptr(piv.X) if piv is not None else None,
                                           _ptr(piv.norm) if piv is not None else None, self.ldx, _ptr(self.sel_dev),
                                           _ptr(self.sel.X) if piv is not None else None,
                                           _ptr(self.sel.norm) if piv is not None else None), "rpc_build_panel_dev")

    def update(self, c, s, acc_dev, idx_dev, panel_mode=0, Lmat=None, panel_done=False, spec_rows=False):
        """rows[c:c+s] = A[S,:]; G[c:c+s] = L^{-1}(rows - P^T G[:c]); diag update (rpcholesky.py:101-107,128-129 /
        205-209) on this rank's columns.  S = idx[acc] (acc None: all of idx); the selected pivot points are
        compacted out of the (already exchanged) payload, so no further communication is needed."""
        k_pad = _round_up(c + s, KT)
        ldat = int(self.lib.rpc_panel_ld(s))          # stride that lets the kernel fetch a panel stage in one bulk copy
        Lm = self.L if Lmat is None else Lmat
        if not panel_done:
            with _timed("panel", 1.0 * s * s * c + s ** 3 / 3.0, 8.0 * c * s):
                check(self.lib.rpc_build_panel(self.ctx.handle, _stream(), _ptr(self.P), self.m_max, c, _ptr(acc_dev), s,
                                               _ptr(Lm), Lm.stride(0), panel_mode, _ptr(self.At), ldat, k_pad),
                      "rpc_build_panel")
        A = self.A
        if panel_done:
            sel_idx, sel_piv = self.sel_dev, self.sel
        elif acc_dev is not None:
            piv = self.piv
            check(self.lib.rpc_compact_selected(self.ctx.handle, _stream(), _ptr(acc_dev), s, _ptr(idx_dev),
                                                _ptr(piv.X) if piv is not None else None,
                                                _ptr(piv.norm) if piv is not None else None, self.ldx, _ptr(self.sel_dev),
                                                _ptr(self.sel.X) if piv is not None else None,
                                                _ptr(self.sel.norm) if piv is not None else None), "rpc_compact_selected")
            sel_idx, sel_piv = self.sel_dev, self.sel
        else:
            sel_idx, sel_piv = idx_dev, self.piv
        rows_ptr = C.c_void_p(self.rows.data_ptr() + 8 * c * self.n_pad)
        n, d = self.n_loc, getattr(A, "d", 0)
        if spec_rows:
            # the rows of all proposals were evaluated on the side stream; the accepted ones are read through acc
            torch.cuda.current_stream().wait_event(self.ev_eval)
            A._count(s * self.n)
            self._schur(c, s, ldat, k_pad, _ptr(self.rows_all), _ptr(self.acc), panel_mode == 0,
                        rows_out=C.c_void_p(self.rows.data_ptr() + 8 * c * self.n_pad))
            return
        with _timed("eval", 2.0 * s * n * d, 8.0 * n * (s + d)):
            A._dev_rows_into(sel_idx, s, sel_piv, rows_ptr, self.n_pad)
        A._count(s * self.n)
        self._schur(c, s, ldat, k_pad, rows_ptr, None, panel_mode == 0)

    def _schur(self, c, s, ldat, k_pad, rows_src, rowmap, panel_lower=True, rows_out=None):
        """G[c:c+s] = At^T [G[:c]; rows_new], diag update; new row i is rows_src[rowmap[i]] (rowmap None: rows_src[i]).
        Sharded with peer windows: the kernel also stores the new diagonal into every rank's copy and publishes the epoch.
        Algorithmic work: s (2c + s) N flops; reads G[:c] and rows_new, writes G_new."""
        n = self.n_loc
        peers, epoch = None, 0
        if self.peer is not None:
            self.peer.epoch_diag += 1
            peers, epoch = C.byref(self.peer.peers), self.peer.epoch_diag
        # algorithmic flops: Schur product 2 s c N + triangular solve s^2 N; with a dense panel (eigh fallback)
        # the s x s block is a full product
        flops = (1.0 * s * (2 * c + s) if panel_lower else 2.0 * s * (c + s)) * n
        with _timed("schur", flops, 8.0 * n * (c + 2 * s + 2), (c, s, self.n_pad) if panel_lower else None):
            if rows_out is not None:
                # speculative rows: the kernel also writes rows[c:c+s] = rows_all[acc] from the stages it multiplies
                check(self.lib.rpc_schur_update_rows(self.ctx.handle, _stream(), _ptr(self.G), self.n_pad, c, s, _ptr(self.At), ldat,
                                                     k_pad, rows_src, self.n_pad, rowmap, rows_out, self.n_pad, _ptr(self.diag),
                                                     self.n_pad, peers, epoch, 1 if panel_lower else 0), "rpc_schur_update_rows")
            else:
                check(self.lib.rpc_schur_update_ex(self.ctx.handle, _stream(), _ptr(self.G), self.n_pad, c, s, _ptr(self.At), ldat,
                                                   k_pad, rows_src, self.n_pad, rowmap, _ptr(self.diag), self.n_pad, peers, epoch,
                                                   1 if panel_lower else 0),
                      "rpc_schur_update_ex")
        self.diag_pending = self.peer is not None

    def result(self, count, arr_idx):
        # sharded: keep the whole padded shard (equal widths on all ranks, so `.G` can all-gather); else cut to N
        nl = self.n_pad if self.info.world > 1 else self.n_loc
        if self.peer is not None and int(self.info_dev[3].item()) != 0:
            raise RuntimeError("peer-memory exchange timed out waiting for another rank's epoch")
        return PSDLowRank(self.G[:count, :nl], n=self.n, idx=arr_idx, rows=self.rows[:count, :nl], shard_info=self.info)


def _checked_block(b):
    """The reference accepts any block size; the single-CTA b x b kernels hold one block in shared memory."""
    b = int(b)
    if b < 1:
        raise ValueError("block size b must be positive")
    if b > _lib.MAX_BLOCK:
        raise ValueError("block size b = %d exceeds the device limit RPC_MAX_BLOCK = %d (include/rpc_b200.h)" % (b, _lib.MAX_BLOCK))
    return b


def _raise_if_invalid(stats_host):
    if stats_host[2] != 0.0:
        raise RuntimeError("internal error: exact-scan binade bound violated")
    if stats_host[1] != 0.0:
        # what Generator.choice raises when diags/sum(diags) is not a distribution (callers rely on it:
        # experiments/comparison.py:51)
        raise ValueError("probabilities contain NaN" if not (stats_host[0] == stats_host[0]) or stats_host[0] == 0.0
                         else "probabilities are not non-negative")


# --------------------------------------------------------------------------------------------------
# simple RPCholesky: cholesky_helper(A, k, 'rp', stoptol) -- rpcholesky.py:12-50
# --------------------------------------------------------------------------------------------------
@_on_operator_device
def cholesky_helper(A, k, alg, stoptol=0):
    if alg not in ("rp", "greedy", "rgreedy"):
        raise RuntimeError("Algorithm '{}' not recognized".format(alg))
    A = _as_operator(A)
    if stoptol is None:
        stoptol = 0
    if A.shard_info.world > 1:
        raise NotImplementedError("simple RPCholesky is not row-sharded (one pivot per step is latency bound); "
                                  "use block or accelerated RPCholesky on a sharded operator")
    st = _State(A, k, 1)
    lib, ctx = st.lib, st.ctx
    rng = np.random.default_rng()
    spec, X, xn, d, ldx, Ad, lda = A._simple_step_args()
    spec_ref = C.byref(spec) if spec is not None else None

    idx_all = torch.empty((k + 1,), dtype=torch.int64, device=st.dev)
    stats_all = torch.zeros((k + 1, 4), dtype=torch.float64, device=st.dev)
    # one uniform per step (rng.choice(range(n), p=...), rpcholesky.py:31); the stream is state independent,
    # so all k are drawn and uploaded at once and the loop below never waits for the device
    u_all = torch.from_numpy(rng.random(k) if alg == "rp" else np.zeros(1)).to(st.dev)
    count = k
    arr_len = k
    check_every_step = stoptol > 0
    orig_trace = None

    def sample_into(i, m):
        check(lib.rpc_sample_pivots(ctx.handle, _stream(), _ptr(st.diag), st.n_pad,
                                    C.c_void_p(u_all.data_ptr() + 8 * i) if m else None, m,
                                    C.c_void_p(idx_all.data_ptr() + 8 * i) if m else None,
                                    C.c_void_p(stats_all.data_ptr() + 32 * i)), "rpc_sample_pivots")

    def pick_greedy(i):
        """np.argmax(diags) (:35) or, for 'rgreedy', rng.choice over the tied maxima (:33, drawn on the host)."""
        gstats = stats_all[k]                  # scratch row: [max, #ties]
        check(lib.rpc_argmax(ctx.handle, _stream(), _ptr(st.diag), st.n, C.c_void_p(idx_all.data_ptr() + 8 * i),
                             _ptr(gstats)), "rpc_argmax")
        if alg == "rgreedy":
            mx, ties = (float(v) for v in gstats[:2].cpu().numpy())
            if ties > 1:
                cand = torch.nonzero(st.diag[:st.n] == mx).flatten().cpu().numpy()
                idx_all[i] = int(rng.choice(cand))

    fast = (not check_every_step) and alg in ("rp", "greedy")
    if fast:
        # no stoptol test and no host-side tie break: the whole loop is enqueued by one native call
        # algorithmic traffic of the k steps together (per step: stream G[:i], read X, write G[i] and rows[i], diag r/w, and
        # the sampler's two passes over the diagonal): the bracket covers sampler + step launches of the whole loop
        with _timed("simple", 2.0 * st.n * (k * (k - 1) / 2 + k * d), 8.0 * st.n * (k * (k - 1) / 2 + k * (d + 7))):
            check(lib.rpc_simple_run(ctx.handle, _stream(), 1 if alg == "greedy" else 0, spec_ref, _ptr(X), _ptr(xn), d, ldx,
                                     _ptr(Ad), lda, st.n, _ptr(u_all), _ptr(idx_all), _ptr(stats_all), 0, k, _ptr(st.G), st.n_pad,
                                     _ptr(st.rows), st.n_pad, _ptr(st.diag), st.n_pad), "rpc_simple_run")
        A._count(k * st.n)
    for i in range(0 if not fast else k, k):
        if alg == "rp":
            sample_into(i, 1)
        else:
            if check_every_step or i == 0:
                sample_into(i, 0)               # python sum(diags) for the stoptol test (exact, like the reference's)
            pick_greedy(i)
        if check_every_step:
            sh = stats_all[i].cpu().numpy()
            if alg == "rp":
                _raise_if_invalid(sh)
            if i == 0:
                orig_trace = float(sh[0])
            elif float(sh[0]) <= stoptol * orig_trace:         # rpcholesky.py:45-48 (drops row i-1, keeps its index)
                count = i - 1
                arr_len = i
                break
        # algorithmic traffic of one step: stream G[:i], read X, write G[i] and rows[i], diag read + write
        with _timed("simple", 2.0 * st.n * (i + d), 8.0 * st.n * (i + d + 5)):
            check(lib.rpc_simple_step(ctx.handle, _stream(), spec_ref, _ptr(X), _ptr(xn), d, ldx, _ptr(Ad), lda, st.n,
                                      _ptr(idx_all), i, _ptr(st.G), st.n_pad, _ptr(st.rows), st.n_pad, _ptr(st.diag),
                                      st.n_pad), "rpc_simple_step")
        A._count(st.n)
    else:
        if check_every_step:
            sample_into(k, 0)
            sh = stats_all[k].cpu().numpy()
            if float(sh[0]) <= stoptol * orig_trace:
                count = k - 1
    if alg == "rp":
        stats_host = stats_all[:arr_len].cpu().numpy()
        for r in range(arr_len):
            _raise_if_invalid(stats_host[r])
    arr_idx = [int(v) for v in idx_all[:arr_len].cpu().numpy()]
    return st.result(count, arr_idx)


# --------------------------------------------------------------------------------------------------
# block RPCholesky: block_cholesky_helper(A, k, b, 'rp', strategy='regularize') -- rpcholesky.py:65-138
# --------------------------------------------------------------------------------------------------
@_on_operator_device
def block_cholesky_helper(A, k, b, alg, stoptol=1e-14, strategy="regularize", rbrp_atol=0.0, rbrp_rtol="1/b"):
    if alg not in ("rp", "greedy"):
        raise RuntimeError("Algorithm '{}' not recognized".format(alg))
    if strategy in ("regularize", "regularized"):
        rbrp = False
    elif strategy in ("rbrp", "pivoting", "pivoted"):
        rbrp = True
    else:
        raise ValueError("'{}' is not a valid strategy for block RPCholesky".format(strategy))
    A = _as_operator(A)
    if stoptol is None:
        stoptol = 1e-14
    b = _checked_block(b)
    if rbrp_rtol == "1/b":                                                    # rpcholesky.py:75-76
        rbrp_rtol = 1.0 / b
    use_tol = rbrp_rtol is not None and rbrp_atol is not None                 # rpcholesky.py:56
    st = _State(A, k, 2 * min(b, k))
    lib, ctx = st.lib, st.ctx
    scale = 2.0 * float(torch.max(rdist.gather_diag(st.info, st.diag, st.diag_full)))   # rpcholesky.py:72
    rng = np.random.default_rng()
    arr_idx = []
    counter = 0
    orig_trace = None
    final_sum = None
    count = k
    while counter < k:
        block_size = min(k - counter, b)
        if alg == "rp":
            up = rng.random(2 * block_size)                    # rng.choice(..., size=2*block_size) (:91)
            st.upload_uniforms(up)
            st.wait_uniforms()
            st.sample(2 * block_size)
        else:
            # np.argpartition(diags, -block_size)[-block_size:] (:94-95): the set of the block_size largest residuals
            st.sample(0)                                       # python sum(diags) for the stoptol test
            full = st.diag_full if st.info.world > 1 else st.diag   # gathered by st.sample
            check(lib.rpc_select_largest(ctx.handle, _stream(), _ptr(full), st.n, block_size, _ptr(st.idx_dev)),
                  "rpc_select_largest")
        ndraw = 2 * block_size if alg == "rp" else block_size
        stats_host, _, _, drawn = st.readback(ndraw, with_chol=False)                              # one sync
        if alg == "rp":
            _raise_if_invalid(stats_host)
        if orig_trace is None:
            orig_trace = float(stats_host[0])
        elif stoptol > 0 and float(stats_host[0]) <= stoptol * orig_trace:     # :133-136, checked after the previous block
            count = counter
            final_sum = float(stats_host[0])
            break
        idx = np.unique(drawn)[:block_size] if alg == "rp" else drawn          # :92
        block_size = len(idx)
        if rbrp:
            # :115-123: H = A[idx,idx] - G[:c,idx]^T G[:c,idx]; pivoted Cholesky; keep idx[piv[:rank]-1]
            st.sel_host[:block_size] = torch.from_numpy(idx)
            st.idx_dev[:block_size].copy_(st.sel_host[:block_size], non_blocking=True)
            st.gather(st.idx_dev, block_size, counter)
            st.residual_block(st.idx_dev, block_size, counter, count=True)
            check(lib.rpc_pivoted_cholesky(ctx.handle, _stream(), _ptr(st.H), block_size, st.m_max, 1 if use_tol else 0,
                                           float(rbrp_rtol) if use_tol else 0.0, float(rbrp_atol) if use_tol else 0.0,
                                           _ptr(st.L), st.m_max, _ptr(st.acc), _ptr(st.info_dev)), "rpc_pivoted_cholesky")
            _, info, acc_host, _ = st.readback(block_size)
            rank = int(info[0])
            if rank < 1:
                raise RuntimeError("pivoted Cholesky of the residual block found no positive pivot")
            idx = idx[acc_host[:rank]]
            arr_idx.extend(idx)
            st.update(counter, rank, st.acc, st.idx_dev)
            counter += rank
            continue
        arr_idx.extend(idx)
        st.sel_host[:block_size] = torch.from_numpy(idx)
        st.sel_dev[:block_size].copy_(st.sel_host[:block_size], non_blocking=True)
        st.gather(st.sel_dev, block_size, counter)
        st.residual_block(st.sel_dev, block_size, counter, count=False)       # C of :103
        Hcopy = st.H[:block_size, :block_size].clone()
        st.cholesky(block_size, 1, block_size, EPS * b * scale, 0)            # :106
        _, info, _, _ = st.readback(0)
        if info[1] != 0:                                                       # :109-114
            warn("Cholesky failed in block partial Cholesky. Falling back to eigendecomposition")
            evals, evecs = torch.linalg.eigh(Hcopy)
            pos = evals > 0
            scal = torch.where(pos, evals.clamp_min(1e-300) ** -0.5, torch.zeros_like(evals))
            scal = torch.where(evals < 0, torch.zeros_like(evals), scal)
            T = (scal[:, None] * evecs.T).contiguous()
            st.update(counter, block_size, None, st.sel_dev, panel_mode=1, Lmat=T)
        else:
            st.update(counter, block_size, None, st.sel_dev)
        counter += block_size
    else:
        if stoptol > 0:
            st.sample(0)
            sh = st.stats.cpu().numpy()
            final_sum = float(sh[0])
            if float(sh[0]) <= stoptol * orig_trace:
                count = counter
    res = st.result(min(count, counter), arr_idx)
    if final_sum is not None and min(count, counter) == counter:
        res._trace_hint = orig_trace - final_sum               # trace(A) - sum(residual diagonal) = ||G||_F^2
    return res


# --------------------------------------------------------------------------------------------------
# accelerated RPCholesky -- rpcholesky.py:159-232 (rejection_cholesky :140-157 runs on the device)
# --------------------------------------------------------------------------------------------------
@_on_operator_device
def accelerated_rpcholesky(A, k, b="auto", stoptol=1e-13, verbose=False):
    A = _as_operator(A)
    if stoptol is None:
        stoptol = 1e-13
    auto_b = False
    if b == "auto":                                            # :169-173
        b = int(np.ceil(k / 10))
        auto_b = True
    b = _checked_block(b)
    # the auto-tuned b may grow to ceil(k/3) (:216); the device kernels hold one block per CTA: capped at RPC_MAX_BLOCK
    b_cap = min(max(b, int(math.ceil(k / 3)), 10), _lib.MAX_BLOCK) if auto_b else b
    st = _State(A, k, b_cap)
    st.enable_speculation()
    rng = np.random.default_rng()
    arr_idx = np.zeros(k, dtype=int)
    counter = 0
    orig_trace = None
    count = k
    rounds = []
    final_sum = None
    # Row-sharded runs with a fixed b draw (and upload + broadcast) round t+1's uniforms at the start of round t.  The two numpy
    # streams are state independent, so the values are the ones the reference would draw; if round t+1 never happens the legacy
    # stream is put back where the reference leaves it.
    prefetch = st.info.world > 1 and not auto_b and USE_PREFETCH
    ahead = None                                               # (slot offset, legacy state before the draw) of the next round

    def draw_round(slot):
        up = rng.random(b)                                     # :184
        state = np.random.get_state()
        ur = np.random.random_sample(b)                        # the b np.random.rand() calls of :151
        return st.upload_uniforms(up, ur, slot), state

    slot = 0
    while counter < k:
        t0 = time.perf_counter()
        if ahead is None:
            u_off, legacy_state = draw_round(slot)
        else:
            u_off, legacy_state = ahead
        st.wait_uniforms(slot)
        if prefetch:
            ahead = draw_round(slot ^ 1)
        with _timed("sample", 0.0, 16.0 * st.n):
            st.sample(b, u_off)
        with _timed("gram", 2.0 * b * b * counter, 8.0 * counter * b):
            st.gather(st.idx_dev, b, counter)
            st.residual_block(st.idx_dev, b, counter)              # :189
        if st.spec:
            # side stream, forked behind the (many-CTA) Gram kernels: the 147-CTA evaluation then runs beside the
            # single-CTA rejection kernel only
            st.launch_eval_all(b)
        with _timed("reject", b ** 3 / 3.0, 8.0 * b * b):
            st.cholesky(b, 0, k - counter, 0.0, u_off + b)         # :190 + truncation :193-196
        st.publish(b)
        with _timed("panel", 1.0 * b * b * counter, 8.0 * counter * b):
            st.prelaunch_panel(counter, b)                         # runs while the host wakes up
        stats_host, info, acc_host, idx_host = st.collect(b)                        # the round's only sync
        _raise_if_invalid(stats_host)
        if orig_trace is None:
            orig_trace = float(stats_host[0])
        elif stoptol > 0 and float(stats_host[0]) <= stoptol * orig_trace:          # :224-227
            np.random.set_state(legacy_state)                  # this round never happens in the reference
            ahead = None
            A.queries -= b * b
            count = counter
            final_sum = float(stats_host[0])
            break
        if info[2] != 0:
            raise RuntimeError("rejection_cholesky requires a strictly positive trace")      # :145
        num_sel = int(info[0])
        accepted = acc_host[:num_sel]
        idx = idx_host[accepted]
        arr_idx[counter:counter + num_sel] = idx
        t1 = time.perf_counter()
        if num_sel > 0:
            st.update(counter, num_sel, st.acc, st.idx_dev, panel_done=True, spec_rows=st.spec)
        counter += num_sel
        slot ^= 1
        rounds.append(num_sel)
        if auto_b:                                             # :211-220, on device-synchronised wall time
            torch.cuda.synchronize(st.dev)
            t2 = time.perf_counter()
            rejection_time, process_time = t1 - t0, t2 - t1
            target = int(np.ceil(b * process_time / (4 * max(rejection_time, 1e-9))))
            b = max([min([target, int(np.ceil(1.5 * b)), int(np.ceil(k / 3))]), int(np.ceil(b / 3)), 10])
            b = min(b, b_cap)
            b = rdist.agree_on_int(st.info, b, st.dev)         # sharded: wall-clock differs per rank, rank 0 decides
        if verbose:
            print("Accepted {} / {}".format(num_sel, len(accepted)))
    else:
        if stoptol > 0:
            st.sample(0)
            sh = st.stats.cpu().numpy()
            final_sum = float(sh[0])
            if float(sh[0]) <= stoptol * orig_trace:
                count = counter
    if ahead is not None:
        np.random.set_state(ahead[1])                          # the prefetched round does not exist in the reference
    res = st.result(min(count, counter), arr_idx)
    if final_sum is not None and min(count, counter) == counter:
        res._trace_hint = orig_trace - final_sum               # trace(A) - sum(residual diagonal) = ||G||_F^2
    res.rounds = rounds
    return res


# --------------------------------------------------------------------------------------------------
# dispatch -- rpcholesky.py:234-257
# --------------------------------------------------------------------------------------------------
def rpcholesky(A, k, b=None, accelerated=True, **kwargs):
    if b is None:
        if accelerated:
            return accelerated_rpcholesky(A, k, **kwargs)
        return cholesky_helper(A, k, "rp", **kwargs)
    if accelerated:
        return accelerated_rpcholesky(A, k, b, **kwargs)
    return block_cholesky_helper(A, k, b, "rp", **kwargs)


def simple_rpcholesky(A, k, **kwargs):
    return rpcholesky(A, k, b=None, accelerated=False, **kwargs)


def block_rpcholesky(A, k, b=100, **kwargs):
    return rpcholesky(A, k, b=b, accelerated=False, **kwargs)


def greedy(A, k, randomized_tiebreaking=False, b=1, **kwargs):
    """rpcholesky.py:245-251.  b == 1: greedy / randomized-greedy pivoting through the simple loop."""
    if b == 1:
        return cholesky_helper(A, k, "rgreedy" if randomized_tiebreaking else "greedy", **kwargs)
    if randomized_tiebreaking:
        warn("Randomized tiebreaking not implemented for block greedy method")
    # np.argpartition's order inside the selected block (and its choice among entries tied at the threshold) is an
    # implementation detail of numpy's introselect: the device returns the same SET in ascending index order
    return block_cholesky_helper(A, k, b, "greedy", **kwargs)


def block_greedy(A, k, b=100, stoptol=None):
    """rpcholesky.py:259-260 (the reference's version raises NameError on an undefined `kwargs`; this is its intent)."""
    return block_cholesky_helper(A, k, b, "greedy", stoptol=stoptol)
