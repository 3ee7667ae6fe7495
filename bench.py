#!/usr/bin/env python3
"""bench.py -- rank-k RPCholesky wall time and kernel entries/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tstar|c3|c2|c1] [--impl reference]

A "step" is one complete factorisation (one pass of the hot path) of the named workload on synthetic
points X = default_rng(0).standard_normal((N, d)); every step replays the same seeded uniforms
(`numpy.random.default_rng` -> PCG64(1234), `np.random.seed(1234)`: the parity harness of SURVEY.md 8c).
Prints ONE JSON line (rank 0).  `value` = A.num_queries() / wall with X resident in HBM; `e2e` = the same
through the public API starting from a pinned HOST copy of X (H2D inside the timed region) and reading
the pivots and the trace-norm error back; `roofline` = the fused Schur-update DMMA kernel against the
measured FP64 peak; `cpu_baseline` = the numpy oracle on a bounded sample on this box's host cores.
`--impl reference` times the reference's own CPU implementation of the path: the UNMODIFIED reference modules copied
into oracle/_ref/ by oracle/make_ref.py (kind "reference"; git-ignored, travels with gpurun) under the replay patch, or,
when that copy is absent, the numpy port oracle/rpc_oracle.py (kind "port", pinned to the reference by tests/golden).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from unittest import mock

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: alg, kernel, N, d, k, b, bandwidth rule
    "tstar": dict(alg="accelerated", kernel="gaussian", N=1_000_000, d=50, k=2000, b=200, bw="sqrtd",
                  text="T*: accelerated_rpcholesky b=200, Gaussian sigma=sqrt(d), N=1e6 d=50 k=2000, fp64"),
    "c3": dict(alg="block", kernel="laplace", N=1_000_000, d=50, k=2000, b=200, bw="d",
               text="configs[2]: block_rpcholesky b=200, Laplace sigma=d, N=1e6 d=50 k=2000, fp64"),
    "c2": dict(alg="accelerated", kernel="gaussian", N=100_000, d=20, k=1000, b=120, bw="sqrtd",
               text="configs[1]: accelerated_rpcholesky b=120, Gaussian sigma=sqrt(d), N=1e5 d=20 k=1000, fp64"),
    "c4": dict(alg="accelerated", kernel="matern", kw={"nu": 2.5}, N=4_000_000, d=3, k=4000, b=400, bw="sqrtd",
               text="configs[3]: accelerated_rpcholesky b=400, Matern-5/2 sigma=sqrt(d), N=4e6 d=3 k=4000, fp64 "
                    "(256 GB of factors: needs >= 2 GPUs, meant for 8)"),
    "c5": dict(alg="krr", kernel="laplace", N=100_000, d=750, k=1000, b=100, bw=5120.0, nt=10_000, lamb=1e-8,
               text="configs[4]: KRR_Nystrom (accelerated_rpcholesky b=100 + normal-equation solve + prediction of 1e4 points), "
                    "Laplace sigma=5120, N=1e5 d=750 k=1000, synthetic QM9-shaped features, fp64"),
    "c1": dict(alg="simple", kernel="gaussian", N=10_000, d=10, k=100, b=None, bw="sqrtd",
               text="configs[0]: simple_rpcholesky, Gaussian sigma=sqrt(d), N=1e4 d=10 k=100, fp64"),
}
SEED = 1234
FP64_PEAK_FALLBACK_TFLOPS = 36.9        # profiles/fp64_peaks_r01.json (measured DMMA stream on this pool)


def bandwidth(w):
    if isinstance(w["bw"], float):
        return w["bw"]
    return float(np.sqrt(w["d"])) if w["bw"] == "sqrtd" else float(w["d"])


def krr_targets(X):
    """Synthetic regression target for the KRR workload (smooth function of a few features)."""
    return np.sin(X[:, :8].sum(1)) + 0.1 * (X[:, 8:16] ** 2).sum(1)


def replay(seed=SEED):
    np.random.seed(seed)
    return mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed)))


def make_points(n, d):
    return np.random.default_rng(0).standard_normal((n, d))


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (numpy/OpenBLAS with all host threads)
# ------------------------------------------------------------------------------------------------
_REF = None


def reference_modules():
    """The unmodified reference from oracle/_ref (oracle/make_ref.py), or None."""
    global _REF
    if _REF is None:
        try:
            from oracle import make_ref
            _REF = make_ref.load() or False
        except Exception:
            _REF = False
    return _REF or None


def cpu_kind(w):
    return "reference" if (reference_modules() and w["alg"] != "krr") else "port"


def run_reference(w, n_sample):
    """One factorisation by the UNMODIFIED reference (oracle/_ref) on the replayed draws: (queries, seconds)."""
    ref = reference_modules()
    X = make_points(n_sample, w["d"])
    A = ref["matrix"].KernelMatrix(X, kernel=w["kernel"], bandwidth=bandwidth(w), **w.get("kw", {}))
    drv = ref["rpcholesky"]
    with replay():
        t0 = time.perf_counter()
        if w["alg"] == "accelerated":
            drv.accelerated_rpcholesky(A, w["k"], b=w["b"])
        elif w["alg"] == "block":
            drv.block_rpcholesky(A, w["k"], b=w["b"])
        else:
            drv.simple_rpcholesky(A, w["k"])
        dt = time.perf_counter() - t0
    return A.num_queries(), dt


def run_cpu(w, n_sample):
    return run_reference(w, n_sample) if cpu_kind(w) == "reference" else run_oracle(w, n_sample)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def run_oracle(w, n_sample):
    from oracle import rpc_oracle as orc
    X = make_points(n_sample, w["d"])
    A = orc.OracleKernelMatrix(X, kernel=w["kernel"], bandwidth=bandwidth(w), **w.get("kw", {}))
    rng, legacy = orc.make_streams(SEED)
    if w["alg"] == "krr":
        from oracle import krr_oracle
        nt = max(100, w["nt"] * n_sample // w["N"])
        Xts = np.random.default_rng(1).standard_normal((nt, w["d"]))
        t0 = time.perf_counter()
        sol, idx, _, q = krr_oracle.fit_nystrom(X, krr_targets(X), w["lamb"], w["k"], w["kernel"], bandwidth(w), "accel", w["b"], SEED)
        krr_oracle.predict_nystrom(Xts, X, idx, sol, w["kernel"], bandwidth(w))
        return q, time.perf_counter() - t0
    t0 = time.perf_counter()
    if w["alg"] == "accelerated":
        orc.accelerated_rpcholesky(A, w["k"], w["b"], rng, legacy)
    elif w["alg"] == "block":
        orc.block_rpcholesky(A, w["k"], w["b"], rng)
    else:
        orc.simple_rpcholesky(A, w["k"], rng)
    dt = time.perf_counter() - t0
    return A.num_queries(), dt


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def sample_size(w, target_seconds):
    """Points for a CPU run of about `target_seconds`: a short probe of the same workload is timed and scaled (the cost
    is linear in N); BLAS-bound and cdist-bound kernels differ by 100x, so a fixed guess does not work."""
    n0 = int(min(w["N"], max(2 * w["k"], 4000)))
    run_cpu(w, n0)                               # imports, BLAS thread pool start-up
    _, t0 = run_cpu(w, n0)
    n = int(min(w["N"], max(n0, n0 * target_seconds / max(t0, 1e-3))))
    return max(n0, n // 1000 * 1000)


def reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    n_s = sample_size(w, 8.0)
    for _ in range(args.warmup):
        run_cpu(w, n_s)
    t_tot, q_tot = 0.0, 0
    for _ in range(args.steps):
        q, dt = run_cpu(w, n_s)
        t_tot += dt
        q_tot += q
    value = q_tot / t_tot
    kind = cpu_kind(w)
    sample = "same workload at N=%d (k, b, d, kernel unchanged; cost is linear in N); %s" % (
        n_s, "unmodified reference modules (oracle/_ref)" if kind == "reference" else "numpy port oracle/rpc_oracle.py")
    out = {
        "impl": "reference", "metric": "kernel entries/s (rank-k RPCholesky, A.num_queries()/wall)", "value": value,
        "unit": "entries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": w["text"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "entries/s", "cores": cpu_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


PLANE_TEXT = {"peer": "the producing kernels' stores into NVLink peer windows (CUDA IPC) + release/acquire epochs, no NCCL in the data path",
              "nccl": "NCCL all-gather + exact sum all-reduce (peer windows unavailable or RPC_B200_PEER=0)"}


def schur_executed_flops(c, s, n_cols, tri=True):
    """Flops the Schur kernel's DMMAs execute for one update of s rows on c previous rows over n_cols columns: 8-row
    fragments x 4-deep k steps, rows padded to the tile (csrc/big_kernels.cu tile_rows_for), and -- with the triangular skip
    -- only the fragments at or below the diagonal of the L^{-T} block (csrc/dmma_gemm.cuh consume_stage_tri)."""
    total = 0
    for m0 in range(0, s, 256):
        m = min(256, s - m0)
        mt = (m + 7) // 8 * 8 if m > 128 else next(t for t in (16, 32, 64, 96, 128) if m <= t)
        nfrag = mt // 8
        for kk0 in range(0, (c + s + 3) // 4 * 4, 4):
            if tri and kk0 - c > 7:
                first = max(0, -(-(kk0 - c - 7 - m0) // 8))        # first fragment whose last row reaches panel row kk0
                total += max(0, nfrag - first) * 8 * 4 * 2
            else:
                total += nfrag * 8 * 4 * 2
    return float(total) * n_cols


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[6]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class _KrrResult:
    """Adapter so that the KRR workload looks like a factorisation to the timing loop."""

    def __init__(self, model, preds, k):
        self.model, self.preds, self.idx, self._k = model, preds, np.asarray(model.sample_idx), k

    def rank(self):
        return self._k

    def trace(self):
        return float("nan")


def run_krr(rpc, X, y, Xts, w, dev, sharded):
    model = rpc.KRR_Nystrom(kernel=w["kernel"], bandwidth=bandwidth(w))
    with replay():
        model.fit_Nystrom(X, y, w["lamb"], w["k"], lambda A, k: rpc.accelerated_rpcholesky(A, k, b=w["b"]), sharded=sharded)
    preds = model.predict_Nystrom(Xts)
    return model, preds


def factorise(rpc, A, w):
    with replay():
        if w["alg"] == "accelerated":
            return rpc.accelerated_rpcholesky(A, w["k"], b=w["b"])
        if w["alg"] == "block":
            return rpc.block_rpcholesky(A, w["k"], b=w["b"])
        return rpc.simple_rpcholesky(A, w["k"])


def gpu_arm(args, w):
    import torch
    import torch.distributed as dist
    import randomly_pivoted_cholesky_b200 as rpc
    import importlib
    drv = importlib.import_module("randomly_pivoted_cholesky_b200.rpcholesky")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries the ONE JSON line: anything a library prints there (NCCL's "NCCL version ..." banner when NCCL_DEBUG is set in
    # the environment, compiler chatter) goes to stderr -- file descriptor 1 is pointed at stderr for the run and the line is
    # written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rpc.load()

    # N > 1: the SAME problem row-sharded over the ranks (strong scaling); every rank holds N/world points
    n, d = w["N"], w["d"]
    X = make_points(n, d)
    Xpin = torch.from_numpy(X).pin_memory()
    sharded = world > 1
    is_krr = w["alg"] == "krr"
    if is_krr:
        # the reference API (KRR_Nystrom.fit_Nystrom(Xtr, Ytr, ...)) takes HOST arrays and builds the operator itself,
        # so this workload has no "inputs resident" variant: value and e2e are the same measurement
        y = krr_targets(X)
        Xts = torch.from_numpy(np.random.default_rng(1).standard_normal((w["nt"], d))).pin_memory()

        class _KrrOperator:                      # what the timing loop needs from `A`
            q = 0
            def reset(self): self.q = 0
            def num_queries(self): return self.q
        A = _KrrOperator()

        def step_fn(rpc_, A_, w_):               # one step = fit + predict
            model, preds = run_krr(rpc_, Xpin, y, Xts, w_, dev, sharded)       # pinned host inputs
            A_.q = model.queries
            return _KrrResult(model, preds, w_["k"])
    else:
        need_gb = 16e-9 * w["k"] * n / world
        if need_gb > 150:
            raise SystemExit("workload %s needs %.0f GB of factors per GPU at %d GPU(s): run it with more GPUs" % (args.workload, need_gb, world))
        A = rpc.KernelMatrix(Xpin, kernel=w["kernel"], bandwidth=bandwidth(w), device=dev, sharded=sharded, **w.get("kw", {}))
        step_fn = factorise

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        A.reset()
        lra = step_fn(rpc, A, w)
        del lra
    sync_all()

    # ---- timed region: K steps, inputs resident -------------------------------------------------
    from randomly_pivoted_cholesky_b200 import _lib
    clocks = ClockSampler(local)
    clocks.start()
    drv.KERNEL_TIMERS = []
    l0 = _lib.total_launches()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    queries = 0
    for _ in range(args.steps):
        A.reset()
        lra = step_fn(rpc, A, w)
        queries += A.num_queries()
        rank_k = lra.rank()
        rounds = getattr(lra, "rounds", None)
        del lra
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    data_plane = getattr(drv, "LAST_DATA_PLANE", "single")      # what the timed factorisations actually ran on
    launches = _lib.total_launches() - l0
    timers = drv.KERNEL_TIMERS
    drv.KERNEL_TIMERS = None
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = queries / (ms * 1e-3)          # A.num_queries() counts the entries of the WHOLE (global) matrix

    # ---- roofline of the dominant kernel (fused Schur update), from the CUDA events of the timed region
    phases = {}
    for ev0, ev1, phase, flops, nbytes, *_ in timers:
        p = phases.setdefault(phase, [0.0, 0.0, 0.0, 0])
        p[0] += ev0.elapsed_time(ev1) * 1e-3
        p[1] += flops
        p[2] += nbytes
        p[3] += 1
    # FP64 roofline denominator: the DMMA stream peak measured NOW on this device (rpc_probe_fp64_peak, best of 5 launches of
    # ~10 ms, after the timed region); the tracked round-1 figure is only the cross-check / fallback
    tracked_tf = FP64_PEAK_FALLBACK_TFLOPS
    try:
        tracked_tf = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))["fp64_peak_tflops"]
    except Exception:
        pass
    try:
        from randomly_pivoted_cholesky_b200 import ops
        peak_tf = float(ops.fp64_peak_tflops(dev))
        peak_src = "DMMA.8x8x4 stream measured in this run on this device (rpc_probe_fp64_peak: %d CTAs x 8 warps x 16 accumulators, best of 5); " \
                   "cross-check profiles/fp64_peaks_r01.json = %.1f" % (torch.cuda.get_device_properties(dev).multi_processor_count, tracked_tf)
    except Exception as exc:                      # pragma: no cover
        peak_tf, peak_src = tracked_tf, "profiles/fp64_peaks_r01.json (in-run probe failed: %s)" % exc
    hbm = None
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    roofline = None
    traffic = None
    if args.workload == "tstar" and world == 1:
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this command
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))["schur"]["avg_dram_bytes_per_launch"]
        except Exception:
            traffic = None
    if "schur" in phases:
        sec, flops, nbytes, cnt = phases["schur"]
        ach = flops / sec / 1e12
        # flops the DMMA pipe actually executes: row-tile and k padding, and the fragment staircase along the diagonal of the
        # triangular block (whose zero half is skipped) -- from the launch shapes of the timed region
        executed = sum(schur_executed_flops(t[5][0], t[5][1], t[5][2]) for t in timers if t[2] == "schur" and len(t) > 5 and t[5])
        roofline = {"bound": "tensor", "kernel": "dg::gemm_kernel<MODE_SCHUR> (FP64 DMMA: Schur product + triangular solve + diag update fused)",
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                    "traffic_source": "profiles/traffic_r02.json (ncu dram__bytes_read+write over the 12 Schur launches of this command)" if traffic else None,
                    "peak_source": peak_src, "peak_tracked_r01": tracked_tf,
                    "launches": cnt, "avg_launch_ms": 1e3 * sec / cnt, "share_of_step": sec / (ms * 1e-3),
                    "algorithmic_flops_per_launch": flops / cnt, "algorithmic_bytes_per_launch": nbytes / cnt,
                    "algorithmic_flops_rule": "SURVEY 8(d): s (2c + s) N_loc per launch = Schur product 2 s c N + triangular solve s^2 N",
                    "executed_flops_per_launch": executed / cnt if executed else None,
                    "executed_tflops": executed / sec / 1e12 if executed else None,
                    "hbm_gbs_achieved": nbytes / sec / 1e9, "hbm_peak_gbs": hbm}
    elif "simple" in phases:
        sec, flops, nbytes, cnt = phases["simple"]
        ach = nbytes / sec / 1e9
        roofline = {"bound": "hbm", "kernel": "simple_step_kernel (fused GEMV over G[:i] + kernel row + diag update) + exact sampler, whole k-step loop per launch bracket",
                    "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm if hbm else None, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs", "launches": cnt, "avg_launch_ms": 1e3 * sec / cnt,
                    "share_of_step": sec / (ms * 1e-3), "algorithmic_bytes_per_launch": nbytes / cnt}
    phase_ms = {k: 1e3 * v[0] / args.steps for k, v in phases.items()}
    # GPU idle time between the end of the rejection kernel and the start of the panel solve: the round's host sync
    gap = 0.0
    for a, b2 in zip(timers, timers[1:]):
        if a[2] == "reject" and b2[2] == "panel":
            gap += a[1].elapsed_time(b2[0])
    phase_ms["host_sync_gap"] = gap / args.steps
    # per-launch detail of the dominant kernel over the LAST timed step: (ms, TFLOP/s)
    per_step = max(1, len([1 for t in timers if t[2] == "schur"]) // args.steps)
    detail = [[round(t[0].elapsed_time(t[1]), 3), round(t[3] / t[0].elapsed_time(t[1]) / 1e9, 2)]
              for t in timers if t[2] == "schur"][-per_step:] if "schur" in phases else []

    # ---- e2e: host X -> H2D -> factorise -> D2H pivots + trace-norm error, through the public API ----
    q2 = 0
    d2h = 0
    e2 = torch.cuda.Event(enable_timing=True)
    e3 = torch.cuda.Event(enable_timing=True)
    for step in range(-min(args.warmup, 1), args.steps):        # one untimed e2e step first (allocator, first touch)
        if step == 0:
            sync_all()
            q2 = 0
            e2.record()
        if is_krr:
            A.reset()
            lra = step_fn(rpc, A, w)
            idx, err = lra.idx, lra.model.reltrace_err
            q2 += A.num_queries()
            d2h = idx.nbytes + 8 + lra.preds.nbytes + lra.model.sol.nbytes
            del lra
            continue
        A2 = rpc.KernelMatrix(Xpin, kernel=w["kernel"], bandwidth=bandwidth(w), device=dev, sharded=sharded, **w.get("kw", {}))
        lra = step_fn(rpc, A2, w)
        idx = np.asarray(lra.idx)
        err = (n - lra.trace()) / n                      # utils.approximation_error with trace(A) = N
        q2 += A2.num_queries()
        d2h = idx.nbytes + 8
        del lra, A2
    e3.record()
    sync_all()
    ms2 = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms2], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = float(t.item())
    e2e = {"value": q2 / (ms2 * 1e-3), "unit": "entries/s",
           "h2d_bytes_per_step": int(X.nbytes + (y.nbytes + Xts.numel() * 8 if is_krr else 0)),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": ms2 / args.steps, "trace_norm_error": float(err),
           "note": "factors G, rows stay in HBM (PSDLowRank is device backed); result read back = pivots + trace error "
                   "(trace(A) - sum of the residual diagonal, the driver's exact account: no pass over G)"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            use_all_host_threads()
            n_s = sample_size(w, 12.0)
            q, dt = run_cpu(w, n_s)
            cpu = {"value": q / dt, "unit": "entries/s", "cores": cpu_threads(), "kind": cpu_kind(w),
                   "sample": "same workload at N=%d, one run (%.1f s); cost is linear in N" % (n_s, dt)}
        out = {
            "metric": "kernel entries/s (rank-k RPCholesky, A.num_queries()/wall)", "value": value, "unit": "entries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["text"], "rank": rank_k, "accepted_per_round": rounds, "l2": "inputs (G: %.1f GB) larger than L2" % (8e-9 * w["k"] * n),
                       "data_plane": data_plane,
                       "parallelism": ("row-sharded x%d (N/%d points per GPU; per round: residual diagonal + pivot payload exchanged by %s; "
                                       "the round's uniforms broadcast from rank 0)" % (world, world, PLANE_TEXT.get(data_plane, data_plane)))
                       if world > 1 else "single GPU"},
            "wall_time_s": ms * 1e-3 / args.steps, "phase_ms_per_step": phase_ms, "dominant_kernel_launches_ms_tflops": detail,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="tstar", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, w)
    else:
        gpu_arm(args, w)


if __name__ == "__main__":
    main()
