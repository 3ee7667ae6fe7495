#!/usr/bin/env python3
"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference seeds nothing, so it is run under the replay patch of SURVEY.md section 8(c):
`numpy.random.default_rng` -> Generator(PCG64(seed)) and `np.random.seed(seed)`.  Inputs are
regenerated from their seeds by the tests (`golden_inputs` below is imported by tests/), outputs
are stored in tests/golden/*.npz.  Nothing under tests/ reads /root/reference at test time.
"""
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (algorithm, operator spec, k, b, seed, extra kwargs)
CASES = {
    "simple_gauss":     dict(alg="simple", N=500, d=5, kernel="gaussian", bw="sqrtd", k=30, seed=11),
    "simple_laplace":   dict(alg="simple", N=400, d=7, kernel="laplace", bw="d", k=25, seed=12),
    "block_laplace":    dict(alg="block", N=800, d=8, kernel="laplace", bw="d", k=60, b=16, seed=13),
    "block_gauss":      dict(alg="block", N=600, d=6, kernel="gaussian", bw="sqrtd", k=50, b=20, seed=14),
    "accel_gauss":      dict(alg="accel", N=1000, d=10, kernel="gaussian", bw="sqrtd", k=80, b=24, seed=15),
    "accel_matern25":   dict(alg="accel", N=700, d=3, kernel="matern", nu=2.5, bw="sqrtd", k=50, b=15, seed=16),
    "accel_matern15":   dict(alg="accel", N=300, d=4, kernel="matern", nu=1.5, bw="sqrtd", k=30, b=12, seed=17),
    "accel_laplace":    dict(alg="accel", N=650, d=9, kernel="laplace", bw="d", k=64, b=20, seed=18),
    "accel_gauss_stab": dict(alg="accel", N=300, d=2, kernel="gaussian", bw=0.5, extra_stability=True, k=40, b=10, seed=19),
    "accel_truncate":   dict(alg="accel", N=400, d=6, kernel="gaussian", bw="sqrtd", k=33, b=32, seed=20),
    "simple_dense":     dict(alg="simple", N=120, dense_rank=40, k=20, seed=21),
    "block_dense":      dict(alg="block", N=150, dense_rank=60, k=30, b=8, seed=22),
    "accel_dense":      dict(alg="accel", N=150, dense_rank=60, k=30, b=10, seed=23),
    "accel_stoptol":    dict(alg="accel", N=200, dense_rank=12, k=40, b=10, seed=24),
    "block_stoptol":    dict(alg="block", N=200, dense_rank=12, k=40, b=10, seed=25),
    "simple_stoptol":   dict(alg="simple", N=100, dense_rank=6, k=20, seed=26, stoptol=1e-10),
    # greedy / randomized-greedy pivoting through the same loop (rpcholesky.py:32-35, 245-251)
    "greedy_gauss":     dict(alg="greedy", N=450, d=5, kernel="gaussian", bw="sqrtd", k=28, seed=27),
    "greedy_dense":     dict(alg="greedy", N=130, dense_rank=50, k=24, seed=28),
    "rgreedy_laplace":  dict(alg="rgreedy", N=380, d=6, kernel="laplace", bw="d", k=22, seed=29),
    # block strategies / block greedy (rpcholesky.py:94-95, 115-123): RBRP = pivoted Cholesky of the residual block
    "rbrp_gauss":       dict(alg="block", N=700, d=6, kernel="gaussian", bw="sqrtd", k=60, b=20, seed=30,
                             blockkw=dict(strategy="rbrp")),
    "rbrp_laplace":     dict(alg="block", N=500, d=5, kernel="laplace", bw="d", k=45, b=16, seed=31,
                             blockkw=dict(strategy="pivoting", rbrp_rtol=1e-3)),
    "rbrp_matern_wide": dict(alg="block", N=900, d=3, kernel="matern", nu=1.5, bw="sqrtd", k=150, b=70, seed=32,
                             blockkw=dict(strategy="rbrp", rbrp_rtol=1e-6, rbrp_atol=1e-9)),
    "rbrp_lowrank":     dict(alg="block", N=200, dense_rank=12, k=40, b=10, seed=33, blockkw=dict(strategy="rbrp")),
    "bgreedy_dense":    dict(alg="bgreedy", N=150, dense_rank=60, k=30, b=8, seed=34),
    "bgreedy_rbrp":     dict(alg="bgreedy", N=160, dense_rank=70, k=36, b=12, seed=35, blockkw=dict(strategy="rbrp")),
}

KERNEL_CASES = {
    "kern_gauss":        dict(kernel="gaussian", bw=1.7, d=6),
    "kern_gauss_stab":   dict(kernel="gaussian", bw=1.7, d=6, extra_stability=True),
    "kern_laplace":      dict(kernel="laplace", bw=4.0, d=9),
    "kern_matern05":     dict(kernel="matern", nu=0.5, bw=2.0, d=3),
    "kern_matern15":     dict(kernel="matern", nu=1.5, bw=2.0, d=3),
    "kern_matern25":     dict(kernel="matern", nu=2.5, bw=2.0, d=3),
    "kern_matern25_stab": dict(kernel="matern", nu=2.5, bw=2.0, d=3, extra_stability=True),
    "kern_gauss_d50":    dict(kernel="gaussian", bw=7.0710678118654755, d=50),
}


def bandwidth_of(spec):
    bw = spec["bw"]
    if bw == "sqrtd":
        return float(np.sqrt(spec["d"]))
    if bw == "d":
        return float(spec["d"])
    return float(bw)


def golden_inputs(spec):
    """Regenerate the input of a case from its seed (shared with the tests)."""
    rng = np.random.default_rng(1000 + spec["seed"])
    if "dense_rank" in spec:
        B = rng.standard_normal((spec["dense_rank"], spec["N"]))
        return {"dense": B.T @ B}
    X = rng.standard_normal((spec["N"], spec["d"]))
    kw = {k: spec[k] for k in ("nu", "extra_stability") if k in spec}
    return {"X": X, "kernel": spec["kernel"], "bandwidth": bandwidth_of(spec), "kw": kw}


def kernel_inputs(spec):
    rng = np.random.default_rng(77)
    X = rng.standard_normal((300, spec["d"]))
    X[5] = X[4]                       # exact duplicate point
    X[6] = X[4] + 1e-9                # near-duplicate (cancellation case)
    idx = np.array([4, 0, 5, 17, 6, 299, 42], dtype=np.int64)
    kw = {k: spec[k] for k in ("nu", "extra_stability") if k in spec}
    return X, idx, kw


def main():
    sys.path.insert(0, "/root/reference")
    import rpcholesky as ref
    from matrix import KernelMatrix

    def run(spec):
        inp = golden_inputs(spec)
        A = inp["dense"] if "dense" in inp else KernelMatrix(inp["X"], kernel=inp["kernel"], bandwidth=inp["bandwidth"], **inp["kw"])
        seed = spec["seed"]
        kw = {"stoptol": spec["stoptol"]} if "stoptol" in spec else {}
        with mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed))):
            np.random.seed(seed)
            if spec["alg"] in ("greedy", "rgreedy"):
                lra = ref.greedy(A, spec["k"], randomized_tiebreaking=spec["alg"] == "rgreedy", **kw)
            elif spec["alg"] == "simple":
                lra = ref.simple_rpcholesky(A, spec["k"], **kw)
            elif spec["alg"] == "block":
                lra = ref.block_rpcholesky(A, spec["k"], b=spec["b"], **kw, **spec.get("blockkw", {}))
            elif spec["alg"] == "bgreedy":
                lra = ref.greedy(A, spec["k"], b=spec["b"], **kw, **spec.get("blockkw", {}))
            else:
                lra = ref.accelerated_rpcholesky(A, spec["k"], b=spec["b"], **kw)
        out = {"idx": np.asarray(lra.idx, dtype=np.int64), "G": lra.G, "rows": lra.rows, "trace": np.float64(lra.trace())}
        if not isinstance(A, np.ndarray):
            out["queries"] = np.int64(A.num_queries())
        return out

    only = set(sys.argv[1:])
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        out = run(spec)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, out["G"].shape, out["idx"][:6])

    if only:
        return
    for name, spec in KERNEL_CASES.items():
        X, idx, kw = kernel_inputs(spec)
        A = KernelMatrix(X, kernel=spec["kernel"], bandwidth=spec["bw"], **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rows=A[idx, :], block=A[idx, idx], diag=A.diag(), one_row=A[int(idx[3]), :])
        print(name)

    # sampler fixtures straight from Generator.choice
    rng0 = np.random.default_rng(5)
    samp = {}
    for tag, n in (("uniform", 1000), ("random", 4097), ("sparse", 513)):
        if tag == "uniform":
            diag = np.ones(n)
        elif tag == "random":
            diag = rng0.random(n) ** 4
        else:
            diag = rng0.random(n) * (rng0.random(n) < 0.1)
        g = np.random.Generator(np.random.PCG64(99))
        idx = g.choice(range(n), size=257, p=diag / sum(diag), replace=True)
        u = np.random.Generator(np.random.PCG64(99)).random(257)
        samp[tag + "_diag"] = diag; samp[tag + "_u"] = u; samp[tag + "_idx"] = idx.astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **samp)
    print("sampler")


if __name__ == "__main__":
    main()
