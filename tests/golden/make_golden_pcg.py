#!/usr/bin/env python3
"""Fixtures pinning Nystrom-preconditioned CG: the preconditioner against the reference's
CompactEigenvalueDecomposition.from_G(G).krr (lra.py:51-55, 78-80), and the WHOLE loop -- preconditioner, blocked mat-vec, CG
recurrences, stopping rule -- against the reference's own statements (block_experiments/pes_code.py:463-524), which are read
from /root/reference at generation time and executed unchanged on a dense kernel matrix of the reference's KernelMatrix.
      python tests/golden/make_golden_pcg.py     # needs /root/reference; build container only"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def pcg_precond_inputs():
    rng = np.random.default_rng(4242)
    G = rng.standard_normal((25, 300)) * (0.7 ** np.arange(25))[:, None]
    v = rng.standard_normal(300)
    return G, v, 1e-3


PCG_CASES = {
    # name: kernel, kw, bandwidth, N, d, k, b, mu, tol, max_iters, seed
    "pcg_loop_gauss": dict(kernel="gaussian", kw={}, bw=2.0, N=700, d=5, k=160, b=32, mu=1e-1, tol=1e-9, max_iters=60, seed=11),
    "pcg_loop_laplace": dict(kernel="laplace", kw={}, bw=6.0, N=520, d=6, k=60, b=12, mu=5e-2, tol=1e-8, max_iters=80, seed=12),
    "pcg_loop_noprecond": dict(kernel="matern", kw={"nu": 1.5}, bw=1.0, N=400, d=3, k=0, b=0, mu=2.0, tol=1e-8, max_iters=120, seed=13),
}


def pcg_loop_inputs(spec):
    rng = np.random.default_rng(spec["seed"])
    X = rng.standard_normal((spec["N"], spec["d"]))
    y = np.sin(X[:, 0]) + 0.3 * rng.standard_normal(spec["N"])
    return X, y


def reference_pcg_lines():
    """The Nystrom-PCG statements of the reference's experiment script, block_experiments/pes_code.py (the script as a whole
    needs MD17 data, `tables` and matplotlib and cannot be imported): from '# prepare preconditioner' to the end of the CG loop,
    dedented by the one level of the script's `for j` loop.  The text is read from /root/reference at generation time and
    executed unchanged; nothing of it is stored in this repository."""
    import textwrap
    src = open("/root/reference/block_experiments/pes_code.py").read().splitlines()
    start = next(i for i, ln in enumerate(src) if ln.strip() == "# prepare preconditioner")
    end = next(i for i, ln in enumerate(src) if i > start and ln.startswith("# close kernel matrix"))
    return textwrap.dedent("\n".join(src[start:end]))


def run_reference_pcg(spec):
    """Executes the reference's own lines on a dense kernel matrix and a factor from the reference's accelerated RPCholesky."""
    import contextlib
    import io
    import time
    from unittest import mock
    from scipy.linalg import qr
    from matrix import KernelMatrix
    from rpcholesky import accelerated_rpcholesky
    X, y = pcg_loop_inputs(spec)
    n = spec["N"]
    A = KernelMatrix(X, kernel=spec["kernel"], bandwidth=spec["bw"], **spec["kw"])
    if spec["k"] > 0:
        np.random.seed(spec["seed"])
        with mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(spec["seed"]))):
            lra = accelerated_rpcholesky(A, spec["k"], b=spec["b"])
        G, j = np.array(lra.G), 0                      # j <= 4: the script's Nystrom-preconditioned branches
    else:
        G, j = None, 5                                 # j == 5: no preconditioner
    K = A[:, :]
    mi = spec["max_iters"]
    ns = dict(np=np, time=time, qr=qr, j=j, G=G, mu=spec["mu"], tol=spec["tol"], max_iters=mi, n_train=n, Y_train=y, A=K,
              blocks=np.unique(np.append(np.arange(0, n, 256), n)), prep=np.zeros((6, 2)), resid=np.full((6, mi + 1), np.nan),
              times=np.zeros((6, mi + 1)), mae=np.zeros((6, mi + 1)), rmse=np.zeros((6, mi + 1)), ave=0.0, tall=np.zeros((1, n)),
              Y_test=np.zeros(1))
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(reference_pcg_lines(), "pes_code.py[pcg]", "exec"), ns)
    resid = ns["resid"][j]
    resid = resid[~np.isnan(resid)]
    probe = np.cos(np.arange(n) * 0.37)
    # (the factor itself is not stored: the oracle's accelerated_rpcholesky reproduces it from the same seed, tests/test_oracle.py)
    return dict(x=ns["x"], resid=resid, precond_probe=ns["precondition"](probe))


def main():
    sys.path.insert(0, "/root/reference")
    from lra import CompactEigenvalueDecomposition
    G, v, mu = pcg_precond_inputs()
    out = CompactEigenvalueDecomposition.from_G(G).krr(v, mu)
    np.savez_compressed(os.path.join(HERE, "pcg_precond.npz"), out=out)
    print("pcg_precond", out[:4])
    for name, spec in PCG_CASES.items():
        got = run_reference_pcg(spec)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **got)
        print(name, "iterations", len(got["resid"]) - 1, "final resid", got["resid"][-1])


if __name__ == "__main__":
    main()
