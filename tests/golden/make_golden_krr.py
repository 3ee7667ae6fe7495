#!/usr/bin/env python3
"""Golden fixtures for the first "next" row: KRR_Nystrom.fit_Nystrom of the UNMODIFIED reference
(KRR_Nystrom.py:34-58) driven by its own RPCholesky under the replay patch, plus the evident intent of the broken
predict_Nystrom (KRR_Nystrom.py:60-66 references an undefined `KtM`): NonsymmetricKernelMatrix(Xts, Xtr[S])[:,:] @ sol.

    python tests/golden/make_golden_krr.py          # needs /root/reference (build container only)
"""
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

KRR_CASES = {
    "krr_laplace_accel": dict(kernel="laplace", bw=6.0, N=600, d=6, nt=90, k=40, b=12, alg="accel", lamb=1e-4, seed=51),
    "krr_gauss_block":   dict(kernel="gaussian", bw=2.0, N=500, d=4, nt=70, k=36, b=9, alg="block", lamb=1e-5, seed=52),
}


def krr_inputs(spec):
    rng = np.random.default_rng(2000 + spec["seed"])
    Xtr = rng.standard_normal((spec["N"], spec["d"]))
    Xts = rng.standard_normal((spec["nt"], spec["d"]))
    w = rng.standard_normal(spec["d"])
    f = lambda X: np.sin(X @ w) + 0.1 * (X ** 2).sum(1)
    return Xtr, f(Xtr), Xts, f(Xts)


def main():
    sys.path.insert(0, "/root/reference")
    import rpcholesky as ref
    from KRR_Nystrom import KRR_Nystrom
    from matrix import NonsymmetricKernelMatrix
    for name, spec in KRR_CASES.items():
        Xtr, ytr, Xts, yts = krr_inputs(spec)
        if spec["alg"] == "accel":
            method = lambda A, k: ref.accelerated_rpcholesky(A, k, b=spec["b"])
        else:
            method = lambda A, k: ref.block_rpcholesky(A, k, b=spec["b"])
        model = KRR_Nystrom(kernel=spec["kernel"], bandwidth=spec["bw"])
        seed = spec["seed"]
        with mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed))):
            np.random.seed(seed)
            model.fit_Nystrom(Xtr, ytr, spec["lamb"], spec["k"], method)
        S = np.asarray(model.sample_idx, dtype=np.int64)
        Kp = NonsymmetricKernelMatrix(Xts, Xtr[S, :], kernel=spec["kernel"], bandwidth=spec["bw"])
        preds = Kp[:, :] @ model.sol
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sample_idx=S, sol=model.sol, preds=preds,
                            reltrace_err=np.float64(model.reltrace_err), queries=np.int64(model.queries))
        print(name, S[:6], float(model.reltrace_err), float(np.abs(preds - yts).mean()))


if __name__ == "__main__":
    main()
