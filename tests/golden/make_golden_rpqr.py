#!/usr/bin/env python3
"""Golden fixtures of accelerated RPQR from the UNMODIFIED reference (rpqr.py), run under the replay patch.
    python tests/golden/make_golden_rpqr.py        # needs /root/reference (read-only); build container only"""
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

RPQR_CASES = {
    "rpqr_decay":  dict(m=80, n=400, k=24, b=8, seed=41, kind="decay"),
    "rpqr_gauss":  dict(m=120, n=300, k=40, b=12, seed=42, kind="gauss"),
    "rpqr_trunc":  dict(m=64, n=500, k=19, b=16, seed=43, kind="decay"),
}


def rpqr_inputs(spec):
    rng = np.random.default_rng(2000 + spec["seed"])
    m, n = spec["m"], spec["n"]
    if spec["kind"] == "gauss":
        return rng.standard_normal((m, n))
    r = min(m, n)
    U, _ = np.linalg.qr(rng.standard_normal((m, r)))
    V, _ = np.linalg.qr(rng.standard_normal((n, r)))
    return (U * (0.8 ** np.arange(r))) @ V.T


def main():
    sys.path.insert(0, "/root/reference")
    import rpqr as ref
    for name, spec in RPQR_CASES.items():
        B = rpqr_inputs(spec)
        seed = spec["seed"]
        with mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed))):
            np.random.seed(seed)
            Q, F, idx = ref.accelerated_rpqr(B, spec["k"], b=spec["b"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), Q=Q, F=F, idx=np.asarray(idx, dtype=np.int64))
        print(name, Q.shape, F.shape, idx[:6], np.linalg.norm(B - Q @ F.T) / np.linalg.norm(B))


if __name__ == "__main__":
    main()
