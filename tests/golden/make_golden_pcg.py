#!/usr/bin/env python3
"""Fixture pinning the Nystrom preconditioner to the reference's CompactEigenvalueDecomposition.from_G(G).krr
(lra.py:51-55, 78-80).      python tests/golden/make_golden_pcg.py     # needs /root/reference; build container only"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def pcg_precond_inputs():
    rng = np.random.default_rng(4242)
    G = rng.standard_normal((25, 300)) * (0.7 ** np.arange(25))[:, None]
    v = rng.standard_normal(300)
    return G, v, 1e-3


def main():
    sys.path.insert(0, "/root/reference")
    from lra import CompactEigenvalueDecomposition
    G, v, mu = pcg_precond_inputs()
    out = CompactEigenvalueDecomposition.from_G(G).krr(v, mu)
    np.savez_compressed(os.path.join(HERE, "pcg_precond.npz"), out=out)
    print("pcg_precond", out[:4])


if __name__ == "__main__":
    main()
