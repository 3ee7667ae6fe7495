"""B200-native RPCholesky hot path: a drop-in for the reference's rpcholesky.py / matrix.py / lra.py.

    from randomly_pivoted_cholesky_b200 import KernelMatrix, rpcholesky
    A = KernelMatrix(X, kernel="gaussian", bandwidth=1.0)
    lra = rpcholesky(A, k)            # accelerated RPCholesky, PSDLowRank result

All numerical work is done by hand-written sm_100a CUDA kernels behind the C ABI in include/rpc_b200.h
(librpc_b200.so).  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .lra import AbstractPSDLowRank, CompactEigenvalueDecomposition, NystromExtension, PSDLowRank
from .matrix import AbstractPSDMatrix, KernelMatrix, NonsymmetricKernelMatrix, PSDMatrix
from .KRR_Nystrom import KRR_Nystrom
from .rpcholesky import (accelerated_rpcholesky, block_cholesky_helper, block_greedy, block_rpcholesky, cholesky_helper,
                         greedy, rpcholesky, simple_rpcholesky)
from .rpqr import accelerated_rpqr, rpqr
from ._lib import build, load, total_launches

__all__ = [
    "AbstractPSDLowRank", "CompactEigenvalueDecomposition", "NystromExtension", "PSDLowRank", "AbstractPSDMatrix", "KernelMatrix",
    "PSDMatrix", "NonsymmetricKernelMatrix", "KRR_Nystrom", "accelerated_rpcholesky", "block_cholesky_helper", "block_rpcholesky", "cholesky_helper", "greedy", "block_greedy", "accelerated_rpqr", "rpqr",
    "rpcholesky", "simple_rpcholesky", "build", "load", "total_launches",
]
