"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol declared in
include/rpc_b200.h, the PSDLowRank container keeps the reference's semantics (modelled on the reference's
tests/test_lra.py), the operator classes refuse to compute without a GPU, and bench.py's reference arm runs."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from randomly_pivoted_cholesky_b200 import _lib
    header = open(os.path.join(ROOT, "include", "rpc_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rpc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    path = _lib.build()                      # nvcc cross-compiles without a GPU; no-op when up to date
    lib = ctypes.CDLL(path)
    for name in sorted(declared):
        assert hasattr(lib, name), "include/rpc_b200.h declares %s but librpc_b200.so does not export it" % name
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    loaded = _lib.load()
    assert loaded.rpc_abi_version() == 1     # no device needed
    assert loaded.rpc_panel_ld(200) % 16 == 4 and loaded.rpc_panel_ld(200) >= 200
    assert loaded.rpc_last_error() is not None
    # the ctypes mirrors of the structs that cross the ABI have the C layout
    assert loaded.rpc_abi_sizeof(0) == ctypes.sizeof(_lib.KernelSpec)
    assert loaded.rpc_abi_sizeof(1) == ctypes.sizeof(_lib.Peers)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import randomly_pivoted_cholesky_b200 as rpc
    with pytest.raises(RuntimeError):
        rpc.KernelMatrix(np.zeros((8, 2)))
    with pytest.raises(RuntimeError):
        rpc.rpcholesky(np.eye(4), 2)


def test_driver_argument_errors_match_the_reference():
    """Errors raised before any device work (rpcholesky.py:37, 97, 126 of the reference): same types, same messages."""
    import randomly_pivoted_cholesky_b200 as rpc
    with pytest.raises(RuntimeError, match="Algorithm 'nope' not recognized"):
        rpc.block_cholesky_helper(np.eye(4), 2, 2, "nope")
    with pytest.raises(ValueError, match="'nope' is not a valid strategy for block RPCholesky"):
        rpc.block_cholesky_helper(np.eye(4), 2, 2, "rp", strategy="nope")
    with pytest.raises(RuntimeError, match="Algorithm 'nope' not recognized"):
        rpc.cholesky_helper(np.eye(4), 2, "nope")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "randomly_pivoted_cholesky_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "rpc_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


@pytest.fixture
def lowrank_case():
    rng = np.random.default_rng(42)
    N, r = 50, 10
    G = rng.standard_normal((r, N))
    return rng, N, r, G, G.T @ G


def test_psdlowrank_container(lowrank_case):
    from randomly_pivoted_cholesky_b200 import CompactEigenvalueDecomposition, PSDLowRank
    rng, N, r, G, A = lowrank_case
    obj = PSDLowRank(G)
    assert obj.shape == (N, N) and obj.rank() == r
    assert obj.trace() == pytest.approx(np.linalg.norm(G, "fro") ** 2)
    x = rng.standard_normal((N, 5))
    np.testing.assert_allclose(obj @ x, A @ x)
    np.testing.assert_allclose(obj.matrix(), A)
    ced = obj.eigenvalue_decomposition()
    assert isinstance(ced, CompactEigenvalueDecomposition)
    np.testing.assert_allclose(ced.matrix(), A, atol=1e-10)
    assert ced.trace() == pytest.approx(np.trace(A))
    b = rng.standard_normal(N)
    np.testing.assert_allclose(ced.krr(b, 0.1), np.linalg.solve(A + 0.1 * np.eye(N), b), atol=1e-5)
    scaling = np.ones(N) * 2.0
    Gs = G * scaling[None, :]
    np.testing.assert_allclose(obj.scale(scaling).matrix(), Gs.T @ Gs)
    np.testing.assert_array_equal(obj.get_left_factor(), G.T)
    np.testing.assert_array_equal(obj.get_right_factor(), G)


def test_psdlowrank_metadata(lowrank_case):
    from randomly_pivoted_cholesky_b200 import PSDLowRank
    _, N, r, G, A = lowrank_case
    obj = PSDLowRank(G)
    with pytest.raises(RuntimeError):
        obj.get_indices()
    with pytest.raises(RuntimeError):
        obj.get_rows()
    idx = list(range(r))
    obj = PSDLowRank(G, idx=idx, rows=A[idx, :])
    assert obj.get_indices() == idx
    np.testing.assert_array_equal(obj.get_rows(), A[idx, :])
    assert obj.G is G and obj.G_device is None and obj.rows_device is None


def test_bench_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    # kind "reference" when oracle/make_ref.py's copy of the unmodified reference is present, else the numpy port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "rpcholesky.py"))
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "entries/s"
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)       # every host core, whatever OMP_NUM_THREADS said


def test_reference_copy_is_unmodified_and_agrees_with_the_port():
    """oracle/_ref (git-ignored, made by oracle/make_ref.py) is a byte-for-byte copy of the reference's modules, and the numpy
    port reproduces it on replayed draws: the two CPU arms of bench.py measure the same computation."""
    import hashlib
    from unittest import mock
    sys.path.insert(0, ROOT)
    from oracle import make_ref, rpc_oracle as orc
    if not os.path.isdir(make_ref.REF_ROOT):
        pytest.skip("/root/reference is not present here")
    assert make_ref.make(verbose=False)
    man = json.load(open(os.path.join(make_ref.DEST, "MANIFEST.json")))
    for name, digest in man["sha256"].items():
        assert hashlib.sha256(open(os.path.join(make_ref.DEST, name), "rb").read()).hexdigest() == digest
        assert open(os.path.join(make_ref.DEST, name), "rb").read() == open(os.path.join(make_ref.REF_ROOT, name), "rb").read()
    ref = make_ref.load()
    X = np.random.default_rng(0).standard_normal((3000, 6))
    A = ref["matrix"].KernelMatrix(X, kernel="gaussian", bandwidth=2.0)
    np.random.seed(5)
    with mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(5))):
        got = ref["rpcholesky"].accelerated_rpcholesky(A, 150, b=30)
    rng, legacy = orc.make_streams(5)
    want = orc.accelerated_rpcholesky(orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=2.0), 150, 30, rng, legacy)
    assert np.array_equal(got.idx, want.idx)
    np.testing.assert_allclose(got.G, want.G, rtol=0, atol=1e-12)
    assert A.num_queries() == 3000 + 150 * 3000 + 30 * 30 * len(want.rounds)


def test_reference_lra_unittests_pass_against_this_package():
    """The reference's own 15 tests of lra.py (tests/test_lra.py) run UNCHANGED against this
    package's lra module: a caller that imports `lra` finds PSDLowRank, CompactEigenvalueDecomposition and
    NystromExtension with the reference's behaviour."""
    import importlib.util
    import unittest
    ref_test = "/root/reference/tests/test_lra.py"
    if not os.path.exists(ref_test):
        pytest.skip("/root/reference is not present here")
    from randomly_pivoted_cholesky_b200 import lra as ours
    saved = sys.modules.get("lra")
    sys.modules["lra"] = ours
    try:
        spec = importlib.util.spec_from_file_location("reference_test_lra", ref_test)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        result = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, "w")).run(suite)
    finally:
        if saved is None:
            sys.modules.pop("lra", None)
        else:
            sys.modules["lra"] = saved
    assert result.testsRun == 15
    assert result.wasSuccessful(), [str(f[0]) + ": " + f[1][-400:] for f in result.failures + result.errors]


def test_nystrom_extension_and_krr_vec():
    """Restatement of the NystromExtension / krr_vec contract (lra.py:82-85, 118-171) that also runs where the reference
    is absent (the GPU box)."""
    from randomly_pivoted_cholesky_b200 import CompactEigenvalueDecomposition, NystromExtension, PSDLowRank
    rng = np.random.default_rng(42)
    N, r = 50, 10
    G = rng.standard_normal((r, N))
    A = G.T @ G
    idx = np.arange(r)
    core, rows = A[np.ix_(idx, idx)], A[idx, :]
    with pytest.raises(RuntimeError):
        NystromExtension(core=core)
    ny = NystromExtension(core=core, rows=rows, idx=idx)
    assert ny.shape == (N, N) and ny.rank() == r and ny.get_factor().shape == (r, N)
    np.testing.assert_allclose(ny.matrix(), rows.T @ np.linalg.pinv(core) @ rows, atol=1e-5)
    x = rng.standard_normal((N, 3))
    np.testing.assert_allclose(ny @ x, ny.matrix() @ x, atol=1e-8)
    assert abs(ny.trace() - np.trace(ny.matrix())) <= 1e-8 * ny.trace()
    sc = rng.random(N) + 0.5
    ny2 = ny.scale(sc)
    np.testing.assert_allclose(ny2.get_rows(), rows * sc[None, :])
    np.testing.assert_allclose(ny2.get_factor(), ny.get_factor() * np.sqrt(sc)[None, :])
    eig = ny.eigenvalue_decomposition()
    np.testing.assert_allclose(eig.matrix() if hasattr(eig, "matrix") else eig @ np.eye(N), ny.matrix(), atol=1e-8)
    # krr_vec: (A + lam I)^{-1} B for a block of right-hand sides
    ced = CompactEigenvalueDecomposition.from_G(G)
    B = rng.standard_normal((N, 4))
    np.testing.assert_allclose(ced.krr_vec(B, 0.1), np.linalg.solve(A + 0.1 * np.eye(N), B), atol=1e-8)
    np.testing.assert_allclose(ced.krr(B[:, 0], 0.1), np.linalg.solve(A + 0.1 * np.eye(N), B[:, 0]), atol=1e-8)
    assert PSDLowRank(G).eigenvalue_decomposition().rank() == r


def test_bench_flop_accounting():
    """bench.py's roofline bookkeeping: algorithmic flops are SURVEY 8(d)'s s (2c + s) N; the executed count replays the
    kernel's fragment skip and must lie between the algorithmic and the dense count."""
    sys.path.insert(0, ROOT)
    import bench
    rounds = [88, 161, 184, 187, 189, 197, 193, 196, 197, 193, 192, 23]          # T*, seed 1234
    c = 0
    alg = ex = dense = 0.0
    for s_ in rounds:
        alg += s_ * (2 * c + s_)
        dense += 2 * s_ * (c + s_)
        ex += bench.schur_executed_flops(c, s_, 1.0)
        assert bench.schur_executed_flops(c, s_, 1.0, tri=False) >= 2 * s_ * (c + s_)      # row / k padding only adds
        c += s_
    assert c == 2000 and alg == 2000 ** 2                                          # N k^2 per unit column
    assert alg < ex < dense
    assert ex / alg < 1.05                                                         # staircase + padding: a few per cent
    # s > 256 runs in row chunks; the skip applies inside each chunk
    assert bench.schur_executed_flops(100, 300, 1.0) < bench.schur_executed_flops(100, 300, 1.0, tri=False)
