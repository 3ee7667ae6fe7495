"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures that
the unmodified reference produced (tests/golden/make_golden.py).  Run on the B200 box with `-m gpu`.

Tolerances (BASELINE.json north_star): pivots bit-exact; F (= G^T) and A[S,:] within 1e-10 relative
Frobenius norm; equal trace-norm error.
"""
import os
from unittest import mock

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from make_golden import CASES, KERNEL_CASES, golden_inputs, kernel_inputs  # noqa: E402
from test_oracle import SET_PARITY, check_block_set_parity  # noqa: E402
from oracle import rpc_oracle as orc  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F_TOL = 1e-10


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def relerr(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def rpc():
    import randomly_pivoted_cholesky_b200 as pkg
    pkg.load()
    return pkg


def replay(seed):
    """The replay recipe of SURVEY.md 8(c): same patch the golden generator applied to the reference."""
    np.random.seed(seed)
    return mock.patch("numpy.random.default_rng", lambda *a, **k: np.random.Generator(np.random.PCG64(seed)))


def make_operator(rpc, inp):
    if "dense" in inp:
        return inp["dense"]
    return rpc.KernelMatrix(inp["X"], kernel=inp["kernel"], bandwidth=inp["bandwidth"], **inp["kw"])


def run_gpu(rpc, spec):
    inp = golden_inputs(spec)
    A = make_operator(rpc, inp)
    kw = {"stoptol": spec["stoptol"]} if "stoptol" in spec else {}
    with replay(spec["seed"]):
        if spec["alg"] in ("greedy", "rgreedy"):
            lra = rpc.greedy(A, spec["k"], randomized_tiebreaking=spec["alg"] == "rgreedy", **kw)
        elif spec["alg"] == "simple":
            lra = rpc.simple_rpcholesky(A, spec["k"], **kw)
        elif spec["alg"] == "block":
            lra = rpc.block_rpcholesky(A, spec["k"], b=spec["b"], **kw, **spec.get("blockkw", {}))
        elif spec["alg"] == "bgreedy":
            lra = rpc.greedy(A, spec["k"], b=spec["b"], **kw, **spec.get("blockkw", {}))
        else:
            lra = rpc.accelerated_rpcholesky(A, spec["k"], b=spec["b"], **kw)
    return A, lra


# rank-deficient cases: after the residual is exhausted the reference divides round-off by round-off, so
# only the rows computed while the residual is still resolved are comparable
ILL_CONDITIONED = {"accel_stoptol", "block_stoptol", "simple_stoptol", "rbrp_lowrank"}


@pytest.mark.parametrize("name", sorted(CASES))
def test_factorisation_matches_reference(rpc, name):
    spec = CASES[name]
    gold = load(name)
    A, lra = run_gpu(rpc, spec)
    idx = np.asarray(lra.idx, dtype=np.int64)
    if name in SET_PARITY:          # block greedy: numpy's argpartition order inside a block is not reproducible
        check_block_set_parity(spec, idx, lra.G, gold)
        return
    assert np.array_equal(idx, gold["idx"]), "pivots must be bit-exact"
    G, rows = lra.G, lra.rows
    assert G.shape == gold["G"].shape and rows.shape == gold["rows"].shape
    assert relerr(rows, gold["rows"]) <= F_TOL
    if name in ILL_CONDITIONED:
        r = spec["dense_rank"] - 2
        assert relerr(G[:r], gold["G"][:r]) <= 1e-7
        assert abs(lra.trace() - float(gold["trace"])) <= 1e-8 * abs(float(gold["trace"]))
    else:
        assert relerr(G, gold["G"]) <= F_TOL
        assert abs(lra.trace() - float(gold["trace"])) <= 1e-11 * abs(float(gold["trace"]))
    if "queries" in gold:
        assert A.num_queries() == int(gold["queries"])


@pytest.mark.parametrize("name", sorted(KERNEL_CASES))
def test_kernel_values_match_reference(rpc, name):
    spec = KERNEL_CASES[name]
    gold = load(name)
    X, idx, kw = kernel_inputs(spec)
    A = rpc.KernelMatrix(X, kernel=spec["kernel"], bandwidth=spec["bw"], **kw)
    # Matern nu=1/2 is first order in the sqrt of the round-off self-distance
    atol = 5e-7 if name == "kern_matern05" else 1e-13
    np.testing.assert_allclose(A[idx, :], gold["rows"], rtol=1e-12, atol=atol)
    np.testing.assert_allclose(A[idx, idx], gold["block"], rtol=1e-12, atol=atol)
    np.testing.assert_array_equal(A.diag(), gold["diag"])
    np.testing.assert_allclose(A[int(idx[3]), :], gold["one_row"], rtol=1e-12, atol=atol)
    assert A.num_queries() == 7 * 300 + 49 + 300 + 300


def test_laplace_l1_distance_is_bit_exact(rpc):
    """Sequential feature order == scipy cdist 'cityblock'; only the exponential's last bits may differ: the device evaluates
    exp(l1 * (-1/bw)) with its table exponential (<= 1.2 ulp) instead of exp(-l1/bw) with libm (one more rounding in the
    argument): a few ulp, i.e. 1e-15 relative.  A summation order other than cdist's would show up at 1e-14 and above."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((700, 50))
    A = rpc.KernelMatrix(X, kernel="laplace", bandwidth=50.0)
    idx = np.arange(0, 700, 53)
    got = A[idx, :]
    want = orc.laplace_mtx(X[idx], X, bandwidth=50.0)
    np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)


@pytest.mark.parametrize("d", [3, 20, 50, 130])
@pytest.mark.parametrize("s", [1, 7, 33, 120, 137, 167, 200, 250, 300])
def test_dmma_evaluator_shapes(rpc, d, s):
    """Gaussian rows against the oracle over the row tiles of both DMMA evaluators: the panel-resident kernel
    (csrc/eval_panel.cu: 8-row-step tiles, short contractions) and the streamed one it falls back to (d = 130 with a
    tall panel does not fit in shared memory)."""
    rng = np.random.default_rng(d * 1000 + s)
    n = 1000 + 37 * s
    X = rng.standard_normal((n, d))
    A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=float(np.sqrt(d)))
    idx = rng.integers(0, n, s)
    want = orc.gaussian_mtx(X[idx], X, bandwidth=float(np.sqrt(d)))
    got = A[idx, :]
    got = got.reshape(want.shape)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("tag", ["uniform", "random", "sparse"])
def test_sampler_golden(rpc, tag):
    from randomly_pivoted_cholesky_b200 import ops
    gold = load("sampler")
    idx, S = ops.sample_pivots(gold[tag + "_diag"], gold[tag + "_u"])
    assert np.array_equal(idx.cpu().numpy(), gold[tag + "_idx"])
    assert S == sum(gold[tag + "_diag"])


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 5000, 100003, 1000000])
@pytest.mark.parametrize("kind", ["ones", "random", "heavy", "sparse", "dyadic", "tiny"])
def test_sampler_bit_exact_vs_numpy(rpc, n, kind):
    """Exact sequential-cumsum reproduction: idx and python sum(diag) equal numpy's, bit for bit."""
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(n + len(kind))
    if kind == "ones":
        d = np.ones(n)
    elif kind == "random":
        d = rng.random(n)
    elif kind == "heavy":
        d = rng.random(n) ** 12
    elif kind == "sparse":
        d = rng.random(n) * (rng.random(n) < 0.05)
        d[rng.integers(0, n)] = 1.0
    elif kind == "dyadic":
        d = np.ldexp(rng.integers(1, 8, n).astype(float), rng.integers(-20, 3, n))   # many rounding ties
    else:
        d = rng.random(n) * 1e-200
    m = 300
    u = rng.random(m)
    u[0] = 0.0
    u[1] = 1.0 - 2.0 ** -53
    idx, S = ops.sample_pivots(d, u)
    assert S == orc.seq_sum(d)
    want = orc.sample_pivots(d, u)
    assert np.array_equal(idx.cpu().numpy(), want)


def test_sampler_raises_like_choice(rpc):
    from randomly_pivoted_cholesky_b200 import ops
    with pytest.raises(ValueError):
        ops.sample_pivots(np.zeros(10), [0.5])
    with pytest.raises(ValueError):
        ops.sample_pivots(np.array([1.0, np.nan, 2.0]), [0.5])
    with pytest.raises(ValueError):
        ops.sample_pivots(np.array([1.0, -0.5, 2.0]), [0.5])


@pytest.mark.parametrize("b", [1, 5, 64, 200, 228, 260])
def test_rejection_cholesky_matches_oracle(rpc, b):
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(b)
    r = max(1, b // 2)
    B = rng.standard_normal((b, r + 3))
    H = B @ B.T + 1e-3 * np.diag(rng.random(b))
    u = rng.random(b)
    Lw, accw = orc.rejection_cholesky(H.copy(), u)
    L, acc = ops.rejection_cholesky(H, u)
    assert np.array_equal(acc, accw)
    np.testing.assert_allclose(L.cpu().numpy(), Lw, rtol=1e-9, atol=1e-11)
    # truncation (rpcholesky.py:193-196)
    if len(accw) > 2:
        L2, acc2 = ops.rejection_cholesky(H, u, max_accept=len(accw) - 2)
        assert np.array_equal(acc2, accw[:-2])
        np.testing.assert_allclose(L2.cpu().numpy(), Lw[:-2, :-2], rtol=1e-9, atol=1e-11)
    with pytest.raises(RuntimeError):
        ops.rejection_cholesky(-np.eye(3), np.zeros(3))


@pytest.mark.parametrize("m,r", [(1, 1), (7, 7), (64, 64), (65, 20), (200, 200), (300, 90), (400, 400)])
def test_pivoted_cholesky_matches_lapack(rpc, m, r):
    """rpc_pivoted_cholesky against `_greedy_cholesky` (dpstrf + rank rule, rpcholesky.py:52-63) on full-rank and
    rank-deficient blocks, with and without the tolerance rule."""
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(100 + m)
    B = rng.standard_normal((m, r)) * np.exp(rng.standard_normal(m))[:, None]
    H = B @ B.T
    for rtol, atol in ((None, None), (1.0 / m, 0.0), (1e-6, 1e-9), (0.3, 0.0)):
        Lw, pivw, rankw = orc._greedy_cholesky(H.copy(), rtol=rtol, atol=atol)
        L, piv, rank = ops.pivoted_cholesky(H, rtol=rtol, atol=atol)
        if rtol is None and r < m:
            # dpstrf's own stop (pivot <= m*eps*max diag) sits at round-off level: allow the rank to differ by a few there
            assert abs(rank - rankw) <= 3 and rank >= r
            q = min(rank, rankw, r)
        else:
            assert rank == rankw
            q = rank if r == m else min(rank, r)
        assert np.array_equal(piv[:q], pivw[:q] - 1)
        np.testing.assert_allclose(L.cpu().numpy()[:q, :q], Lw[:q, :q], rtol=1e-8, atol=1e-10 * np.abs(Lw).max())


@pytest.mark.parametrize("n,b", [(10, 3), (1000, 1), (5000, 5000), (100_000, 200), (1_000_003, 2000)])
def test_select_largest_matches_argpartition_set(rpc, n, b):
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(n)
    for kind in ("random", "ties", "ones"):
        if kind == "random":
            d = rng.random(n) ** 3
        elif kind == "ties":
            d = np.round(rng.random(n) * 20) / 20           # many ties, also at the threshold
        else:
            d = np.ones(n)
        got = ops.select_largest(d, b).cpu().numpy()
        assert len(np.unique(got)) == b and np.all(np.diff(got) > 0)
        ref = np.argpartition(d, -b)[-b:]
        thr = d[ref].min()
        assert np.array_equal(np.sort(d[got]), np.sort(d[ref]))                      # same multiset of values
        above = np.flatnonzero(d > thr)
        tied = np.flatnonzero(d == thr)[:b - len(above)]                             # lowest indices among the ties
        assert np.array_equal(got, np.sort(np.concatenate([above, tied])))


def test_shifted_cholesky_and_failure_flag(rpc):
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(3)
    B = rng.standard_normal((40, 60))
    Cm = B @ B.T
    L, info = ops.shifted_cholesky(Cm, 1e-9)
    assert info == 0
    np.testing.assert_allclose(L.cpu().numpy(), np.linalg.cholesky(Cm + 1e-9 * np.eye(40)), rtol=1e-10, atol=1e-10)
    bad = Cm.copy()
    bad[7, 7] = -1.0
    _, info = ops.shifted_cholesky(bad, 0.0)
    assert info == 8


@pytest.mark.parametrize("c,s", [(0, 1), (0, 40), (5, 3), (37, 16), (100, 64), (130, 130), (64, 200), (300, 256), (50, 300),
                                 (20, 23), (10, 50), (33, 72), (70, 88), (12, 104), (99, 118), (7, 56), (0, 248), (3, 136)])
def test_schur_update_matches_numpy(rpc, c, s):
    """G_new = L^{-1}(rows - P^T G[:c]) and the fused diagonal update, against numpy (fp64)."""
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(c * 1000 + s)
    n_pad = 128 * 9
    k = c + s
    Gh = np.zeros((k, n_pad))
    Gh[:c] = rng.standard_normal((c, n_pad))
    rows = rng.standard_normal((s, n_pad))
    P = rng.standard_normal((c, s))
    Lh = np.tril(rng.standard_normal((s, s))) + 4 * np.sqrt(s) * np.eye(s)
    diag = rng.random(n_pad) * 50
    want = np.linalg.solve(Lh, rows - P.T @ Gh[:c])
    wdiag = np.clip(diag - np.sum(want ** 2, axis=0), 0, None)
    dev = torch.device("cuda")
    G = torch.from_numpy(Gh).to(dev)
    dg = torch.from_numpy(diag).to(dev)
    ops.schur_update(G, c, torch.from_numpy(rows).to(dev), torch.from_numpy(P).to(dev), torch.from_numpy(Lh).to(dev), dg)
    got = G[c:].cpu().numpy()
    assert relerr(got, want) <= 1e-12
    np.testing.assert_allclose(dg.cpu().numpy(), wdiag, rtol=1e-11, atol=1e-10)
    assert np.array_equal(G[:c].cpu().numpy(), Gh[:c])


@pytest.mark.parametrize("c,s,left", [(40, 167, 30), (0, 136, 2), (96, 200, 40), (16, 256, 36), (30, 129, 58)])
def test_schur_update_tail_of_the_tile_waves(rpc, c, s, left):
    """Shapes whose column tiles do not fill the last wave of the persistent grid: the leftover tiles go through the
    row-chunked tail launch + finishing kernel; results must equal the plain path's."""
    from randomly_pivoted_cholesky_b200 import ops
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rng = np.random.default_rng(c * 1000 + s)
    n_pad = (4 * sms + left) * 64
    assert n_pad % 128 == 0
    k = c + s
    Gh = np.zeros((k, n_pad))
    Gh[:c] = rng.standard_normal((c, n_pad))
    rows = rng.standard_normal((s, n_pad))
    P = rng.standard_normal((c, s))
    Lh = np.tril(rng.standard_normal((s, s))) + 4 * np.sqrt(s) * np.eye(s)
    diag = rng.random(n_pad) * 50
    want = np.linalg.solve(Lh, rows - P.T @ Gh[:c])
    wdiag = np.clip(diag - np.sum(want ** 2, axis=0), 0, None)
    dev = torch.device("cuda")
    G = torch.from_numpy(Gh).to(dev)
    dg = torch.from_numpy(diag).to(dev)
    ops.schur_update(G, c, torch.from_numpy(rows).to(dev), torch.from_numpy(P).to(dev), torch.from_numpy(Lh).to(dev), dg)
    got = G[c:].cpu().numpy()
    assert relerr(got, want) <= 1e-12
    # every tail column separately (a chunk that dropped or double-counted rows would show up here)
    np.testing.assert_allclose(dg.cpu().numpy(), wdiag, rtol=1e-11, atol=1e-10)
    assert np.array_equal(G[:c].cpu().numpy(), Gh[:c])


@pytest.mark.parametrize("shape", ["tail", "row_chunks"])
def test_speculative_rows_and_tail_launch_against_oracle(rpc, monkeypatch, shape):
    """The >= 4-GPU code path forced on one GPU: all proposals evaluated on the side stream, accepted rows read through the
    row map and written to rows[c:c+s] by the Schur kernel itself, on a shape whose Schur update also takes the row-chunked tail
    launch, and on one with more than 256 accepted pivots per round (row chunks: only the first chunk writes the rows)."""
    import importlib
    drivers = importlib.import_module("randomly_pivoted_cholesky_b200.rpcholesky")     # the module, not the function
    monkeypatch.setattr(drivers, "USE_SPEC", "1")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n, d, k, b = ((4 * sms + 30) * 64 - 5, 9, 500, 200) if shape == "tail" else (24000, 6, 1400, 640)
    X = np.random.default_rng(17).standard_normal((n, d))
    bw = float(np.sqrt(d))
    A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=bw)
    with replay(23):
        lra = rpc.accelerated_rpcholesky(A, k, b=b)
    rng, legacy = orc.make_streams(23)
    ref = orc.accelerated_rpcholesky(orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=bw), k, b, rng, legacy)
    assert np.array_equal(np.asarray(lra.idx), ref.idx)
    assert relerr(lra.G, ref.G) <= F_TOL and relerr(lra.rows, ref.rows) <= F_TOL
    assert A.num_queries() == n + k * n + b * b * len(ref.rounds)


def test_configs_c1_c2_against_oracle(rpc):
    """BASELINE.json configs[0] in full and configs[1] at reduced N: GPU vs oracle on replayed draws."""
    # C1: simple_rpcholesky, Gaussian, N=10,000 d=10 k=100
    X = np.random.default_rng(0).standard_normal((10000, 10))
    bw = float(np.sqrt(10))
    with replay(1234):
        lra = rpc.simple_rpcholesky(rpc.KernelMatrix(X, kernel="gaussian", bandwidth=bw), 100)
    rng, legacy = orc.make_streams(1234)
    A0 = orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=bw)
    ref = orc.simple_rpcholesky(A0, 100, rng)
    assert list(lra.idx) == list(ref.idx)
    assert relerr(lra.G, ref.G) <= F_TOL and relerr(lra.rows, ref.rows) <= F_TOL
    tr = 10000.0
    assert abs((tr - lra.trace()) / tr - (tr - ref.trace()) / tr) <= 1e-12
    # C2 shape at N=20,000: accelerated b=120, Gaussian d=20, k=1000
    X = np.random.default_rng(1).standard_normal((20000, 20))
    bw = float(np.sqrt(20))
    with replay(4321):
        lra = rpc.accelerated_rpcholesky(rpc.KernelMatrix(X, kernel="gaussian", bandwidth=bw), 1000, b=120)
    rng, legacy = orc.make_streams(4321)
    ref = orc.accelerated_rpcholesky(orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=bw), 1000, 120, rng, legacy)
    assert np.array_equal(np.asarray(lra.idx), ref.idx)
    assert lra.rounds == ref.rounds
    assert relerr(lra.G, ref.G) <= F_TOL and relerr(lra.rows, ref.rows) <= F_TOL
    # trace(): the driver's account trace(A) - sum(residual diagonal) (no pass over G) against ||G||_F^2 and the oracle
    assert lra._trace_hint is not None
    assert abs(lra.trace() - np.linalg.norm(lra.G, "fro") ** 2) <= 1e-12 * lra.trace()
    assert abs(lra.trace() - ref.trace()) <= 1e-11 * ref.trace()


def test_block_laplace_against_oracle(rpc):
    """configs[2] shape (block b=200, Laplace, d=50) at N=20,000, k=600."""
    X = np.random.default_rng(2).standard_normal((20000, 50))
    with replay(99):
        lra = rpc.block_rpcholesky(rpc.KernelMatrix(X, kernel="laplace", bandwidth=50.0), 600, b=200)
    rng, _ = orc.make_streams(99)
    ref = orc.block_rpcholesky(orc.OracleKernelMatrix(X, kernel="laplace", bandwidth=50.0), 600, 200, rng)
    assert [int(i) for i in lra.idx] == [int(i) for i in ref.idx]
    assert relerr(lra.G, ref.G) <= F_TOL and relerr(lra.rows, ref.rows) <= F_TOL
    assert lra._trace_hint is not None
    assert abs(lra.trace() - np.linalg.norm(lra.G, "fro") ** 2) <= 1e-12 * lra.trace()
    assert abs(lra.trace() - ref.trace()) <= 1e-11 * ref.trace()


def test_dispatch_and_errors(rpc):
    X = np.random.default_rng(3).standard_normal((600, 4))
    A = rpc.KernelMatrix(X, kernel="matern", bandwidth=2.0, nu=2.5)
    with replay(5):
        a = rpc.rpcholesky(A, 30, b=10)                 # accelerated with b
    with replay(5):
        b_ = rpc.accelerated_rpcholesky(A, 30, b=10)
    assert np.array_equal(a.idx, b_.idx)
    with replay(6):
        c = rpc.rpcholesky(A, 30, b=10, accelerated=False)
    with replay(6):
        d = rpc.block_rpcholesky(A, 30, b=10)
    assert list(c.idx) == list(d.idx)
    lra = rpc.rpcholesky(A, 40)                        # b="auto"
    assert lra.G.shape == (40, 600) and lra.rank() == 40
    with pytest.raises(RuntimeError):
        rpc.KernelMatrix(X, kernel="cauchy")
    with pytest.raises(ValueError):
        rpc.block_rpcholesky(A, 10, b=5, strategy="nope")
    with pytest.raises(RuntimeError):
        rpc.cholesky_helper(A, 10, "bogus")
    # exhausted matrix: Generator.choice raises ValueError in the reference (comparison.py:51 relies on it)
    Z = np.zeros((50, 50))
    Z[:3, :3] = np.eye(3)
    with pytest.raises(ValueError):
        rpc.simple_rpcholesky(Z, 10)


def test_reconstruction_property_large(rpc):
    """Size-independent properties at N = 200,000: A[S,:] rows reproduce, F F^T interpolates A on S,
    diag(A - F F^T) >= 0 and the trace identity holds."""
    n, d, k, b = 200000, 8, 300, 60
    X = np.random.default_rng(7).standard_normal((n, d))
    A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=float(np.sqrt(d)))
    with replay(8):
        lra = rpc.accelerated_rpcholesky(A, k, b=b)
    S = np.asarray(lra.idx)
    assert len(set(S.tolist())) == k
    G = lra.G_device
    rows = lra.rows_device
    St = torch.as_tensor(S, device=G.device)
    # Nystrom interpolation on the pivot rows: (F F^T)[S,:] == A[S,:]
    rec = G[:, St].T @ G
    err = float(torch.linalg.norm(rec - rows) / torch.linalg.norm(rows))
    assert err <= 1e-9
    resid = 1.0 - (G * G).sum(0)
    assert float(resid.min()) >= -1e-12
    assert abs(float(resid.clamp_min(0).sum()) - (n - lra.trace())) <= 1e-6 * n


@pytest.mark.parametrize("kernel,kw,b,k", [("matern", {"nu": 2.5}, 400, 900), ("gaussian", {}, 640, 1400),
                                           ("laplace", {}, 288, 600), ("laplace", {}, 1000, 2200), ("matern", {"nu": 1.5}, 1536, 3000)])
def test_large_proposal_blocks_against_oracle(rpc, kernel, kw, b, k):
    """configs[3] shape (accelerated, b = 400 > one 256-row tile, Matern-5/2, d = 3) at N = 24,000 and two more
    kernels: exercises the row-chunked evaluator / Schur update and the global-memory rejection path."""
    d = 3 if kernel == "matern" else 6
    X = np.random.default_rng(11).standard_normal((24000, d))
    bw = float(d) if kernel == "laplace" else float(np.sqrt(d))
    with replay(31):
        lra = rpc.accelerated_rpcholesky(rpc.KernelMatrix(X, kernel=kernel, bandwidth=bw, **kw), k, b=b)
    rng, legacy = orc.make_streams(31)
    ref = orc.accelerated_rpcholesky(orc.OracleKernelMatrix(X, kernel=kernel, bandwidth=bw, **kw), k, b, rng, legacy)
    assert np.array_equal(np.asarray(lra.idx), ref.idx)
    assert lra.rounds == ref.rounds
    if b > 600:
        assert max(ref.rounds) > 256          # more accepted pivots than one 256-row tile: row-chunked update
    # G: 1e-10 unless a round's triangular factor is ill conditioned (oracle.g_tolerance states the argument); rows: 1e-10
    eg = relerr(lra.G, ref.G)
    assert eg <= orc.g_tolerance(ref), (eg, max(ref.cond_L))
    assert relerr(lra.rows, ref.rows) <= F_TOL


def test_block_large_b_against_oracle(rpc):
    X = np.random.default_rng(12).standard_normal((20000, 5))
    with replay(41):
        lra = rpc.block_rpcholesky(rpc.KernelMatrix(X, kernel="gaussian", bandwidth=float(np.sqrt(5))), 700, b=320)
    rng, _ = orc.make_streams(41)
    ref = orc.block_rpcholesky(orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=float(np.sqrt(5))), 700, 320, rng)
    assert [int(i) for i in lra.idx] == [int(i) for i in ref.idx]
    eg = relerr(lra.G, ref.G)
    assert eg <= orc.g_tolerance(ref), (eg, max(ref.cond_L))
    assert relerr(lra.rows, ref.rows) <= F_TOL


# ---- first "next" row: KRR_Nystrom on top of the hot path ------------------------------------------
from make_golden_krr import KRR_CASES, krr_inputs  # noqa: E402


@pytest.mark.parametrize("name", sorted(KRR_CASES))
def test_krr_nystrom_matches_reference(rpc, name):
    spec = KRR_CASES[name]
    gold = load(name)
    Xtr, ytr, Xts, _ = krr_inputs(spec)
    if spec["alg"] == "accel":
        method = lambda A, k: rpc.accelerated_rpcholesky(A, k, b=spec["b"])
    else:
        method = lambda A, k: rpc.block_rpcholesky(A, k, b=spec["b"])
    model = rpc.KRR_Nystrom(kernel=spec["kernel"], bandwidth=spec["bw"])
    with replay(spec["seed"]):
        model.fit_Nystrom(Xtr, ytr, spec["lamb"], spec["k"], method)
    assert np.array_equal(np.asarray(model.sample_idx, dtype=np.int64), gold["sample_idx"])
    assert model.queries == int(gold["queries"])
    assert abs(model.reltrace_err - float(gold["reltrace_err"])) <= 1e-12
    preds = model.predict_Nystrom(Xts)
    # the normal equations square the condition number of K_MM: compare what the model predicts, not beta itself
    np.testing.assert_allclose(preds, gold["preds"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(model.sol, gold["sol"], rtol=1e-4, atol=1e-5 * np.abs(gold["sol"]).max())


def test_exact_krr_small(rpc):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((300, 4)); y = np.sin(X[:, 0]) + X[:, 1] ** 2
    Xt = rng.standard_normal((40, 4))
    model = rpc.KRR_Nystrom(kernel="gaussian", bandwidth=1.5)
    model.fit(X, y, 1e-4)
    Kx = orc.gaussian_mtx(X, X, bandwidth=1.5)
    want = orc.gaussian_mtx(Xt, X, bandwidth=1.5) @ np.linalg.solve(Kx + 1e-4 * 300 * np.eye(300), y)
    np.testing.assert_allclose(model.predict(Xt), want, rtol=1e-7, atol=1e-8)


# ---- edge cases: tiny and ragged problems, exhausted matrices, error behaviour equal to the reference's ----------------
@pytest.mark.parametrize("n,d,k,b", [(5, 2, 3, 2), (130, 3, 1, 1), (129, 4, 40, 200), (257, 2, 30, 7), (1000, 6, 64, 64)])
def test_tiny_and_ragged_sizes(rpc, n, d, k, b):
    X = np.random.default_rng(n).standard_normal((n, d))
    bw = 1.3
    for alg in ("accel", "block", "simple"):
        A = rpc.KernelMatrix(X, kernel="gaussian", bandwidth=bw)
        A0 = orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=bw)
        rng, legacy = orc.make_streams(3)
        with replay(3):
            if alg == "accel":
                lra = rpc.accelerated_rpcholesky(A, k, b=b)
                ref = orc.accelerated_rpcholesky(A0, k, b, rng, legacy)
            elif alg == "block":
                lra = rpc.block_rpcholesky(A, k, b=b)
                ref = orc.block_rpcholesky(A0, k, b, rng)
            else:
                lra = rpc.simple_rpcholesky(A, k)
                ref = orc.simple_rpcholesky(A0, k, rng)
        assert [int(i) for i in lra.idx] == [int(i) for i in ref.idx], alg
        assert lra.G.shape == ref.G.shape
        assert relerr(lra.rows, ref.rows) <= F_TOL
        if alg == "simple":
            tol = F_TOL                      # one pivot per step: a division by sqrt(diag), no triangular solve
        else:
            tol = orc.g_tolerance(ref)       # 1e-10 unless a round's factor is ill conditioned (tiny N: pivots crowd together)
        eg = relerr(lra.G, ref.G)
        assert eg <= tol, (alg, eg, getattr(ref, "cond_L", None))
        assert relerr(lra.G.T @ lra.G, ref.G.T @ ref.G) <= 2 * tol
        assert A.num_queries() == A0.num_queries()


def test_exhausted_matrix_raises_like_the_reference(rpc):
    """20 distinct points repeated 10 times: the kernel matrix has rank 20.  The reference stops early through stoptol
    (block / accelerated) or runs into Generator.choice's ValueError (simple, stoptol = 0); so must we."""
    base = np.random.default_rng(5).standard_normal((20, 3))
    X = np.tile(base, (10, 1))
    A0 = orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=1.0)
    rng, legacy = orc.make_streams(9)
    ref = orc.accelerated_rpcholesky(A0, 60, 15, rng, legacy)
    with replay(9):
        lra = rpc.accelerated_rpcholesky(rpc.KernelMatrix(X, kernel="gaussian", bandwidth=1.0), 60, b=15)
    assert lra.G.shape == ref.G.shape and lra.G.shape[0] < 60          # early stop at the same round
    n_ok = 18
    assert np.array_equal(np.asarray(lra.idx)[:n_ok], ref.idx[:n_ok])
    # greedy on a dense rank-deficient matrix keeps going on round-off, like the reference; simple 'rp' with an exactly
    # exhausted diagonal raises ValueError
    Z = np.zeros((40, 40)); Z[:4, :4] = np.diag([4.0, 9.0, 16.0, 1.0])        # perfect squares: the residual hits exactly 0
    with pytest.raises(ValueError):
        rpc.simple_rpcholesky(Z, 8)
    with pytest.raises(ValueError):
        orc.simple_rpcholesky(orc.OracleDenseMatrix(Z), 8, np.random.default_rng(0))
    g = rpc.greedy(Z, 4)
    assert list(g.idx) == [2, 1, 0, 3]


def test_input_forms(rpc):
    """float32 / non-contiguous / torch inputs, ndarray auto-wrapping, list and slice indexing of the operator."""
    rng = np.random.default_rng(8)
    X64 = rng.standard_normal((300, 5))
    A = rpc.KernelMatrix(np.asfortranarray(X64), kernel="laplace", bandwidth=2.0)
    B = rpc.KernelMatrix(torch.from_numpy(X64), kernel="laplace", bandwidth=2.0)
    want = orc.laplace_mtx(X64, X64, bandwidth=2.0)
    np.testing.assert_allclose(A[:, :], want, rtol=1e-13)
    np.testing.assert_allclose(B[[3, 5], :], want[[3, 5], :], rtol=1e-13)
    np.testing.assert_allclose(A[7, 9], want[7, 9], rtol=1e-13)
    np.testing.assert_allclose(A[np.array([1, 2]), slice(10, 20)], want[1:3, 10:20], rtol=1e-13)
    np.testing.assert_allclose(A[4, :], want[4, :], rtol=1e-13)
    with pytest.raises(RuntimeError):
        A["a", 0]
    # numpy index semantics of the reference's `self.data[vec, :]` (matrix.py:189): negative indices wrap, out-of-range raise
    np.testing.assert_allclose(A[-1, :], want[-1, :], rtol=1e-13)
    np.testing.assert_allclose(A[[0, -300], [-2, 299]], want[np.ix_([0, -300], [-2, 299])], rtol=1e-13)
    for bad in (300, -301):
        with pytest.raises(IndexError):
            A[bad, :]
        with pytest.raises(IndexError):
            A[[1, 2], [0, bad]]
    with pytest.raises(IndexError):
        A.out_of_sample(X64[:3], [0, 300])
    assert A.trace() == 300.0 and isinstance(A.to_matrix(), rpc.PSDMatrix)
    D = X64 @ X64.T
    with replay(2):
        a = rpc.rpcholesky(D.astype(np.float32).astype(np.float64), 4, b=2)
    assert a.G.shape == (4, 300)
    # median bandwidths (matrix.py:168-176)
    from scipy.spatial.distance import cdist
    M = rpc.KernelMatrix(X64, kernel="gaussian", bandwidth="median")
    assert abs(M.bandwidth - np.median(cdist(X64, X64))) <= 1e-12
    M1 = rpc.KernelMatrix(X64, kernel="laplace", bandwidth="approx_median")
    assert abs(M1.bandwidth - np.median(cdist(X64, X64, "cityblock"))) <= 1e-12


def test_nonsymmetric_kernel_matrix_and_out_of_sample(rpc):
    """matrix.py:191-239: the cross-kernel operator used for prediction."""
    rng = np.random.default_rng(21)
    X = rng.standard_normal((70, 5)); Y = rng.standard_normal((310, 5))
    for kernel, kw, ref in [("gaussian", {}, orc.gaussian_mtx), ("laplace", {}, orc.laplace_mtx),
                            ("matern", {"nu": 1.5}, orc.matern_mtx)]:
        K = rpc.NonsymmetricKernelMatrix(X, Y, kernel=kernel, bandwidth=1.7, **kw)
        want = ref(X, Y, bandwidth=1.7, **kw)
        assert K.shape == (70, 310)
        np.testing.assert_allclose(K[:, :], want, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(K[[3, 9], :], want[[3, 9], :], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(K[5, 7], want[5, 7], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(K[:, np.array([1, 300])], want[:, [1, 300]], rtol=1e-12, atol=1e-14)
        A = rpc.KernelMatrix(Y, kernel=kernel, bandwidth=1.7, **kw)
        np.testing.assert_allclose(A.out_of_sample(X, [4, 8, 15]), want[:, [4, 8, 15]], rtol=1e-12, atol=1e-14)


# ---- accelerated RPQR: sampler + rejection Cholesky reused, QR / GEMM from the libraries ------
from make_golden_rpqr import RPQR_CASES, rpqr_inputs  # noqa: E402


@pytest.mark.parametrize("name", sorted(RPQR_CASES))
def test_rpqr_matches_reference(rpc, name):
    from randomly_pivoted_cholesky_b200.rpqr import accelerated_rpqr, rpqr
    spec = RPQR_CASES[name]
    gold = load(name)
    B = rpqr_inputs(spec)
    with replay(spec["seed"]):
        Q, F, idx = accelerated_rpqr(B, spec["k"], b=spec["b"])
    assert np.array_equal(np.asarray(idx, dtype=np.int64), gold["idx"]), "pivots must be bit-exact"
    assert relerr(Q, gold["Q"]) <= 1e-9 and relerr(F, gold["F"]) <= 1e-9
    np.testing.assert_allclose(Q.T @ Q, np.eye(spec["k"]), atol=1e-12)
    # auto b (wall-clock dependent, like the reference): structural properties only
    Q2, F2, idx2 = rpqr(B, spec["k"])
    np.testing.assert_allclose(Q2.T @ Q2, np.eye(spec["k"]), atol=1e-12)
    np.testing.assert_allclose(F2, B.T @ Q2, rtol=1e-10, atol=1e-12)
    assert len(set(int(i) for i in idx2)) == spec["k"]


def test_rpqr_large_against_oracle(rpc):
    from oracle import rpqr_oracle
    from randomly_pivoted_cholesky_b200.rpqr import accelerated_rpqr
    rng0 = np.random.default_rng(9)
    m, n, k, b = 300, 20000, 120, 40
    B = rng0.standard_normal((m, 60)) @ rng0.standard_normal((60, n)) + 0.05 * rng0.standard_normal((m, n))
    with replay(5):
        Q, F, idx = accelerated_rpqr(B, k, b=b)
    rng, legacy = orc.make_streams(5)
    Qw, Fw, idxw, _ = rpqr_oracle.accelerated_rpqr(B, k, b, rng, legacy)
    assert np.array_equal(np.asarray(idx, dtype=np.int64), np.asarray(idxw, dtype=np.int64))
    assert relerr(Q, Qw) <= 1e-8 and relerr(F, Fw) <= 1e-8


# ---- Nystrom-preconditioned CG --------------------------------------------------------
@pytest.mark.parametrize("kernel,kw,bw", [("gaussian", {}, 2.0), ("laplace", {}, 6.0), ("matern", {"nu": 1.5}, 2.5)])
def test_nystrom_pcg_matches_oracle(rpc, kernel, kw, bw):
    from make_golden_pcg import pcg_precond_inputs
    from oracle import pcg_oracle
    from randomly_pivoted_cholesky_b200.pcg import NystromPreconditioner, kernel_matvec, pcg
    import torch
    rng0 = np.random.default_rng(21)
    n, d, k, b, mu = 1500, 5, 300, 50, 0.05
    X = rng0.standard_normal((n, d))
    y = np.sin(X[:, 0]) + 0.1 * rng0.standard_normal(n)
    A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=bw, **kw)
    with replay(31):
        lra = rpc.accelerated_rpcholesky(A, k, b=b)
    A0 = orc.OracleKernelMatrix(X, kernel=kernel, bandwidth=bw, **kw)
    K = A0.rows(np.arange(n))
    # the blocked device mat-vec against the dense product
    p = rng0.standard_normal(n)
    got = kernel_matvec(A, torch.from_numpy(p).to(A.device), block_rows=200).cpu().numpy()
    assert relerr(got, K @ p) <= 1e-12
    # preconditioner: the reference's eigen-form solve (lra.py:78-80) on the fixture, and the oracle on this factor
    Gf, vf, muf = pcg_precond_inputs()
    pre_f = NystromPreconditioner(torch.from_numpy(Gf).to(A.device), muf)
    assert relerr(pre_f(torch.from_numpy(vf).to(A.device)).cpu().numpy(), load("pcg_precond")["out"]) <= 1e-9
    pre = NystromPreconditioner(lra, mu)
    v = rng0.standard_normal(n)
    assert relerr(pre(torch.from_numpy(v).to(A.device)).cpu().numpy(), pcg_oracle.nystrom_preconditioner(lra.G, mu)(v)) <= 1e-9
    # the solve, iterate for iterate
    x, resid = pcg(A, y, mu, pre, tol=1e-10, max_iters=80, block_rows=256)
    xw, residw = pcg_oracle.pcg(K, y, mu, pcg_oracle.nystrom_preconditioner(lra.G, mu), tol=1e-10, max_iters=80)
    assert resid[-1] < 1e-10 and abs(len(resid) - len(residw)) <= 1
    q = min(8, len(resid), len(residw))
    np.testing.assert_allclose(resid[:q], residw[:q], rtol=1e-6)          # CG amplifies round-off later on
    exact = np.linalg.solve(K + mu * np.eye(n), y)
    assert relerr(x, exact) <= 1e-8 and relerr(xw, exact) <= 1e-8
    x0, resid0 = pcg(A, y, mu, None, tol=1e-10, max_iters=80)
    assert len(resid) < len(resid0) or resid[-1] < resid0[-1]              # the preconditioner pays


@pytest.mark.parametrize("name", ["pcg_loop_gauss", "pcg_loop_laplace", "pcg_loop_noprecond"])
def test_pcg_loop_matches_the_reference_statements(rpc, name):
    """The whole Nystrom-PCG loop against the iterates of the reference's OWN statements (block_experiments/pes_code.py:463-524,
    executed unchanged at fixture-generation time): same iteration count, same residual history, same solution."""
    from make_golden_pcg import PCG_CASES, pcg_loop_inputs
    from randomly_pivoted_cholesky_b200.pcg import NystromPreconditioner, pcg
    spec, gold = PCG_CASES[name], load(name)
    X, y = pcg_loop_inputs(spec)
    A = rpc.KernelMatrix(X, kernel=spec["kernel"], bandwidth=spec["bw"], **spec["kw"])
    pre = None
    if spec["k"] > 0:
        with replay(spec["seed"]):
            lra = rpc.accelerated_rpcholesky(A, spec["k"], b=spec["b"])
        pre = NystromPreconditioner(lra, spec["mu"])
        probe = np.cos(np.arange(spec["N"]) * 0.37)
        assert relerr(pre(torch.from_numpy(probe).to(A.device)).cpu().numpy(), gold["precond_probe"]) <= 1e-9
    x, resid = pcg(A, y, spec["mu"], pre, tol=spec["tol"], max_iters=spec["max_iters"])
    # CG amplifies round-off from iterate to iterate (the mat-vec sums in another order than numpy's): the history is compared
    # digit for digit while the residual is large, then by convergence behaviour; the solution itself to 1e-7
    want = gold["resid"]
    assert abs(len(resid) - len(want)) <= 1
    q = min(len(resid), len(want))
    np.testing.assert_allclose(resid[:8], want[:8], rtol=1e-6)
    big = want[:q] > 1e-3
    np.testing.assert_allclose(np.asarray(resid[:q])[big], want[:q][big], rtol=1e-4)
    assert np.all(np.abs(np.log10(np.asarray(resid[:q])) - np.log10(want[:q])) <= 1.0)
    assert relerr(x, gold["x"]) <= 1e-7


@pytest.mark.parametrize("k,n", [(1, 1000), (37, 5000), (128, 20000), (300, 33331), (1000, 100000)])
def test_factor_row_products_match_numpy(rpc, k, n):
    """rpc_syrk_rows (DMMA + tensor-map TMA), rpc_rows_dot, rpc_rows_tdot against numpy on a padded (k, n_pad) factor whose pad
    columns hold garbage (they must not enter: the tensor map ends at n columns)."""
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(k + n)
    n_pad = (n + 127) // 128 * 128
    Rh = rng.standard_normal((k, n_pad))
    R = torch.from_numpy(Rh).cuda()
    y, w = rng.standard_normal(n), rng.standard_normal(k)
    M = ops.syrk_rows(R, n).cpu().numpy()
    want = Rh[:, :n] @ Rh[:, :n].T
    assert relerr(M, want) <= 1e-13 and np.array_equal(M, M.T)
    assert relerr(ops.rows_dot(R, torch.from_numpy(y).cuda(), n).cpu().numpy(), Rh[:, :n] @ y) <= 1e-13
    assert relerr(ops.rows_tdot(R, torch.from_numpy(w).cuda(), n).cpu().numpy(), Rh[:, :n].T @ w) <= 1e-13
    # deterministic: fixed-order reductions
    assert np.array_equal(ops.syrk_rows(R, n).cpu().numpy(), M)


@pytest.mark.parametrize("k,m", [(1, 1), (31, 2), (32, 1), (100, 3), (1000, 1)])
def test_chol_solve_matches_scipy(rpc, k, m):
    from randomly_pivoted_cholesky_b200 import ops
    import scipy.linalg
    rng = np.random.default_rng(k)
    B = rng.standard_normal((k, k + 5))
    S = B @ B.T + k * np.eye(k)
    L = np.linalg.cholesky(S)
    b = rng.standard_normal((k, m)) if m > 1 else rng.standard_normal(k)
    got = ops.chol_solve(torch.from_numpy(L).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, scipy.linalg.solve(S, b, assume_a="pos"), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("kernel,kw,d", [("gaussian", {}, 50), ("gaussian", {}, 3), ("matern", {"nu": 2.5}, 7), ("laplace", {}, 12),
                                         ("laplace", {}, 750), ("gaussian", {"extra_stability": True}, 9)])
@pytest.mark.parametrize("m", [1, 40, 248, 300, 611])
def test_fused_kernel_times_vector(rpc, kernel, kw, d, m):
    """rpc_kernel_rows_wsum (no kernel value written to memory) against the oracle's dense block: both evaluator families, one and
    several row blocks, the chunked fallback for d = 750, accumulation into an existing vector."""
    from randomly_pivoted_cholesky_b200 import ops
    rng = np.random.default_rng(d * 1000 + m)
    n = 3000 if d < 100 else 900
    X = rng.standard_normal((n, d))
    bw = float(d) if kernel == "laplace" else float(np.sqrt(d))
    A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=bw, **kw)
    idx = rng.integers(0, n, m)
    w = rng.standard_normal(m)
    Kb = orc.OracleKernelMatrix(X, kernel=kernel, bandwidth=bw, **kw).rows(idx)          # (m, n)
    piv = A._dev_new_payload(m)
    A._dev_gather(torch.as_tensor(idx, device=A.device), m, piv)
    wd = torch.from_numpy(w).to(A.device)
    got = ops.kernel_weighted_sum(A, piv.X, piv.norm, m, wd)
    assert relerr(got[:n].cpu().numpy(), w @ Kb) <= 1e-12
    base = torch.from_numpy(rng.standard_normal(A.n_pad)).to(A.device)
    acc = ops.kernel_weighted_sum(A, piv.X, piv.norm, m, wd, out=base.clone(), accumulate=True)
    assert relerr(acc[:n].cpu().numpy(), base[:n].cpu().numpy() + w @ Kb) <= 1e-12


@pytest.mark.parametrize("kernel,n,d", [("gaussian", 1000, 5), ("laplace", 129, 12), ("matern", 256, 3)])
def test_padding_contract_is_checkable(rpc, kernel, n, d):
    """rpc_check_padding: the operators the host builds satisfy the padding contract of include/rpc_b200.h, and a binding that
    breaks it (non-zero pad feature, pad point, pad diagonal entry, NaN) is caught."""
    X = np.random.default_rng(n).standard_normal((n, d))
    A = rpc.KernelMatrix(X, kernel=kernel, bandwidth=1.5, **({"nu": 1.5} if kernel == "matern" else {}))
    diag = torch.zeros((A.n_pad,), dtype=torch.float64, device=A.device)
    A._dev_diag_into(diag)
    assert A.padding_violations(diag) == 0
    if A.ldx > d:
        A._Xd[3, d] = 1e-300
        assert A.padding_violations() == 1
        A._Xd[3, d] = 0.0
    if A.n_pad > n:
        A._Xd[n, 0] = float("nan")
        diag[A.n_pad - 1] = 2.0
        assert A.padding_violations(diag) == 2
        A._Xd[n, 0] = 0.0
        assert A.padding_violations() == 0
