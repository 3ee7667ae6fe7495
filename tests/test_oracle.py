"""The oracle (oracle/rpc_oracle.py) against the golden fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from make_golden import CASES, KERNEL_CASES, golden_inputs, kernel_inputs
from oracle import rpc_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def relerr(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# block greedy with the 'regularize' strategy: numpy's argpartition order inside a block is an implementation detail of
# its introselect, so parity is the pivot SET of every block plus the quantities that do not depend on the order
# inside a block (G^T G, hence the residual and the trace)
SET_PARITY = {"bgreedy_dense"}


def check_block_set_parity(spec, idx, G, gold, tol=1e-10):
    idx = np.asarray(idx, dtype=np.int64)
    assert idx.shape == gold["idx"].shape and G.shape == gold["G"].shape
    for c in range(0, spec["k"], spec["b"]):
        assert np.array_equal(np.sort(idx[c:c + spec["b"]]), np.sort(gold["idx"][c:c + spec["b"]])), "pivot set of block %d" % c
    assert relerr(G.T @ G, gold["G"].T @ gold["G"]) <= tol
    assert abs(np.sum(G * G) - float(gold["trace"])) <= 1e-11 * abs(float(gold["trace"]))


def run_oracle(spec):
    inp = golden_inputs(spec)
    A = orc.OracleDenseMatrix(inp["dense"]) if "dense" in inp else orc.OracleKernelMatrix(
        inp["X"], kernel=inp["kernel"], bandwidth=inp["bandwidth"], **inp["kw"])
    rng, legacy = orc.make_streams(spec["seed"])
    kw = {"stoptol": spec["stoptol"]} if "stoptol" in spec else {}
    if spec["alg"] in ("simple", "greedy", "rgreedy"):
        res = orc.simple_rpcholesky(A, spec["k"], rng, alg={"simple": "rp"}.get(spec["alg"], spec["alg"]), **kw)
    elif spec["alg"] in ("block", "bgreedy"):
        res = orc.block_rpcholesky(A, spec["k"], spec["b"], rng, alg="rp" if spec["alg"] == "block" else "greedy", **kw,
                                   **spec.get("blockkw", {}))
    else:
        res = orc.accelerated_rpcholesky(A, spec["k"], spec["b"], rng, legacy, **kw)
    return A, res


@pytest.mark.parametrize("name", sorted(CASES))
def test_factorisation_matches_reference(name):
    spec = CASES[name]
    gold = load(name)
    A, res = run_oracle(spec)
    if name in SET_PARITY:
        check_block_set_parity(spec, res.idx, res.G, gold)
        return
    assert np.array_equal(np.asarray(res.idx, dtype=np.int64), gold["idx"])          # pivots bit-exact
    assert res.G.shape == gold["G"].shape and res.rows.shape == gold["rows"].shape
    assert relerr(res.G, gold["G"]) <= 1e-12
    assert relerr(res.rows, gold["rows"]) <= 1e-13
    assert abs(res.trace() - float(gold["trace"])) <= 1e-12 * abs(float(gold["trace"]))
    if "queries" in gold:
        assert A.num_queries() == int(gold["queries"])


@pytest.mark.parametrize("name", sorted(KERNEL_CASES))
def test_kernel_values_match_reference(name):
    spec = KERNEL_CASES[name]
    gold = load(name)
    X, idx, kw = kernel_inputs(spec)
    A = orc.OracleKernelMatrix(X, kernel=spec["kernel"], bandwidth=spec["bw"], **kw)
    # same BLAS/numpy calls in the same order -> expect (near) bit equality
    np.testing.assert_allclose(A.rows(idx), gold["rows"], rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(A.block(idx), gold["block"], rtol=1e-14, atol=1e-16)
    np.testing.assert_array_equal(A.diag(), gold["diag"])
    np.testing.assert_allclose(A.rows(int(idx[3]))[0], gold["one_row"], rtol=1e-14, atol=1e-16)


def test_manhattan_is_sequential_cdist():
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(3)
    X = rng.standard_normal((9, 50)); Y = rng.standard_normal((200, 50))
    assert np.array_equal(orc.manhattan_distances_loop(X, Y), cdist(X, Y, "cityblock"))
    assert np.array_equal(orc.manhattan_distances(X, Y), cdist(X, Y, "cityblock"))


def test_euclidean_matches_sklearn():
    from sklearn.metrics.pairwise import euclidean_distances
    rng = np.random.default_rng(4)
    X = rng.standard_normal((11, 20)); Y = rng.standard_normal((300, 20))
    assert np.array_equal(orc.euclidean_distances(X, Y), euclidean_distances(X, Y))


@pytest.mark.parametrize("tag", ["uniform", "random", "sparse"])
def test_sampler_is_generator_choice(tag):
    gold = load("sampler")
    idx = orc.sample_pivots(gold[tag + "_diag"], gold[tag + "_u"])
    assert np.array_equal(idx, gold[tag + "_idx"])


def test_sampler_live_against_numpy():
    rng = np.random.default_rng(8)
    for n in (1, 2, 100, 5000):
        diag = rng.random(n)
        a = np.random.Generator(np.random.PCG64(1)).choice(range(n), size=64, p=diag / sum(diag), replace=True)
        b = orc.sample_pivots(diag, np.random.Generator(np.random.PCG64(1)).random(64))
        assert np.array_equal(a, b)
        # scalar draw = one double of the stream
        g1 = np.random.Generator(np.random.PCG64(2)); g2 = np.random.Generator(np.random.PCG64(2))
        assert g1.choice(range(n), p=diag / sum(diag)) == orc.sample_pivots(diag, [g2.random()])[0]
        assert g1.random() == g2.random()


def test_seq_sum_is_builtin_sum():
    rng = np.random.default_rng(9)
    for n in (1, 17, 1000, 100003):
        v = rng.random(n) ** 3
        assert orc.seq_sum(v) == sum(v)


def test_sampler_raises_like_choice():
    with pytest.raises(ValueError):
        orc.sample_pivots(np.zeros(10), [0.5])
    with pytest.raises(ValueError):
        orc.sample_pivots(np.array([1.0, np.nan]), [0.5])


def test_rejection_cholesky_properties():
    rng = np.random.default_rng(10)
    B = rng.standard_normal((30, 12)); H = B @ B.T      # rank 12 < 30
    H0 = H.copy()
    L, acc = orc.rejection_cholesky(H, np.zeros(30))    # u = 0 -> accept while the residual is positive
    assert len(acc) >= 12
    L2, acc2 = orc.rejection_cholesky(H0.copy(), rng.random(30))
    np.testing.assert_allclose(L2 @ L2.T, H0[np.ix_(acc2, acc2)], atol=1e-10)
    assert np.allclose(L2, np.tril(L2))
    with pytest.raises(RuntimeError):
        orc.rejection_cholesky(-np.eye(3), np.zeros(3))
    with pytest.raises(RuntimeError):
        orc.rejection_cholesky(np.ones((2, 3)), np.zeros(2))


# ---- first "next" row: KRR_Nystrom (oracle/krr_oracle.py vs the live reference) ---------------------------------
from make_golden_krr import KRR_CASES, krr_inputs  # noqa: E402
from oracle import krr_oracle  # noqa: E402


@pytest.mark.parametrize("name", sorted(KRR_CASES))
def test_krr_oracle_matches_reference(name):
    spec = KRR_CASES[name]
    gold = load(name)
    Xtr, ytr, Xts, _ = krr_inputs(spec)
    sol, idx, err, queries = krr_oracle.fit_nystrom(Xtr, ytr, spec["lamb"], spec["k"], spec["kernel"], spec["bw"], spec["alg"],
                                                    spec["b"], spec["seed"])
    assert np.array_equal(idx, gold["sample_idx"])
    assert queries == int(gold["queries"])
    assert abs(err - float(gold["reltrace_err"])) <= 1e-13
    np.testing.assert_allclose(sol, gold["sol"], rtol=1e-7, atol=1e-9)
    preds = krr_oracle.predict_nystrom(Xts, Xtr, idx, sol, spec["kernel"], spec["bw"])
    np.testing.assert_allclose(preds, gold["preds"], rtol=1e-8, atol=1e-10)


# ---- accelerated RPQR -----------------------------------------------------------------
from make_golden_rpqr import RPQR_CASES, rpqr_inputs  # noqa: E402


@pytest.mark.parametrize("name", sorted(RPQR_CASES))
def test_rpqr_oracle_matches_reference(name):
    from oracle import rpqr_oracle
    spec = RPQR_CASES[name]
    gold = load(name)
    B = rpqr_inputs(spec)
    rng, legacy = orc.make_streams(spec["seed"])
    Q, F, idx, _ = rpqr_oracle.accelerated_rpqr(B, spec["k"], spec["b"], rng, legacy)
    assert np.array_equal(np.asarray(idx, dtype=np.int64), gold["idx"])
    assert relerr(Q, gold["Q"]) <= 1e-12 and relerr(F, gold["F"]) <= 1e-12


# ---- Nystrom-preconditioned CG ---------------------------------------------------------
def test_pcg_oracle_preconditioner_matches_reference_krr():
    from make_golden_pcg import pcg_precond_inputs
    from oracle import pcg_oracle
    G, v, mu = pcg_precond_inputs()
    out = pcg_oracle.nystrom_preconditioner(G, mu)(v)
    assert relerr(out, load("pcg_precond")["out"]) <= 1e-10
    # it is the inverse of G^T G + mu I
    np.testing.assert_allclose((G.T @ G + mu * np.eye(G.shape[1])) @ out, v, rtol=1e-8, atol=1e-8)


def test_pcg_oracle_solves_the_system():
    from oracle import pcg_oracle
    rng = np.random.default_rng(8)
    X = rng.standard_normal((200, 4))
    A = orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=2.0)
    K = A.rows(np.arange(200))
    y = rng.standard_normal(200)
    mu = 1e-2
    rs, lg = orc.make_streams(3)
    res = orc.accelerated_rpcholesky(orc.OracleKernelMatrix(X, kernel="gaussian", bandwidth=2.0), 40, 10, rs, lg)
    x, resid = pcg_oracle.pcg(K, y, mu, pcg_oracle.nystrom_preconditioner(res.G, mu), tol=1e-10, max_iters=200)
    np.testing.assert_allclose(x, np.linalg.solve(K + mu * np.eye(200), y), rtol=1e-6, atol=1e-8)
    x0, resid0 = pcg_oracle.pcg(K, y, mu, None, tol=1e-10, max_iters=200)
    assert len(resid) < len(resid0)            # the preconditioner pays


def pcg_case_operands(spec):
    """Dense kernel matrix and Nystrom factor of a PCG fixture case, from the oracle (pinned to the reference elsewhere in this file)."""
    from make_golden_pcg import pcg_loop_inputs
    X, y = pcg_loop_inputs(spec)
    A = orc.OracleKernelMatrix(X, kernel=spec["kernel"], bandwidth=spec["bw"], **spec["kw"])
    K = A.rows(np.arange(spec["N"]))
    G = None
    if spec["k"] > 0:
        rs, lg = orc.make_streams(spec["seed"])
        G = orc.accelerated_rpcholesky(orc.OracleKernelMatrix(X, kernel=spec["kernel"], bandwidth=spec["bw"], **spec["kw"]),
                                       spec["k"], spec["b"], rs, lg).G
    return X, y, K, G


@pytest.mark.parametrize("name", ["pcg_loop_gauss", "pcg_loop_laplace", "pcg_loop_noprecond"])
def test_pcg_oracle_matches_the_reference_loop(name):
    """The CG loop is pinned: the fixtures hold the iterates of the reference's own statements (pes_code.py:463-524, executed
    unchanged by tests/golden/make_golden_pcg.py); the numpy restatement reproduces the iteration count, every residual and x."""
    from make_golden_pcg import PCG_CASES
    from oracle import pcg_oracle
    spec, gold = PCG_CASES[name], load(name)
    X, y, K, G = pcg_case_operands(spec)
    pre = pcg_oracle.nystrom_preconditioner(G, spec["mu"]) if G is not None else None
    x, resid = pcg_oracle.pcg(K, y, spec["mu"], pre, tol=spec["tol"], max_iters=spec["max_iters"])
    assert len(resid) == len(gold["resid"])
    np.testing.assert_allclose(resid, gold["resid"], rtol=1e-5, atol=1e-14)
    assert relerr(x, gold["x"]) <= 1e-7
    if pre is not None:
        probe = np.cos(np.arange(spec["N"]) * 0.37)
        assert relerr(pre(probe), gold["precond_probe"]) <= 1e-9
