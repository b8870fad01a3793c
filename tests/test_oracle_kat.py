"""Known-answer tests that pin the CPU oracle (the reference ships no numeric goldens and R is not
installed here -- SURVEY.md 8c).  Every expected value below is derived by hand from the R
semantics the reference relies on."""
import json
import os

import numpy as np
import pytest

import oracle
from helpers import design, markers

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_gb_2x3_by_hand():
    X = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    # tcrossprod(X)/3 = [[14,32],[32,77]]/3
    assert np.array_equal(oracle.gb_kernel(X), np.array([[14.0, 32.0], [32.0, 77.0]]) / 3.0)


def test_gk_three_points_by_hand():
    X = np.array([[0.0, 0.0], [3.0, 4.0], [6.0, 8.0]])        # distances 5, 10, 5
    D = oracle.gk_distance(X)
    assert np.array_equal(D, np.array([[0.0, 25.0, 100.0], [25.0, 0.0, 25.0], [100.0, 25.0, 0.0]]))
    # all 9 entries sorted: 0,0,0,25,25,25,25,100,100 ; median (type 7, h = 4) = 25
    assert oracle.quantile7(D, 0.5) == 25.0
    K = oracle.gk_kernel(D, 1.0, 25.0)
    assert np.all(np.diag(K) == 1.0)
    assert K[0, 1] == np.exp(-1.0) and K[0, 2] == np.exp(-4.0)
    K2 = oracle.gk_kernel(D, 0.5, 25.0)
    assert K2[0, 2] == np.exp(-2.0)


def test_quantile_type7_interpolation():
    # R: quantile(c(1,2,3,4), 0.4) = 2.2 ; quantile(1:10, 0.33) = 3.97 ; ends are min / max
    assert oracle.quantile7([4, 1, 3, 2], 0.4) == pytest.approx(2.2, abs=1e-15)
    assert oracle.quantile7(np.arange(1, 11), 0.33) == pytest.approx(3.97, abs=1e-14)
    assert oracle.quantile7([5, 7, 9], 0.0) == 5 and oracle.quantile7([5, 7, 9], 1.0) == 9
    x = np.random.default_rng(0).gamma(2, 3, size=1001)
    for pr in (0.1, 0.5, 0.77):
        assert oracle.quantile7(x, pr) == pytest.approx(np.quantile(x, pr), rel=1e-14)


def test_expansion_is_incidence_product():
    rng = np.random.default_rng(1)
    L, n = 7, 19
    K0 = rng.standard_normal((L, L))
    gid = rng.integers(0, L, n)
    Zg = np.zeros((n, L))
    Zg[np.arange(n), gid] = 1.0
    assert np.array_equal(oracle.expand(K0, gid), Zg @ (K0 @ Zg.T))     # single non-zero term: exact


def test_balanced_kronecker_structure():
    L, E, p = 9, 3, 20
    X = markers(L, p, 3)
    Y, names = design(L, E)
    K0 = oracle.gb_kernel(X)
    K = oracle.getK(Y, X, kernel="GB", model="MDs", rownames=names)
    assert list(K) == ["G", "GE"] and K["G"]["Type"] == "D" and K["GE"]["Type"] == "BD"
    assert np.array_equal(K["G"]["Kernel"], np.kron(np.ones((E, E)), K0))
    assert np.array_equal(K["GE"]["Kernel"], np.kron(np.eye(E), K0))
    Ke = oracle.getK(Y, X, kernel="GB", model="MDe", rownames=names)
    assert list(Ke) == ["G", "env01", "env02", "env03"]
    for k in range(E):
        M = Ke[f"env{k + 1:02d}"]["Kernel"]
        mask = np.zeros((E, E))
        mask[k, k] = 1
        assert np.array_equal(M, np.kron(mask, K0))
        assert np.mean(np.diag(M)) == pytest.approx(np.mean(np.diag(K0)) * L / (L * E))


def test_names_and_errors_follow_reference():
    L, E = 6, 2
    X = markers(L, 10, 0)
    Y, names = design(L, E)
    K = oracle.getK(Y, X, kernel="GK", model="MDe", bandwidth=[1, 2], rownames=names, intercept_random=True)
    assert list(K) == ["G_1", "G_2", "env01_1", "env02_1", "env01_2", "env02_2", "Gi"]      # R/getK.R:185,221,233
    assert np.array_equal(K["Gi"]["Kernel"], np.kron(np.ones((E, E)), np.eye(L)))
    with pytest.raises(ValueError, match="Single model"):
        oracle.getK(Y, X, kernel="GB", model="SM", rownames=names)                          # test.R:46-47
    Y1, _ = design(L, 1)
    with pytest.raises(ValueError):
        oracle.getK(Y1, X, kernel="GB", model="MM", rownames=names)                         # test.R:16-17
    with pytest.raises(ValueError, match="Genotype names are missing"):
        oracle.getK(Y, X, kernel="GB", model="MM")
    with pytest.raises(ValueError, match="Not all genotypes"):
        oracle.getK(Y, X[:-1], kernel="GB", model="MM", rownames=names[:-1])
    with pytest.raises(ValueError, match="kernel selected is not available"):
        oracle.getK(Y, X, kernel="RBF", model="MM", rownames=names)
    with pytest.raises(ValueError, match="Model selected is not available"):
        oracle.getK(Y, X, kernel="GB", model="Cov", rownames=names)


def test_rank_of_kernels():
    L, p = 20, 60
    X = markers(L, p, 4)
    s, _ = oracle.eig_block(oracle.gb_kernel(X), 1e-10)
    assert s.shape[0] == L - 1                      # centred markers: rank L-1
    D = oracle.gk_distance(X)
    s, _ = oracle.eig_block(oracle.gk_kernel(D, 1.0, oracle.quantile7(D, 0.5)), 1e-10)
    assert s.shape[0] == L and np.all(np.diff(s) <= 0)     # full rank, descending


def test_setdec_block_handling():
    L, E = 5, 3
    K0 = np.eye(L) + 0.5
    Kbd = np.kron(np.diag([1.0, 0.0, 1.0]), K0)            # middle block has no eigenvalue > tol -> skipped (:162)
    Kbd[0, L * 2] = 123.0                                   # off-block entries are ignored (:160)
    dec = oracle.setDEC({"GE": {"Kernel": Kbd, "Type": "BD"}}, [L] * E, 1e-10)
    assert [b["pos"] for b in dec[0]] == [(0, L), (2 * L, 3 * L)]
    for b in dec[0]:
        assert np.allclose((b["U"] * b["s"]) @ b["U"].T, K0, atol=1e-12)
    with pytest.raises(ValueError):
        oracle.setDEC({"GE": {"Kernel": Kbd, "Type": "BD"}}, None, 1e-10)
    with pytest.raises(ValueError):
        oracle.setDEC({"GE": {"Kernel": Kbd, "Type": "XX"}}, [L] * E, 1e-10)


def test_zero_tape_first_iteration_by_hand():
    """z = 0 and chisq = df make the sampler a deterministic fixed-point map; check iteration 1 on
    K = I (U = I, s = 1) against the closed form from R/BGGE.R:245-367."""
    n = 4
    y = np.array([1.0, 2.0, 4.0, 7.0])
    K = {"K1": {"Kernel": np.eye(n), "Type": "D"}}
    eig = [[{"s": np.ones(n), "U": np.eye(n), "pos": None}]]
    nz, nc = oracle.tape_sizes(n, [n], 1)
    chis = np.array([n + 3.0, n + 3.0])
    tr = {"keep_u": True}
    oracle.BGGE(y, K, ite=1, burn=0, thin=1, draws=oracle.TapeDraws(np.zeros(nz), chis), eig=eig, trace=tr)
    mu0 = y.mean()
    vy = y.var(ddof=1)
    nu = 3.0
    Sce, Sc = (nu + 2) * 0.5 * vy, (nu + 2) * 0.5 * vy / 1.0
    # mu step with z=0: mu = mean(temp + mu0) = mu0 ; kernel step: v = 0.2/(1+0.2*0.01), b = tau*v*(y-mu0)
    v = 0.2 / (1 + 0.2 * 0.01)
    b = 0.01 * v * (y - mu0)
    assert tr["mu"][0] == pytest.approx(mu0, rel=1e-15)
    assert np.allclose(tr["u"][0][0], b, rtol=1e-14)
    assert tr["varU"][0][0] == pytest.approx((np.sum(b * b) + nu * Sc) / (n + nu), rel=1e-14)
    e = (y - mu0) - b
    assert tr["varE"][0] == pytest.approx((e @ e + Sce) / (n + nu), rel=1e-14)


def test_default_chain_counts_and_fields():
    L, E = 8, 2
    X = markers(L, 12, 0)
    Y, names = design(L, E)
    K = oracle.getK(Y, X, kernel="GB", model="MM", rownames=names)
    y = np.random.default_rng(0).standard_normal(L * E)
    fit = oracle.BGGE(y, K, ne=[L] * E, draws=oracle.RandomDraws(1))
    assert list(fit)[:9] == ["yHat", "varE", "varE.sd", "K", "chain", "ite", "burn", "thin", "y"]   # "List of 9"
    assert fit["chain"]["mu"].shape == (333,)
    its = np.arange(1, 1001)
    assert int(np.sum(its[its % 3 == 0] > 200)) == 267
    nz, nc = oracle.tape_sizes(L * E, [L - 1], 1000)
    assert (nz, nc) == (1 * L * E + 1000 * (1 + (L - 1)), 2000)


def test_golden_fixture_regression():
    """Committed fixture (tests/golden/make_golden.py): guards the oracle itself against drift."""
    with open(os.path.join(GOLD, "oracle_small.json")) as f:
        g = json.load(f)
    X = np.array(g["X"])
    Y = [tuple(r) for r in g["Y"]]
    K = oracle.getK(Y, X, kernel="GK", model="MDs", rownames=g["names"])
    assert np.allclose(K["G"]["Kernel"], np.array(g["G"]), rtol=0, atol=1e-13)
    assert np.allclose(K["GE"]["Kernel"], np.array(g["GE"]), rtol=0, atol=1e-13)
    eig = [[{"s": np.array(b["s"]), "U": np.array(b["U"]), "pos": None if b["pos"] is None else tuple(b["pos"])}
            for b in blocks] for blocks in g["eig"]]
    fit = oracle.BGGE(np.array(g["y"], dtype=float), K, ne=g["ne"], ite=g["ite"], burn=g["burn"], thin=g["thin"],
                      draws=oracle.TapeDraws(np.array(g["z"]), np.array(g["c"])), eig=eig)
    assert np.allclose(fit["yHat"], np.array(g["yHat"]), rtol=1e-11)
    assert np.allclose(fit["chain"]["varE"], np.array(g["chain_varE"]), rtol=1e-11)
    assert np.allclose(fit["chain"]["K"]["GE"]["varU"], np.array(g["chain_varU_GE"]), rtol=1e-11)
