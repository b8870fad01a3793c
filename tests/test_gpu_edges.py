"""Edge cases the reference's behaviour implies: minimal sizes, ragged shapes, all-but-two phenotypes
missing, kernels of rank 1, blocks dropped by tol, errors surfaced through the ABI."""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import design, markers, relmax

pytestmark = pytest.mark.gpu


def test_minimal_and_ragged_getk(lib):
    import bgge_b200
    for L, p, E in [(2, 1, 2), (3, 17, 2), (127, 15, 2), (129, 1, 3)]:
        rng = np.random.default_rng(L)
        X = rng.standard_normal((L, p))
        Y, names = design(L, E)
        for kernel in ("GB", "GK"):
            got = bgge_b200.getK(Y, X, kernel=kernel, model="MDs", rownames=names)
            ref = oracle.getK(Y, X, kernel=kernel, model="MDs", rownames=names)
            for nm in ref:
                assert relmax(got[nm]["Kernel"], ref[nm]["Kernel"]) <= 1e-12, (L, p, kernel, nm)


def test_gk_with_identical_lines_has_exact_ones(lib):
    """Duplicate marker rows: distance exactly 0 off the diagonal too -> kernel entry exactly 1."""
    import bgge_b200
    X = markers(20, 33, 1)
    X[7] = X[3]
    Y, names = design(20, 2)
    K = bgge_b200.getK(Y, X, kernel="GK", model="MM", rownames=names)["G"]["Kernel"]
    assert K[3, 7] == 1.0 and K[7, 3] == 1.0 and K[3 + 20, 7] == 1.0


def test_rank_one_and_dropped_blocks():
    import bgge_b200
    n = 24
    rng = np.random.default_rng(2)
    v = rng.standard_normal(8)
    blk = np.outer(v, v)                                       # rank 1 -> a single kept eigenpair (R drops U to a vector)
    Kbd = np.zeros((n, n))
    Kbd[:8, :8] = blk
    Kbd[16:, 16:] = blk * 2                                    # middle block all zero -> dropped (R/BGGE.R:162)
    K = {"A": {"Kernel": np.eye(n) * 0.5 + 0.5, "Type": "D"}, "B": {"Kernel": Kbd, "Type": "BD"}}
    ne = [8, 8, 8]
    Ei = bgge_b200.setDEC(K, ne, 1e-10)
    assert [b["pos"] for b in Ei[1]] == [(0, 8), (16, 24)] and all(b["s"].shape == (1,) for b in Ei[1])
    y = rng.standard_normal(n)
    ite = 15
    nz, nc = oracle.tape_sizes(n, [n, 2], ite)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([n + 3, 2 + 3, n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, K, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    for mode in (1, 2, 0):
        got = bgge_b200.BGGE(y, K, ne=ne, ite=ite, burn=0, thin=1, draws=(z, c), eig=Ei, mode=mode)
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9
        assert relmax(got["K"]["B"]["u"], ref["K"]["B"]["u"]) <= 1e-9
        assert np.all(got["K"]["B"]["u"][8:16] == 0.0)          # rows of the dropped block stay exactly zero


def test_almost_everything_missing():
    import bgge_b200
    L, E = 12, 2
    X = markers(L, 20, 3)
    Y, names = design(L, E)
    K = oracle.getK(Y, X, kernel="GB", model="MM", rownames=names)
    y = np.full(L * E, np.nan)
    y[[1, 17]] = [0.3, -1.2]                                   # two observations: var(y) still defined
    Ei = bgge_b200.setDEC(K, None, 1e-10)
    ite = 10
    nz, nc = oracle.tape_sizes(L * E, [Ei[0][0]["s"].shape[0]], ite, n_na=L * E - 2)
    rng = np.random.default_rng(4)
    z, c = rng.standard_normal(nz), rng.chisquare(np.tile(np.array([Ei[0][0]["s"].shape[0] + 3, L * E + 3], dtype=float), ite))
    ref = oracle.BGGE(y, K, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    got = bgge_b200.BGGE(y, K, ite=ite, burn=0, thin=1, draws=(z, c), eig=Ei)
    assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9
    assert relmax(got["y"], ref["y"]) <= 1e-9


def test_abi_errors_are_reported_not_thrown(lib):
    h = C.c_void_p()
    lib.call("bgge_gibbs_create", 10, 1, 0, None, C.byref(h))
    with pytest.raises(lib.BGGEError, match="run before init"):
        lib.call("bgge_gibbs_run", h, 1)
    with pytest.raises(lib.BGGEError, match="y not set|no eigen blocks"):
        lib.call("bgge_gibbs_init", h, 10, 0, 1, 0.5, 0, 0, 0)
    s = np.array([1.0, -1.0])
    U = np.zeros((10, 2), order="F")
    with pytest.raises(lib.BGGEError, match="not positive"):
        lib.call("bgge_gibbs_add_block", h, 0, 0, 10, 2, lib.ptr(s), lib.ptr(U), 10, 0)
    with pytest.raises(lib.BGGEError, match="bad block"):
        lib.call("bgge_gibbs_add_block", h, 0, 5, 10, 2, lib.ptr(np.ones(2)), lib.ptr(U), 10, 0)
    lib.load().bgge_gibbs_destroy(h)
    # tape too short
    import bgge_b200
    K = {"A": {"Kernel": np.eye(6), "Type": "D"}}
    with pytest.raises(lib.BGGEError, match="tape too short"):
        bgge_b200.BGGE(np.arange(6.0), K, ite=5, draws=(np.zeros(10), np.ones(3)))
    with pytest.raises(lib.BGGEError, match="outside"):
        out = np.empty((4, 4), order="F")
        lib.call("bgge_getk_expand", lib.ptr(np.eye(2)), 2, lib.ptr(np.array([0, 1, 2, 0], dtype=np.int32)), None, 4, 0, 0,
                 lib.ptr(out))
