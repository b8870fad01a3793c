"""Factorised eigen route: getK(..., factored=True) + bgge_gibbs_add_expanded_kernel must
describe exactly the same kernels / eigen blocks as the dense route, without ever building n x n."""
import numpy as np
import pytest

import oracle
from helpers import design, markers, phenotype, relmax

pytestmark = pytest.mark.gpu


def _designs():
    L, E = 31, 3
    yield "balanced", design(L, E), [L] * E
    # unbalanced, env-sorted: 7 cells dropped, a few genotypes repeated inside an environment
    Y, names = design(L, E, seed=3, drop=7)
    Y = sorted(Y + [Y[2], Y[2], Y[40]], key=lambda r: r[0])
    envs = [r[0] for r in Y]
    ne = [envs.count(e) for e in sorted(set(envs))]
    yield "unbalanced+repeats", (Y, names), ne
    # rows not sorted by environment: BD blocks see mixed masks -> direct fallback inside the library
    Y, names = design(L, E, seed=5, shuffle=True)
    yield "shuffled", (Y, names), [L] * E


@pytest.mark.parametrize("kernel,model", [("GB", "MDs"), ("GK", "MDe"), ("GB", "MM")])
def test_factored_blocks_reconstruct_dense_kernels(kernel, model):
    import bgge_b200
    import warnings
    for label, (Y, names), ne in _designs():
        X = markers(len(names), 45, 1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Kf = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names, factored=True, intercept_random=True)
            Kd = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names, intercept_random=True)
        assert list(Kf) == list(Kd)
        for nm in Kd:
            assert Kf[nm]["Type"] == Kd[nm]["Type"]
            assert relmax(bgge_b200.materialize(Kf[nm]), Kd[nm]["Kernel"]) <= 1e-12, (label, nm)
        Ef = bgge_b200.setDEC(Kf, ne, 1e-10)
        Ed = oracle.setDEC(Kd, ne, 1e-10)
        for nm, bf, bd in zip(Kd, Ef, Ed):
            assert [b["pos"] for b in bf] == [b["pos"] for b in bd], (label, nm)
            for f, d in zip(bf, bd):
                assert f["s"].shape == d["s"].shape, (label, nm)
                assert relmax(f["s"], d["s"]) <= 1e-10, (label, nm)
                U = f["U"]
                assert np.max(np.abs(U.T @ U - np.eye(U.shape[1]))) <= 1e-11, (label, nm)
                a, b = (0, Kd[nm]["Kernel"].shape[0]) if f["pos"] is None else f["pos"]
                blk = Kd[nm]["Kernel"][a:b, a:b]
                rec = (U * f["s"]) @ U.T
                # the truncated part (values <= tol) is what is missing from the reconstruction
                assert np.max(np.abs(rec - blk)) <= 1e-9 * max(1.0, np.max(np.abs(blk))), (label, nm)


@pytest.mark.parametrize("mode", [1, 2, 0])
@pytest.mark.parametrize("L,E", [(28, 3), (700, 4)])
def test_fit_from_factored_kernels_matches_oracle_per_iteration(L, E, mode):
    """mode 1: the G kernel stays factorised in the sweep (streams W, L rows, instead of U, L*E rows);
    mode 0: the two-pass kernels on the expanded U.  Both against the oracle on the dense kernels."""
    import bgge_b200
    p = 50
    X = markers(L, p, 2)
    Y, names = design(L, E)
    ne = [L] * E
    Kf = bgge_b200.getK(Y, X, kernel="GK", model="MDs", rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel="GK", model="MDs", rownames=names)
    y = phenotype(oracle.gb_kernel(X), L, E, 3)
    y[[5, 40]] = np.nan
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    ite = 20 if L < 100 else 6
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    nz, nc = oracle.tape_sizes(L * E, nrs, ite, n_na=2)
    rng = np.random.default_rng(4)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [L * E + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    got = bgge_b200.BGGE(y, Kf, ne=ne, ite=ite, burn=0, thin=1, draws=(z, c), mode=mode)   # decomposes inside the library
    assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9
    assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9
    for nm in ref["K"]:
        assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= 1e-9
    assert relmax(got["yHat"], ref["yHat"]) <= 1e-9


@pytest.mark.parametrize("E,drop_env", [(2, None), (6, None), (5, 2)])
def test_shared_blocks_multi_rhs_sweep(E, drop_env):
    """Identical diagonal blocks (balanced environments) share one decomposition and are swept as a
    multi-right-hand-side panel: E = 2 -> one group of 2, E = 6 -> groups of 4 + 2, and a design where one
    environment misses genotypes (its block differs, so it must NOT be shared).  Per-iteration parity against
    the oracle run on the dense kernels with the same eigen blocks."""
    import bgge_b200
    L, p = 40, 64
    X = markers(L, p, 9)
    Y, names = design(L, E)
    if drop_env is not None:
        Y = [r for k, r in enumerate(Y) if not (r[0] == f"env{drop_env + 1:02d}" and k % 7 == 0)]
    envs = [r[0] for r in Y]
    ne = [envs.count(e) for e in sorted(set(envs))]
    n = len(Y)
    Kf = bgge_b200.getK(Y, X, kernel="GB", model="MDs", rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel="GB", model="MDs", rownames=names)
    rng = np.random.default_rng(E)
    y = rng.standard_normal(n)
    y[rng.choice(n, 5, replace=False)] = np.nan
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    assert len(Ei[1]) == E
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    ite = 12
    nz, nc = oracle.tape_sizes(n, nrs, ite, n_na=5)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    for mode in (1, 2, 0):
        got = bgge_b200.BGGE(y, Kf, ne=ne, ite=ite, burn=0, thin=1, draws=(z, c), mode=mode)
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, mode
        for nm in ref["K"]:
            assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= 1e-9, (mode, nm)
            assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (mode, nm)
        assert relmax(got["yHat"], ref["yHat"]) <= 1e-9, mode


@pytest.mark.parametrize("E,drop_env,kernel", [(3, None, "GB"), (6, None, "GK"), (5, 1, "GB")])
def test_mde_environment_kernels_share_one_decomposition_and_sweep_together(E, drop_env, kernel):
    """MDe: one kernel per environment, each one block.  In a balanced trial they are the same L x L matrix: the
    handle decomposes it once and the resident kernel sweeps up to 4 of them as ONE step (they touch disjoint
    rows, so this equals the sequential sweep).  E = 6 -> steps of 4 + 2; a design where one environment misses
    genotypes (its matrix differs: not shared, not merged).  Per-iteration parity with the oracle on the dense
    kernels, for the resident (1), streaming (2) and two-pass (0) paths, plus thinning / burn-in on mode 1."""
    import bgge_b200
    L, p = 36, 48
    X = markers(L, p, 5)
    Y, names = design(L, E)
    if drop_env is not None:
        Y = [r for k, r in enumerate(Y) if not (r[0] == f"env{drop_env + 1:02d}" and k % 5 == 0)]
    envs = [r[0] for r in Y]
    ne = [envs.count(e) for e in sorted(set(envs))]
    n = len(Y)
    Kf = bgge_b200.getK(Y, X, kernel=kernel, model="MDe", rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel=kernel, model="MDe", rownames=names)
    assert len(Kf) == E + 1
    rng = np.random.default_rng(10 + E)
    y = rng.standard_normal(n)
    y[rng.choice(n, 4, replace=False)] = np.nan
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    ite = 14
    nz, nc = oracle.tape_sizes(n, nrs, ite, n_na=4)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    for burn, thin, modes in ((0, 1, (1, 2, 0)), (4, 3, (1,))):
        ref = oracle.BGGE(y, Kd, ne=ne, ite=ite, burn=burn, thin=thin, draws=oracle.TapeDraws(z, c), eig=Ei)
        for mode in modes:
            got = bgge_b200.BGGE(y, Kf, ne=ne, ite=ite, burn=burn, thin=thin, draws=(z, c), mode=mode)
            assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, mode
            assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9, mode
            for nm in ref["K"]:
                assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= 1e-9, (mode, nm)
                assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (mode, nm)
                assert relmax(got["K"][nm]["u.sd"], ref["K"][nm]["u.sd"]) <= 1e-7, (mode, nm)
            assert relmax(got["yHat"], ref["yHat"]) <= 1e-9, mode
            assert relmax(got["y"], ref["y"]) <= 1e-9, mode


def test_mde_merged_sweeps_with_fixed_effects():
    """Merged environment kernels together with XF (environment means) and missing phenotypes: resident kernel
    (merged steps + the scalar XF step), streaming path (xf kernels + merged units), two-pass, against the oracle."""
    import bgge_b200
    L, E, p = 30, 4, 40
    X = markers(L, p, 8)
    Y, names = design(L, E)
    ne = [L] * E
    n = L * E
    Kf = bgge_b200.getK(Y, X, kernel="GB", model="MDe", rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel="GB", model="MDe", rownames=names)
    rng = np.random.default_rng(21)
    y = rng.standard_normal(n)
    y[rng.choice(n, 6, replace=False)] = np.nan
    XF = np.zeros((n, E))
    for e in range(E):
        XF[e * L:(e + 1) * L, e] = 1.0
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    ite = 16
    nz, nc = oracle.tape_sizes(n, nrs, ite, q=E, n_na=6)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, XF=XF, ne=ne, ite=ite, burn=2, thin=2, draws=oracle.TapeDraws(z, c), eig=Ei)
    for mode in (1, 2, 0):
        got = bgge_b200.BGGE(y, Kf, XF=XF, ne=ne, ite=ite, burn=2, thin=2, draws=(z, c), mode=mode)
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, mode
        assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9, mode
        for nm in ref["K"]:
            assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= 1e-9, (mode, nm)
            assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (mode, nm)
        assert relmax(np.asarray(got["yHat"]).ravel(), np.asarray(ref["yHat"]).ravel()) <= 1e-9, mode


def test_shared_kernel_handles_reproduce_owner(lib):
    """bgge_gibbs_share_kernel: chains that borrow a decomposition (factorised + shared blocks included) give
    the same chain as a handle that built it itself."""
    import ctypes as C
    import bgge_b200
    L, E, p = 36, 4, 48
    X = markers(L, p, 3)
    Y, names = design(L, E)
    Kf = bgge_b200.getK(Y, X, kernel="GK", model="MDs", rownames=names, factored=True)
    n = L * E
    y = np.random.default_rng(1).standard_normal(n)
    want = bgge_b200.BGGE(y, Kf, ne=[L] * E, ite=30, burn=6, thin=2, seed=21, chain=2)
    ne = np.array([L] * E, dtype=np.int64)
    h0 = C.c_void_p()
    lib.call("bgge_gibbs_create", n, 2, 0, None, C.byref(h0))
    for j, k in enumerate(Kf.values()):
        f = k["factor"]
        lib.call("bgge_gibbs_add_expanded_kernel", h0, j, 0 if k["Type"] == "D" else 1, lib.ptr(lib.f64(f["K0"])), L, L, 0,
                 lib.ptr(f["gid"]), lib.ptr(f["env"]), n, f["mode"], f["env_k"], lib.ptr(ne), E, 1e-10, 1)
    for rep in range(2):
        h = C.c_void_p()
        lib.call("bgge_gibbs_create", n, 2, 0, None, C.byref(h))
        lib.call("bgge_gibbs_set_y", h, lib.ptr(y), None)
        for j in range(2):
            lib.call("bgge_gibbs_share_kernel", h, j, h0, j)
        lib.call("bgge_gibbs_init", h, 30, 6, 2, 0.5, lib.RNG_PHILOX, 21, 2)
        lib.call("bgge_gibbs_run", h, 30)
        cv = np.empty(15)
        lib.call("bgge_gibbs_chain", h, None, lib.ptr(cv), None, None)
        lib.load().bgge_gibbs_destroy(h)
        assert np.array_equal(cv, want["chain"]["varE"]), rep
    lib.load().bgge_gibbs_destroy(h0)
