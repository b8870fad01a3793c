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


@pytest.mark.parametrize("mode", [1, 0])
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
