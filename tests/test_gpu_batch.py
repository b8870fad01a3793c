"""Batched chains / folds (config C4): every chain of the batched sampler must reproduce the single-chain
sampler (itself parity-tested against the oracle) run with the same RNG stream."""
import numpy as np
import pytest

import oracle
from helpers import design, markers, phenotype, relmax

pytestmark = pytest.mark.gpu


def _folds(y, B, seed):
    """B phenotype vectors: different NA masks (folds) over the same data, plus one complete chain."""
    rng = np.random.default_rng(seed)
    Y = np.tile(y, (B, 1))
    perm = rng.permutation(y.shape[0])
    for b in range(1, B):
        Y[b, perm[(perm % max(2, B - 1)) == (b - 1) % max(2, B - 1)][: y.shape[0] // 5]] = np.nan
    return Y


@pytest.mark.parametrize("model,kernel,B", [("MM", "GB", 3), ("MDs", "GK", 5), ("MDe", "GB", 4), ("MDs", "GB", 70)])
def test_batch_reproduces_single_chains(model, kernel, B):
    import bgge_b200
    L, E, p = 33, 3, 60
    X = markers(L, p, 6)
    Y0, names = design(L, E)
    K = bgge_b200.getK(Y0, X, kernel=kernel, model=model, rownames=names)
    y = phenotype(oracle.gb_kernel(X), L, E, 7)
    Y = _folds(y, B, 1)
    ne = [L] * E
    kw = dict(ne=ne, ite=40, burn=10, thin=2, seed=11)
    fits = bgge_b200.BGGE_batch(Y, K, chain0=3, **kw)
    assert len(fits) == B
    for b in ([0, 1, B - 1] if B > 5 else range(B)):
        ref = bgge_b200.BGGE(Y[b], K, chain=3 + b, **kw)
        got = fits[b]
        assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9, b
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, b
        for nm in ref["K"]:
            assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= 1e-9, (b, nm)
            assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (b, nm)
            assert relmax(got["K"][nm]["u.sd"], ref["K"][nm]["u.sd"]) <= 1e-7, (b, nm)
        assert relmax(got["yHat"], ref["yHat"]) <= 1e-9, b
        assert relmax(got["y"], ref["y"]) <= 1e-9, b
        assert abs(got["varE"] - ref["varE"]) <= 1e-9 * ref["varE"]


def test_batch_large_blocks_split_k():
    """Shapes that force K splits and several M tiles (rows and nr beyond one 128 tile, ragged edges)."""
    import bgge_b200
    rng = np.random.default_rng(2)
    n, B = 1500, 6
    spec = [("D", [(0, n, 333)]), ("BD", [(0, 700, 150), (800, 700, 129)])]
    eig, K = [], {}
    for j, (typ, blocks) in enumerate(spec):
        bl = []
        for pos0, rows, nr in blocks:
            Q, _ = np.linalg.qr(rng.standard_normal((rows, nr)))
            bl.append({"s": np.sort(rng.gamma(2.0, 1.0, nr) + 0.05)[::-1].copy(), "U": np.asfortranarray(Q),
                       "pos": None if typ == "D" else (pos0, pos0 + rows)})
        eig.append(bl)
        K[f"K{j}"] = {"Type": typ, "mean_diag": 0.8}
    y = rng.standard_normal(n)
    Y = _folds(y, B, 3)
    kw = dict(ne=[1, 1], ite=12, burn=2, thin=1, seed=5, eig=eig)
    fits = bgge_b200.BGGE_batch(Y, K, chain0=0, **kw)
    for b in range(B):
        ref = bgge_b200.BGGE(Y[b], K, chain=b, **kw)
        assert relmax(fits[b]["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, b
        for nm in ref["K"]:
            assert relmax(fits[b]["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (b, nm)
        assert relmax(fits[b]["yHat"], ref["yHat"]) <= 1e-9, b


@pytest.mark.parametrize("model,B", [("MM", 5), ("MDs", 9)])
def test_batched_chains_match_the_oracle_draw_for_draw(model, B):
    """the batched sampler against the ORACLE, not against the single-chain CUDA sampler.  Chain b draws from
    the Philox stream (seed, chain0 + b); helpers.PhiloxDraws regenerates exactly that stream on the host and feeds it
    to the oracle in the reference's call order.  Folds (different NA masks per chain), D and BD kernels."""
    import bgge_b200
    from helpers import PhiloxDraws
    L, E, p = 50, 3, 70
    X = markers(L, p, 13)
    Y, names = design(L, E)
    n = L * E
    ne = [L] * E
    Kf = bgge_b200.getK(Y, X, kernel="GB", model=model, rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel="GB", model=model, rownames=names)
    rng = np.random.default_rng(14)
    y = rng.standard_normal(n)
    perm = rng.permutation(L)
    Yb = np.empty((B, n))
    for b in range(B):
        Yb[b] = y
        Yb[b, np.tile(perm % 5 == b % 5, E)] = np.nan      # fold b % 5 masked in every environment
    Yb[0] = y                                              # one chain without missing values
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(bk["s"].shape[0] for bk in bl) for bl in Ei]
    ite, burn, thin, seed, chain0 = 12, 2, 2, 77, 3
    fits = bgge_b200.BGGE_batch(Yb, Kf, ne=ne, ite=ite, burn=burn, thin=thin, seed=seed, chain0=chain0, eig=Ei)
    for b in range(B):
        draws = PhiloxDraws(seed, chain0 + b, n, nrs, n_na=int(np.isnan(Yb[b]).sum()))
        ref = oracle.BGGE(Yb[b], Kd, ne=ne, ite=ite, burn=burn, thin=thin, draws=draws, eig=Ei)
        got = fits[b]
        assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9, b
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, b
        for nm in ref["K"]:
            assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= 1e-9, (b, nm)
            assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (b, nm)
        assert relmax(got["yHat"], ref["yHat"]) <= 1e-9, b
        assert relmax(got["y"], ref["y"]) <= 1e-9, b


def test_single_chain_device_rng_matches_the_oracle_draw_for_draw():
    """The same for the single-chain sampler with XF: device Philox draws == the host mirror, so a device-RNG run can be
    checked against the oracle exactly, not only statistically."""
    import bgge_b200
    from helpers import PhiloxDraws
    L, E, p = 40, 3, 60
    X = markers(L, p, 2)
    Y, names = design(L, E)
    n = L * E
    ne = [L] * E
    Kf = bgge_b200.getK(Y, X, kernel="GK", model="MDs", rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel="GK", model="MDs", rownames=names)
    rng = np.random.default_rng(3)
    y = rng.standard_normal(n)
    y[rng.choice(n, 6, replace=False)] = np.nan
    XF = np.zeros((n, E))
    for e in range(E):
        XF[e * L:(e + 1) * L, e] = 1.0
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(bk["s"].shape[0] for bk in bl) for bl in Ei]
    for mode in (1, 2):
        got = bgge_b200.BGGE(y, Kf, XF=XF, ne=ne, ite=15, burn=3, thin=2, seed=9, chain=4, eig=Ei, mode=mode)
        ref = oracle.BGGE(y, Kd, XF=XF, ne=ne, ite=15, burn=3, thin=2, draws=PhiloxDraws(9, 4, n, nrs, q=E, n_na=6), eig=Ei)
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, mode
        assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9, mode
        assert relmax(np.asarray(got["yHat"]).ravel(), np.asarray(ref["yHat"]).ravel()) <= 1e-9, mode


def _env_design(L, E):
    XF = np.zeros((L * E, E))
    for e in range(E):
        XF[e * L:(e + 1) * L, e] = 1.0
    return XF


@pytest.mark.parametrize("model,B,q_extra", [("MM", 4, 0), ("MDs", 9, 1), ("MDe", 41, 0)])
def test_batched_chains_with_fixed_effects_match_the_oracle_draw_for_draw(model, B, q_extra):
    """XF in the batched sampler (R/BGGE.R:208-212, 284-290, 375, 414-415): environment indicators (+ one continuous
    covariate) shared by all chains, folds with different NA masks, every chain against the oracle fed the chain's own
    Philox stream.  B = 41 uses the 16-chain thread tiles with a ragged last tile."""
    import bgge_b200
    from helpers import PhiloxDraws
    L, E, p = 36, 3, 50
    X = markers(L, p, 21)
    Y, names = design(L, E)
    n = L * E
    ne = [L] * E
    Kf = bgge_b200.getK(Y, X, kernel="GB", model=model, rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel="GB", model=model, rownames=names)
    rng = np.random.default_rng(22)
    XF = _env_design(L, E)
    if q_extra:
        XF = np.hstack([XF, rng.standard_normal((n, q_extra))])
    q = XF.shape[1]
    y = rng.standard_normal(n) + XF @ rng.standard_normal(q)
    perm = rng.permutation(L)
    Yb = np.tile(y, (B, 1))
    for b in range(1, B):
        Yb[b, np.tile(perm % 6 == b % 6, E)] = np.nan
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(bk["s"].shape[0] for bk in bl) for bl in Ei]
    ite, burn, thin, seed, chain0 = 10, 2, 2, 5, 1
    fits = bgge_b200.BGGE_batch(Yb, Kf, XF=XF, ne=ne, ite=ite, burn=burn, thin=thin, seed=seed, chain0=chain0, eig=Ei)
    for b in (range(B) if B <= 9 else [0, 1, 15, 16, 31, 32, B - 1]):
        draws = PhiloxDraws(seed, chain0 + b, n, nrs, q=q, n_na=int(np.isnan(Yb[b]).sum()))
        ref = oracle.BGGE(Yb[b], Kd, XF=XF, ne=ne, ite=ite, burn=burn, thin=thin, draws=draws, eig=Ei)
        got = fits[b]
        assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9, b
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, b
        assert relmax(got["chain"]["Bet"], ref["_Bet_chain"]) <= 1e-9, b
        for nm in ref["K"]:
            assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= 1e-9, (b, nm)
            assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (b, nm)
        assert relmax(got["yHat"], ref["yHat"]) <= 1e-9, b
        assert relmax(got["y"], ref["y"]) <= 1e-9, b


def test_batch_with_fixed_effects_reproduces_single_chains():
    import bgge_b200
    L, E, p, B = 33, 4, 60, 6
    X = markers(L, p, 6)
    Y0, names = design(L, E)
    K = bgge_b200.getK(Y0, X, kernel="GK", model="MDs", rownames=names)
    y = phenotype(oracle.gb_kernel(X), L, E, 7)
    Y = _folds(y, B, 1)
    XF = _env_design(L, E)
    kw = dict(XF=XF, ne=[L] * E, ite=30, burn=6, thin=3, seed=4)
    fits = bgge_b200.BGGE_batch(Y, K, chain0=2, **kw)
    for b in range(B):
        ref = bgge_b200.BGGE(Y[b], K, chain=2 + b, mode=1, **kw)
        got = fits[b]
        assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= 1e-9, b
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= 1e-9, b
        for nm in ref["K"]:
            assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= 1e-9, (b, nm)
        assert relmax(got["yHat"], np.asarray(ref["yHat"]).ravel()) <= 1e-9, b
        assert relmax(got["y"], ref["y"]) <= 1e-9, b


def test_batch_rejects_too_many_or_singular_fixed_effects():
    import bgge_b200
    L, E = 20, 2
    X = markers(L, 30, 1)
    Y0, names = design(L, E)
    K = bgge_b200.getK(Y0, X, kernel="GB", model="MM", rownames=names)
    Y = np.random.default_rng(0).standard_normal((2, L * E))
    with pytest.raises(Exception, match="outside"):
        bgge_b200.BGGE_batch(Y, K, XF=np.random.default_rng(1).standard_normal((L * E, 17)), ne=[L] * E, ite=2, burn=0, thin=1)
    XF = np.ones((L * E, 2))
    with pytest.raises(Exception, match="singular|positive definite"):
        bgge_b200.BGGE_batch(Y, K, XF=XF, ne=[L] * E, ite=2, burn=0, thin=1)
