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
