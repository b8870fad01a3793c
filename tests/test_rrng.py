"""R-compatible RNG stream: known first outputs of R after set.seed(k), tape layout, and (GPU)
BGGE(rng="R") against the oracle fed with the same tapes."""
import math

import numpy as np
import pytest

from bgge_b200.rrng import RStream, qnorm, r_tapes
from oracle import bgge_oracle as oracle

# set.seed(k); runif(3) / rnorm(5) / rexp(3) in R (default Mersenne-Twister, Inversion), as printed by R
R_RUNIF = {1: [0.2655087, 0.3721239, 0.5728534], 42: [0.9148060, 0.9370754, 0.2861395], 123: [0.2875775, 0.7883051, 0.4089769]}
R_RNORM = {1: [-0.6264538, 0.1836433, -0.8356286, 1.5952808, 0.3295078],
           42: [1.37095845, -0.56469817, 0.36312841, 0.63286260, 0.40426832],
           123: [-0.56047565, -0.23017749, 1.55870831, 0.07050839, 0.12928774]}
R_REXP = {1: [0.7551818, 1.1816428, 0.1457067], 42: [0.1983368, 0.6608953, 0.2834910], 123: [0.84345726, 0.57661027, 1.32905487]}


@pytest.mark.parametrize("seed", [1, 42, 123])
def test_known_first_outputs_of_r(seed):
    assert np.allclose(RStream(seed).unif(3), R_RUNIF[seed], rtol=0, atol=5e-8)
    assert np.allclose(RStream(seed).norm(5), R_RNORM[seed], rtol=0, atol=5e-8)
    r = RStream(seed)
    assert np.allclose([r.exp1() for _ in range(3)], R_REXP[seed], rtol=0, atol=5e-8)


def test_stream_is_sequential_not_per_call():
    a = RStream(5)
    x = np.concatenate([a.norm(3), a.norm(4)])
    assert np.array_equal(x, RStream(5).norm(7))
    b = RStream(5)
    b.unif(1)
    assert not np.array_equal(b.norm(2), RStream(5).norm(2))


def test_qnorm_matches_scipy():
    from scipy.stats import norm
    p = np.concatenate([np.random.default_rng(0).random(20000), [1e-300, 1e-20, 1e-9, 0.075, 0.925, 1 - 1e-9, 0.5]])
    ref = norm.ppf(p)
    assert np.max(np.abs(qnorm(p) - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-14
    assert qnorm(0.5)[0] == 0.0


def test_gamma_first_branch_and_moments():
    # GD accepts immediately when the normal deviate t >= 0: value = (sqrt(a - 1/2) + t / 2)^2
    for seed in range(1, 40):
        t = RStream(seed).norm1()
        if t >= 0:
            a = 7.5
            assert RStream(seed).gamma1(a) == (math.sqrt(a - 0.5) + 0.5 * t) ** 2
    for a in (2.0, 5.5, 301.5):
        r = RStream(int(a * 10))
        g = np.array([r.gamma1(a) for _ in range(20000)])
        assert abs(g.mean() - a) <= 5 * math.sqrt(a / g.size)
        assert abs(g.var() - a) <= 0.06 * a
    with pytest.raises(ValueError):
        RStream(1).gamma1(0.5)


def test_tape_layout_follows_the_reference_call_order():
    n, nrs, q, n_na, ite = 7, [3, 5], 2, 2, 4
    z, c = r_tapes(11, n, nrs, q, n_na, ite)
    zn, cn = oracle.tape_sizes(n, nrs, ite, q, n_na)
    assert z.shape[0] == zn and c.shape[0] == cn
    # replay the stream by hand
    r = RStream(11)
    zz = [r.norm(2 * n)]
    cc = []
    for _ in range(ite):
        zz.append(r.norm(1))
        zz.append(r.norm(q))
        for nr in nrs:
            zz.append(r.norm(nr))
            cc.append(2.0 * r.gamma1(0.5 * (nr + 3.0)))
        cc.append(2.0 * r.gamma1(0.5 * (n + 3.0)))
        zz.append(r.norm(n_na))
    assert np.array_equal(z, np.concatenate(zz)) and np.array_equal(c, np.array(cc))
    assert np.all(c > 0)


@pytest.mark.gpu
def test_bgge_with_r_stream_matches_oracle_on_the_same_tapes():
    import bgge_b200
    rng = np.random.default_rng(3)
    L, E, p = 25, 3, 40
    X = rng.standard_normal((L, p))
    names = [f"g{i}" for i in range(L)]
    Y = [(f"e{e}", names[i], 0.0) for e in range(E) for i in range(L)]
    K = oracle.getK(Y, X, kernel="GB", model="MDs", rownames=names)
    n = L * E
    y = rng.standard_normal(n)
    y[[3, 17, 40]] = np.nan
    XF = np.column_stack([np.ones(n), rng.standard_normal(n)])[:, 1:]
    ne = np.full(E, L)
    ite = 30
    Ei = bgge_b200.setDEC(K, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in blocks) for blocks in Ei]
    z, c = r_tapes(42, n, nrs, 1, 3, ite)
    ref = oracle.BGGE(y, K, XF=XF, ne=ne, ite=ite, burn=6, thin=2, draws=oracle.TapeDraws(z, c), eig=Ei)
    got = bgge_b200.BGGE(y, K, XF=XF, ne=ne, ite=ite, burn=6, thin=2, eig=Ei, rng="R", seed=42)
    assert np.max(np.abs(got["chain"]["varE"] - ref["chain"]["varE"]) / ref["chain"]["varE"]) <= 1e-9
    assert np.max(np.abs(got["yHat"].ravel() - ref["yHat"].ravel())) <= 1e-9 * np.max(np.abs(ref["yHat"]))
    with pytest.raises(ValueError):
        bgge_b200.BGGE(y, K, XF=XF, ne=ne, ite=ite, eig=Ei, rng="R", draws=(z, c))
