"""Parity against the ORACLE (not against another CUDA path) at the shapes the benchmark runs:
the C5 row shape (L = 10000, E = 4: factorised G on 2-CTA clusters, the four identical GE blocks as one
multi-right-hand-side panel on 4-CTA clusters), the C3 row shape (L = 3000, E = 6, MDe: merged units 4 + 2 on the
streaming path), long column pipelines of the multi-RHS kernel, and the exact C1 / C2 configurations of
BASELINE.json (599 x 1279 x 4, ite = 1000, burn = 200, thin = 3).  The base kernels are low rank (few markers) so the
oracle's dense eigen blocks stay small; the device kernels see the full row counts.  R/BGGE.R:274-403."""
import os
from collections import OrderedDict

import numpy as np
import pytest

import oracle
from helpers import design, markers, phenotype, relmax

pytestmark = pytest.mark.gpu


def _mean_diag(f):
    d = np.ones(f["L"]) if f["K0"] is None else np.diag(f["K0"])
    v = d[f["gid"]].astype(np.float64)
    if f["mode"] == 2:
        v = v * (f["env"] == f["env_k"])
    return float(v.mean())


def _tapes(n, nrs, ite, n_na, q=0, seed=7):
    nz, nc = oracle.tape_sizes(n, nrs, ite, q=q, n_na=n_na)
    rng = np.random.default_rng(seed)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    assert c.shape[0] == nc
    return z, c


def _check(got, ref, tol=1e-9, usd=False):
    assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= tol
    assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= tol
    for nm in ref["K"]:
        assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= tol, nm
        assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= tol, nm
        if usd:
            assert relmax(got["K"][nm]["u.sd"], ref["K"][nm]["u.sd"]) <= 1e-7, nm
    assert relmax(np.asarray(got["yHat"]).ravel(), np.asarray(ref["yHat"]).ravel()) <= tol
    assert relmax(got["y"], ref["y"]) <= tol


def _fit_both(L, E, p, kernel, model, ite, mode, n_na, seed, env=None, burn=0, thin=1):
    """getK(factored) -> setDEC on the device -> oracle.BGGE on the fetched eigen blocks vs bgge_b200.BGGE (which
    decomposes again inside its own handle), same tapes."""
    import bgge_b200
    X = markers(L, p, seed)
    Y, names = design(L, E)
    ne = [L] * E
    n = L * E
    Kf = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names, factored=True)
    rng = np.random.default_rng(seed + 1)
    y = rng.standard_normal(n) + np.repeat(rng.standard_normal(E), L)
    if n_na:
        y[rng.choice(n, n_na, replace=False)] = np.nan
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    z, c = _tapes(n, nrs, ite, n_na, seed=seed + 2)
    Kor = OrderedDict((nm, {"Type": k["Type"], "mean_diag": _mean_diag(k["factor"])}) for nm, k in Kf.items())
    ref = oracle.BGGE(y, Kor, ne=ne, ite=ite, burn=burn, thin=thin, draws=oracle.TapeDraws(z, c), eig=Ei)
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        info = {}
        got = bgge_b200.BGGE(y, Kf, ne=ne, ite=ite, burn=burn, thin=thin, draws=(z, c), mode=mode, info=info)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return got, ref, info, nrs


def test_c5_row_shape_shared_blocks_on_4cta_clusters_match_oracle():
    """n = 40000 (L = 10000 x E = 4), MDs: G factorised (W 10000 x nr on 2-CTA clusters), GE = four identical blocks
    swept as one 4-right-hand-side panel by the register variant sweep_mrhs_kernel<7, 4, 2, false> on 4-CTA clusters
    (BGGE_MRHS_TMEM=0) -- five tape-injected iterations against the oracle."""
    got, ref, info, nrs = _fit_both(10000, 4, 192, "GB", "MDs", ite=5, mode=None, n_na=300, seed=31,
                                    env={"BGGE_MRHS_TMEM": "0"})
    assert not info["resident"]
    g, ge = info["kernels"]
    assert g["path"] == 1 and g["cluster_size"] == 2 and g["factored_blocks"] == 1
    assert ge["path"] == 2 and ge["cluster_size"] == 4 and ge["factored_blocks"] == 4
    assert ge["basis_bytes"] == 8 * 10000 * (nrs[1] // 4)          # W streamed once for the four blocks
    _check(got, ref)


def test_c5_row_shape_with_tensor_memory_residuals_on_2cta_clusters():
    """The same model as the DEFAULT plans it -- exactly the kernel instances the headline number comes from: the
    residual slices of the 4-member panel live in tensor memory (tcgen05.st / tcgen05.ld), so 2-CTA clusters hold the
    10000-row block (sweep_mrhs_kernel<14, 4, 1, true>, 74 clusters = all 148 SMs)."""
    got, ref, info, nrs = _fit_both(10000, 4, 192, "GB", "MDs", ite=5, mode=None, n_na=300, seed=31)
    g, ge = info["kernels"]
    assert g["path"] == 1 and g["cluster_size"] == 2 and g["factored_blocks"] == 1
    assert ge["path"] == 2 and ge["cluster_size"] == 2 and ge["factored_blocks"] == 4
    _check(got, ref)


@pytest.mark.parametrize("E,env,want_cs", [
    (4, {"BGGE_MRHS_CS": "4", "BGGE_FUSED_MAX_CLUSTERS": "3"}, 4),   # ~830 columns per cluster: ring, exchange ring and draw batches wrap
    (4, {}, 1),                                                       # single-CTA clusters, 7 row steps
    (2, {"BGGE_FUSED_MAX_CLUSTERS": "5"}, 1),                         # two right-hand sides
    (3, {"BGGE_MRHS_CS": "2"}, 2),                                    # three members in a 4-wide panel (one idle slot)
    (4, {"BGGE_MRHS_TMEM": "2", "BGGE_FUSED_MAX_CLUSTERS": "3"}, 1),  # residual slices in tensor memory, one column per group
    (4, {"BGGE_MRHS_TMEM": "2", "BGGE_MRHS_CS": "2"}, 2),             # the same on 2-CTA clusters
])
def test_multi_rhs_kernel_long_pipelines_match_oracle(E, env, want_cs):
    """sweep_mrhs_kernel with hundreds of column groups per cluster (the C5 launch runs ~140 groups per cluster): a
    full-rank Gaussian kernel at L = 2500, streaming path forced (mode 2), against the oracle."""
    L = 2500
    got, ref, info, nrs = _fit_both(L, E, 300, "GK", "MDs", ite=4, mode=2, n_na=57, seed=40 + E, env=env)
    ge = info["kernels"][1]
    assert ge["path"] == 2 and ge["cluster_size"] == want_cs and ge["factored_blocks"] == E
    if "BGGE_FUSED_MAX_CLUSTERS" in env:
        assert ge["clusters"] <= int(env["BGGE_FUSED_MAX_CLUSTERS"])
    assert nrs[1] == E * L                                            # full rank: 2500 columns per block
    _check(got, ref)


def test_c3_row_shape_merged_units_match_oracle():
    """n = 18000 (L = 3000 x E = 6), MDe: 1 + 6 kernels; the environment kernels share one decomposition and are
    swept as merged units of 4 + 2 members on the streaming path from iteration 2 on (iteration 1 unmerged)."""
    got, ref, info, nrs = _fit_both(3000, 6, 160, "GB", "MDe", ite=6, mode=2, n_na=200, seed=52, burn=1, thin=2)
    assert not info["resident"]
    assert info["merged_units"] == 2 and info["kernels_in_units"] == 6
    assert info["kernels"][0]["factored_blocks"] == 1
    _check(got, ref, usd=True)


@pytest.mark.parametrize("kernel,model", [("GB", "MM"), ("GK", "MDs")])
def test_exact_baseline_configs_c1_c2_match_oracle(kernel, model):
    """BASELINE.json configs[0] / configs[1]: 599 lines x 1279 markers x 4 environments, getK + BGGE with the
    reference's defaults ite = 1000, burn = 200, thin = 3 (333 recorded, 267 kept draws), through the public
    getK() / BGGE() pair on the resident kernel, against the oracle on the dense kernels with the same tapes."""
    import bgge_b200
    L, p, E = 599, 1279, 4
    X = markers(L, p, 20260000)
    Y, names = design(L, E)
    ne = [L] * E
    n = L * E
    K = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names)
    Kd = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names)
    for nm in Kd:
        assert relmax(K[nm]["Kernel"], Kd[nm]["Kernel"]) <= 1e-12, nm
    y = phenotype(oracle.gb_kernel(X), L, E, 5)
    y[np.random.default_rng(6).choice(n, 120, replace=False)] = np.nan
    Ei = bgge_b200.setDEC(K, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    z, c = _tapes(n, nrs, 1000, 120, seed=8)
    ref = oracle.BGGE(y, Kd, ne=ne, ite=1000, burn=200, thin=3, draws=oracle.TapeDraws(z, c), eig=Ei)
    info = {}
    got = bgge_b200.BGGE(y, K, ne=ne, ite=1000, burn=200, thin=3, draws=(z, c), info=info)
    assert info["resident"]
    assert got["chain"]["mu"].shape[0] == 333
    _check(got, ref, usd=True)
    assert abs(got["varE"] - ref["varE"]) <= 1e-9 * ref["varE"]
    assert abs(got["varE.sd"] - ref["varE.sd"]) <= 1e-7 * ref["varE.sd"]
