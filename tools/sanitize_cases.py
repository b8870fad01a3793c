"""Small-shape runs of every hand-synchronised kernel for compute-sanitizer (tools/sanitize.sh): the mbarrier / DSMEM /
grid-barrier protocols of sweep_fused_kernel (cs = 1, 2, 8; NB = 1, 4), sweep_mrhs_kernel (cs = 1, 2, 4; NB = 2, 4),
resident_sweep_kernel (XF, NA, merged steps), gram_dmma_kernel (K split on and off) and dgemm_dmma_kernel (batched
chains).  Every case is checked against the oracle as well, so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle                                   # noqa: E402
import bgge_b200                                # noqa: E402
from helpers import design, markers, relmax     # noqa: E402


def fit_case(label, L, E, p, kernel, model, mode, env, ite=3, with_xf=False, n_na=3, want_path=None):
    X = markers(L, p, 1)
    Y, names = design(L, E)
    n = L * E
    ne = [L] * E
    Kf = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names)
    rng = np.random.default_rng(2)
    y = rng.standard_normal(n)
    if n_na:
        y[rng.choice(n, n_na, replace=False)] = np.nan
    XF = None
    if with_xf:
        XF = np.zeros((n, E))
        for e in range(E):
            XF[e * L:(e + 1) * L, e] = 1.0
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    nz, nc = oracle.tape_sizes(n, nrs, ite, q=(E if with_xf else 0), n_na=n_na)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, XF=XF, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        info = {}
        got = bgge_b200.BGGE(y, Kf, XF=XF, ne=ne, ite=ite, burn=0, thin=1, draws=(z, c), mode=mode, info=info)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    err = relmax(got["chain"]["varE"], ref["chain"]["varE"])
    paths = [(k["path"], k["cluster_size"]) for k in info["kernels"]]
    print(f"{label:34s} resident={info['resident']} units={info['merged_units']} paths={paths} err={err:.2e}", flush=True)
    assert err <= 1e-9, label
    assert np.all(np.isfinite(got["chain"]["varE"])) and np.ptp(got["chain"]["varE"]) > 0, label
    if want_path is not None:          # the case must run on the engine its label names
        assert paths[-1][0] == want_path, (label, paths)


def main():
    # streaming kernels, old engine: dense blocks NB = 1 (cs 1 / 2 / 8), shared blocks NB = 4 on 8-CTA clusters
    fit_case("fused nb=1 cs=1", 96, 2, 60, "GK", "MM", 2, {})
    fit_case("fused nb=1 cs=2", 96, 2, 60, "GK", "MM", 2, {"BGGE_FUSED_CS": "2"})
    fit_case("fused nb=4 cs=8 (old engine)", 96, 4, 60, "GK", "MDs", 2, {"BGGE_MRHS": "0", "BGGE_FUSED_CS": "8"})
    fit_case("fused nb=2 cs=2 (old engine)", 96, 2, 60, "GK", "MDs", 2, {"BGGE_MRHS": "0", "BGGE_FUSED_CS": "2"})
    # multi-RHS engine forced on small blocks: NB = 4 / 2, cs = 1 / 2 / 4, few clusters -> long pipelines
    fit_case("mrhs nb=4 cs=1", 200, 4, 260, "GK", "MDs", 2, {"BGGE_MRHS": "2", "BGGE_FUSED_MAX_CLUSTERS": "2"}, want_path=2)
    fit_case("mrhs nb=4 cs=4", 200, 4, 260, "GK", "MDs", 2, {"BGGE_MRHS": "2", "BGGE_MRHS_CS": "4", "BGGE_FUSED_MAX_CLUSTERS": "1"},
             want_path=2)
    fit_case("mrhs nb=2 cs=2", 2900, 2, 300, "GK", "MDs", 2, {"BGGE_MRHS": "2", "BGGE_MRHS_CS": "2", "BGGE_FUSED_MAX_CLUSTERS": "4"},
             want_path=2)
    # residual slices in tensor memory (tcgen05.alloc / st / ld / dealloc), 1- and 2-CTA clusters
    fit_case("mrhs nb=4 tmem cs=1", 2500, 4, 300, "GK", "MDs", 2, {"BGGE_MRHS_TMEM": "2", "BGGE_FUSED_MAX_CLUSTERS": "3"}, want_path=2)
    fit_case("mrhs nb=4 tmem cs=2", 2500, 4, 300, "GK", "MDs", 2, {"BGGE_MRHS_TMEM": "2", "BGGE_MRHS_CS": "2"}, want_path=2)
    fit_case("mrhs merged units 4+2 (MDe)", 120, 6, 150, "GK", "MDe", 2, {"BGGE_MRHS": "2"}, ite=4)
    # resident kernel: XF + NA, merged steps
    fit_case("resident MDs + XF + NA", 64, 3, 80, "GK", "MDs", 1, {}, ite=4, with_xf=True)
    fit_case("resident MDe merged steps", 48, 5, 60, "GB", "MDe", 1, {}, ite=4)
    # two-pass kernels
    fit_case("two-pass proj/back", 96, 3, 60, "GK", "MDs", 0, {})
    # Gram kernel: direct and K-split scheduling; GK epilogues
    lib = bgge_b200._lib
    for (L, p, split) in ((300, 260, "1"), (300, 2100, "3")):
        X = markers(L, p, 4)
        os.environ["BGGE_GRAM_SPLIT"] = split
        out = np.empty((1, L, L))
        lib.call("bgge_getk_base", lib.ptr(lib.f64(X)), L, p, lib.GK, lib.ptr(np.ones(1)), 1, 0.5, lib.ptr(out), None)
        del os.environ["BGGE_GRAM_SPLIT"]
        D = oracle.gk_distance(X)
        err = relmax(out[0], oracle.gk_kernel(D, 1.0, oracle.quantile7(D.ravel(), 0.5)))
        print(f"gram L={L} p={p} split={split}              err={err:.2e}", flush=True)
        assert err <= 1e-12
    # batched chains: the skinny DMMA GEMMs
    L, E, p, B = 80, 2, 100, 9
    X = markers(L, p, 6)
    Y, names = design(L, E)
    Kf = bgge_b200.getK(Y, X, kernel="GB", model="MM", rownames=names, factored=True)
    Yb = np.random.default_rng(7).standard_normal((B, L * E))
    fits = bgge_b200.BGGE_batch(Yb, Kf, ne=[L] * E, ite=6, burn=0, thin=1, seed=5)
    one = bgge_b200.BGGE(Yb[3], Kf, ne=[L] * E, ite=6, burn=0, thin=1, seed=5, chain=3, mode=2)
    err = relmax(fits[3]["chain"]["varE"], one["chain"]["varE"])
    print(f"batched chains B={B}                   err={err:.2e}", flush=True)
    assert err <= 1e-9
    XF = np.zeros((L * E, E))
    for e in range(E):
        XF[e * L:(e + 1) * L, e] = 1.0
    Yb[2, 5:9] = np.nan
    fits = bgge_b200.BGGE_batch(Yb, Kf, XF=XF, ne=[L] * E, ite=6, burn=0, thin=1, seed=5)
    one = bgge_b200.BGGE(Yb[2], Kf, XF=XF, ne=[L] * E, ite=6, burn=0, thin=1, seed=5, chain=2, mode=2)
    err = relmax(fits[2]["chain"]["varE"], one["chain"]["varE"])
    print(f"batched chains B={B} with XF, NA        err={err:.2e}", flush=True)
    assert err <= 1e-9
    print("sanitize cases ok")


if __name__ == "__main__":
    main()
