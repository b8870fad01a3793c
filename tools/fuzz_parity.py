"""Randomised parity sweep: random designs (balanced / unbalanced / shuffled), models, kernels, missing phenotypes, fixed
effects, thinning, and all three sweep implementations, each compared with the oracle on injected draw tapes
(<= 1e-9 per recorded draw).  usage: python tools/fuzz_parity.py [n_cases] [seed]   (needs a B200)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle                                   # noqa: E402
import bgge_b200                                # noqa: E402
from helpers import design, markers, relmax     # noqa: E402


def one_case(rng, idx):
    L = int(rng.choice([7, 16, 33, 64, 97, 130, 257, 400, 700]))
    E = int(rng.integers(1, 7))
    p = int(rng.choice([5, 40, 150, 600, 1500]))
    kernel = str(rng.choice(["GB", "GK"]))
    model = str(rng.choice(["SM", "MM", "MDs", "MDe"], p=[0.1, 0.2, 0.35, 0.35])) if E > 1 else "SM"
    balanced = bool(rng.random() < 0.6) or model == "SM"
    drop = 0 if balanced else int(rng.integers(1, max(2, L * E // 5)))
    shuffle = (not balanced) and bool(rng.random() < 0.5)
    X = markers(L, p, int(rng.integers(1 << 30)))
    if model == "SM":
        Y, names = design(L, 1)
        E = 1
    else:
        Y, names = design(L, E, seed=int(rng.integers(1 << 30)), shuffle=shuffle, drop=drop)
    n = len(Y)
    env = np.array([r[0] for r in Y])
    # BGGE(ne=) needs environment-sorted rows for BD kernels; shuffled designs only make sense for the kernels, so sort
    order = np.argsort(env, kind="stable")
    Y = [Y[i] for i in order]
    env = env[order]
    ne = [int((env == e).sum()) for e in sorted(set(env))]
    with_xf = bool(rng.random() < 0.35) and len(ne) > 1 and min(ne) > 2
    n_na = int(rng.integers(0, max(1, n // 6)))
    ite = int(rng.integers(2, 7))
    thin = int(rng.integers(1, 3))
    burn = int(rng.integers(0, 2))
    mode = int(rng.choice([0, 1, 2]))
    desc = f"#{idx} L={L} E={E} p={p} {kernel} {model} n={n} drop={drop} shuffle={shuffle} xf={with_xf} na={n_na} ite={ite} thin={thin} mode={mode}"
    Kf = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names)
    for nm in Kd:
        dense = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names)[nm]["Kernel"] if idx % 5 == 0 else None
        if dense is not None:
            assert relmax(dense, Kd[nm]["Kernel"]) <= 1e-12, (desc, nm)
    y = rng.standard_normal(n)
    if n_na:
        y[rng.choice(n, n_na, replace=False)] = np.nan
    XF = None
    if with_xf:
        XF = np.zeros((n, len(ne)))
        pos = 0
        for e, m in enumerate(ne):
            XF[pos:pos + m, e] = 1.0
            pos += m
    types = [k["Type"] for k in Kd.values()]
    ne_arg = ne if ("BD" in types) else None
    if ne_arg is not None and len(ne_arg) == 1:
        return desc + "  (skipped: BD with one block)"
    Ei = bgge_b200.setDEC(Kf, ne_arg, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    if any(nr == 0 for nr in nrs):
        return desc + "  (skipped: a kernel kept no eigenvalue)"
    q = 0 if XF is None else XF.shape[1]
    nz, nc = oracle.tape_sizes(n, nrs, ite, q=q, n_na=n_na)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, XF=XF, ne=ne_arg, ite=ite, burn=burn, thin=thin, draws=oracle.TapeDraws(z, c), eig=Ei)
    # planner hooks: force the multi-right-hand-side engine / tensor-memory residuals / few clusters on shapes that
    # would not pick them by themselves, so that the fuzz reaches every engine
    hooks = {}
    if model in ("MDs", "MDe") and balanced and L >= 97 and mode == 2 and rng.random() < 0.6:
        hooks["BGGE_MRHS"] = "2"
        cs = rng.choice(["", "1", "2", "4"])
        if cs:
            hooks["BGGE_MRHS_CS"] = str(cs)
        tm = rng.choice(["", "0", "2"])
        if tm:
            hooks["BGGE_MRHS_TMEM"] = str(tm)
    if mode == 2 and rng.random() < 0.3:
        hooks["BGGE_FUSED_MAX_CLUSTERS"] = str(rng.integers(1, 5))
    desc += "".join(f" {k[5:]}={v}" for k, v in hooks.items())
    saved = {k: os.environ.get(k) for k in hooks}
    os.environ.update(hooks)
    info = {}
    # no eig= here: BGGE() decomposes again inside its own handle (deterministic, equal to setDEC's blocks), which is
    # what lets identical blocks share one W (multi-RHS panels) and the environment kernels of MDe form merged units
    try:
        got = bgge_b200.BGGE(y, Kf, XF=XF, ne=ne_arg, ite=ite, burn=burn, thin=thin, draws=(z, c), mode=mode, info=info)
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    errs = [relmax(got["chain"]["mu"], ref["chain"]["mu"]), relmax(got["chain"]["varE"], ref["chain"]["varE"]),
            relmax(np.asarray(got["yHat"]).ravel(), np.asarray(ref["yHat"]).ravel()), relmax(got["y"], ref["y"])]
    for nm in ref["K"]:
        errs.append(relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]))
        errs.append(relmax(got["K"][nm]["u"], ref["K"][nm]["u"]))
    worst = max(errs) if ref["chain"]["mu"].size else 0.0
    assert worst <= 1e-9, (desc, errs)
    paths = "".join(str(k["path"]) for k in info["kernels"])
    return f"{desc} resident={info['resident']} units={info['merged_units']} paths={paths} err={worst:.1e}"


def batch_case(rng, idx):
    """The batched sampler (skinny DMMA GEMMs, every N-tile width and both thread-tile shapes): random chain counts,
    folds, shared fixed effects; a few chains of each batch against the oracle fed the chain's own Philox stream."""
    from helpers import PhiloxDraws
    L = int(rng.choice([12, 33, 64, 130, 300]))
    E = int(rng.integers(2, 5))
    p = int(rng.choice([20, 80, 400]))
    B = int(rng.choice([1, 2, 7, 8, 9, 16, 23, 40, 41, 48, 49, 64, 77, 80, 96, 100, 128, 130]))
    model = str(rng.choice(["MM", "MDs", "MDe"]))
    kernel = str(rng.choice(["GB", "GK"]))
    X = markers(L, p, int(rng.integers(1 << 30)))
    Y, names = design(L, E)
    n = L * E
    ne = [L] * E
    q = int(rng.choice([0, 0, 1, E]))
    XF = None
    if q == E:
        XF = np.zeros((n, E))
        for e in range(E):
            XF[e * L:(e + 1) * L, e] = 1.0
    elif q == 1:
        XF = rng.standard_normal((n, 1))
    y = rng.standard_normal(n)
    Yb = np.tile(y, (B, 1))
    for b in range(1, B):
        k = int(rng.integers(0, max(1, n // 5)))
        if k:
            Yb[b, rng.choice(n, k, replace=False)] = np.nan
    ite, burn, thin = int(rng.integers(3, 8)), int(rng.integers(0, 2)), int(rng.integers(1, 3))
    seed, chain0 = int(rng.integers(1 << 40)), int(rng.integers(0, 50))
    desc = f"#{idx} batch B={B} L={L} E={E} p={p} {kernel} {model} q={q} ite={ite} thin={thin}"
    Kf = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names)
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(bk["s"].shape[0] for bk in bl) for bl in Ei]
    fits = bgge_b200.BGGE_batch(Yb, Kf, XF=XF, ne=ne, ite=ite, burn=burn, thin=thin, seed=seed, chain0=chain0, eig=Ei)
    worst = 0.0
    for b in sorted(set([0, B - 1, int(rng.integers(0, B)), int(rng.integers(0, B))])):
        draws = PhiloxDraws(seed, chain0 + b, n, nrs, q=q, n_na=int(np.isnan(Yb[b]).sum()))
        ref = oracle.BGGE(Yb[b], Kd, XF=XF, ne=ne, ite=ite, burn=burn, thin=thin, draws=draws, eig=Ei)
        got = fits[b]
        errs = [relmax(got["chain"]["mu"], ref["chain"]["mu"]), relmax(got["chain"]["varE"], ref["chain"]["varE"]),
                relmax(got["yHat"], np.asarray(ref["yHat"]).ravel()), relmax(got["y"], ref["y"])]
        for nm in ref["K"]:
            errs.append(relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]))
            errs.append(relmax(got["K"][nm]["u"], ref["K"][nm]["u"]))
        if ref["chain"]["mu"].size:
            worst = max(worst, max(errs))
    assert worst <= 1e-9, (desc, worst)
    return f"{desc} err={worst:.1e}"


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    kind = sys.argv[3] if len(sys.argv) > 3 else "single"
    rng = np.random.default_rng(seed)
    for i in range(ncases):
        print((batch_case if kind == "batch" else one_case)(rng, i), flush=True)
    print(f"fuzz parity ok: {ncases} {kind} cases, seed {seed}")


if __name__ == "__main__":
    main()
