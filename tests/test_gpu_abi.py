"""The remaining C-ABI entry points exercised directly through ctypes: one-shot bgge_fit, the device-pointer
getK pipeline, chunked runs (verbose semantics), the device-memory cache."""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import design, markers, phenotype, relmax

pytestmark = pytest.mark.gpu


def vp(t):
    return C.c_void_p(t.data_ptr())


def test_bgge_fit_one_shot_matches_wrapper(lib):
    """bgge_fit (what the .Call shim calls) == handle API driven by the Python wrapper, same tape."""
    import bgge_b200
    L, E, p = 21, 3, 40
    X = markers(L, p, 1)
    Y, names = design(L, E)
    K = oracle.getK(Y, X, kernel="GK", model="MDs", rownames=names)
    y = phenotype(oracle.gb_kernel(X), L, E, 2)
    y[[3, 30]] = np.nan
    n, nk, ite, burn, thin = L * E, 2, 30, 6, 2
    ne = np.array([L] * E, dtype=np.int64)
    Ei = bgge_b200.setDEC(K, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    nz, nc = oracle.tape_sizes(n, nrs, ite, n_na=2)
    rng = np.random.default_rng(3)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    want = bgge_b200.BGGE(y, K, ne=ne, ite=ite, burn=burn, thin=thin, draws=(z, c))
    mats = [lib.f64(k["Kernel"]) for k in K.values()]
    kp = (C.c_void_p * nk)(*[m.ctypes.data for m in mats])
    types = np.array([0, 1], dtype=np.int32)
    ncum = ite // thin
    yHat, u, usd, yout = np.empty(n), np.empty((nk, n)), np.empty((nk, n)), np.empty(n)
    varu, varusd = np.empty(nk), np.empty(nk)
    cmu, cve, cvu = np.empty(ncum), np.empty(ncum), np.empty((nk, ncum))
    varE, varEsd = C.c_double(), C.c_double()
    na = np.isnan(y).astype(np.uint8)
    yy = np.where(np.isnan(y), 0.0, y)                       # NA passed as a byte mask, as the R shim does
    lib.call("bgge_fit", lib.ptr(yy), lib.ptr(na), n, kp, lib.ptr(types), nk, None, 0, lib.ptr(ne), E, ite, burn, thin, 1e-10,
             0.5, lib.RNG_TAPE, 0, lib.ptr(z), z.shape[0], lib.ptr(c), c.shape[0], lib.ptr(yHat), C.byref(varE),
             C.byref(varEsd), lib.ptr(u), lib.ptr(usd), lib.ptr(varu), lib.ptr(varusd), lib.ptr(yout), lib.ptr(cmu),
             lib.ptr(cve), lib.ptr(cvu))
    assert relmax(yHat, want["yHat"]) <= 1e-12
    assert relmax(cve, want["chain"]["varE"]) <= 1e-12
    assert relmax(cvu[1], want["chain"]["K"]["GE"]["varU"]) <= 1e-12
    assert relmax(u[0], want["K"]["G"]["u"]) <= 1e-12
    assert relmax(yout, want["y"]) <= 1e-12
    assert abs(varE.value - want["varE"]) <= 1e-12 * want["varE"]


def test_device_pointer_getk_pipeline_matches_host_entry(lib):
    import torch
    L, p, E = 150, 130, 2
    X = markers(L, p, 5)
    Y, names = design(L, E)
    ref = oracle.getK(Y, X, kernel="GK", model="MDs", rownames=names, bandwidth=[0.7])
    ldx, pp = C.c_int64(), C.c_int64()
    lib.call("bgge_marker_ld", L, C.byref(ldx))
    lib.call("bgge_marker_cols", p, C.byref(pp))
    Xp = torch.zeros((pp.value, ldx.value), dtype=torch.float64, device="cuda")
    Xp[:p, :L] = torch.from_numpy(X.T.copy()).cuda()
    S = torch.empty((L, L), dtype=torch.float64, device="cuda")
    lib.call("bgge_dev_gram", vp(Xp), ldx.value, L, pp.value, 0, (L + 127) // 128, vp(S), L, 0, 1.0, 1, None)
    lib.call("bgge_dev_gk_distance", vp(S), L, L, None)
    q = torch.empty(1, dtype=torch.float64, device="cuda")
    hq = C.c_double()
    lib.call("bgge_dev_quantile7", vp(S), L, L, 0.5, vp(q), C.byref(hq), None)
    K0 = torch.empty_like(S)
    lib.call("bgge_dev_gk_exp", vp(S), L, L, 0.7, vp(q), vp(K0), L, None)
    n = L * E
    gid = torch.arange(L, dtype=torch.int32, device="cuda").repeat(E)
    env = torch.arange(E, dtype=torch.int32, device="cuda").repeat_interleave(L)
    out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    for mode, nm in [(lib.EXPAND_G, "G"), (lib.EXPAND_GE, "GE")]:
        lib.call("bgge_dev_expand", vp(K0), L, L, vp(gid), vp(env), n, mode, 0, vp(out), n, None)
        torch.cuda.synchronize()
        assert relmax(out.cpu().numpy().T, ref[nm]["Kernel"]) <= 1e-12, nm
    D = oracle.gk_distance(X)
    assert abs(hq.value - oracle.quantile7(D, 0.5)) <= 1e-12 * hq.value


def test_chunked_runs_equal_one_run():
    """verbose=k runs the sweep in chunks of k iterations (R/BGGE.R:398-401); the chain must not depend on it."""
    import bgge_b200
    L, E = 18, 2
    X = markers(L, 30, 7)
    Y, names = design(L, E)
    K = bgge_b200.getK(Y, X, kernel="GB", model="MDs", rownames=names)
    y = phenotype(oracle.gb_kernel(X), L, E, 8)
    a = bgge_b200.BGGE(y, K, ne=[L] * E, ite=60, burn=10, thin=3, seed=4)
    b = bgge_b200.BGGE(y, K, ne=[L] * E, ite=60, burn=10, thin=3, seed=4, verbose=7)
    assert np.array_equal(a["chain"]["varE"], b["chain"]["varE"])
    assert np.array_equal(a["yHat"], b["yHat"])


def test_release_cache_and_memory_query(lib):
    free0, tot = C.c_uint64(), C.c_uint64()
    lib.call("bgge_device_memory", C.byref(free0), C.byref(tot))
    assert 0 < free0.value <= tot.value
    out = np.empty((64, 64), order="F")
    lib.call("bgge_getk_base", lib.ptr(lib.f64(np.random.default_rng(0).standard_normal((64, 32)))), 64, 32, lib.GB, None, 1,
             0.5, lib.ptr(out), None)
    lib.call("bgge_release_cache")
    free1 = C.c_uint64()
    lib.call("bgge_device_memory", C.byref(free1), None)
    assert free1.value > 0


def test_bgge_getk_single_call_matches_oracle(lib):
    """bgge_getk: the one-call entry point the .Call shim uses (one upload of X, all kernels of the model)."""
    L, p, E = 37, 29, 3
    X = markers(L, p, 4)
    Y, names = design(L, E)
    ref = oracle.getK(Y, X, kernel="GK", model="MDe", rownames=names, bandwidth=[1.0, 0.3], intercept_random=True)
    n = L * E
    gid = np.tile(np.arange(L, dtype=np.int32), E)
    env = np.repeat(np.arange(E, dtype=np.int32), L)
    bw = np.array([1.0, 0.3])
    n_out = 2 * (1 + E) + 1
    outs = [np.empty((n, n), order="F") for _ in range(n_out)]
    arr = (C.c_void_p * n_out)(*[o.ctypes.data for o in outs])
    q = C.c_double()
    lib.call("bgge_getk", lib.ptr(lib.f64(X)), L, p, lib.GK, lib.ptr(bw), 2, 0.5, lib.ptr(gid), lib.ptr(env), n, E, 2, 1, arr,
             n_out, C.byref(q))
    assert len(ref) == n_out
    for o, (nm, k) in zip(outs, ref.items()):
        assert relmax(o, k["Kernel"]) <= 1e-12, nm


def test_default_getk_output_takes_the_factorised_route():
    """getK() (dense, R-compatible output) + BGGE(): the fit goes through the L x L decomposition; editing a
    kernel matrix afterwards makes BGGE fall back to the dense n x n route for that kernel."""
    import bgge_b200
    from bgge_b200.getk import factor_matches
    L, E = 26, 3
    X = markers(L, 40, 5)
    Y, names = design(L, E)
    K = bgge_b200.getK(Y, X, kernel="GB", model="MDs", rownames=names)
    assert all(factor_matches(k) for k in K.values())
    y = phenotype(oracle.gb_kernel(X), L, E, 6)
    a = bgge_b200.BGGE(y, K, ne=[L] * E, ite=40, burn=10, thin=2, seed=3)
    Kd = {nm: {"Kernel": k["Kernel"], "Type": k["Type"]} for nm, k in K.items()}      # no factor: dense route
    b = bgge_b200.BGGE(y, Kd, ne=[L] * E, ite=40, burn=10, thin=2, seed=3)
    # different eigen routes give the same eigenpairs up to rounding -> same chain up to rounding amplification
    assert relmax(a["chain"]["varE"], b["chain"]["varE"]) <= 1e-6
    K["GE"]["Kernel"] = K["GE"]["Kernel"].copy()
    K["GE"]["Kernel"][0, 0] += 1.0
    assert not factor_matches(K["GE"])
    c = bgge_b200.BGGE(y, K, ne=[L] * E, ite=10, burn=2, thin=2, seed=3)
    assert np.isfinite(c["yHat"]).all()


def _fit_kernels(lib, y, descs, gid, env, ne, ite, burn, thin, z, c, XF=None, chunk=0, progress=None):
    n, nk = y.shape[0], len(descs)
    arr = (lib.KernelDesc * nk)(*descs)
    q = 0 if XF is None else XF.shape[1]
    ncum = ite // thin
    out = {"yHat": np.empty(n), "u": np.empty((nk, n)), "usd": np.empty((nk, n)), "y": np.empty(n), "varu": np.empty(nk),
           "varusd": np.empty(nk), "cmu": np.empty(ncum), "cve": np.empty(ncum), "cvu": np.empty((nk, ncum)),
           "cbet": np.empty((max(q, 1), ncum)).T}
    cbet = np.empty((ncum, max(q, 1)))            # row-major n_cum x q, like bgge_gibbs_chain
    varE, varEsd, nfact = C.c_double(), C.c_double(), C.c_int()
    na = np.isnan(y).astype(np.uint8)
    yy = np.where(np.isnan(y), 0.0, y)
    cb = lib.PROGRESS_FN(progress) if progress is not None else None
    lib.call("bgge_fit_kernels", lib.ptr(yy), lib.ptr(na), n, C.cast(arr, C.c_void_p), nk, lib.ptr(gid), lib.ptr(env),
             None if XF is None else lib.ptr(lib.f64(XF)), q, lib.ptr(ne), ne.shape[0], ite, burn, thin, 1e-10, 0.5,
             lib.RNG_TAPE, 0, lib.ptr(z), z.shape[0], lib.ptr(c), c.shape[0], chunk, cb, None, lib.ptr(out["yHat"]),
             C.byref(varE), C.byref(varEsd), lib.ptr(out["u"]), lib.ptr(out["usd"]), lib.ptr(out["varu"]),
             lib.ptr(out["varusd"]), lib.ptr(out["y"]), lib.ptr(out["cmu"]), lib.ptr(out["cve"]), lib.ptr(out["cvu"]),
             lib.ptr(cbet), C.byref(nfact))
    out.update(varE=varE.value, varEsd=varEsd.value, nfact=nfact.value, cbet=cbet)
    return out


@pytest.mark.parametrize("model,kernel,with_xf", [("MDs", "GK", False), ("MDe", "GB", True), ("MM", "GB", False)])
def test_fit_kernels_one_shot_from_factor_descriptions_matches_oracle(lib, model, kernel, with_xf):
    """bgge_fit_kernels -- what the .Call shim under BGGE() binds -- fed with getK()'s factor descriptions (base kernel
    + codes, the n x n matrices are never built), against the oracle on the dense kernels with the same tapes;
    verbose chunks call the progress hook after every piece and the chain does not depend on the chunking."""
    import bgge_b200
    L, E, p = 60, 4, 80
    X = markers(L, p, 11)
    Y, names = design(L, E)
    n = L * E
    ne = np.array([L] * E, dtype=np.int64)
    Kf = bgge_b200.getK(Y, X, kernel=kernel, model=model, rownames=names, factored=True, intercept_random=(model == "MM"))
    Kd = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names, intercept_random=(model == "MM"))
    rng = np.random.default_rng(12)
    y = rng.standard_normal(n)
    y[rng.choice(n, 7, replace=False)] = np.nan
    XF = None
    if with_xf:
        XF = np.zeros((n, E))
        for e in range(E):
            XF[e * L:(e + 1) * L, e] = 1.0
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    ite, burn, thin = 24, 6, 2
    nz, nc = oracle.tape_sizes(n, nrs, ite, q=(E if with_xf else 0), n_na=7)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, XF=XF, ne=ne, ite=ite, burn=burn, thin=thin, draws=oracle.TapeDraws(z, c), eig=Ei)
    f0 = next(iter(Kf.values()))["factor"]
    keep = []
    descs = []
    for k in Kf.values():
        f = k["factor"]
        K0 = None if f["K0"] is None else lib.f64(f["K0"])
        keep.append(K0)
        descs.append(lib.KernelDesc(0 if k["Type"] == "D" else 1, f["mode"], f["env_k"], 0, f["L"],
                                    None if K0 is None else K0.ctypes.data, None))
    seen = []
    got = _fit_kernels(lib, y, descs, f0["gid"], f0["env"], ne, ite, burn, thin, z, c, XF=XF, chunk=5,
                       progress=lambda done, spi, user: seen.append((done, spi)) or 0)
    assert [d for d, _ in seen] == [5, 10, 15, 20, 24] and all(s > 0 for _, s in seen)
    assert got["nfact"] == len(Kf)
    one = _fit_kernels(lib, y, descs, f0["gid"], f0["env"], ne, ite, burn, thin, z, c, XF=XF)
    for key in ("yHat", "cmu", "cve", "cvu", "u", "y"):
        if with_xf:   # the resident kernel converts r <-> r + XF beta at every launch boundary: rounding-level differences
            assert relmax(got[key], one[key]) <= 1e-12, key
        else:
            assert np.array_equal(got[key], one[key]), key              # chunked == one piece, bit for bit
    assert relmax(got["cve"], ref["chain"]["varE"]) <= 1e-9
    assert relmax(got["cmu"], ref["chain"]["mu"]) <= 1e-9
    for j, nm in enumerate(ref["K"]):
        assert relmax(got["cvu"][j], ref["chain"]["K"][nm]["varU"]) <= 1e-9, nm
        assert relmax(got["u"][j], ref["K"][nm]["u"]) <= 1e-9, nm
    assert relmax(got["yHat"], np.asarray(ref["yHat"]).ravel()) <= 1e-9
    assert relmax(got["y"], ref["y"]) <= 1e-9
    if with_xf:
        assert relmax(got["cbet"], ref["_Bet_chain"]) <= 1e-9


def test_fit_kernels_mixes_dense_and_factored_kernels_and_can_be_interrupted(lib):
    """A K list whose second kernel was edited by the caller (so it must be decomposed as the dense n x n matrix it is,
    R/BGGE.R:144) next to a factored first kernel; and a progress hook that asks to stop."""
    import bgge_b200
    L, E, p = 40, 3, 50
    X = markers(L, p, 4)
    Y, names = design(L, E)
    n = L * E
    ne = np.array([L] * E, dtype=np.int64)
    Kf = bgge_b200.getK(Y, X, kernel="GB", model="MDs", rownames=names)      # dense + factor descriptions
    Kd = oracle.getK(Y, X, kernel="GB", model="MDs", rownames=names)
    for K in (Kf, Kd):                                                       # the caller shrinks GE's first environment block
        K["GE"]["Kernel"][:L, :L] *= 0.5
    assert not bgge_b200.getk.factor_matches(Kf["GE"]) and bgge_b200.getk.factor_matches(Kf["G"])
    rng = np.random.default_rng(5)
    y = rng.standard_normal(n)
    ite = 10
    Ei = [bgge_b200.setDEC({"G": Kf["G"]}, ne, 1e-10)[0], bgge_b200.setDEC({"GE": {"Type": "BD", "Kernel": Kf["GE"]["Kernel"]}}, ne, 1e-10)[0]]
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    nz, nc = oracle.tape_sizes(n, nrs, ite)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    f = Kf["G"]["factor"]
    K0 = lib.f64(f["K0"])
    dense = lib.f64(Kf["GE"]["Kernel"])
    descs = [lib.KernelDesc(0, f["mode"], 0, 0, L, K0.ctypes.data, None),
             lib.KernelDesc(1, lib.KERNEL_DENSE, 0, 0, 0, None, dense.ctypes.data)]
    got = _fit_kernels(lib, y, descs, f["gid"], f["env"], ne, ite, 0, 1, z, c)
    assert got["nfact"] == 1
    assert relmax(got["cve"], ref["chain"]["varE"]) <= 1e-9
    assert relmax(got["cvu"][1], ref["chain"]["K"]["GE"]["varU"]) <= 1e-9
    assert relmax(got["yHat"], ref["yHat"]) <= 1e-9
    with pytest.raises(lib.BGGEError, match="interrupted after 4 of 10"):
        _fit_kernels(lib, y, descs, f["gid"], f["env"], ne, ite, 0, 1, z, c, chunk=2, progress=lambda d, s, u: int(d >= 4))
