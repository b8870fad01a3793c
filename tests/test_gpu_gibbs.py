"""GPU parity: the device Gibbs sweep vs the CPU oracle with an injected draw tape and a shared
eigenbasis.  Tolerance: per-iteration samples <= 1e-9 relative (north_star)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import design, markers, phenotype, relmax

pytestmark = pytest.mark.gpu

TOL_IT = 1e-9


def _problem(L, E, p, model, kernel="GB", seed=0, n_na=0, q=0):
    X = markers(L, p, seed)
    Y, names = design(L, E)
    K = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names)
    K0 = oracle.gb_kernel(X)
    y = phenotype(K0, L, E, seed + 1)
    rng = np.random.default_rng(seed + 2)
    if n_na:
        y[rng.choice(y.shape[0], size=n_na, replace=False)] = np.nan
    XF = None
    if q:
        XF = np.c_[np.ones(y.shape[0]), rng.standard_normal((y.shape[0], q - 1))] if q > 1 else rng.standard_normal((y.shape[0], 1))
    return K, y, XF, [L] * E


def _tape(n, nrs, ite, q, n_na, seed):
    nz, nc = oracle.tape_sizes(n, nrs, ite, q=q, n_na=n_na)
    rng = np.random.default_rng(seed)
    z = rng.standard_normal(nz)
    dfs = np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=np.float64), ite)
    c = rng.chisquare(dfs)
    return z, c


CASES = [
    # L, E, p, model, kernel, n_na, q
    (30, 1, 40, "SM", "GB", 0, 0),
    (25, 3, 40, "MM", "GB", 0, 0),
    (25, 3, 40, "MDs", "GK", 0, 0),
    (17, 4, 30, "MDe", "GB", 0, 0),
    (25, 3, 40, "MDs", "GB", 11, 0),
    (25, 3, 40, "MM", "GK", 0, 2),
    (23, 3, 40, "MDs", "GK", 9, 3),
    (301, 2, 100, "MDs", "GB", 20, 0),     # rows not a multiple of the 256-row tile, odd block starts
]


@pytest.mark.parametrize("mode", [1, 2, 0])
@pytest.mark.parametrize("L,E,p,model,kernel,n_na,q", CASES)
def test_per_iteration_parity(L, E, p, model, kernel, n_na, q, mode):
    import bgge_b200
    K, y, XF, ne = _problem(L, E, p, model, kernel=kernel, n_na=n_na, q=q)
    n = y.shape[0]
    ite = 25
    Ei = bgge_b200.setDEC(K, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in blocks) for blocks in Ei]
    z, c = _tape(n, nrs, ite, q, n_na, 7)
    tr = {"keep_u": True}
    ref = oracle.BGGE(y, K, XF=XF, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei, trace=tr)
    got = bgge_b200.BGGE(y, K, XF=XF, ne=ne, ite=ite, burn=0, thin=1, draws=(z, c), eig=Ei, mode=mode)
    # every iteration's scalar samples (thin=1 records all of them)
    assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= TOL_IT
    assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= TOL_IT
    for nm in ref["K"]:
        assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= TOL_IT, nm
        assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= TOL_IT, nm
        assert relmax(got["K"][nm]["u.sd"], ref["K"][nm]["u.sd"]) <= 1e-7, nm
        assert abs(got["K"][nm]["varu"] - ref["K"][nm]["varu"]) <= TOL_IT * abs(ref["K"][nm]["varu"])
    assert relmax(np.asarray(got["yHat"]).ravel(), ref["yHat"]) <= TOL_IT
    assert abs(got["varE"] - ref["varE"]) <= TOL_IT * ref["varE"]
    assert abs(got["varE.sd"] - ref["varE.sd"]) <= 1e-7 * ref["varE.sd"]
    assert relmax(got["y"], ref["y"]) <= TOL_IT
    assert list(got.keys()) == ["yHat", "varE", "varE.sd", "K", "chain", "ite", "burn", "thin", "y"]


@pytest.mark.parametrize("model,q,n_na", [("MDs", 3, 6), ("MM", 1, 0), ("MDe", 2, 5)])
def test_fixed_effects_in_the_resident_kernel_and_chunked_runs(model, q, n_na, capsys):
    """XF models run in the resident kernel too (r kept without the fixed-effect part, the XF step is scalar work).
    One run, a run cut into chunks of 4 iterations (verbose = 4: state handed back and converted each time), the
    streaming path and the oracle agree per iteration; the chain of fixed effects enters yHat."""
    import bgge_b200
    K, y, XF, ne = _problem(22, 3, 30, model, kernel="GB", n_na=n_na, q=q, seed=2)
    n = y.shape[0]
    ite = 18
    Ei = bgge_b200.setDEC(K, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in blocks) for blocks in Ei]
    z, c = _tape(n, nrs, ite, q, n_na, 5)
    ref = oracle.BGGE(y, K, XF=XF, ne=ne, ite=ite, burn=3, thin=2, draws=oracle.TapeDraws(z, c), eig=Ei)
    runs = {"resident": dict(mode=1), "resident, chunks of 4": dict(mode=1, verbose=4), "streaming": dict(mode=2)}
    for label, kw in runs.items():
        got = bgge_b200.BGGE(y, K, XF=XF, ne=ne, ite=ite, burn=3, thin=2, draws=(z, c), eig=Ei, **kw)
        assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= TOL_IT, label
        assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= TOL_IT, label
        for nm in ref["K"]:
            assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= TOL_IT, (label, nm)
            assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= TOL_IT, (label, nm)
        assert relmax(np.asarray(got["yHat"]).ravel(), np.asarray(ref["yHat"]).ravel()) <= TOL_IT, label
        assert relmax(got["y"], ref["y"]) <= TOL_IT, label
    capsys.readouterr()


def test_state_after_each_iteration(lib):
    """Step the handle one sweep at a time and compare u, e (R's temp), sigb, sigsq, mu."""
    import bgge_b200
    K, y, XF, ne = _problem(20, 3, 30, "MDs", n_na=5)
    n, nk, ite = y.shape[0], len(K), 6
    Ei = bgge_b200.setDEC(K, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in blocks) for blocks in Ei]
    z, c = _tape(n, nrs, ite, 0, 5, 3)
    h = C.c_void_p()
    lib.call("bgge_gibbs_create", n, nk, 0, None, C.byref(h))
    lib.call("bgge_gibbs_set_y", h, lib.ptr(y), None)
    for j, k in enumerate(K.values()):
        lib.call("bgge_gibbs_set_kernel", h, j, 0 if k["Type"] == "D" else 1, float(np.mean(np.diag(k["Kernel"]))))
        for blk in Ei[j]:
            U = lib.f64(blk["U"])
            lib.call("bgge_gibbs_add_block", h, j, 0 if blk["pos"] is None else blk["pos"][0], U.shape[0], U.shape[1],
                     lib.ptr(blk["s"]), lib.ptr(U), U.shape[0], 0)
    lib.call("bgge_gibbs_set_tape", h, lib.ptr(z), z.shape[0], lib.ptr(c), c.shape[0])
    lib.call("bgge_gibbs_init", h, ite, 0, 1, 0.5, lib.RNG_TAPE, 0, 0)
    for it in range(1, ite + 1):
        tr = {"keep_u": True}
        oracle.BGGE(y, K, ne=ne, ite=it, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei, trace=tr)
        lib.call("bgge_gibbs_run", h, 1)
        mu, s2 = C.c_double(), C.c_double()
        sigb, u, e = np.empty(nk), np.empty((nk, n)), np.empty(n)
        lib.call("bgge_gibbs_state", h, C.byref(mu), C.byref(s2), lib.ptr(sigb), lib.ptr(u), lib.ptr(e), None)
        assert abs(mu.value - tr["mu"][-1]) <= TOL_IT * abs(tr["mu"][-1])
        assert abs(s2.value - tr["varE"][-1]) <= TOL_IT * tr["varE"][-1]
        assert relmax(sigb, tr["varU"][-1]) <= TOL_IT
        assert relmax(u, tr["u"][-1]) <= TOL_IT
        assert relmax(e, tr["e"]) <= TOL_IT
    lib.load().bgge_gibbs_destroy(h)


def test_default_thinning_counts():
    """ite=1000, burn=200, thin=3 -> 333 recorded, 267 kept."""
    import bgge_b200
    K, y, XF, ne = _problem(15, 2, 20, "MM")
    fit = bgge_b200.BGGE(y, K, ne=ne, seed=1)
    assert fit["chain"]["mu"].shape == (333,)
    assert np.all(fit["chain"]["varE"] > 0)
    assert np.isfinite(fit["yHat"]).all() and fit["yHat"].shape == y.shape
    assert str(fit).startswith("Model Fitted with")


def test_eig_matches_lapack(lib):
    rng = np.random.default_rng(5)
    m = 150
    A = rng.standard_normal((m, 40))
    K = A @ A.T / 40                      # rank 40 -> tol truncation matters
    s, U = np.empty(m), np.empty((m, m), order="F")
    nr = C.c_int64()
    lib.call("bgge_eig", lib.ptr(lib.f64(K)), m, m, 1e-10, 1, lib.ptr(s), lib.ptr(U), C.byref(nr))
    sr, Ur = oracle.eig_block(K, 1e-10)
    assert nr.value == sr.shape[0] == 40
    assert relmax(s[:40], sr) <= 1e-12
    U = U[:, :40]
    assert np.max(np.abs(U.T @ U - np.eye(40))) <= 1e-12
    assert relmax((U * s[:40]) @ U.T, K) <= 1e-12
    # canonical sign: largest-|.| entry of every vector is positive
    assert np.all(U[np.argmax(np.abs(U), axis=0), np.arange(40)] > 0)


def test_philox_known_answers(lib):
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kats:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        lib.call("bgge_test_philox", lib.ptr(c), lib.ptr(k), lib.ptr(o))
        assert tuple(int(v) for v in o) == exp


def test_device_draw_distributions(lib):
    from scipy import stats
    n = 200000
    z, c = np.empty(n), np.empty(n)
    lib.call("bgge_test_draws", 123, 4, n, lib.ptr(z), 37.0, lib.ptr(c))
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert stats.kstest(c, "chi2", args=(37.0,)).pvalue > 1e-3
    z2 = np.empty(n)
    lib.call("bgge_test_draws", 123, 5, n, lib.ptr(z2), 37.0, None)
    assert abs(np.corrcoef(z, z2)[0, 1]) < 0.01          # chains are independent streams


def test_statistical_agreement_with_oracle():
    """Device RNG: posterior means within Monte-Carlo error of the oracle's own chains."""
    import bgge_b200
    K, y, XF, ne = _problem(60, 3, 200, "MDs", seed=4)
    kw = dict(ne=ne, ite=3000, burn=500, thin=2)
    dev = [bgge_b200.BGGE(y, K, seed=s, **kw) for s in (1, 2)]
    ref = [oracle.BGGE(y, K, draws=oracle.RandomDraws(s), **kw) for s in (11, 12)]

    def mcse(fit, getter):
        x = getter(fit["chain"])[250:]
        nb = 25
        bm = x[:len(x) // nb * nb].reshape(nb, -1).mean(axis=1)
        return bm.std(ddof=1) / np.sqrt(nb)

    for getter, key in [(lambda ch: ch["varE"], "varE"), (lambda ch: ch["K"]["G"]["varU"], None),
                        (lambda ch: ch["K"]["GE"]["varU"], None), (lambda ch: ch["mu"], None)]:
        dm = np.mean([getter(f["chain"])[250:].mean() for f in dev])
        rm = np.mean([getter(f["chain"])[250:].mean() for f in ref])
        se = np.sqrt(np.mean([mcse(f, getter) ** 2 for f in dev + ref]))
        assert abs(dm - rm) <= 6 * se + 1e-3 * abs(rm), (dm, rm, se)
    yd = np.mean([f["yHat"] for f in dev], axis=0)
    yr = np.mean([f["yHat"] for f in ref], axis=0)
    assert np.corrcoef(yd, yr)[0, 1] > 0.995


def _synthetic_blocks(spec, seed):
    """Eigen blocks without a kernel matrix: spec = [(type, [(pos0, rows, nr), ...]), ...]."""
    rng = np.random.default_rng(seed)
    eig, K = [], {}
    for j, (typ, blocks) in enumerate(spec):
        bl = []
        for pos0, rows, nr in blocks:
            Q, _ = np.linalg.qr(rng.standard_normal((rows, nr)))
            bl.append({"s": np.sort(rng.gamma(2.0, 1.0, nr) + 0.05)[::-1].copy(), "U": np.asfortranarray(Q),
                       "pos": None if typ == "D" else (pos0, pos0 + rows)})
        eig.append(bl)
        K[f"K{j}"] = {"Type": typ, "mean_diag": 0.7 + 0.1 * j}
    return K, eig


BIG = [
    # n, spec  -- shapes that exercise the cluster paths (rows per block > 4096 -> CS 2/4/8), gaps, ragged slices
    (9001, [("D", [(0, 9001, 37)])]),                                               # CS=4, odd rows
    (12000, [("D", [(0, 12000, 19)]), ("BD", [(0, 5000, 23), (7000, 5000, 9)])]),   # CS=2 blocks + gap rows
    (40000, [("D", [(0, 40000, 21)]), ("BD", [(k * 10000, 10000, 14) for k in range(4)])]),   # the C5 shapes: CS=8 / CS=4
    (4500, [("BD", [(0, 1500, 300), (1500, 1500, 280), (3000, 1500, 310)])]),       # many columns, small slices (G=4)
]


@pytest.mark.parametrize("mode", [1, 2, 0])
@pytest.mark.parametrize("n,spec", BIG)
def test_parity_on_large_blocks(n, spec, mode):
    import bgge_b200
    K, Ei = _synthetic_blocks(spec, 11)
    rng = np.random.default_rng(5)
    y = rng.standard_normal(n) * 1.3 + 0.4
    na = rng.choice(n, size=25, replace=False)
    y[na] = np.nan
    ite = 8
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    z, c = _tape(n, nrs, ite, 0, 25, 9)
    ne = [1, 1]     # only consulted by the validation of BD kernels
    ref = oracle.BGGE(y, K, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    got = bgge_b200.BGGE(y, K, ne=ne, ite=ite, burn=0, thin=1, draws=(z, c), eig=Ei, mode=mode)
    assert relmax(got["chain"]["mu"], ref["chain"]["mu"]) <= TOL_IT
    assert relmax(got["chain"]["varE"], ref["chain"]["varE"]) <= TOL_IT
    for nm in ref["K"]:
        assert relmax(got["chain"]["K"][nm]["varU"], ref["chain"]["K"][nm]["varU"]) <= TOL_IT, nm
        assert relmax(got["K"][nm]["u"], ref["K"][nm]["u"]) <= TOL_IT, nm
    assert relmax(got["yHat"], ref["yHat"]) <= TOL_IT


def test_fused_and_two_pass_agree_with_device_rng():
    """Same seed, both sweep implementations: identical draws, results equal to rounding."""
    import bgge_b200
    K, y, XF, ne = _problem(40, 3, 60, "MDs", seed=8, n_na=7)
    a = bgge_b200.BGGE(y, K, ne=ne, ite=200, burn=50, thin=2, seed=5, mode=1)
    for mode in (2, 0):
        b = bgge_b200.BGGE(y, K, ne=ne, ite=200, burn=50, thin=2, seed=5, mode=mode)
        assert relmax(a["chain"]["varE"], b["chain"]["varE"]) <= 1e-8, mode
        assert relmax(a["yHat"], b["yHat"]) <= 1e-8, mode


def test_small_models_run_resident_and_chunked_runs_continue(lib):
    """A model whose basis fits on chip runs in the resident kernel (mode 1); splitting the run into several
    bgge_gibbs_run calls (state handed back through device memory each time) gives the same chain as one call,
    and mode 2 on the same handle layout is not resident."""
    K, y, XF, ne = _problem(30, 3, 40, "MDs", seed=4, n_na=3)
    n, nk, ite = y.shape[0], len(K), 24
    import bgge_b200
    Ei = bgge_b200.setDEC(K, ne, 1e-10)

    def make(mode):
        h = C.c_void_p()
        lib.call("bgge_gibbs_create", n, nk, 0, None, C.byref(h))
        lib.call("bgge_gibbs_set_y", h, lib.ptr(y), None)
        for j, k in enumerate(K.values()):
            lib.call("bgge_gibbs_set_kernel", h, j, 0 if k["Type"] == "D" else 1, float(np.mean(np.diag(k["Kernel"]))))
            for blk in Ei[j]:
                U = lib.f64(blk["U"])
                lib.call("bgge_gibbs_add_block", h, j, 0 if blk["pos"] is None else blk["pos"][0], U.shape[0], U.shape[1],
                         lib.ptr(blk["s"]), lib.ptr(U), U.shape[0], 0)
        lib.call("bgge_gibbs_set_mode", h, mode)
        lib.call("bgge_gibbs_init", h, ite, 4, 2, 0.5, lib.RNG_PHILOX, 11, 0)
        en, ctas, smem = C.c_int(), C.c_int(), C.c_int64()
        lib.call("bgge_gibbs_resident_info", h, C.byref(en), C.byref(ctas), C.byref(smem))
        return h, en.value, ctas.value, smem.value

    def chains(h):
        ncum = ite // 2
        mu, ve, vu = np.empty(ncum), np.empty(ncum), np.empty((nk, ncum))
        lib.call("bgge_gibbs_sync", h)
        lib.call("bgge_gibbs_chain", h, lib.ptr(mu), lib.ptr(ve), lib.ptr(vu), None)
        return mu, ve, vu

    h1, en1, ctas1, smem1 = make(1)
    assert en1 == 1 and ctas1 >= 1 and 0 < smem1 <= 200 * 1024
    lib.call("bgge_gibbs_run", h1, ite)
    one = chains(h1)
    h2, en2, _, _ = make(1)
    for k in (1, 7, 3, 13):
        lib.call("bgge_gibbs_run", h2, k)
    parts = chains(h2)
    for a, b in zip(one, parts):
        assert np.array_equal(a, b)
    h3, en3, ctas3, _ = make(2)
    assert en3 == 0 and ctas3 == 0
    lib.call("bgge_gibbs_run", h3, ite)
    for a, b in zip(one, chains(h3)):
        assert relmax(a, b) <= 1e-9
    for h in (h1, h2, h3):
        lib.load().bgge_gibbs_destroy(h)
