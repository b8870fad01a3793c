"""Column-sharded single chain.  One GPU is enough to test it: `world` handles in one process sweep
their column slices in lockstep and the "all-reduce" is a sum of their partial vectors."""
import ctypes as C

import numpy as np
import pytest

from oracle import bgge_oracle as oracle
from helpers import markers, design, phenotype, relmax

pytestmark = pytest.mark.gpu


def _handles(lib, world, y, K, Ei, tapes, ite, burn, thin, factored_K=None, ne=None):
    n, nk = y.shape[0], len(K)
    hs = []
    for rank in range(world):
        h = C.c_void_p()
        lib.call("bgge_gibbs_create", n, nk, 0, None, C.byref(h))
        lib.call("bgge_gibbs_set_y", h, lib.ptr(y), None)
        if factored_K is not None:
            from bgge_b200.bgge import _add_factored
            ne_arr = np.ascontiguousarray(ne, dtype=np.int64)
            for j, k in enumerate(factored_K.values()):
                _add_factored(h, j, k, ne_arr, 1e-10, True)
        else:
            for j, k in enumerate(K.values()):
                lib.call("bgge_gibbs_set_kernel", h, j, 0 if k["Type"] == "D" else 1, float(np.mean(np.diag(k["Kernel"]))))
                for blk in Ei[j]:
                    U = lib.f64(blk["U"])
                    lib.call("bgge_gibbs_add_block", h, j, 0 if blk["pos"] is None else blk["pos"][0], U.shape[0], U.shape[1],
                             lib.ptr(blk["s"]), lib.ptr(U), U.shape[0], 0)
        lib.call("bgge_gibbs_set_shard", h, rank, world)
        z, c = tapes
        lib.call("bgge_gibbs_set_tape", h, lib.ptr(z), z.shape[0], lib.ptr(c), c.shape[0])
        lib.call("bgge_gibbs_init", h, ite, burn, thin, 0.5, lib.RNG_TAPE, 0, 0)
        hs.append(h)
    return hs


def _lockstep(lib, hs, nk, ite):
    import torch
    from bgge_b200.multigpu import DeviceVector
    vec, cnt = C.c_void_p(), C.c_int64()
    for _ in range(ite):
        for j in range(nk):
            ts = []
            for h in hs:
                lib.call("bgge_gibbs_shard_begin", h, j, C.byref(vec), C.byref(cnt))
                ts.append(torch.as_tensor(DeviceVector(vec.value, cnt.value), device="cuda"))
            torch.cuda.synchronize()
            tot = ts[0].clone()
            for t in ts[1:]:
                tot += t
            for t in ts:
                t.copy_(tot)
            torch.cuda.synchronize()
            for h in hs:
                lib.call("bgge_gibbs_shard_end", h, j)
        for h in hs:
            lib.call("bgge_gibbs_shard_iteration_end", h)


def _summary(lib, h, n, nk, ncum):
    yHat, u, usd, yout = np.empty(n), np.empty((nk, n)), np.empty((nk, n)), np.empty(n)
    varu, varusd = np.empty(nk), np.empty(nk)
    ve, vesd = C.c_double(), C.c_double()
    lib.call("bgge_gibbs_summary", h, lib.ptr(yHat), C.byref(ve), C.byref(vesd), lib.ptr(u), lib.ptr(usd), lib.ptr(varu),
             lib.ptr(varusd), lib.ptr(yout))
    cmu, cve, cvu = np.empty(ncum), np.empty(ncum), np.empty((nk, ncum))
    lib.call("bgge_gibbs_chain", h, lib.ptr(cmu), lib.ptr(cve), lib.ptr(cvu), None)
    return {"yHat": yHat, "u": u, "y": yout, "mu": cmu, "varE": cve, "varU": cvu}


@pytest.mark.parametrize("world,L,E,model,factored", [(2, 40, 3, "MDs", False), (3, 64, 2, "MDs", True), (2, 30, 3, "MDe", True),
                                                      (4, 21, 1, "SM", False)])
def test_column_shards_reproduce_the_single_chain(lib, world, L, E, model, factored):
    import bgge_b200
    p = 48
    X = markers(L, p, 3)
    Y, names = design(L, E)
    ne = [L] * E
    Kd = oracle.getK(Y, X, kernel="GB", model=model, rownames=names)
    Kf = bgge_b200.getK(Y, X, kernel="GB", model=model, rownames=names, factored=True) if factored else None
    n, nk = L * E, len(Kd)
    rng = np.random.default_rng(world)
    y = phenotype(oracle.gb_kernel(X), L, E, 5)
    y[rng.choice(n, 4, replace=False)] = np.nan
    Ei = bgge_b200.setDEC(Kf if factored else Kd, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    ite, burn, thin = 12, 2, 2
    nz, nc = oracle.tape_sizes(n, nrs, ite, n_na=4)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, ne=ne, ite=ite, burn=burn, thin=thin, draws=oracle.TapeDraws(z, c), eig=Ei)
    hs = _handles(lib, world, y, Kd, Ei, (z, c), ite, burn, thin, factored_K=Kf, ne=ne)
    try:
        _lockstep(lib, hs, nk, ite)
        outs = [_summary(lib, h, n, nk, ite // thin) for h in hs]
    finally:
        for h in hs:
            lib.load().bgge_gibbs_destroy(h)
    for o in outs[1:]:   # the state is replicated: every process holds the same chain, bit for bit
        for key in o:
            assert np.array_equal(o[key], outs[0][key]), key
    got = outs[0]
    assert relmax(got["mu"], ref["chain"]["mu"]) <= 1e-9
    assert relmax(got["varE"], ref["chain"]["varE"]) <= 1e-9
    for j, nm in enumerate(ref["K"]):
        assert relmax(got["varU"][j], ref["chain"]["K"][nm]["varU"]) <= 1e-9, nm
        assert relmax(got["u"][j], ref["K"][nm]["u"]) <= 1e-9, nm
    assert relmax(got["yHat"], ref["yHat"]) <= 1e-9
    assert relmax(got["y"], ref["y"]) <= 1e-9


def test_shard_api_guards(lib):
    import bgge_b200
    X = markers(12, 20, 1)
    Y, names = design(12, 2)
    K = oracle.getK(Y, X, kernel="GB", model="MM", rownames=names)
    y = np.random.default_rng(0).standard_normal(24)
    Ei = bgge_b200.setDEC(K, [12, 12], 1e-10)
    z, c = np.zeros(10000), np.ones(1000)
    (h,) = _handles(lib, 1, y, K, Ei, (z, c), 3, 0, 1)[:1]
    try:   # world = 1: an ordinary handle, the shard calls refuse
        assert lib.load().bgge_gibbs_shard_begin(h, 0, None, None) != 0
        lib.call("bgge_gibbs_run", h, 1)
    finally:
        lib.load().bgge_gibbs_destroy(h)
    hs = _handles(lib, 2, y, K, Ei, (z, c), 3, 0, 1)
    try:
        assert lib.load().bgge_gibbs_run(hs[0], 1) != 0                    # a shard is driven step by step
        assert lib.load().bgge_gibbs_shard_end(hs[0], 0) != 0              # end without begin
        lib.call("bgge_gibbs_shard_begin", hs[0], 0, None, None)
        assert lib.load().bgge_gibbs_shard_begin(hs[0], 0, None, None) != 0   # step still open
        assert lib.load().bgge_gibbs_shard_iteration_end(hs[0]) != 0
    finally:
        for h in hs:
            lib.load().bgge_gibbs_destroy(h)
