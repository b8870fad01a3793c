"""Multi-GPU inside the library (include/bgge_b200.h "multi-GPU inside the library"): communicators, the row-panel
getK build with the in-place exchange, and the column-sharded chain whose all-reduce runs on the handle's stream.
Everything goes through ctypes -- no torch, no torch.distributed on these paths.  The two-device cases need a box with
>= 2 GPUs (gpurun --gpus 2) and skip elsewhere; the world-size-1 forms run everywhere."""
import ctypes as C
import sys
import threading

import numpy as np
import pytest

import oracle
from helpers import design, markers, relmax

pytestmark = pytest.mark.gpu


def _comms(lib, ndev):
    arr = (C.c_void_p * ndev)()
    lib.call("bgge_comm_init_all", ndev, None, C.cast(arr, C.POINTER(C.c_void_p)))
    return arr


def _free(lib, arr):
    for c in arr:
        lib.load().bgge_comm_destroy(c)


@pytest.mark.parametrize("ndev", [1, 2, 4])
@pytest.mark.parametrize("kind,L,p", [("GB", 700, 900), ("GK", 513, 333), ("GB", 1500, 2100)])
def test_getk_base_multi_matches_one_gpu_and_oracle(lib, ndev, kind, L, p):
    if lib.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    X = markers(L, p, 3)
    bw = np.array([1.0, 0.5]) if kind == "GK" else np.ones(1)
    nb = bw.shape[0]
    code = lib.GB if kind == "GB" else lib.GK
    one = np.empty((nb, L, L))
    lib.call("bgge_getk_base", lib.ptr(lib.f64(X)), L, p, code, lib.ptr(bw), nb, 0.5, lib.ptr(one), None)
    comms = _comms(lib, ndev)
    try:
        got = np.empty((nb, L, L))
        q = C.c_double()
        lib.call("bgge_getk_base_multi", C.cast(comms, C.POINTER(C.c_void_p)), ndev, lib.ptr(lib.f64(X)), L, p, code,
                 lib.ptr(bw), nb, 0.5, lib.ptr(got), C.byref(q))
    finally:
        _free(lib, comms)
    lib.call("bgge_set_device", 0)
    if kind == "GK":
        D = oracle.gk_distance(X)
        qq = oracle.quantile7(D.ravel(), 0.5)
        assert abs(q.value - qq) <= 1e-12 * qq
    for b in range(nb):
        ref = oracle.gb_kernel(X) if kind == "GB" else oracle.gk_kernel(D, bw[b], qq)
        assert relmax(got[b], ref) <= 1e-12
        assert relmax(got[b], one[b]) <= 1e-12
        assert np.array_equal(got[b], got[b].T)
        if kind == "GK":
            assert np.all(np.diag(got[b]) == 1.0)


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_chain_with_the_all_reduce_inside_the_library(lib, world):
    """bgge_gibbs_set_comm + bgge_gibbs_run_sharded: one handle per device, driven from one host thread per device,
    NCCL all-reduce of the per-step vector on the handle's own (non-blocking) stream -- no host synchronisation
    between the sweep, the collective and the finalize pass.  Chains of all ranks are bitwise identical and match
    the oracle to 1e-9."""
    import bgge_b200
    if lib.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    L, E, p = 300, 3, 400
    X = markers(L, p, 5)
    Y, names = design(L, E)
    ne = [L] * E
    n = L * E
    Kf = bgge_b200.getK(Y, X, kernel="GK", model="MDs", rownames=names, factored=True)
    Kd = oracle.getK(Y, X, kernel="GK", model="MDs", rownames=names)
    rng = np.random.default_rng(2)
    y = rng.standard_normal(n)
    y[rng.choice(n, 9, replace=False)] = np.nan
    Ei = bgge_b200.setDEC(Kf, ne, 1e-10)
    nrs = [sum(b["s"].shape[0] for b in bl) for bl in Ei]
    ite = 8
    nz, nc = oracle.tape_sizes(n, nrs, ite, n_na=9)
    z = rng.standard_normal(nz)
    c = rng.chisquare(np.tile(np.array([nr + 3 for nr in nrs] + [n + 3], dtype=float), ite))
    ref = oracle.BGGE(y, Kd, ne=ne, ite=ite, burn=0, thin=1, draws=oracle.TapeDraws(z, c), eig=Ei)
    comms = _comms(lib, world)
    ne_arr = np.array(ne, dtype=np.int64)
    out, errs = [None] * world, []

    def rank_main(r):
        try:
            lib.call("bgge_set_device", r)
            h = C.c_void_p()
            lib.call("bgge_gibbs_create", n, 2, 0, None, C.byref(h))
            lib.call("bgge_gibbs_set_y", h, lib.ptr(y), None)
            lib.call("bgge_gibbs_set_mode", h, 2)
            lib.call("bgge_gibbs_set_comm", h, comms[r])
            for j, k in enumerate(Kf.values()):
                f = k["factor"]
                lib.call("bgge_gibbs_add_expanded_kernel", h, j, 0 if k["Type"] == "D" else 1, lib.ptr(lib.f64(f["K0"])), L,
                         L, 0, lib.ptr(f["gid"]), lib.ptr(f["env"]), n, f["mode"], f["env_k"], lib.ptr(ne_arr), E, 1e-10, 1)
            lib.call("bgge_gibbs_set_tape", h, lib.ptr(z), z.shape[0], lib.ptr(c), c.shape[0])
            lib.call("bgge_gibbs_init", h, ite, 0, 1, 0.5, lib.RNG_TAPE, 0, 0)
            lib.call("bgge_gibbs_run_sharded", h, 3)
            lib.call("bgge_gibbs_run_sharded", h, ite - 3)
            lib.call("bgge_gibbs_sync", h)
            cm, cv, cu = np.empty(ite), np.empty(ite), np.empty((2, ite))
            lib.call("bgge_gibbs_chain", h, lib.ptr(cm), lib.ptr(cv), lib.ptr(cu), None)
            lib.load().bgge_gibbs_destroy(h)
            out[r] = (cm, cv, cu)
        except Exception as e:   # noqa: BLE001
            errs.append((r, e))

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    _free(lib, comms)
    lib.call("bgge_set_device", 0)
    assert not errs, errs
    for r in range(world):
        assert relmax(out[r][0], ref["chain"]["mu"]) <= 1e-9
        assert relmax(out[r][1], ref["chain"]["varE"]) <= 1e-9
        for j, nm in enumerate(ref["K"]):
            assert relmax(out[r][2][j], ref["chain"]["K"][nm]["varU"]) <= 1e-9
        for a, b in zip(out[r], out[0]):
            assert np.array_equal(a, b)


def test_gram_panels_single_rank_equals_one_shot(lib):
    """world = 1: bgge_dev_gram_panels is the plain mirror build (same bits as bgge_getk_base)."""
    L, p = 300, 200
    X = markers(L, p, 1)
    one = np.empty((1, L, L))
    lib.call("bgge_getk_base", lib.ptr(lib.f64(X)), L, p, lib.GB, None, 1, 0.5, lib.ptr(one), None)
    comms = _comms(lib, 1)
    try:
        got = np.empty((1, L, L))
        lib.call("bgge_getk_base_multi", C.cast(comms, C.POINTER(C.c_void_p)), 1, lib.ptr(lib.f64(X)), L, p, lib.GB, None, 1,
                 0.5, lib.ptr(got), None)
    finally:
        _free(lib, comms)
    assert np.array_equal(got, one)
