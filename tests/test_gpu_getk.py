"""GPU parity: getK through the C ABI vs the CPU oracle (bit-exact expansions, <=1e-12 kernels)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import design, markers, relmax

pytestmark = pytest.mark.gpu

TOL_K = 1e-12   # north_star: K matches the reference to <= 1e-12 relative (max|dK| / max|K|)


@pytest.mark.parametrize("L,p", [(5, 7), (37, 50), (128, 16), (129, 333), (300, 1279)])
def test_gb_base_kernel(lib, L, p):
    X = markers(L, p, 11)
    out = np.empty((L, L), order="F")
    lib.call("bgge_getk_base", lib.ptr(lib.f64(X)), L, p, lib.GB, None, 1, 0.5, lib.ptr(out), None)
    ref = oracle.gb_kernel(X)
    assert relmax(out, ref) <= TOL_K
    assert np.array_equal(out, out.T)          # tcrossprod is exactly symmetric


@pytest.mark.parametrize("L,p,quantil", [(6, 9, 0.5), (37, 50, 0.5), (130, 257, 0.25), (200, 64, 0.9), (64, 40, 0.0),
                                         (64, 40, 1.0)])
def test_gk_base_kernel(lib, L, p, quantil):
    X = markers(L, p, 5)
    bw = np.array([1.0, 0.2, 5.0])
    out = np.empty((3, L, L))
    outs = np.empty(3 * L * L)
    q = C.c_double()
    lib.call("bgge_getk_base", lib.ptr(lib.f64(X)), L, p, lib.GK, lib.ptr(bw), 3, quantil, lib.ptr(outs), C.byref(q))
    D = oracle.gk_distance(X)
    qref = oracle.quantile7(D, quantil)
    if qref > 0:
        assert abs(q.value - qref) <= 1e-12 * qref
    for b in range(3):
        got = outs[b * L * L:(b + 1) * L * L].reshape(L, L, order="F")
        if qref == 0:
            continue
        ref = oracle.gk_kernel(D, bw[b], qref)
        assert relmax(got, ref) <= TOL_K
        assert np.all(np.diag(got) == 1.0)     # exact-zero distance on the diagonal
        assert np.array_equal(got, got.T)


def test_quantile_exact_selection(lib):
    """The device quantile is an exact order statistic: feed a matrix, compare to sort."""
    import torch
    rng = np.random.default_rng(3)
    for L, prob in [(33, 0.5), (100, 0.123), (257, 0.75), (64, 0.5)]:
        A = rng.gamma(2.0, 50.0, size=(L, L))
        A = (A + A.T) / 2
        np.fill_diagonal(A, 0.0)
        A[3, 5] = A[5, 3] = A[7, 9] = A[9, 7]      # ties
        d = torch.from_numpy(np.asfortranarray(A).T.copy()).cuda()   # column-major bytes
        hq = C.c_double()
        lib.call("bgge_dev_quantile7", C.c_void_p(d.data_ptr()), L, L, prob, None, C.byref(hq), None)
        assert hq.value == oracle.quantile7(A, prob)


MODELS = [("SM", 1), ("MM", 3), ("MDs", 3), ("MDe", 3)]


@pytest.mark.parametrize("kernel", ["GB", "GK"])
@pytest.mark.parametrize("model,E", MODELS)
def test_getk_models_match_oracle(kernel, model, E):
    import bgge_b200
    L, p = 41, 60
    X = markers(L, p, 2)
    Y, names = design(L, E, seed=4, shuffle=(model != "SM"), drop=(5 if E > 1 else 0))
    bw = [1.0, 0.5] if kernel == "GK" else 1
    got = bgge_b200.getK(Y, X, kernel=kernel, model=model, bandwidth=bw, rownames=names, intercept_random=True)
    ref = oracle.getK(Y, X, kernel=kernel, model=model, bandwidth=bw, rownames=names, intercept_random=True)
    assert list(got.keys()) == list(ref.keys())
    for nm in ref:
        assert got[nm]["Type"] == ref[nm]["Type"]
        assert got[nm]["Kernel"].shape == ref[nm]["Kernel"].shape
        assert relmax(got[nm]["Kernel"], ref[nm]["Kernel"]) <= TOL_K, nm
        # structural zeros / ones of the expansions are exact
        assert np.array_equal(got[nm]["Kernel"] == 0, ref[nm]["Kernel"] == 0), nm


def test_expansion_is_bit_exact_gather(lib):
    rng = np.random.default_rng(0)
    L, n, E = 23, 101, 4
    K0 = rng.standard_normal((L, L))
    gid = rng.integers(0, L, size=n).astype(np.int32)
    env = np.sort(rng.integers(0, E, size=n)).astype(np.int32)
    for mode, k in [(lib.EXPAND_G, 0), (lib.EXPAND_GE, 0), (lib.EXPAND_ENV, 2), (lib.EXPAND_GI, 0)]:
        out = np.empty((n, n), order="F")
        lib.call("bgge_getk_expand", lib.ptr(lib.f64(K0)), L, lib.ptr(gid), lib.ptr(env), n, mode, k, lib.ptr(out))
        G = K0[np.ix_(gid, gid)]
        if mode == lib.EXPAND_GE:
            G = G * (env[:, None] == env[None, :])
        elif mode == lib.EXPAND_ENV:
            G = G * ((env == k)[:, None] & (env == k)[None, :])
        elif mode == lib.EXPAND_GI:
            G = (gid[:, None] == gid[None, :]).astype(np.float64)
        assert np.array_equal(out, G)


@pytest.mark.parametrize("n,sort_min", [(2500, None), (2501, None), (101, "1"), (2500, "1000000")])
def test_expansion_in_genotype_order_is_bit_exact(lib, monkeypatch, n, sort_min):
    """Large outputs are written column by column in genotype order (counting sort of the row codes on the device,
    16-byte streaming stores); odd n takes the 8-byte kernel.  Shuffled, unbalanced codes, some genotypes absent."""
    if sort_min is not None:
        monkeypatch.setenv("BGGE_EXPAND_SORT_MIN", sort_min)
    rng = np.random.default_rng(n)
    L, E = 300, 5
    K0 = rng.standard_normal((L, L))
    gid = rng.integers(0, L - 7, size=n).astype(np.int32)
    env = rng.integers(0, E, size=n).astype(np.int32)
    for mode, k in [(lib.EXPAND_G, 0), (lib.EXPAND_GE, 0), (lib.EXPAND_ENV, 3), (lib.EXPAND_GI, 0)]:
        out = np.full((n, n), np.nan, order="F")
        lib.call("bgge_getk_expand", lib.ptr(lib.f64(K0)), L, lib.ptr(gid), lib.ptr(env), n, mode, k, lib.ptr(out))
        G = K0[np.ix_(gid, gid)]
        if mode == lib.EXPAND_GE:
            G = G * (env[:, None] == env[None, :])
        elif mode == lib.EXPAND_ENV:
            G = G * ((env == k)[:, None] & (env == k)[None, :])
        elif mode == lib.EXPAND_GI:
            G = (gid[:, None] == gid[None, :]).astype(np.float64)
        assert np.array_equal(out, G), mode


def test_staged_host_copies_are_exact(lib, monkeypatch):
    """Large pageable host buffers travel through the threaded pinned-staging pipeline (csrc/staging.cu): marker upload
    (pitched destination), base-kernel download, n x n expansion download.  Same bits as the plain cudaMemcpy path."""
    rng = np.random.default_rng(5)
    L, p, E = 700, 3000, 3
    X = np.asfortranarray(rng.standard_normal((L, p)))
    n = L * E
    gid = np.tile(np.arange(L, dtype=np.int32), E)
    env = np.repeat(np.arange(E, dtype=np.int32), L)

    def run():
        K0 = np.empty((1, L, L))
        lib.call("bgge_getk_base", lib.ptr(X), L, p, lib.GB, None, 0, 0.5, lib.ptr(K0), None)
        out = np.full((n, n), np.nan, order="F")
        lib.call("bgge_getk_expand", lib.ptr(np.asfortranarray(K0[0].T)), L, lib.ptr(gid), lib.ptr(env), n, lib.EXPAND_GE, 0,
                 lib.ptr(out))
        return K0[0].copy(), out

    monkeypatch.setenv("BGGE_NO_STAGED_COPY", "1")
    K_plain, G_plain = run()
    monkeypatch.delenv("BGGE_NO_STAGED_COPY")
    monkeypatch.setenv("BGGE_STAGED_MIN_BYTES", str(1 << 20))
    for threads in ("2", "5", "8"):
        monkeypatch.setenv("BGGE_COPY_THREADS", threads)
        K_st, G_st = run()
        assert np.array_equal(K_st, K_plain), threads
        assert np.array_equal(G_st, G_plain), threads
    assert relmax(K_plain, oracle.gb_kernel(X)) <= 1e-12


@pytest.mark.parametrize("kind,chunks", [("GB", "3"), ("GK", "2"), ("GB", "5")])
def test_host_marker_matrix_in_chunks_matches_one_shot(lib, monkeypatch, kind, chunks):
    """Large host marker matrices are contracted chunk by chunk (upload of chunk c + 1 under the contraction of chunk c,
    planes added in chunk order): same kernel as the one-shot build to rounding, <= 1e-12 to the oracle.  Ragged last
    chunk (p = 3000 is not a multiple of the 16-aligned chunk width)."""
    rng = np.random.default_rng(11)
    L, p = 333, 3000
    X = np.asfortranarray(rng.standard_normal((L, p)))
    bw = np.array([0.7])

    def run():
        K0 = np.empty((1, L, L))
        lib.call("bgge_getk_base", lib.ptr(X), L, p, lib.GB if kind == "GB" else lib.GK, lib.ptr(bw), 1, 0.5, lib.ptr(K0), None)
        return K0[0].T.copy()

    one = run()
    monkeypatch.setenv("BGGE_GRAM_HOST_CHUNKS", chunks)
    monkeypatch.setenv("BGGE_STAGED_MIN_BYTES", "65536")
    got = run()
    if kind == "GB":
        ref = oracle.gb_kernel(X)
    else:
        D = oracle.gk_distance(X)
        ref = oracle.gk_kernel(D, 0.7, oracle.quantile7(D.ravel(), 0.5))
    assert relmax(got, one) <= 1e-13
    assert relmax(got, ref) <= 1e-12
    assert np.array_equal(got, got.T)


def test_setkernel_path_and_errors():
    import bgge_b200
    L, E = 12, 2
    rng = np.random.default_rng(1)
    A = rng.standard_normal((L, L))
    A = A @ A.T
    Y, names = design(L, E)
    perm = rng.permutation(L)
    user = {"A": (A[np.ix_(perm, perm)], [names[i] for i in perm])}
    got = bgge_b200.getK(Y, setKernel=user, model="MDs")
    ref = oracle.getK(Y, setKernel=user, model="MDs")
    assert list(got.keys()) == ["G", "GE"] == list(ref.keys())
    for nm in ref:
        assert np.array_equal(got[nm]["Kernel"], ref[nm]["Kernel"])
    with pytest.raises(ValueError):
        bgge_b200.getK(Y, np.ones((L, 3)), kernel="GB", model="SM", rownames=names)      # >1 env under SM
    Y1, _ = design(L, 1)
    with pytest.raises(ValueError):
        bgge_b200.getK(Y1, np.ones((L, 3)), kernel="GK", model="MM", rownames=names)     # 1 env under MM
    with pytest.raises(ValueError):
        bgge_b200.getK(Y, np.ones((L, 3)), kernel="GB", model="MM")                      # no rownames
    with pytest.raises(ValueError):
        bgge_b200.getK(Y, np.ones((L, 3)), kernel="XX", model="MM", rownames=names)


def test_gram_row_panels_and_symmetrize(lib):
    """Row-panel build (the multi-GPU decomposition) == one-shot build."""
    import torch
    L, p = 300, 160
    X = markers(L, p, 9)
    ldx, pp = C.c_int64(), C.c_int64()
    lib.call("bgge_marker_ld", L, C.byref(ldx))
    lib.call("bgge_marker_cols", p, C.byref(pp))
    Xp = torch.zeros((pp.value, ldx.value), dtype=torch.float64, device="cuda")
    Xp[:p, :L] = torch.from_numpy(X.T.copy()).cuda()
    S = torch.zeros((L, L), dtype=torch.float64, device="cuda")
    T = (L + 127) // 128
    for tr in range(T):   # one tile row per "rank", lower triangle only
        lib.call("bgge_dev_gram", C.c_void_p(Xp.data_ptr()), ldx.value, L, pp.value, tr, tr + 1,
                 C.c_void_p(S.data_ptr()), L, 0, float(p), 0, None)
    lib.call("bgge_dev_symmetrize", C.c_void_p(S.data_ptr()), L, L, None)
    torch.cuda.synchronize()
    full = torch.zeros((L, L), dtype=torch.float64, device="cuda")
    lib.call("bgge_dev_gram", C.c_void_p(Xp.data_ptr()), ldx.value, L, pp.value, 0, T, C.c_void_p(full.data_ptr()), L,
             0, float(p), 1, None)
    torch.cuda.synchronize()
    assert torch.equal(S, full)
    assert relmax(full.cpu().numpy().T, oracle.gb_kernel(X)) <= TOL_K
