"""CPU-only tests: the C-ABI library loads and exports every declared symbol, fails loudly
without a GPU, and the host-side logic (validation, naming, sharding plans) is right."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    from bgge_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "bgge_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+(bgge_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 36
    assert declared == set(built.SIGNATURES), declared ^ set(built.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", built.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (bgge_\w+)", out))
    assert declared <= exported, declared - exported
    lib = built.load()
    assert lib.bgge_version() >= 100
    assert isinstance(lib.bgge_last_error(), bytes)


def test_ctypes_prototypes_have_the_header_arity(built):
    """Every prototype in include/bgge_b200.h has as many parameters as its ctypes signature (ABI drift guard)."""
    hdr = open(os.path.join(ROOT, "include", "bgge_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"^\s*(?:int|void|const char\*)\s+(bgge_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.M | re.S)
    assert len(protos) >= 36
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else len([x for x in params.split(",") if x.strip()])
        assert n == len(built.SIGNATURES[name][1]), (name, n, len(built.SIGNATURES[name][1]))


def test_no_cpu_fallback(built):
    """Without a CUDA device every compute entry point must fail loudly (never compute on the host)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X = np.ones((4, 3))
    out = np.zeros((4, 4))
    with pytest.raises(built.BGGEError) as ei:
        built.call("bgge_getk_base", built.ptr(X), 4, 3, 0, None, 1, 0.5, built.ptr(out), None)
    assert ei.value.code == 5 and "no CPU fallback" in str(ei.value)
    h = C.c_void_p()
    with pytest.raises(built.BGGEError):
        built.call("bgge_gibbs_create", 10, 1, 0, None, C.byref(h))
    with pytest.raises(built.BGGEError):
        built.call("bgge_eig", built.ptr(out), 4, 4, 1e-10, 1, None, None, C.byref(C.c_int64()))
    assert built.device_count() == 0


def test_product_never_imports_oracle():
    for dirpath, _d, files in os.walk(os.path.join(ROOT, "bgge_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_getk_validation_runs_before_native_call():
    import bgge_b200
    L, E = 5, 2
    names = [f"g{i}" for i in range(L)]
    Y = [(f"e{e}", names[i], 0.0) for e in range(E) for i in range(L)]
    X = np.ones((L, 3))
    with pytest.raises(ValueError, match="Single model"):
        bgge_b200.getK(Y, X, kernel="GB", model="SM", rownames=names)
    with pytest.raises(ValueError, match="Genotype names are missing"):
        bgge_b200.getK(Y, X, kernel="GB", model="MM")
    with pytest.raises(ValueError, match="Not all genotypes"):
        bgge_b200.getK(Y, X[:-1], kernel="GB", model="MM", rownames=names[:-1])
    with pytest.raises(ValueError, match="kernel selected is not available"):
        bgge_b200.getK(Y, X, kernel="RBF", model="MM", rownames=names)
    with pytest.raises(ValueError, match="Model selected is not available"):
        bgge_b200.getK(Y, X, kernel="GB", model="Cov", rownames=names)
    with pytest.raises(ValueError, match="EXPR"):
        bgge_b200.getK(Y, X, rownames=names)           # defaults are vectors, no match.arg
    Y1 = [("e0", names[i], 0.0) for i in range(L)]
    with pytest.raises(ValueError):
        bgge_b200.getK(Y1, X, kernel="GB", model="MDs", rownames=names)
    with pytest.warns(UserWarning, match="repeated genotypes"):
        with pytest.raises(Exception):                 # the warning fires, then the native call fails (no GPU) or succeeds
            bgge_b200.getK(Y + [Y[0]], X, kernel="GB", model="MM", rownames=names)
            raise RuntimeError("reached")


def test_bgge_validation_runs_before_native_call():
    import bgge_b200
    n = 6
    K = {"A": {"Kernel": np.eye(n), "Type": "XX"}}
    with pytest.raises(ValueError, match="types BD or D"):
        bgge_b200.BGGE(np.zeros(n), K, ne=[3, 3])
    K = {"A": {"Kernel": np.eye(n), "Type": "BD"}}
    with pytest.raises(ValueError, match="block diagonal"):
        bgge_b200.BGGE(np.zeros(n), K, ne=[6])
    with pytest.raises(ValueError, match="number of subjects"):
        bgge_b200.BGGE(np.zeros(n), K)


def test_factor_coding_matches_oracle():
    import oracle
    from bgge_b200.getk import _factor
    col = np.array(["b", "a", "c", "a", "10", "9"], dtype=object)
    lv, codes = _factor(col)
    lo, co = oracle.design_codes(np.array([str(v) for v in col]))
    assert lv == lo and np.array_equal(codes, co)
    lv, codes = _factor(np.array([3, 1, 2, 1]))
    assert lv == [1, 2, 3] and list(codes) == [2, 0, 1, 0]


def test_shard_units_partitions():
    from bgge_b200.multigpu import shard_units
    for n_units, world in [(320, 8), (320, 1), (7, 4), (3, 8), (0, 2)]:
        seen = []
        for r in range(world):
            seen += list(shard_units(n_units, world, r))
        assert seen == list(range(n_units))
        sizes = [len(shard_units(n_units, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_column_shards_partition_the_columns():
    """column_range (the host mirror of the sharded sweep's planner): disjoint, covering, on column-group boundaries."""
    from bgge_b200.multigpu import column_range
    for nr in (1, 7, 599, 2999, 10000):
        for world in (1, 2, 3, 4, 8):
            for group in (1, 2, 4, 6):
                cur = 0
                for r in range(world):
                    lo, hi = column_range(nr, world, r, group)
                    assert lo == cur and hi >= lo
                    assert r == world - 1 or hi % group == 0
                    cur = hi
                assert cur == nr
    with pytest.raises(ValueError):
        column_range(10, 2, 2)


def test_panel_plan_is_balanced_and_covers():
    from bgge_b200 import multigpu
    for L, world in [(10000, 8), (10000, 2), (599, 4), (3000, 8), (130, 2)]:
        T = (L + 127) // 128
        rows = []
        for c in range(2 * world):
            t0, t1 = multigpu.chunk_tile_rows(L, world, c)
            rows += list(range(t0, t1))
        assert rows == list(range(T))
        tiles = multigpu.tiles_per_rank(L, world)
        assert sum(tiles) == T * (T + 1) // 2
        if T >= 8 * world:
            assert max(tiles) <= 1.1 * (sum(tiles) / world)     # the slowest rank sets the time


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from bgge_b200 import multigpu

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
L, p = 700, 40
rng = np.random.default_rng(0)
X = rng.standard_normal((L, p))
full = (X @ X.T) / p
ch, R, owners = multigpu.panel_plan(L, world)
mine = torch.zeros((2, L, R), dtype=torch.float64)
for k, c in enumerate(owners[rank]):
    t0, t1 = multigpu.chunk_tile_rows(L, world, c)
    for ti in range(t0, t1):                     # numpy stand-in for bgge_dev_gram: lower-triangle tiles only
        i0, i1 = ti * 128, min(L, (ti + 1) * 128)
        blk = full[i0:i1, :i1].copy()
        for tj in range(ti + 1):
            j0, j1 = tj * 128, min(L, (tj + 1) * 128)
            mine[k, j0:j1, i0 - t0 * 128:i1 - t0 * 128] = torch.from_numpy(blk[:, j0:j1].T.copy())
gathered = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
S = torch.zeros((L, L), dtype=torch.float64)
multigpu.assemble_panels(S, gathered, L, world)
S = S.numpy().T                                  # S[j, i] storage -> matrix[i, j]
low = np.tril(S)
S = low + np.tril(S, -1).T
assert np.array_equal(S, full), np.max(np.abs(S - full))
units = list(multigpu.shard_units(320, world, rank))
cnt = torch.tensor([len(units)], dtype=torch.int64)
dist.all_reduce(cnt)
assert int(cnt.item()) == 320
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_row_panel_plan_over_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_factor_matches_compares_every_entry():
    """an edited off-diagonal entry must send BGGE()/setDEC() to the dense route (the reference always
    decomposes the matrix it is given, R/BGGE.R:144,160); the check is exact, not sampled."""
    from bgge_b200 import _lib
    from bgge_b200.getk import factor_matches
    rng = np.random.default_rng(0)
    L, E = 23, 3
    A = rng.standard_normal((L, L))
    K0 = A @ A.T
    gid = np.tile(np.arange(L, dtype=np.int32), E)
    env = np.repeat(np.arange(E, dtype=np.int32), L)
    for mode, k in ((_lib.EXPAND_G, 0), (_lib.EXPAND_GE, 0), (_lib.EXPAND_ENV, 1), (_lib.EXPAND_GI, 0)):
        base = (gid[:, None] == gid[None, :]).astype(float) if mode == _lib.EXPAND_GI else K0[np.ix_(gid, gid)]
        if mode == _lib.EXPAND_GE:
            base = base * (env[:, None] == env[None, :])
        if mode == _lib.EXPAND_ENV:
            base = base * ((env == k)[:, None] & (env == k)[None, :])
        ent = {"Type": "D", "Kernel": np.asfortranarray(base),
               "factor": {"K0": None if mode == _lib.EXPAND_GI else K0, "gid": gid, "env": env, "mode": mode, "env_k": k, "L": L}}
        for chunk in (1 << 24, 97):               # one chunk, and many ragged column chunks
            assert factor_matches(ent, chunk=chunk)
        for (i, j) in ((5, 40), (L * E - 1, 0), (L + 2, L + 3)):
            old = ent["Kernel"][i, j]
            ent["Kernel"][i, j] = old + 1e-9
            assert not factor_matches(ent, chunk=97), (mode, i, j)
            ent["Kernel"][i, j] = old
        assert factor_matches(ent)
        # row-major storage goes through the library as the transpose (K0 handed over in the same order)
        entc = {"Type": "D", "Kernel": np.ascontiguousarray(base).copy(), "factor": ent["factor"]}
        assert factor_matches(entc)
        entc["Kernel"][7, 1] += 1e-12
        assert not factor_matches(entc)
        # a NaN never matches, -0.0 equals 0.0
        entn = {"Type": "D", "Kernel": np.asfortranarray(base.copy()), "factor": ent["factor"]}
        zi = np.argwhere(entn["Kernel"] == 0.0)
        if len(zi):
            entn["Kernel"][tuple(zi[0])] = -0.0
            assert factor_matches(entn)
        entn["Kernel"][3, 3] = np.nan
        assert not factor_matches(entn)
    assert factor_matches({"Type": "D", "factor": ent["factor"]})          # factored-only entry
    assert not factor_matches({"Type": "D", "Kernel": base})               # no description to match


def test_numeric_factor_levels_sort_numerically():
    """factor() on a numeric column sorts numerically ('10' must not come before '2')."""
    from bgge_b200.getk import _factor
    lev, codes = _factor(np.array([10, 2, 33, 2, 10], dtype=object))
    assert lev == [2, 10, 33] and codes.tolist() == [1, 0, 2, 0, 1]
    lev, codes = _factor(np.array(["e10", "e2", "e2"], dtype=object))
    assert lev == ["e10", "e2"] and codes.tolist() == [0, 1, 1]
    lev, codes = _factor(np.array([1.5, 0.25, 1.5], dtype=object))
    assert lev == [0.25, 1.5] and codes.tolist() == [1, 0, 1]


def test_factored_kernel_must_describe_the_rows_of_y():
    """a factored description with a different row count than y is an error, not a truncation."""
    from bgge_b200.bgge import _check_factor_rows
    f = {"gid": np.zeros(12, dtype=np.int32), "env": np.zeros(12, dtype=np.int32)}
    _check_factor_rows({"factor": f}, 12, "G")
    for n in (11, 13):
        with pytest.raises(ValueError, match="non-conformable"):
            _check_factor_rows({"factor": f}, n, "G")


def test_r_shim_type_checks_against_the_header():
    """R is not in the image, so the .Call shim cannot be built; it is at least type-checked: every libbgge_b200 call
    in r-shim/src/bgge_shim.c must match include/bgge_b200.h (arity, pointer types, struct fields) and every R API
    call a mock of Rinternals.h with R 4.x's prototypes (tests/mock_r).  Catches ABI drift without R."""
    cmd = ["gcc", "-fsyntax-only", "-std=c11", "-Wall", "-Wextra", "-Wno-cast-function-type", "-Werror",
           "-I", os.path.join(ROOT, "tests", "mock_r"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "r-shim", "src", "bgge_shim.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the routines the R code calls are the ones the shim registers, with the registered arities
    shim = open(os.path.join(ROOT, "r-shim", "src", "bgge_shim.c")).read()
    rsrc = open(os.path.join(ROOT, "r-shim", "R", "native.R")).read()
    reg = dict(re.findall(r'\{"(bgge_R_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', shim))
    assert set(re.findall(r'\.Call\("(bgge_R_\w+)"', rsrc)) == set(reg)
    for name, nargs in reg.items():
        sig = re.search(r"SEXP %s\(([^)]*)\)" % name, shim).group(1)
        assert sig.count("SEXP") == int(nargs), name
    # the header is plain C (the shim includes it from a C translation unit)
    r = subprocess.run(["gcc", "-fsyntax-only", "-std=c99", "-Wall", "-Werror", "-x", "c", os.path.join(ROOT, "include", "bgge_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_factor_matches_entry_point_with_pitch_and_bad_codes():
    """bgge_factor_matches through ctypes (host-only, no device needed): a pitched K (ldk > n), a mismatch in the last
    column, out-of-range genotype codes rejected with the library's error."""
    import ctypes as C
    from bgge_b200 import _lib
    rng = np.random.default_rng(3)
    L, n, ldk = 9, 31, 40
    K0 = np.asfortranarray(rng.standard_normal((L, L)))
    gid = rng.integers(0, L, size=n).astype(np.int32)
    env = rng.integers(0, 3, size=n).astype(np.int32)
    Kp = np.full((ldk, n), np.nan, order="F")
    Kp[:n, :] = K0[np.ix_(gid, gid)] * (env[:, None] == env[None, :])
    ok = C.c_int(-1)
    _lib.call("bgge_factor_matches", _lib.ptr(Kp), n, ldk, _lib.ptr(K0), L, _lib.ptr(gid), _lib.ptr(env), _lib.EXPAND_GE, 0, C.byref(ok))
    assert ok.value == 1
    Kp[n - 1, n - 1] += 1.0
    _lib.call("bgge_factor_matches", _lib.ptr(Kp), n, ldk, _lib.ptr(K0), L, _lib.ptr(gid), _lib.ptr(env), _lib.EXPAND_GE, 0, C.byref(ok))
    assert ok.value == 0
    bad = gid.copy()
    bad[4] = L
    with pytest.raises(_lib.BGGEError, match="outside"):
        _lib.call("bgge_factor_matches", _lib.ptr(Kp), n, ldk, _lib.ptr(K0), L, _lib.ptr(bad), _lib.ptr(env), _lib.EXPAND_GE, 0, C.byref(ok))
