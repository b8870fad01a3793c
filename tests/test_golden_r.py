"""Consumer of tests/golden/r_*.json -- files produced by the REFERENCE ITSELF (tests/golden/make_golden.R, which
needs R and cannot run in this repository's image).  For every file present: getK against the recorded kernels
(<= 1e-12) and BGGE on the recorded draws and R's own eigen() output against the recorded chains (<= 1e-9), for the
oracle (CPU) and for the CUDA path (-m gpu).  No file present -> the reference-pinned tests skip and parity stays
"unpinned"; the loader itself is exercised on a stand-in file with the same schema written by the oracle (which
proves the plumbing, not the parity)."""
import glob
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

import oracle
from helpers import relmax

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "r_*.json")))


def load(path):
    g = json.load(open(path))
    g["X"] = np.array(g["X"], dtype=np.float64)
    g["y"] = np.array([np.nan if v is None else v for v in g["y"]], dtype=np.float64)
    g["XF"] = None if g.get("XF") is None else np.array(g["XF"], dtype=np.float64)
    g["Y"] = list(zip(g["env"], g["gid"], [0.0] * len(g["gid"])))
    g["Kd"] = OrderedDict((nm, {"Type": g["K"][nm]["Type"], "Kernel": np.array(g["K"][nm]["Kernel"], dtype=np.float64)})
                          for nm in g["kernel_names"])
    g["z"] = np.array(g["tape"]["normals"], dtype=np.float64)
    g["c"] = np.array(g["tape"]["chisq"], dtype=np.float64)
    return g


def eigen_blocks(g):
    """setDEC()'s structure (R/BGGE.R:119-182) from the eigen() calls the reference made, in call order: one call per
    type-D kernel, one per `ne` block of a type-BD kernel; values > tol kept, blocks keeping none dropped (:162)."""
    calls = iter(g["eigen"])
    ne = np.asarray(g["ne"], dtype=np.int64)
    posf = np.cumsum(ne)
    posi = posf - ne
    out = []
    for nm in g["kernel_names"]:
        spans = [None] if g["K"][nm]["Type"] == "D" else [(int(a), int(b)) for a, b in zip(posi, posf)]
        blocks = []
        for sp in spans:
            e = next(calls)
            w = np.atleast_1d(np.array(e["values"], dtype=np.float64))
            V = np.array(e["vectors"], dtype=np.float64).reshape(w.shape[0], -1)
            keep = w > g["tol"]
            if keep.any():
                blocks.append({"s": w[keep].copy(), "U": np.asfortranarray(V[:, keep]), "pos": sp})
        out.append(blocks)
    return out


def bw_arg(g):
    b = g.get("bandwidth", 1)
    return b if isinstance(b, list) else [b]


def check_fit(got, g, tol=1e-9):
    f = g["fit"]
    assert relmax(got["chain"]["mu"], f["chain"]["mu"]) <= tol
    assert relmax(got["chain"]["varE"], f["chain"]["varE"]) <= tol
    for nm in g["kernel_names"]:
        assert relmax(got["chain"]["K"][nm]["varU"], f["chain"]["K"][nm]["varU"]) <= tol, nm
        assert relmax(got["K"][nm]["u"], f["K"][nm]["u"]) <= tol, nm
        assert relmax(got["K"][nm]["u.sd"], f["K"][nm]["u_sd"]) <= 1e-7, nm
        assert abs(got["K"][nm]["varu"] - f["K"][nm]["varu"]) <= tol * abs(f["K"][nm]["varu"]), nm
    assert relmax(np.asarray(got["yHat"]).ravel(), f["yHat"]) <= tol
    assert abs(got["varE"] - f["varE"]) <= tol * f["varE"]
    assert relmax(got["y"], f["y"]) <= tol


def run_oracle(g):
    K = oracle.getK(g["Y"], g["X"], kernel=g["kernel"], model=g["model"], rownames=g["rownames"], bandwidth=bw_arg(g),
                    intercept_random=bool(g.get("intercept_random", False)))
    assert list(K) == g["kernel_names"]
    for nm in K:
        assert K[nm]["Type"] == g["K"][nm]["Type"]
        assert relmax(K[nm]["Kernel"], g["Kd"][nm]["Kernel"]) <= 1e-12, nm
    fit = oracle.BGGE(g["y"], g["Kd"], XF=g["XF"], ne=g["ne"], ite=g["ite"], burn=g["burn"], thin=g["thin"], tol=g["tol"],
                      R2=g["R2"], draws=oracle.TapeDraws(g["z"], g["c"]), eig=eigen_blocks(g))
    check_fit(fit, g)


def run_cuda(g):
    import bgge_b200
    K = bgge_b200.getK(g["Y"], g["X"], kernel=g["kernel"], model=g["model"], rownames=g["rownames"], bandwidth=bw_arg(g),
                       intercept_random=bool(g.get("intercept_random", False)))
    assert list(K) == g["kernel_names"]
    for nm in K:
        assert K[nm]["Type"] == g["K"][nm]["Type"]
        assert relmax(K[nm]["Kernel"], g["Kd"][nm]["Kernel"]) <= 1e-12, nm
    for mode in (1, 2, 0):
        fit = bgge_b200.BGGE(g["y"], g["Kd"], XF=g["XF"], ne=g["ne"], ite=g["ite"], burn=g["burn"], thin=g["thin"],
                             tol=g["tol"], R2=g["R2"], draws=(g["z"], g["c"]), eig=eigen_blocks(g), mode=mode)
        check_fit(fit, g)


@pytest.mark.skipif(not FILES, reason="no tests/golden/r_*.json: run tests/golden/make_golden.R where R is installed")
@pytest.mark.parametrize("path", FILES or ["-"])
def test_oracle_matches_the_reference_run(path):
    run_oracle(load(path))


@pytest.mark.gpu
@pytest.mark.skipif(not FILES, reason="no tests/golden/r_*.json: run tests/golden/make_golden.R where R is installed")
@pytest.mark.parametrize("path", FILES or ["-"])
def test_cuda_path_matches_the_reference_run(path):
    run_cuda(load(path))


# ---------------------------------------------------------------------------------------------------------------------
# the loader on a stand-in file (schema of make_golden.R, contents from the oracle): plumbing only, NOT a pin
# ---------------------------------------------------------------------------------------------------------------------
class _Recorder(oracle.RandomDraws):
    def __init__(self, seed):
        super().__init__(seed)
        self.z, self.c = [], []

    def normals(self, n):
        v = super().normals(n)
        self.z.extend(np.atleast_1d(v).tolist())
        return v

    def chisq(self, df):
        v = super().chisq(df)
        self.c.append(float(v))
        return v


def _stand_in(tmp_path, kernel, model, with_xf, n_na):
    from helpers import design, markers
    L, p, E = 14, 20, 3
    X = markers(L, p, 7)
    Y, names = design(L, E)
    n = L * E
    rng = np.random.default_rng(3)
    y = rng.standard_normal(n)
    if n_na:
        y[rng.choice(n, n_na, replace=False)] = np.nan
    XF = None
    if with_xf:
        XF = np.zeros((n, E))
        for e in range(E):
            XF[e * L:(e + 1) * L, e] = 1.0
    K = oracle.getK(Y, X, kernel=kernel, model=model, rownames=names)
    ne = [L] * E
    Ei = oracle.setDEC(K, ne, 1e-10)
    rec = _Recorder(5)
    fit = oracle.BGGE(y, K, XF=XF, ne=ne, ite=12, burn=2, thin=2, draws=rec, eig=Ei)
    eig_calls = []
    for nm, blocks in zip(K, Ei):
        spans = [None] if K[nm]["Type"] == "D" else [(e * L, (e + 1) * L) for e in range(E)]
        for sp in spans:
            sub = K[nm]["Kernel"] if sp is None else K[nm]["Kernel"][sp[0]:sp[1], sp[0]:sp[1]]
            w, V = np.linalg.eigh(sub)            # full spectrum, descending, like eigen(): the loader truncates
            blk = [b for b in blocks if b["pos"] == sp]
            w, V = w[::-1].copy(), V[:, ::-1].copy()
            if blk:                                # keep the oracle's own vectors for the kept part (sign / rotation)
                k = blk[0]["s"].shape[0]
                w[:k], V[:, :k] = blk[0]["s"], blk[0]["U"]
                w[k:] = np.minimum(w[k:], 1e-12)
            eig_calls.append({"values": w.tolist(), "vectors": V.tolist()})
    obj = {"tag": "standin", "L": L, "p": p, "E": E, "kernel": kernel, "model": model, "bandwidth": 1,
           "intercept_random": False, "ite": 12, "burn": 2, "thin": 2, "ne": ne, "tol": 1e-10, "R2": 0.5,
           "X": X.tolist(), "rownames": names, "env": [r[0] for r in Y], "gid": [r[1] for r in Y],
           "y": [None if np.isnan(v) else float(v) for v in y], "XF": None if XF is None else XF.tolist(),
           "K": {nm: {"Type": k["Type"], "Kernel": k["Kernel"].tolist()} for nm, k in K.items()},
           "kernel_names": list(K), "tape": {"normals": rec.z, "chisq": rec.c}, "eigen": eig_calls,
           "fit": {"yHat": np.asarray(fit["yHat"]).ravel().tolist(), "varE": fit["varE"], "varE_sd": fit["varE.sd"],
                   "K": {nm: {"u": k["u"].tolist(), "u_sd": k["u.sd"].tolist(), "varu": k["varu"], "varu_sd": k["varu.sd"]}
                         for nm, k in fit["K"].items()},
                   "chain": {"mu": fit["chain"]["mu"].tolist(), "varE": fit["chain"]["varE"].tolist(),
                             "K": {nm: {"varU": k["varU"].tolist()} for nm, k in fit["chain"]["K"].items()}},
                   "y": fit["y"].tolist()}}
    path = os.path.join(str(tmp_path), "r_standin.json")
    json.dump(obj, open(path, "w"))
    return path


@pytest.mark.parametrize("kernel,model,with_xf,n_na", [("GK", "MDs", False, 3), ("GB", "MDe", True, 2)])
def test_loader_round_trip_on_a_stand_in_file(tmp_path, kernel, model, with_xf, n_na):
    run_oracle(load(_stand_in(tmp_path, kernel, model, with_xf, n_na)))


@pytest.mark.gpu
def test_loader_round_trip_on_a_stand_in_file_cuda(tmp_path):
    run_cuda(load(_stand_in(tmp_path, "GK", "MDs", False, 3)))
