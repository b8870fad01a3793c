"""getK(): kernel construction for G x E models -- host mirror of R/getK.R:95-237.

Signature, argument meaning, returned structure, names and error behaviour follow the
reference; all numeric work (R/getK.R:136-149, 138/148/173, 199-201, 207-218, 232) runs
in libbgge_b200 on the GPU.  What stays here is what stays in R: factor coding,
validation, naming.
"""
from __future__ import annotations

import ctypes as C
import warnings
from collections import OrderedDict

import numpy as np

from . import _lib

_MODELS = {"SM": 0, "MM": 0, "MDs": 1, "MDe": 2}


def _factor(col):
    """factor(): sorted unique levels, 0-based codes (R/getK.R:101-104)."""
    col = np.asarray(col)
    if col.dtype == object:
        # factor() on a numeric column sorts numerically; only genuinely non-numeric columns are compared as strings
        # (codepoint order -- R collates by locale; only the order of the MDe environment kernels could differ)
        if all(isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool) for v in col):
            col = np.array([float(v) for v in col])
            if np.all(col == np.round(col)):
                col = col.astype(np.int64)
        else:
            col = np.array([str(v) for v in col])
    levels, codes = np.unique(col, return_inverse=True)
    return [v.item() if hasattr(v, "item") else v for v in levels], np.ascontiguousarray(codes, dtype=np.int32)


def _columns(Y):
    if hasattr(Y, "iloc"):
        return np.asarray(Y.iloc[:, 0]), np.asarray(Y.iloc[:, 1])
    Ya = np.asarray(Y, dtype=object)
    if Ya.ndim != 2 or Ya.shape[1] < 2:
        raise ValueError("Y must have at least two columns (environment, genotype)")
    return Ya[:, 0], Ya[:, 1]


def _named_matrix(km):
    if isinstance(km, tuple):
        mat, names = km
        return np.asarray(mat, dtype=np.float64), [str(v) for v in names]
    if hasattr(km, "index") and hasattr(km, "to_numpy"):
        return km.to_numpy(dtype=np.float64), [str(v) for v in km.index]
    return None, None


def factor_matches(entry, chunk=1 << 24):
    """True when entry["Kernel"] still IS the expansion its "factor" describes (or when there is no dense matrix).
    A caller may have edited the matrix after getK() -- the reference always decomposes the matrix it is given
    (R/BGGE.R:144,160) -- so EVERY entry is compared, bit for bit, in column chunks (an n^2 pass over host memory,
    negligible next to the eigendecomposition it gates).  Any difference sends BGGE()/setDEC() to the dense route."""
    f, K = entry.get("factor"), entry.get("Kernel")
    if f is None or K is None:
        return f is not None
    gid, env, mode, k = f["gid"], f["env"], f["mode"], f["env_k"]
    n = gid.shape[0]
    K = np.asarray(K)
    if K.shape != (n, n):
        return False
    if K.dtype == np.float64 and (K.flags.f_contiguous or K.flags.c_contiguous) and chunk == 1 << 24:
        # the library's threaded host pass (a 40000^2 kernel: a fraction of a second instead of ~10 s of numpy).  A
        # C-ordered K is read as its transpose, so K0 is handed over in the same order: (K')[i,j] vs (K0')[g_i,g_j]
        # is K[j,i] vs K0[g_j,g_i], the same set of comparisons
        K0 = f["K0"]
        if K0 is not None:
            K0 = np.asfortranarray(K0, dtype=np.float64) if K.flags.f_contiguous else np.ascontiguousarray(K0, dtype=np.float64)
        gid32 = np.ascontiguousarray(gid, dtype=np.int32)
        env32 = np.ascontiguousarray(env, dtype=np.int32)
        if gid32.size and (gid32.min() < 0 or gid32.max() >= f["L"]):
            return False
        ok = C.c_int(0)
        _lib.call("bgge_factor_matches", _lib.ptr(K), n, n, None if K0 is None else _lib.ptr(K0), int(f["L"]), _lib.ptr(gid32),
                  _lib.ptr(env32), int(mode), int(k), C.byref(ok))
        return bool(ok.value)
    cols = max(1, min(n, chunk // max(1, n)))
    for c0 in range(0, n, cols):
        c1 = min(n, c0 + cols)
        gj = gid[c0:c1]
        base = (gid[:, None] == gj[None, :]).astype(np.float64) if f["K0"] is None else f["K0"][np.ix_(gid, gj)]
        if mode == _lib.EXPAND_GE:
            base = base * (env[:, None] == env[None, c0:c1])
        elif mode == _lib.EXPAND_ENV:
            base = base * ((env == k)[:, None] & (env[c0:c1] == k)[None, :])
        if not np.array_equal(K[:, c0:c1], base):
            return False
    return True


def materialize(entry):
    """Dense n x n matrix of a factored kernel entry (see getK(..., factored=True))."""
    if "Kernel" in entry and entry["Kernel"] is not None:
        return entry["Kernel"]
    f = entry["factor"]
    n, L = f["gid"].shape[0], f["L"]
    out = np.empty((n, n), dtype=np.float64, order="F")
    _lib.call("bgge_getk_expand", _lib.ptr(f["K0"]), L, _lib.ptr(f["gid"]), _lib.ptr(f["env"]), n, f["mode"], f["env_k"],
              _lib.ptr(out))
    return out


def getK(Y, X=None, kernel=("GK", "GB"), setKernel=None, bandwidth=1, model=("SM", "MM", "MDs", "MDe"), quantil=0.5,
         intercept_random=False, rownames=None, factored=False):
    """Create kernels for the genomic G x E models of BGGE().

    Y        : rows of (environment, genotype, ...) -- DataFrame, 2-D array or list of tuples.
    X        : L x p marker matrix (pre-scaled by the user); row names via ``rownames`` or a
               DataFrame index.  Not used when ``setKernel`` is given.
    kernel   : "GB" (X X'/p) or "GK" (Gaussian, exp(-bandwidth * D / quantile(D, quantil))).
    setKernel: dict/list of user kernels, each a (matrix, names) tuple or a DataFrame.
    model    : "SM", "MM", "MDs" or "MDe".
    Returns an OrderedDict name -> {"Kernel": n x n float64 (Fortran order), "Type": "D"|"BD"} (every entry
    also carries "factor", the L x L description it was expanded from).

    factored=True (extension): the n x n matrices are not built; each entry carries
    {"Type", "factor": {"K0": L x L base kernel, "gid", "env", "mode", "env_k", "L"}} instead, which BGGE()
    decomposes through the L x L problem (bgge_gibbs_add_expanded_kernel).  materialize(entry) gives the
    dense matrix on demand.
    """
    if not isinstance(model, str) or (setKernel is None and not isinstance(kernel, str)):
        # the reference has no match.arg: switch() on a length-2 vector errors;
        # `kernel` is only looked at when setKernel is NULL (R/getK.R:124,133)
        raise ValueError("EXPR must be a length 1 vector: pass `kernel` and `model` explicitly")
    ecol, gcol = _columns(Y)
    env_levels, env = _factor(ecol)
    subjects, gid = _factor(gcol)
    n_env, L, n = len(env_levels), len(subjects), gid.shape[0]

    # check for repeated genotypes (R/getK.R:108-109)
    cell = env.astype(np.int64) * L + gid
    if np.unique(cell).shape[0] != n:
        warnings.warn("There are repeated genotypes in some environment. They were kept")

    if model == "SM":
        if n_env > 1:
            raise ValueError("Single model choosen, but more than one environment is in the phenotype file")
    elif model in ("MM", "MDs", "MDe"):
        if n_env < 2:   # model.matrix() on a one-level factor (R/getK.R:120; tests/testthat/test.R:16-17)
            raise ValueError("contrasts can be applied only to factors with 2 or more levels")
    else:
        raise ValueError("Model selected is not available ")
    model_code = _MODELS[model]

    if setKernel is None:
        if hasattr(X, "index") and hasattr(X, "to_numpy") and rownames is None:
            rownames = list(X.index)
            X = X.to_numpy()
        if rownames is None:
            raise ValueError("Genotype names are missing")
        pos = {str(r): i for i, r in enumerate(rownames)}
        if not all(str(s) in pos for s in subjects):
            raise ValueError("Not all genotypes presents in the phenotypic file are in marker matrix")
        Xs = _lib.f64(np.asarray(X, dtype=np.float64)[[pos[str(s)] for s in subjects], :])   # X[subjects,]
        p = Xs.shape[1]
        if kernel == "GB":
            kind, bw = _lib.GB, np.ones(1)
        elif kernel == "GK":
            kind, bw = _lib.GK, np.ascontiguousarray(np.atleast_1d(bandwidth), dtype=np.float64)
        else:
            raise ValueError("kernel selected is not available. Please choose one method available or make "
                             "available other kernel through argument K")
        nbase = 1 if kind == _lib.GB else bw.shape[0]
        # base kernels over the L genotypes (R/getK.R:136, 142-146) ...
        K0s = np.empty((nbase, L, L), dtype=np.float64)
        _lib.call("bgge_getk_base", _lib.ptr(Xs), L, p, kind, _lib.ptr(bw), nbase, float(quantil), _lib.ptr(K0s), None)
        base_f = [np.asfortranarray(K0s[b].T) for b in range(nbase)]   # symmetric: layout only
        fres = _factored_result(base_f, [str(i + 1) for i in range(nbase)], gid, env, env_levels, L, model_code,
                                intercept_random)
        if factored:
            return fres
        # ... and their n x n expansions (R/getK.R:138, 148, 199-201, 207-218, 232).  Every dense entry also keeps
        # its factored description, so BGGE() can decompose through the L x L problem (it verifies first that the
        # matrix still matches the description).
        for ent in fres.values():
            ent["Kernel"] = materialize(ent)
        return fres
    else:
        items = list(setKernel.items()) if hasattr(setKernel, "items") else [(None, k) for k in setKernel]
        base = []
        for _nm, km in items:
            mat, names = _named_matrix(km)
            if mat is None:
                raise ValueError("Genotype names are missing in some kernel")
            pos = {r: i for i, r in enumerate(names)}
            if not all(str(s) in pos for s in subjects):
                raise ValueError("Not all genotypes presents in phenotypic file are in the kernel matrix.\n"
                                 "             Please check dimnames")
            idx = [pos[str(s)] for s in subjects]
            base.append(_lib.f64(mat[np.ix_(idx, idx)]))
        tmp_names = [str(k) for k, _ in items] if hasattr(setKernel, "items") else [str(i + 1) for i in range(len(base))]
        nbase = len(base)
        fres = _factored_result(base, tmp_names, gid, env, env_levels, L, model_code, intercept_random)
        if not factored:
            for ent in fres.values():
                ent["Kernel"] = materialize(ent)
        return fres


def _factored_result(base, tmp_names, gid, env, env_levels, L, model_code, intercept_random):
    """Entries in getK()'s order and with its names (R/getK.R:184-185, 202, 220-224, 233), factored form."""
    def entry(k0, typ, mode, k=0):
        return {"Type": typ, "factor": {"K0": k0, "gid": gid, "env": env, "mode": mode, "env_k": k, "L": L}}
    multi = len(base) > 1
    out = OrderedDict()
    for t, k0 in zip(tmp_names, base):
        out["G_" + t if multi else "G"] = entry(k0, "D", _lib.EXPAND_G)
    if model_code == 1:
        for t, k0 in zip(tmp_names, base):
            out["GE_" + t if multi else "GE"] = entry(k0, "BD", _lib.EXPAND_GE)
    if model_code == 2:
        for t, k0 in zip(tmp_names, base):
            for k, lev in enumerate(env_levels):
                out[f"{lev}_{t}" if multi else str(lev)] = entry(k0, "BD", _lib.EXPAND_ENV, k)
    if intercept_random:
        out["Gi"] = entry(None, "D", _lib.EXPAND_GI)
    return out
