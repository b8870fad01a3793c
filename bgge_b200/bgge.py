"""BGGE(): Gibbs sampler for genomic G x E models -- host mirror of R/BGGE.R:84-444.

The eigendecompositions (setDEC, R/BGGE.R:119-182), every sweep (R/BGGE.R:274-403) and the
posterior summaries (R/BGGE.R:408-433) run in libbgge_b200 on the GPU; this module only
validates arguments, drives the handle and assembles the returned object with the
reference's field names.
"""
from __future__ import annotations

import ctypes as C
import time
from collections import OrderedDict

import numpy as np

from . import _lib
from .getk import factor_matches
from .multigpu import run_sharded
from .rrng import r_tapes


class BGGEFit(OrderedDict):
    """The "BGGE" S3 object: yHat, varE, varE.sd, K, chain, ite, burn, thin, y."""

    def __str__(self):   # print.BGGE, R/utils.R:9-18
        head = "  ".join(f"{v:.3g}" for v in np.asarray(self["yHat"]).ravel()[:10])
        return ("Model Fitted with: \n " + f"{self['ite']}  Iterations, burning the first  {self['burn']}  and thining every  "
                f"{self['thin']} \n\n Some predicted Values: \n{head}\n\n Use str() function to found more datailed information.\n")


def _factor_mean_diag(f):
    """mean(diag(K)) of an expanded kernel from its factored description (R/BGGE.R:249)."""
    gid, env = f["gid"], f["env"]
    d = np.ones(f["L"]) if f["K0"] is None else np.diag(f["K0"])
    v = d[gid].astype(np.float64)
    if f["mode"] == _lib.EXPAND_ENV:
        v = v * (env == f["env_k"])
    return float(v.mean())


def _plan_info(h, nk):
    """Execution plan of an initialised handle (which kernels run, how they are tiled)."""
    en, ct, sm = C.c_int(), C.c_int(), C.c_int64()
    _lib.call("bgge_gibbs_resident_info", h, C.byref(en), C.byref(ct), C.byref(sm))
    un, ku = C.c_int(), C.c_int()
    _lib.call("bgge_gibbs_merged_info", h, C.byref(un), C.byref(ku))
    ks = []
    for j in range(nk):
        f_, cs_, ncl_, nl_ = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.call("bgge_gibbs_kernel_info", h, j, C.byref(f_), C.byref(cs_), C.byref(ncl_), C.byref(nl_))
        bb_, nf_ = C.c_int64(), C.c_int()
        _lib.call("bgge_gibbs_kernel_bytes", h, j, C.byref(bb_), C.byref(nf_))
        ks.append({"path": f_.value, "cluster_size": cs_.value, "clusters": ncl_.value, "launches": nl_.value,
                   "basis_bytes": bb_.value, "factored_blocks": nf_.value})
    return {"resident": bool(en.value), "resident_ctas": ct.value, "merged_units": un.value,
            "kernels_in_units": ku.value, "kernels": ks}


def _check_factor_rows(k, n, name):
    """The factored description must describe exactly the n rows of y (the reference fails with non-conformable
    arguments when K and y disagree)."""
    f = k["factor"]
    if f["gid"].shape[0] != n or f["env"].shape[0] != n:
        raise ValueError(f"kernel {name} describes {f['gid'].shape[0]} rows, expected {n} (non-conformable arguments)")


def _add_factored(h, j, k, ne_arr, tol, canonical_sign):
    f = k["factor"]
    tcode = _lib.TYPE_D if k["Type"] == "D" else _lib.TYPE_BD
    K0 = None if f["K0"] is None else _lib.f64(f["K0"])
    _lib.call("bgge_gibbs_add_expanded_kernel", h, j, tcode, _lib.ptr(K0), f["L"], f["L"], 0, _lib.ptr(f["gid"]),
              _lib.ptr(f["env"]), int(f["gid"].shape[0]), f["mode"], f["env_k"], None if ne_arr is None else _lib.ptr(ne_arr),
              0 if ne_arr is None else ne_arr.shape[0], float(tol), int(canonical_sign))


def setDEC(K, ne=None, tol=1e-10, canonical_sign=True):
    """setDEC() of R/BGGE.R:119-182 on the device: per kernel a list of blocks
    {"s", "U", "pos"} (pos = (start, stop) 0-based half-open, None for type D).  Dense entries go through
    bgge_eig per block; factored entries (getK(..., factored=True)) through the L x L route."""
    types = [k["Type"] for k in K.values()]
    if not all(t in ("BD", "D") for t in types):
        raise ValueError("Matrix should be of types BD or D")
    if ne is None:
        if "BD" in types:
            raise ValueError("For type BD, number of subjects in each sub matrices should be provided")
    elif np.size(ne) <= 1 and "BD" in types:
        raise ValueError("ne invalid. For type BD, number of subjects in each sub matrices should be provided")
    ne_arr = None if ne is None else np.ascontiguousarray(np.atleast_1d(ne), dtype=np.int64)
    out = []
    for k in K.values():
        if "factor" in k and factor_matches(k):
            n = k["factor"]["gid"].shape[0]
            h = C.c_void_p()
            _lib.call("bgge_gibbs_create", n, 1, 0, None, C.byref(h))
            try:
                _add_factored(h, 0, k, ne_arr, tol, canonical_sign)
                nb = C.c_int()
                _lib.call("bgge_gibbs_block_count", h, 0, C.byref(nb))
                blocks = []
                for b in range(nb.value):
                    pos0, rows, nr = C.c_int64(), C.c_int64(), C.c_int64()
                    _lib.call("bgge_gibbs_fetch_block", h, 0, b, C.byref(pos0), C.byref(rows), C.byref(nr), None, None, 0)
                    s = np.empty(nr.value)
                    U = np.empty((rows.value, nr.value), order="F")
                    _lib.call("bgge_gibbs_fetch_block", h, 0, b, None, None, None, _lib.ptr(s), _lib.ptr(U), rows.value)
                    blocks.append({"s": s, "U": U, "pos": None if k["Type"] == "D" else (pos0.value, pos0.value + rows.value)})
            finally:
                _lib.load().bgge_gibbs_destroy(h)
            out.append(blocks)
            continue
        Km = _lib.f64(k["Kernel"])
        n = Km.shape[0]
        spans = [(0, n)] if k["Type"] == "D" else list(zip(np.cumsum(ne) - np.asarray(ne), np.cumsum(ne)))
        blocks = []
        for a, b in spans:
            a, b = int(a), int(b)
            m = b - a
            sub = _lib.f64(Km[a:b, a:b])
            s = np.empty(m)
            U = np.empty((m, m), order="F")
            nr = C.c_int64(0)
            _lib.call("bgge_eig", _lib.ptr(sub), m, m, float(tol), int(canonical_sign), _lib.ptr(s), _lib.ptr(U),
                      C.byref(nr))
            if nr.value > 0:
                blocks.append({"s": s[:nr.value].copy(), "U": np.asfortranarray(U[:, :nr.value]),
                               "pos": None if k["Type"] == "D" else (a, b)})
        if not blocks:
            raise IndexError("subscript out of bounds")   # tmp[[1]] at R/BGGE.R:177
        out.append(blocks)
    return out


def BGGE(y, K, XF=None, ne=None, ite=1000, burn=200, thin=3, verbose=False, tol=1e-10, R2=0.5, *, seed=0, chain=0,
         draws=None, eig=None, device=None, mode=None, rng="philox", shard=None, info=None):
    """Fit the G x E model by Gibbs sampling (signature of R/BGGE.R:84 plus keyword-only extras).

    draws : optional (normals, chisq) tapes -> injected-draw mode.
    eig   : optional precomputed setDEC() result (shared eigenbasis); kernel entries may then carry
            "mean_diag" instead of the n x n "Kernel".
    mode  : None/1 = resident kernel (small models) or fused single-pass sweep kernels where the shapes fit,
            2 = never resident, 0 = two-pass kernels only.
    rng   : "philox" (default) = counter-based device draws keyed by (seed, chain); "R" = the stream of R's
            `set.seed(seed)` (Mersenne-Twister, inversion normals, Ahrens-Dieter gammas) consumed in the
            reference's call order and injected as tapes (bgge_b200.rrng).
    info  : optional dict, filled after initialisation with the execution plan: "resident" (bool), "merged_units",
            and per kernel "path" (0 two-pass, 1 sweep_fused_kernel, 2 sweep_mrhs_kernel), "cluster_size", "clusters",
            "launches", "basis_bytes", "factored_blocks" (bgge_gibbs_kernel_info / _kernel_bytes).
    shard : None, or (rank, world, comm_or_reduce): this process sweeps its slice of the eigen-columns of a single
            chain.  Third element a bgge_b200.multigpu.Comm: the per-step all-reduce runs inside the library on the
            handle's own stream (fast path).  A callable reduce(ptr, count): it must add the per-step partial vectors
            of all processes in place and return when done (multigpu.nccl_allreduce(); the handle's stream is drained
            before each call).  Every process calls BGGE() with the same arguments; all return the same fit.  No XF.
    """
    if rng not in ("philox", "R"):
        raise ValueError("rng must be 'philox' or 'R'")
    if rng == "R" and draws is not None:
        raise ValueError("rng='R' builds its own tapes: do not pass draws")
    y = np.array(y, dtype=np.float64).ravel()
    n = y.shape[0]
    if not hasattr(K, "items"):
        K = OrderedDict((f"K{i + 1}", k) for i, k in enumerate(K))          # R/BGGE.R:203-205
    names = list(K.keys())
    nk = len(names)
    types = [k["Type"] for k in K.values()]
    if not all(t in ("BD", "D") for t in types):
        raise ValueError("Matrix should be of types BD or D")                # R/BGGE.R:220
    if ne is None and "BD" in types:
        raise ValueError("For type BD, number of subjects in each sub matrices should be provided")
    if ne is not None and np.size(ne) == 1 and "BD" in types:
        raise ValueError("Type BD should be used only for block diagonal matrices")   # R/BGGE.R:223
    ne_arr = None if ne is None else np.ascontiguousarray(np.atleast_1d(ne), dtype=np.int64)
    q = 0
    if XF is not None:
        XF = _lib.f64(np.asarray(XF, dtype=np.float64).reshape(n, -1))
        q = XF.shape[1]
    if device is not None:
        _lib.call("bgge_set_device", int(device))

    h = C.c_void_p()
    _lib.call("bgge_gibbs_create", n, nk, q, None, C.byref(h))
    try:
        _lib.call("bgge_gibbs_set_y", h, _lib.ptr(y), None)
        if q:
            _lib.call("bgge_gibbs_set_xf", h, _lib.ptr(XF))
        if mode is not None:
            _lib.call("bgge_gibbs_set_mode", h, int(mode))
        if shard is not None and int(shard[1]) > 1:
            if q:
                raise ValueError("a column-sharded chain does not support XF")
            if hasattr(shard[2], "handle"):
                _lib.call("bgge_gibbs_set_comm", h, shard[2].handle)
            else:
                _lib.call("bgge_gibbs_set_shard", h, int(shard[0]), int(shard[1]))
        for j, k in enumerate(K.values()):
            tcode = _lib.TYPE_D if k["Type"] == "D" else _lib.TYPE_BD
            if eig is None and "factor" in k and factor_matches(k):
                # getK() output: decompose through the L x L problem instead of the n x n matrix
                _check_factor_rows(k, n, names[j])
                _add_factored(h, j, k, ne_arr, tol, True)
                continue
            if eig is not None and ("mean_diag" in k or "factor" in k):
                md = k["mean_diag"] if "mean_diag" in k else _factor_mean_diag(k["factor"])
                _lib.call("bgge_gibbs_set_kernel", h, j, tcode, float(md))
                Km = None
            else:
                Km = _lib.f64(k["Kernel"])
                if Km.shape != (n, n):
                    raise ValueError(f"kernel {names[j]} is {Km.shape}, expected {(n, n)}")
            if eig is None:
                _lib.call("bgge_gibbs_add_kernel_matrix", h, j, tcode, _lib.ptr(Km), n, 0,
                          None if ne_arr is None else _lib.ptr(ne_arr), 0 if ne_arr is None else ne_arr.shape[0],
                          float(tol), 1)
            else:
                if Km is not None:
                    _lib.call("bgge_gibbs_set_kernel", h, j, tcode, float(np.mean(np.diag(Km))))
                for blk in eig[j]:
                    U = _lib.f64(blk["U"])
                    s = np.ascontiguousarray(blk["s"], dtype=np.float64)
                    pos0 = 0 if blk["pos"] is None else int(blk["pos"][0])
                    _lib.call("bgge_gibbs_add_block", h, j, pos0, U.shape[0], U.shape[1], _lib.ptr(s), _lib.ptr(U),
                              U.shape[0], 0)
        mode = _lib.RNG_PHILOX
        if rng == "R":
            nrs = []
            for j in range(nk):
                nb_, tot = C.c_int(), 0
                _lib.call("bgge_gibbs_block_count", h, j, C.byref(nb_))
                for b in range(nb_.value):
                    nr_ = C.c_int64()
                    _lib.call("bgge_gibbs_fetch_block", h, j, b, None, None, C.byref(nr_), None, None, 0)
                    tot += nr_.value
                nrs.append(tot)
            draws = r_tapes(int(seed), n, nrs, q, int(np.isnan(y).sum()), int(ite))
        if draws is not None:
            z = np.ascontiguousarray(draws[0], dtype=np.float64)
            c = np.ascontiguousarray(draws[1], dtype=np.float64)
            _lib.call("bgge_gibbs_set_tape", h, _lib.ptr(z), z.shape[0], _lib.ptr(c), c.shape[0])
            mode = _lib.RNG_TAPE
        _lib.call("bgge_gibbs_init", h, int(ite), int(burn), int(thin), float(R2), mode, int(seed), int(chain))
        if info is not None:
            info.update(_plan_info(h, nk))
        step = int(verbose) if int(verbose) != 0 else int(ite)
        done = 0
        while done < ite:                                                    # R/BGGE.R:274, 398-401
            t0 = time.perf_counter()
            k = min(step, ite - done)
            if shard is not None and int(shard[1]) > 1:
                if hasattr(shard[2], "handle"):
                    _lib.call("bgge_gibbs_run_sharded", h, k)
                else:
                    run_sharded(h, nk, k, shard[2])
            else:
                _lib.call("bgge_gibbs_run", h, k)
            _lib.call("bgge_gibbs_sync", h)
            done += k
            if int(verbose) != 0 and done % int(verbose) == 0:
                print("Iter: ", done, "time: ", round((time.perf_counter() - t0) / k, 3))

        ncum = ite // thin
        yHat, u, usd, yout = np.empty(n), np.empty((nk, n)), np.empty((nk, n)), np.empty(n)
        varu, varusd = np.empty(nk), np.empty(nk)
        varE, varEsd = C.c_double(), C.c_double()
        _lib.call("bgge_gibbs_summary", h, _lib.ptr(yHat), C.byref(varE), C.byref(varEsd), _lib.ptr(u), _lib.ptr(usd),
                  _lib.ptr(varu), _lib.ptr(varusd), _lib.ptr(yout))
        cmu, cve, cvu = np.empty(ncum), np.empty(ncum), np.empty((nk, ncum))
        cbet = np.empty((ncum, max(q, 1)))
        _lib.call("bgge_gibbs_chain", h, _lib.ptr(cmu), _lib.ptr(cve), _lib.ptr(cvu), _lib.ptr(cbet))
    finally:
        _lib.load().bgge_gibbs_destroy(h)

    out = BGGEFit()
    out["yHat"] = yHat.reshape(n, 1) if q else yHat                           # R/BGGE.R:411-419
    out["varE"] = varE.value
    out["varE.sd"] = varEsd.value
    out["K"] = OrderedDict((nm, {"u": u[j].copy(), "u.sd": usd[j].copy(), "varu": float(varu[j]),
                                 "varu.sd": float(varusd[j])}) for j, nm in enumerate(names))
    out["chain"] = {"mu": cmu, "varE": cve,
                    "K": OrderedDict((nm, {"varU": cvu[j].copy()}) for j, nm in enumerate(names))}
    out["ite"], out["burn"], out["thin"] = ite, burn, thin
    out["y"] = yout
    return out


def BGGE_batch(Y, K, XF=None, ne=None, ite=1000, burn=200, thin=3, tol=1e-10, R2=0.5, *, seed=0, chain0=0, eig=None,
               device=None):
    """B independent BGGE() chains / cross-validation folds sharing one eigen decomposition (config C4).

    Y is a (B, n) array: row b is the phenotype vector of chain b (NaN = masked fold / missing).  Chain b
    uses the RNG stream (seed, chain0 + b), i.e. it reproduces BGGE(Y[b], K, ..., seed=seed, chain=chain0 + b).
    The eigen blocks are computed once (setDEC on the device) and the per-kernel U'e / U b passes of all
    chains run as skinny FP64 tensor-core GEMMs.  XF (n x q, q <= 16) is one fixed-effects design shared by all
    chains, as in BGGE(XF=) (R/BGGE.R:208-212, 284-290).  Returns a list of BGGEFit objects.
    """
    Y = np.ascontiguousarray(np.atleast_2d(np.asarray(Y, dtype=np.float64)))
    B, n = Y.shape
    if not hasattr(K, "items"):
        K = OrderedDict((f"K{i + 1}", k) for i, k in enumerate(K))
    names = list(K.keys())
    nk = len(names)
    types = [k["Type"] for k in K.values()]
    if not all(t in ("BD", "D") for t in types):
        raise ValueError("Matrix should be of types BD or D")
    if ne is None and "BD" in types:
        raise ValueError("For type BD, number of subjects in each sub matrices should be provided")
    if ne is not None and np.size(ne) == 1 and "BD" in types:
        raise ValueError("Type BD should be used only for block diagonal matrices")
    ne_arr = None if ne is None else np.ascontiguousarray(np.atleast_1d(ne), dtype=np.int64)
    if device is not None:
        _lib.call("bgge_set_device", int(device))
    h0, hb = C.c_void_p(), C.c_void_p()
    _lib.call("bgge_gibbs_create", n, nk, 0, None, C.byref(h0))      # owns the decomposition
    try:
        _lib.call("bgge_batch_create", n, nk, B, None, C.byref(hb))
        _lib.call("bgge_batch_set_y", hb, _lib.ptr(Y), None)
        if XF is not None:
            XFm = _lib.f64(np.asarray(XF, dtype=np.float64).reshape(n, -1))
            _lib.call("bgge_batch_set_xf", hb, _lib.ptr(XFm), XFm.shape[1])
        for j, k in enumerate(K.values()):
            tcode = _lib.TYPE_D if k["Type"] == "D" else _lib.TYPE_BD
            if eig is not None:
                md = k["mean_diag"] if "mean_diag" in k else (_factor_mean_diag(k["factor"]) if "factor" in k
                                                              else float(np.mean(np.diag(k["Kernel"]))))
                for blk in eig[j]:
                    U = _lib.f64(blk["U"])
                    s = np.ascontiguousarray(blk["s"], dtype=np.float64)
                    pos0 = 0 if blk["pos"] is None else int(blk["pos"][0])
                    _lib.call("bgge_gibbs_add_block", h0, j, pos0, U.shape[0], U.shape[1], _lib.ptr(s), _lib.ptr(U),
                              U.shape[0], 0)
            elif "factor" in k and factor_matches(k):
                _check_factor_rows(k, n, names[j])
                md = _factor_mean_diag(k["factor"])
                _add_factored(h0, j, k, ne_arr, tol, True)
            else:
                Km = _lib.f64(k["Kernel"])
                md = float(np.mean(np.diag(Km)))
                _lib.call("bgge_gibbs_add_kernel_matrix", h0, j, tcode, _lib.ptr(Km), n, 0,
                          None if ne_arr is None else _lib.ptr(ne_arr), 0 if ne_arr is None else ne_arr.shape[0],
                          float(tol), 1)
            _lib.call("bgge_batch_set_kernel", hb, j, tcode, md)
            nb = C.c_int()
            _lib.call("bgge_gibbs_block_count", h0, j, C.byref(nb))
            for b in range(nb.value):
                pos0, rows, nr, ldu = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
                _lib.call("bgge_gibbs_fetch_block", h0, j, b, C.byref(pos0), C.byref(rows), C.byref(nr), None, None, 0)
                s = np.empty(nr.value)
                _lib.call("bgge_gibbs_fetch_block", h0, j, b, None, None, None, _lib.ptr(s), None, 0)
                dptr = C.c_void_p()
                _lib.call("bgge_gibbs_block_device", h0, j, b, C.byref(dptr), C.byref(ldu))
                _lib.call("bgge_batch_add_block", hb, j, pos0.value, rows.value, nr.value, _lib.ptr(s), dptr, ldu.value, 1)
        _lib.call("bgge_batch_init", hb, int(ite), int(burn), int(thin), float(R2), int(seed), int(chain0))
        _lib.call("bgge_batch_run", hb, int(ite))
        _lib.call("bgge_batch_sync", hb)
        ncum = ite // thin
        yHat, yout = np.empty((B, n)), np.empty((B, n))
        u, usd = np.empty((nk, B, n)), np.empty((nk, B, n))
        varE, varEsd, varu, varusd = np.empty(B), np.empty(B), np.empty((nk, B)), np.empty((nk, B))
        _lib.call("bgge_batch_summary", hb, _lib.ptr(yHat), _lib.ptr(varE), _lib.ptr(varEsd), _lib.ptr(u), _lib.ptr(usd),
                  _lib.ptr(varu), _lib.ptr(varusd), _lib.ptr(yout))
        cmu, cve, cvu = np.empty((B, ncum)), np.empty((B, ncum)), np.empty((nk, B, ncum))
        _lib.call("bgge_batch_chain", hb, _lib.ptr(cmu), _lib.ptr(cve), _lib.ptr(cvu))
        cbet = None
        if XF is not None:
            cbet = np.empty((B, ncum, XFm.shape[1]))
            _lib.call("bgge_batch_chain_bet", hb, _lib.ptr(cbet))
    finally:
        if hb:
            _lib.load().bgge_batch_destroy(hb)
        _lib.load().bgge_gibbs_destroy(h0)
    fits = []
    for b in range(B):
        out = BGGEFit()
        out["yHat"] = yHat[b].copy()
        out["varE"] = float(varE[b])
        out["varE.sd"] = float(varEsd[b])
        out["K"] = OrderedDict((nm, {"u": u[j, b].copy(), "u.sd": usd[j, b].copy(), "varu": float(varu[j, b]),
                                     "varu.sd": float(varusd[j, b])}) for j, nm in enumerate(names))
        out["chain"] = {"mu": cmu[b].copy(), "varE": cve[b].copy(),
                        "K": OrderedDict((nm, {"varU": cvu[j, b].copy()}) for j, nm in enumerate(names))}
        if cbet is not None:
            out["chain"]["Bet"] = cbet[b].copy()      # extra (the reference keeps B.mcmc internally)
        out["ite"], out["burn"], out["thin"] = ite, burn, thin
        out["y"] = yout[b].copy()
        fits.append(out)
    return fits
