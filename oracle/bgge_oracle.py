"""CPU restatement of italo-granato/BGGE v0.6.6 (R/getK.R, R/BGGE.R) in numpy/scipy.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It exists so the CUDA path can be
checked against the reference's algorithm.  Nothing under bgge_b200/ may import it.

PARITY UNPINNED: the reference is pure R and ships no numeric golden values
(tests/testthat/test.R:1-85 assert only classes/names; README.md:132-133 is an
unseeded run) and no R interpreter exists in this image, so this restatement cannot
be checked against outputs of the reference itself.  It is pinned instead by
hand-derived known-answer tests (tests/test_oracle_kat.py) and by the R semantics
documented in SURVEY.md Appendix A/B.  The arithmetic the reference delegates to
third-party code (R >= 3.1.1 base/stats + LAPACK/BLAS; DESCRIPTION:15-20, no lockfile)
is restated from the published algorithms: tcrossprod = dsyrk, dist = direct
differences (stats C_Cdist), quantile type 7, eigen = LAPACK dsyevr (values
descending), chol upper, var/sd with n-1, rnorm(n,m,s) = m + s*z,
1/rgamma(1,a,rate=r) = 2r/chisq(2a).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

try:  # same LAPACK driver as R's eigen(symmetric=TRUE)
    from scipy.linalg.lapack import dsyevr as _dsyevr
except Exception:  # pragma: no cover
    _dsyevr = None


# --------------------------------------------------------------------------------------
# draws
# --------------------------------------------------------------------------------------
class TapeDraws:
    """Injected draw tape.  normals/chisq are consumed strictly in the
    reference's call order: R/BGGE.R:254 (init), :279 (mu), :107 (XF), :91 (b), :97
    (sigb), :102 (sigsq), :379 (NA)."""

    def __init__(self, normals, chisq):
        self.z = np.asarray(normals, dtype=np.float64)
        self.c = np.asarray(chisq, dtype=np.float64)
        self.iz = 0
        self.ic = 0

    def normals(self, k):
        out = self.z[self.iz:self.iz + k]
        if out.shape[0] != k:
            raise IndexError("normal tape exhausted")
        self.iz += k
        return out

    def chisq(self, df):
        v = self.c[self.ic]
        self.ic += 1
        return float(v)


class RandomDraws:
    """Free-running draws from numpy (for statistical tests)."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def normals(self, k):
        return self.rng.standard_normal(k)

    def chisq(self, df):
        return float(self.rng.chisquare(df))


def tape_sizes(n, nr_per_kernel, ite, q=0, n_na=0):
    """Number of normals / chi-squares one BGGE() call consumes."""
    nk = len(nr_per_kernel)
    per_it = 1 + q + int(sum(nr_per_kernel)) + n_na
    return nk * n + ite * per_it, ite * (nk + 1)


# --------------------------------------------------------------------------------------
# getK pieces
# --------------------------------------------------------------------------------------
def design_codes(col):
    """factor(): levels = sorted unique values, codes 0-based (R/getK.R:101-104)."""
    col = np.asarray(col)
    levels, codes = np.unique(col, return_inverse=True)
    return [x.item() if hasattr(x, "item") else x for x in levels], codes.astype(np.int32)


def gb_kernel(X):
    """tcrossprod(X) / ncol(X)  (R/getK.R:136)."""
    X = np.asarray(X, dtype=np.float64)
    return (X @ X.T) / X.shape[1]


def gk_distance(X):
    """(as.matrix(dist(X)))^2  (R/getK.R:142): Euclidean distance by direct differences
    (stats C_Cdist: sum of dev*dev over columns, sqrt), then squared again; exact-zero
    diagonal, exactly symmetric."""
    X = np.asarray(X, dtype=np.float64)
    L = X.shape[0]
    D = np.zeros((L, L), dtype=np.float64)
    for i in range(L - 1):
        dev = X[i + 1:] - X[i]
        d = np.sqrt(np.einsum("ij,ij->i", dev, dev))
        d2 = d * d
        D[i + 1:, i] = d2
        D[i, i + 1:] = d2
    return D


def quantile7(x, prob):
    """stats::quantile.default(type = 7) for one probability (R/getK.R:146)."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    nn = x.shape[0]
    index = (nn - 1) * prob          # 0-based
    lo = int(math.floor(index))
    hi = int(math.ceil(index))
    qs = x[lo]
    h = index - lo
    if index > lo and x[hi] != qs:
        qs = (1.0 - h) * qs + h * x[hi]
    return float(qs)


def gk_kernel(D, bw, q):
    """exp(-bandwidth[i] * D / quantile(D, quantil))  (R/getK.R:146); R evaluates
    ((-bw) * D) / q."""
    return np.exp(((-bw) * D) / q)


def expand(K0, gid):
    """Zg %*% tcrossprod(K0, Zg)  (R/getK.R:138,148,173,232) == K0[gid, gid] exactly,
    because Zg is a one-hot incidence matrix (each product has a single non-zero term)."""
    return K0[np.ix_(gid, gid)]


def getK(Y, X=None, kernel=None, setKernel=None, bandwidth=1, model=None, quantil=0.5,
         intercept_random=False, rownames=None):
    """R/getK.R:95-237.  Y: sequence of rows / 2-D array / DataFrame whose first two
    columns are environment and genotype.  X: L x p array (rownames given separately,
    or DataFrame index).  Returns OrderedDict name -> {"Kernel": n x n, "Type": "D"|"BD"}."""
    if hasattr(Y, "iloc"):
        ecol, gcol = np.asarray(Y.iloc[:, 0]), np.asarray(Y.iloc[:, 1])
    else:
        Ya = np.asarray(Y, dtype=object)
        ecol, gcol = Ya[:, 0], Ya[:, 1]
    ecol = np.array([str(v) for v in ecol]) if ecol.dtype == object else ecol
    gcol = np.array([str(v) for v in gcol]) if gcol.dtype == object else gcol
    env_levels, env = design_codes(ecol)            # :104
    subjects, gid = design_codes(gcol)              # :103
    nEnv = len(env_levels)
    L = len(subjects)

    if model == "SM":                               # :112-116
        if nEnv > 1:
            raise ValueError("Single model choosen, but more than one environment is in the phenotype file")
    elif model in ("MM", "MDs", "MDe"):
        # model.matrix(~factor(Y[,1L]) - 1) on a one-level factor errors (:120; test.R:16-17)
        if nEnv < 2:
            raise ValueError("contrasts can be applied only to factors with 2 or more levels")
    else:
        raise ValueError("Model selected is not available ")   # :228

    if setKernel is None:
        if hasattr(X, "index") and rownames is None:
            rownames = list(X.index)
            X = X.to_numpy()
        if rownames is None:
            raise ValueError("Genotype names are missing")     # :126
        rn = {str(r): i for i, r in enumerate(rownames)}
        if not all(str(s) in rn for s in subjects):
            raise ValueError("Not all genotypes presents in the phenotypic file are in marker matrix")  # :129
        Xs = np.asarray(X, dtype=np.float64)[[rn[str(s)] for s in subjects], :]   # :131
        if kernel == "GB":
            base = [gb_kernel(Xs)]
        elif kernel == "GK":
            D = gk_distance(Xs)
            q = quantile7(D, quantil)
            base = [gk_kernel(D, float(bw), q) for bw in np.atleast_1d(bandwidth)]
        else:
            raise ValueError("kernel selected is not available. Please choose one method available or make available other kernel through argument K")
        tmp_names = [str(i + 1) for i in range(len(base))]      # :155
    else:
        base = []
        items = list(setKernel.items()) if hasattr(setKernel, "items") else [(None, k) for k in setKernel]
        for nm, km in items:
            if isinstance(km, tuple):
                mat, names = km
            elif hasattr(km, "index"):
                mat, names = km.to_numpy(), list(km.index)
            else:
                raise ValueError("Genotype names are missing in some kernel")   # :161
            pos = {str(r): i for i, r in enumerate(names)}
            if not all(str(s) in pos for s in subjects):
                raise ValueError("Not all genotypes presents in phenotypic file are in the kernel matrix.")  # :166
            idx = [pos[str(s)] for s in subjects]
            base.append(np.asarray(mat, dtype=np.float64)[np.ix_(idx, idx)])   # :169
        if hasattr(setKernel, "items"):
            tmp_names = [str(k) for k, _ in items]                              # :180
        else:
            tmp_names = [str(i + 1) for i in range(len(base))]                  # :178

    gnames = ["G_" + t for t in tmp_names] if len(base) > 1 else ["G"]          # :185
    G = OrderedDict((nm, {"Kernel": expand(k0, gid), "Type": "D"}) for nm, k0 in zip(gnames, base))

    out = OrderedDict(G)
    if model == "MDs":                                                          # :198-204
        E = (env[:, None] == env[None, :]).astype(np.float64)
        genames = ["GE_" + t for t in tmp_names] if len(base) > 1 else ["GE"]
        for nm, g in zip(genames, G.values()):
            out[nm] = {"Kernel": g["Kernel"] * E, "Type": "BD"}
    elif model == "MDe":                                                        # :206-226
        for j, g in enumerate(G.values()):
            for k in range(nEnv):
                m = ((env == k)[:, None] & (env == k)[None, :]).astype(np.float64)
                nm = f"{env_levels[k]}_{tmp_names[j]}" if len(base) > 1 else str(env_levels[k])
                out[nm] = {"Kernel": g["Kernel"] * m, "Type": "BD"}
    if intercept_random:                                                        # :231-234
        out["Gi"] = {"Kernel": expand(np.eye(L), gid), "Type": "D"}
    return out


# --------------------------------------------------------------------------------------
# BGGE pieces
# --------------------------------------------------------------------------------------
def eig_block(K, tol):
    """eig() R/BGGE.R:112-116: eigen(K) (dsyevr, lower triangle, values descending),
    keep values > tol."""
    K = np.ascontiguousarray(K, dtype=np.float64)
    if _dsyevr is not None:
        w, z, _m, _isuppz, info = _dsyevr(K, compute_v=1, range="A", lower=1)
        if info != 0:
            raise RuntimeError(f"dsyevr failed: {info}")
    else:  # pragma: no cover
        w, z = np.linalg.eigh(K)
    w = w[::-1]
    z = z[:, ::-1]
    keep = w > tol
    return w[keep].copy(), np.ascontiguousarray(z[:, keep])


def setDEC(K, ne, tol):
    """setDEC() R/BGGE.R:119-182.  Returns per kernel a list of blocks
    {"s","U","pos"} with pos = (start, stop) 0-based half-open, or None for D."""
    types = [k["Type"] for k in K.values()]
    if not all(t in ("BD", "D") for t in types):
        raise ValueError("Matrix should be of types BD or D")                   # :125
    ne = None if ne is None else np.atleast_1d(np.asarray(ne, dtype=np.int64))
    if ne is None:
        if "BD" in types:
            raise ValueError("For type BD, number of subjects in each sub matrices should be provided")  # :129
    elif len(ne) <= 1 and "BD" in types:
        raise ValueError("ne invalid. For type BD, number of subjects in each sub matrices should be provided")  # :132
    out = []
    for k in K.values():
        if k["Type"] == "D":                                                    # :144-154
            s, U = eig_block(k["Kernel"], tol)
            out.append([{"s": s, "U": U, "pos": None}])
        else:                                                                   # :156-179
            posf = np.cumsum(ne)
            posi = posf - ne
            blocks = []
            for j in range(len(ne)):
                sub = k["Kernel"][posi[j]:posf[j], posi[j]:posf[j]]
                s, U = eig_block(sub, tol)
                if s.shape[0] != 0:                                             # :162
                    blocks.append({"s": s, "U": U, "pos": (int(posi[j]), int(posf[j]))})
            if not blocks:
                raise IndexError("subscript out of bounds")                     # tmp[[1]] at :177
            out.append(blocks)
    return out


def _var(x):
    return float(np.var(x, ddof=1))


def _sd(x):
    return float(np.std(x, ddof=1)) if len(x) > 1 else float("nan")


def BGGE(y, K, XF=None, ne=None, ite=1000, burn=200, thin=3, verbose=False, tol=1e-10, R2=0.5,
         draws=None, eig=None, trace=None):
    """R/BGGE.R:84-444.  `draws` supplies the normal / chi-square stream (TapeDraws or
    RandomDraws).  `eig` optionally supplies setDEC()'s result (shared eigenbasis for
    injected-draw parity).  `trace`, if a dict, receives per-iteration
    samples: mu, varE, varU (ite x nk) and u (ite x nk x n)."""
    if draws is None:
        draws = RandomDraws(0)
    y = np.array(y, dtype=np.float64).ravel()                                   # :191
    yNA = np.isnan(y)
    whichNa = np.flatnonzero(yNA)
    nNa = whichNa.shape[0]
    mu = float(np.mean(y[~yNA]))                                                # :195
    n = y.shape[0]
    if nNa > 0:
        y[whichNa] = mu                                                         # :199
    if not isinstance(K, (dict, OrderedDict)):
        K = OrderedDict((f"K{i + 1}", k) for i, k in enumerate(K))              # :203-205
    names = list(K.keys())
    nk = len(K)

    if XF is not None:                                                          # :208-212
        XF = np.asarray(XF, dtype=np.float64).reshape(n, -1)
        XtX = XF.T @ XF
        Bet = np.linalg.solve(XtX, XF.T @ y)
        nBet = Bet.shape[0]
        tXX = np.linalg.inv(XtX)
        Lchol = np.linalg.cholesky(tXX)      # chol(sigsq*tXX)^T = sqrt(sigsq) * lower(tXX)

    types = [k["Type"] for k in K.values()]
    if not all(t in ("BD", "D") for t in types):
        raise ValueError("Matrix should be of types BD or D")                   # :220
    if ne is not None and np.size(ne) == 1 and "BD" in types:
        raise ValueError("Type BD should be used only for block diagonal matrices")  # :223
    Ei = setDEC(K, ne, tol) if eig is None else eig                             # :225

    nCum = sum(1 for i in range(1, ite + 1) if i % thin == 0)                   # :229
    chain_mu = np.zeros(nCum)
    chain_varE = np.zeros(nCum)
    chain_varU = np.zeros((nk, nCum))
    cpred = np.full((nk, nCum, n), np.nan)
    nu = 3.0                                                                    # :245
    vy = _var(y)
    Sce = (nu + 2) * (1 - R2) * vy                                              # :246
    # :249 -- a kernel entry may carry "mean_diag" instead of the n x n matrix (tests with supplied eigen blocks)
    Sc = np.array([(nu + 2) * R2 * vy / float(k["mean_diag"] if "mean_diag" in k else np.mean(np.diag(k["Kernel"])))
                   for k in K.values()])
    tau = 0.01                                                                  # :251
    u = [draws.normals(n) * (1.0 / (2 * n)) for _ in range(nk)]                 # :254
    sigsq = vy                                                                  # :256
    sigb = np.full(nk, 0.2)                                                     # :258
    temp = y - mu                                                               # :260
    if XF is not None:
        B_mcmc = np.zeros((nCum, nBet))
        temp = temp - XF @ Bet                                                  # :264
    usum = np.zeros(n)
    for j in range(nk):
        usum = usum + u[j]
    temp = temp - usum                                                          # :267
    nSel = 0

    for i in range(1, ite + 1):                                                 # :274
        temp = temp + mu                                                        # :278
        mu = float(np.mean(temp) + math.sqrt(sigsq / n) * draws.normals(1)[0])  # :279
        temp = temp - mu                                                        # :281
        if XF is not None:                                                      # :284-290
            temp = temp + XF @ Bet
            media = tXX @ (XF.T @ temp)
            Bet = media + math.sqrt(sigsq) * (Lchol @ draws.normals(nBet))
            temp = temp - XF @ Bet
        for j in range(nk):                                                     # :293
            blocks = Ei[j]
            temp = temp + u[j]                                                  # :297 / :314
            if types[j] == "D":
                d = blocks[0]["U"].T @ temp                                     # :298
                s = blocks[0]["s"]
            else:
                d = np.concatenate([b["U"].T @ temp[b["pos"][0]:b["pos"][1]] for b in blocks])  # :321
                s = np.concatenate([b["s"] for b in blocks])                    # :323
            lam = sigb[j]
            vari = s * lam / (1 + s * lam * tau)                                # :302 / :327
            media = tau * vari * d                                              # :303
            nr = s.shape[0]
            b = media + np.sqrt(vari) * draws.normals(nr)                       # :305 via :89-92
            if types[j] == "D":
                u[j] = blocks[0]["U"] @ b                                       # :306
            else:
                ut = np.zeros(n)                                                # :334
                off = 0
                for blk in blocks:
                    k_ = blk["s"].shape[0]
                    ut[blk["pos"][0]:blk["pos"][1]] = blk["U"] @ b[off:off + k_]  # :336
                    off += k_
                u[j] = ut
            temp = temp - u[j]                                                  # :307 / :339
            z = float(np.sum(b ** 2 * (1.0 / s)))                               # :96
            sigb[j] = (z + nu * Sc[j]) / draws.chisq(nr + nu)                   # :97, :360
        sigsq = (float(temp @ temp) + Sce) / draws.chisq(n + nu)                # :102, :366
        tau = 1.0 / sigsq                                                       # :367
        if nNa > 0:                                                             # :370-381
            uhat = np.zeros(n)
            for j in range(nk):
                uhat = uhat + u[j]
            aux = XF[yNA, :] @ Bet if XF is not None else 0.0
            y[whichNa] = aux + mu + uhat[whichNa] + math.sqrt(sigsq) * draws.normals(nNa)
            temp[whichNa] = y[whichNa] - uhat[whichNa] - aux - mu
        if i % thin == 0:                                                       # :384-395
            chain_varE[nSel] = sigsq
            chain_mu[nSel] = mu
            if XF is not None:
                B_mcmc[nSel] = Bet
            for j in range(nk):
                cpred[j, nSel] = u[j]
                chain_varU[j, nSel] = sigb[j]
            nSel += 1
        if trace is not None:
            trace.setdefault("mu", []).append(mu)
            trace.setdefault("varE", []).append(sigsq)
            trace.setdefault("varU", []).append(sigb.copy())
            if trace.get("keep_u", False):
                trace.setdefault("u", []).append(np.array(u))
            trace["e"] = temp.copy()

    its = np.arange(1, ite + 1)
    draw = its[its % thin == 0] > burn                                          # :408
    yHat = np.full(n, float(np.mean(chain_mu[draw])))                           # :410
    if XF is not None:
        yHat = yHat + XF @ B_mcmc[draw].mean(axis=0).reshape(-1)               # :414-415
    out = OrderedDict()
    Kout = OrderedDict()
    for j, nm in enumerate(names):                                              # :426-433
        kept = cpred[j][draw]
        Kout[nm] = {"u": kept.mean(axis=0),
                    "u.sd": kept.std(axis=0, ddof=1) if kept.shape[0] > 1 else np.full(n, np.nan),
                    "varu": float(np.mean(chain_varU[j][draw])),
                    "varu.sd": _sd(chain_varU[j][draw])}
        yHat = yHat + Kout[nm]["u"]                                             # :418-419
    out["yHat"] = yHat
    out["varE"] = float(np.mean(chain_varE[draw]))
    out["varE.sd"] = _sd(chain_varE[draw])
    out["K"] = Kout
    out["chain"] = {"mu": chain_mu, "varE": chain_varE,
                    "K": OrderedDict((nm, {"varU": chain_varU[j]}) for j, nm in enumerate(names))}
    out["ite"], out["burn"], out["thin"] = ite, burn, thin
    out["y"] = y
    if XF is not None:
        out["_Bet_chain"] = B_mcmc       # not part of the reference object (oracle extra)
    return out
