"""Shared synthetic-data builders for the tests (SURVEY.md 8d recipe, small sizes)."""
import numpy as np


def markers(L, p, seed):
    """Binomial(2, maf) genotypes, centred and scaled with the n-1 sd (R's scale())."""
    rng = np.random.default_rng(seed)
    maf = rng.uniform(0.05, 0.5, size=p)
    X = rng.binomial(2, maf, size=(L, p)).astype(np.float64)
    for j in range(p):
        while X[:, j].std() == 0:
            X[:, j] = rng.binomial(2, maf[j], size=L)
    X = (X - X.mean(axis=0)) / X.std(axis=0, ddof=1)
    return X


def design(L, E, seed=None, shuffle=False, drop=0):
    """Env-major balanced design like gl(n=E,k=L) x gl(n=L,k=1); optionally unbalanced/shuffled."""
    env = np.repeat(np.arange(E), L)
    gid = np.tile(np.arange(L), E)
    if drop:
        rng = np.random.default_rng(seed)
        keep = np.sort(rng.choice(L * E, size=L * E - drop, replace=False))
        env, gid = env[keep], gid[keep]
    if shuffle:
        rng = np.random.default_rng(seed)
        perm = rng.permutation(env.shape[0])
        env, gid = env[perm], gid[perm]
    Y = [(f"env{e + 1:02d}", f"{g + 1:05d}", 0.0) for e, g in zip(env, gid)]
    names = [f"{g + 1:05d}" for g in range(L)]
    return Y, names


def phenotype(K0, L, E, seed, h2=0.5):
    rng = np.random.default_rng(seed)
    w, V = np.linalg.eigh(K0 + 1e-8 * np.eye(L))
    root = V * np.sqrt(np.clip(w, 0, None))
    g = root @ rng.standard_normal(L)
    y = []
    for e in range(E):
        ge = np.sqrt(0.5) * (root @ rng.standard_normal(L))
        y.append(rng.normal(0, 0.3) + g + ge)
    y = np.concatenate(y)
    y = y + rng.standard_normal(L * E) * np.sqrt(np.var(y) * (1 - h2) / h2)
    return y


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


# ----------------------------------------------------------------------------------------------------------------------
# Host mirror of the device RNG (bgge_b200/csrc/rng.cuh): Philox4x32-10 keyed by (seed, chain), counter =
# (index, iteration, stream); normals by Box-Muller on 53-bit uniforms, chi-square as 2 Gamma(df/2) by
# Marsaglia-Tsang.  PhiloxDraws hands the ORACLE exactly the draws a device chain consumes, in the reference's call
# order, so that device-RNG runs -- single chains and the batched sampler -- can be compared with the
# oracle draw for draw instead of only statistically.
# ----------------------------------------------------------------------------------------------------------------------
STREAM_MU, STREAM_XF, STREAM_NA, STREAM_CHI_E, STREAM_INIT_U, STREAM_CHI_K = (0xFFFF0001, 0xFFFF0002, 0xFFFF0003,
                                                                              0xFFFF0004, 0xFFFF0005, 0xFFFE0000)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10: counters are uint32 arrays (broadcastable), the key two scalars."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in np.broadcast_arrays(c0, c1, c2, c3))
    mask = np.uint64(0xFFFFFFFF)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & mask
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(hi, lo):
    a = ((hi << np.uint64(32)) | lo) >> np.uint64(11)
    return (a.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


class PhiloxDraws:
    """draws= object for oracle.BGGE reproducing the device stream of (seed, chain) for a model with n rows, kernels
    with nrs[j] kept eigenvalues, q fixed effects and n_na missing phenotypes."""

    def __init__(self, seed, chain, n, nrs, q=0, n_na=0):
        self.n_na = n_na
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = ((seed >> 32) & 0xFFFFFFFF) ^ (chain & 0xFFFFFFFF)
        self.n, self.nrs, self.q, self.nk = n, list(nrs), q, len(nrs)
        self.it = 0                  # iteration whose draws are being consumed (0 = initialisation)
        self.init_left = self.nk     # u_j initialisations still to come
        self.stage = "init"
        self.j = 0

    def _words(self, it, stream, idx):
        idx = np.asarray(idx, dtype=np.uint64)
        return philox4x32_10(idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.uint64(it & 0xFFFFFFFF), np.uint64(stream),
                             self.k0, self.k1)

    def _normal(self, it, stream, idx):
        w = self._words(it, stream, idx)
        u1, u2 = _u53(w[0], w[1]), _u53(w[2], w[3])
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(np.pi * (2.0 * u2))

    def normals(self, k):
        if self.stage == "init":                                   # R/BGGE.R:254: kernel-major, index j n + i
            jdx = self.nk - self.init_left
            out = self._normal(0, STREAM_INIT_U, jdx * self.n + np.arange(k))
            self.init_left -= 1
            if self.init_left == 0:
                self.stage, self.it = "mu", 1
            return out
        if self.stage == "mu":                                     # :279
            assert k == 1
            self.stage = "xf" if self.q else "kernel"
            self.j = 0
            return self._normal(self.it, STREAM_MU, np.zeros(1))
        if self.stage == "xf":                                     # :107
            assert k == self.q
            self.stage = "kernel"
            return self._normal(self.it, STREAM_XF, np.arange(k))
        if self.stage == "kernel":                                 # :91, stream = kernel index, index = eigen-column
            assert k == self.nrs[self.j], (k, self.nrs, self.j)
            self.stage = "chik"
            return self._normal(self.it, self.j, np.arange(k))
        if self.stage == "na":                                     # :379, index = rank of the missing entry
            self.stage, self.it = "mu", self.it + 1
            return self._normal(self.it - 1, STREAM_NA, np.arange(k))
        raise AssertionError(f"unexpected normals({k}) in stage {self.stage}")

    def _chisq(self, stream, df):
        a = 0.5 * df
        d = a - 1.0 / 3.0
        c = 1.0 / np.sqrt(9.0 * d)
        attempt = 0
        while True:
            x = float(self._normal(self.it, stream, np.array([2 * attempt]))[0])
            v = 1.0 + c * x
            if v > 0.0:
                v = v * v * v
                w = self._words(self.it, stream, np.array([2 * attempt + 1]))
                uu = float(_u53(w[0], w[1])[0])
                if np.log(uu) < 0.5 * x * x + d - d * v + d * np.log(v):
                    return 2.0 * d * v
            attempt += 1

    def chisq(self, df):
        if self.stage == "chik":                                   # :97
            v = self._chisq(STREAM_CHI_K + self.j, df)
            self.j += 1
            self.stage = "kernel" if self.j < self.nk else "chie"
            return v
        if self.stage == "chie":                                   # :102; then the imputation draws (:379) or the next iteration
            v = self._chisq(STREAM_CHI_E, df)
            if self.n_na > 0:
                self.stage = "na"
            else:
                self.stage, self.it = "mu", self.it + 1
            return v
        raise AssertionError(f"unexpected chisq in stage {self.stage}")
