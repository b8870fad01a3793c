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
