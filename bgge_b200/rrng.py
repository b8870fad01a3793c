"""R-compatible random number stream and BGGE draw tapes.

The reference draws from R's global generator (`rnorm` at R/BGGE.R:91,107,254,279,379 and `rgamma` at
R/BGGE.R:97,102).  With R's defaults that is: Mersenne-Twister (MT19937) seeded through `set.seed`'s linear
congruential scrambling, `unif_rand` = 32-bit output * 2^-32 clamped into (0, 1), `norm_rand` by INVERSION
(two uniforms -> 53-bit u -> Wichura's AS 241 `qnorm`), `rgamma` by Ahrens & Dieter's GD algorithm for shape >= 1
(the only case BGGE reaches: shape = (nr + 3) / 2) with `exp_rand` by Ahrens & Dieter 1972 in the rejection branch.

This module restates those published algorithms (R itself is not part of /root/reference and is not installed
here) so that `BGGE(..., rng="R", seed=k)` consumes the same stream as `set.seed(k); BGGE(...)` in R, in the
reference's call order, and hands it to the device sampler as injected tapes.

Pinned by the known first outputs of `set.seed(1|42|123)` for `runif`, `rnorm` and `rexp` (tests/test_rrng.py).
`rgamma` has no offline pin: it is checked structurally (its acceptance branches against the restated
normal / uniform / exponential streams) and statistically.  A chain reproduces R's only when the eigenvectors
also carry LAPACK's signs (the draw b = m + sqrt(v) z is not invariant under a sign flip of an eigenvector),
which the device eigensolver does not promise; the marginal law is the same either way.
"""
from __future__ import annotations

import math

import numpy as np

_I2_32M1 = 2.328306437080797e-10          # R: fixup() bound
_MT_SCALE = 2.3283064365386963e-10        # 2^-32
_BIG = 134217728.0                        # 2^27, norm_rand INVERSION

# ---- AS 241 (Wichura 1988), PPND16
_A = (3.3871328727963666080, 1.3314166789178437745e+2, 1.9715909503065514427e+3, 1.3731693765509461125e+4,
      4.5921953931549871457e+4, 6.7265770927008700853e+4, 3.3430575583588128105e+4, 2.5090809287301226727e+3)
_B = (1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3, 2.1213794301586595867e+4,
      3.9307895800092710610e+4, 2.8729085735721942674e+4, 5.2264952788528545610e+3)
_C = (1.42343711074968357734, 4.63033784615654529590, 5.76949722146069140550, 3.64784832476320460504,
      1.27045825245236838258, 2.41780725177450611770e-1, 2.27238449892691845833e-2, 7.74545014278341407640e-4)
_D = (1.0, 2.05319162663775882187, 1.67638483018380384940, 6.89767334985100004550e-1, 1.48103976427480074590e-1,
      1.51986665636164571966e-2, 5.47593808499534494600e-4, 1.05075007164441684324e-9)
_E = (6.65790464350110377720, 5.46378491116411436990, 1.78482653991729133580, 2.96560571828504891230e-1,
      2.65321895265761230930e-2, 1.24266094738807843860e-3, 2.71155556874348757815e-5, 2.01033439929228813265e-7)
_F = (1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
      7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7, 2.04426310338993978564e-15)


def _horner(c, r):
    v = np.full_like(r, c[7])
    for k in range(6, -1, -1):
        v = v * r + c[k]
    return v


def qnorm(p):
    """Standard normal quantile, AS 241 as in R's qnorm5 (vectorised)."""
    p = np.atleast_1d(np.asarray(p, dtype=np.float64))
    q = p - 0.5
    out = np.empty_like(p)
    mid = np.abs(q) <= 0.425
    if mid.any():
        r = 0.180625 - q[mid] * q[mid]
        out[mid] = q[mid] * _horner(_A, r) / _horner(_B, r)
    if (~mid).any():
        qq = q[~mid]
        r = np.where(qq < 0, p[~mid], 1.0 - p[~mid])
        r = np.sqrt(-np.log(r))
        val = np.empty_like(r)
        lo = r <= 5.0
        if lo.any():
            rr = r[lo] - 1.6
            val[lo] = _horner(_C, rr) / _horner(_D, rr)
        if (~lo).any():
            rr = r[~lo] - 5.0
            val[~lo] = _horner(_E, rr) / _horner(_F, rr)
        out[~mid] = np.where(qq < 0, -val, val)
    return out


# exp_rand table: q[k] = sum_{i=1}^{k+1} ln(2)^i / i!
_Q = []
_t, _s = 1.0, 0.0
for _i in range(1, 17):
    _t *= math.log(2.0) / _i
    _s += _t
    _Q.append(_s)
_Q[-1] = 1.0

# rgamma GD constants (Ahrens & Dieter 1982)
_SQRT32 = 5.656854
_Q1, _Q2, _Q3, _Q4, _Q5, _Q6, _Q7 = 0.04166669, 0.02083148, 0.00801191, 0.00144121, -7.388e-5, 2.4511e-4, 2.424e-4
_A1, _A2, _A3, _A4, _A5, _A6, _A7 = 0.3333333, -0.250003, 0.2000062, -0.1662921, 0.1423657, -0.1367177, 0.1233795


class RStream:
    """R's default RNG after `set.seed(seed)`: unif_rand / norm_rand / exp_rand / rgamma in stream order."""

    def __init__(self, seed: int):
        s = int(seed) & 0xFFFFFFFF
        for _ in range(50):                                   # initial scrambling (RNG.c, RNG_Init)
            s = (69069 * s + 1) & 0xFFFFFFFF
        st = np.empty(625, dtype=np.uint32)
        for j in range(625):
            s = (69069 * s + 1) & 0xFFFFFFFF
            st[j] = s
        # i_seed[0] is the position, forced to N = 624 by FixupSeeds; i_seed[1..624] is the state vector
        self._bg = np.random.MT19937()
        self._bg.state = {"bit_generator": "MT19937", "state": {"key": st[1:].copy(), "pos": 624}}
        self.uniforms_used = 0

    # ---- uniforms
    def unif(self, k: int) -> np.ndarray:
        raw = self._bg.random_raw(int(k)).astype(np.float64)
        self.uniforms_used += int(k)
        u = raw * _MT_SCALE
        u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
        return u

    def unif1(self) -> float:
        return float(self.unif(1)[0])

    # ---- normals (INVERSION)
    def norm(self, k: int) -> np.ndarray:
        if k <= 0:
            return np.empty(0)
        u = self.unif(2 * int(k))
        hi = np.floor(_BIG * u[0::2])
        return qnorm((hi + u[1::2]) / _BIG)

    def norm1(self) -> float:
        return float(self.norm(1)[0])

    # ---- exponentials (sexp.c)
    def exp1(self) -> float:
        a = 0.0
        u = self.unif1()
        while u <= 0.0 or u >= 1.0:
            u = self.unif1()
        while True:
            u += u
            if u > 1.0:
                break
            a += _Q[0]
        u -= 1.0
        if u <= _Q[0]:
            return a + u
        i = 0
        ustar = self.unif1()
        umin = ustar
        while True:
            ustar = self.unif1()
            if umin > ustar:
                umin = ustar
            i += 1
            if not (u > _Q[i]):
                break
        return a + umin * _Q[0]

    # ---- gamma(shape a >= 1, scale 1): Ahrens-Dieter GD as in rgamma.c
    def gamma1(self, a: float) -> float:
        if a < 1.0:
            raise ValueError("only shape >= 1 is restated (BGGE uses (nr + 3) / 2)")
        s2 = a - 0.5
        s = math.sqrt(s2)
        d = _SQRT32 - s * 12.0
        t = self.norm1()
        x = s + 0.5 * t
        ret = x * x
        if t >= 0.0:
            return ret
        u = self.unif1()
        if d * u <= t * t * t:
            return ret
        r = 1.0 / a
        q0 = ((((((_Q7 * r + _Q6) * r + _Q5) * r + _Q4) * r + _Q3) * r + _Q2) * r + _Q1) * r
        if a <= 3.686:
            b = 0.463 + s + 0.178 * s2
            si = 1.235
            c = 0.195 / s - 0.079 + 0.16 * s
        elif a <= 13.022:
            b = 1.654 + 0.0076 * s2
            si = 1.68 / s + 0.275
            c = 0.062 / s + 0.024
        else:
            b = 1.77
            si = 0.75
            c = 0.1515 / s

        def qfun(tt):
            v = tt / (s + s)
            if abs(v) <= 0.25:
                return q0 + 0.5 * tt * tt * ((((((_A7 * v + _A6) * v + _A5) * v + _A4) * v + _A3) * v + _A2) * v + _A1) * v
            return q0 - s * tt + 0.25 * tt * tt + (s2 + s2) * math.log(1.0 + v)

        if x > 0.0:
            if math.log(1.0 - u) <= qfun(t):
                return ret
        while True:
            e = self.exp1()
            u = self.unif1()
            u = u + u - 1.0
            t = b - si * e if u < 0.0 else b + si * e
            if t >= -0.71874483771719:
                q = qfun(t)
                if q > 0.0:
                    w = math.expm1(q)
                    if c * abs(u) <= w * math.exp(e - 0.5 * t * t):
                        break
        x = s + 0.5 * t
        return x * x


def r_tapes(seed: int, n: int, nrs, q: int, n_na: int, ite: int, nu: float = 3.0):
    """Normal and chi-square tapes for `ite` iterations in the reference's call order:
    init: n normals per kernel (R/BGGE.R:254); then per iteration: 1 (mu, :279), q (XF, :107), for each kernel its
    nr_j normals (:91) followed by its variance draw rgamma(1, (nr_j + nu)/2, .) (:97), the residual variance draw
    rgamma(1, (n + nu)/2, .) (:102), and n_na normals (:379).  The chi-square tape holds 2 * Gamma(shape, 1), so that
    1 / rgamma(1, shape, rate = S / 2) = S / chi2."""
    rs = RStream(seed)
    nrs = [int(v) for v in nrs]
    nk = len(nrs)
    per_it = 1 + int(q) + sum(nrs) + int(n_na)
    z = np.empty(nk * n + ite * per_it)
    c = np.empty(ite * (nk + 1))
    z[:nk * n] = rs.norm(nk * n)
    zp, cp = nk * n, 0
    for _ in range(int(ite)):
        k = 1 + int(q) + (nrs[0] if nk else 0)                # mu, XF and the first kernel: one contiguous run
        z[zp:zp + k] = rs.norm(k)
        zp += k
        for j in range(nk):
            if j > 0:
                z[zp:zp + nrs[j]] = rs.norm(nrs[j])
                zp += nrs[j]
            c[cp] = 2.0 * rs.gamma1(0.5 * (nrs[j] + nu))
            cp += 1
        c[cp] = 2.0 * rs.gamma1(0.5 * (n + nu))
        cp += 1
        if n_na:
            z[zp:zp + n_na] = rs.norm(n_na)
            zp += n_na
    return z, c
