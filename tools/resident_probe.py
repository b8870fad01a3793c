"""Phase timing of the resident sweep kernel on a wheat-shaped model (debug aid, GPU only).
Usage: BGGE_RESIDENT_TIMING=1 [BGGE_RESIDENT_CTAS=n] python tools/resident_probe.py [MM|MDs|MDe] [iters] [xf]"""
import sys, os, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgge_b200
from bgge_b200 import _lib as lib

model = sys.argv[1] if len(sys.argv) > 1 else "MDs"   # MM | MDs | MDe
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
use_xf = len(sys.argv) > 3 and sys.argv[3] == "xf"   # environment means as fixed effects
L, E, p = 599, 4, 1279
rng = np.random.default_rng(0)
X = rng.integers(0, 3, size=(L, p)).astype(np.float64)
names = [f"g{i}" for i in range(L)]
Y = [(f"e{e}", names[i], 0.0) for e in range(E) for i in range(L)]
K = bgge_b200.getK(Y, X, kernel="GB", model=model, rownames=names, factored=True)
y = rng.standard_normal(L * E)
ne = np.full(E, L, dtype=np.int64)
t0 = time.perf_counter()
XF = None
if use_xf:
    XF = np.zeros((L * E, E))
    for e in range(E):
        XF[e * L:(e + 1) * L, e] = 1.0
fit = bgge_b200.BGGE(y, K, XF=XF, ne=ne, ite=iters, burn=iters // 5, thin=2, seed=1)
dt = time.perf_counter() - t0
print(f"{model}: {iters} iterations, whole BGGE() call {dt:.3f} s; varE {fit['varE']:.4f}")
