"""bgge_getk_base from a pageable host marker matrix at C5 size: wall time against the number of upload/contraction
chunks (BGGE_GRAM_HOST_CHUNKS) and staging threads (BGGE_COPY_THREADS)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from bgge_b200 import _lib as lib  # noqa: E402

L, p = 10000, 100000
rng = np.random.default_rng(0)
X = np.empty((p, L))            # C-order (p, L) == column-major L x p
for i in range(0, p, 10000):
    X[i:i + 10000] = rng.standard_normal((10000, L))
K0 = np.empty((1, L, L))
bw = np.ones(1)
for chunks in ("1", "3", "4", "6", "8", "4"):
    os.environ["BGGE_GRAM_HOST_CHUNKS"] = chunks
    t0 = time.perf_counter()
    lib.call("bgge_getk_base", lib.ptr(X), L, p, lib.GK, lib.ptr(bw), 1, 0.5, lib.ptr(K0), None)
    print(f"chunks {chunks}: {time.perf_counter() - t0:.3f} s  (K0[0,1] = {K0[0, 0, 1]:.12f})", flush=True)
