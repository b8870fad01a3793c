"""Host-side time breakdown of one e2e BGGE() step on the C5 shapes (synthetic eigen blocks)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgge_b200 import _lib as lib
lib.call("bgge_set_device", 0)
L, E = 10000, 4
n = L * E
dev = torch.device("cuda", 0)
UG = torch.randn((L, n), dtype=torch.float64, device=dev) / 200
V = [torch.randn((L, L), dtype=torch.float64, device=dev) / 100 for _ in range(E)]
s = np.sort(np.random.default_rng(0).gamma(2, 1, L) + 0.05)[::-1].copy()
y = np.random.default_rng(1).standard_normal(n)
def vp(t): return C.c_void_p(t.data_ptr())
def step(iters, verbose):
    t = [time.perf_counter()]
    h = C.c_void_p()
    lib.call("bgge_gibbs_create", n, 2, 0, None, C.byref(h)); t.append(time.perf_counter())
    lib.call("bgge_gibbs_set_y", h, lib.ptr(y), None)
    lib.call("bgge_gibbs_set_kernel", h, 0, 0, 1.0)
    lib.call("bgge_gibbs_add_block", h, 0, 0, n, L, lib.ptr(s), vp(UG), n, 1)
    lib.call("bgge_gibbs_set_kernel", h, 1, 1, 1.0)
    for e in range(E):
        lib.call("bgge_gibbs_add_block", h, 1, e * L, L, L, lib.ptr(s), vp(V[e]), L, 1)
    t.append(time.perf_counter())
    lib.call("bgge_gibbs_init", h, iters, iters // 5, 3, 0.5, 0, 1, 0); t.append(time.perf_counter())
    lib.call("bgge_gibbs_run", h, iters); t.append(time.perf_counter())
    lib.call("bgge_gibbs_sync", h); t.append(time.perf_counter())
    yh = np.empty(n); u = np.empty((2, n)); usd = np.empty((2, n)); yo = np.empty(n); vu = np.empty(2); vs = np.empty(2)
    a, b = C.c_double(), C.c_double()
    lib.call("bgge_gibbs_summary", h, lib.ptr(yh), C.byref(a), C.byref(b), lib.ptr(u), lib.ptr(usd), lib.ptr(vu), lib.ptr(vs), lib.ptr(yo))
    t.append(time.perf_counter())
    lib.load().bgge_gibbs_destroy(h); t.append(time.perf_counter())
    if verbose:
        names = ["create", "set/add", "init", "run(launch)", "sync", "summary", "destroy"]
        print(" ".join(f"{nm}={1e3 * (t[i + 1] - t[i]):.1f}ms" for i, nm in enumerate(names)), f"total={1e3 * (t[-1] - t[0]):.1f}ms")
for k in range(4):
    step(200, True)
