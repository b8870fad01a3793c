"""Diagnostic: repeated 200-iteration runs on the C5 shapes; prints per-run wall time and a clock trace."""
import ctypes as C, os, subprocess, sys, time
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
def make(iters):
    h = C.c_void_p()
    lib.call("bgge_gibbs_create", n, 2, 0, None, C.byref(h))
    lib.call("bgge_gibbs_set_y", h, lib.ptr(y), None)
    lib.call("bgge_gibbs_set_kernel", h, 0, 0, 1.0)
    lib.call("bgge_gibbs_add_block", h, 0, 0, n, L, lib.ptr(s), vp(UG), n, 1)
    lib.call("bgge_gibbs_set_kernel", h, 1, 1, 1.0)
    for e in range(E):
        lib.call("bgge_gibbs_add_block", h, 1, e * L, L, L, lib.ptr(s), vp(V[e]), L, 1)
    lib.call("bgge_gibbs_init", h, iters, iters // 5, 3, 0.5, 0, 1, 0)
    return h
mode = sys.argv[1] if len(sys.argv) > 1 else "fresh"
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,pstate", "--format=csv,noheader", "-lms", "100"],
                       stdout=open("/tmp/trace.csv", "w"))
time.sleep(0.5)
h = make(200 * 12) if mode == "reuse" else None
for k in range(12):
    t0 = time.perf_counter()
    hh = h if mode == "reuse" else make(200)
    lib.call("bgge_gibbs_run", hh, 200)
    lib.call("bgge_gibbs_sync", hh)
    t1 = time.perf_counter()
    if mode != "reuse":
        lib.load().bgge_gibbs_destroy(hh)
    print(f"{mode} run {k}: {1e3 * (t1 - t0):.1f} ms", flush=True)
    if mode == "gap":
        time.sleep(0.3)
smi.terminate()
print(open("/tmp/trace.csv").read()[-1500:])
