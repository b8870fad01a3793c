"""Times bgge_dev_gram (library kernel) repeatedly on synthetic device data; prints every repetition."""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bgge_b200 import _lib as lib
lib.call("bgge_set_device", 0)
L, ldx = 10000, 10112
T = (L + 127) // 128
def vp(t): return C.c_void_p(t.data_ptr())
def run(name, X, p, divisor, mirror, reps=6):
    S = torch.empty((L, L), dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        lib.call("bgge_dev_gram", vp(X), ldx, L, p, 0, T, vp(S), L, 0, divisor, mirror, st)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:32s} p={p:6d} ms: " + " ".join(f"{t:7.1f}" for t in ts), flush=True)
p = 100000
X = torch.zeros((p, ldx), dtype=torch.float64, device="cuda")
X[:, :L] = torch.randn((p, L), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
run("randn, mirror, divisor p", X, p, float(p), 1)
run("randn, no mirror, divisor 1", X, p, 1.0, 0)
time.sleep(3)
run("after 3 s idle", X, p, 1.0, 0)
a = torch.randn((8192, 8192), dtype=torch.float64, device="cuda")
for _ in range(12): a @ a
run("right after 12 DGEMMs", X, p, 1.0, 0)
# ---- data / stream dependence
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
del X
Xg, ldx2, pp = bench.build_markers(torch, lib, L, p, 20260004, torch.device("cuda", 0))
torch.cuda.synchronize()
print("marker values sample:", Xg[0, :5].tolist(), "absmax", float(Xg.abs().max()))
run("genotypes, default stream", Xg, p, 1.0, 1)
s2 = torch.cuda.Stream()
with torch.cuda.stream(s2):
    run("genotypes, side stream", Xg, p, 1.0, 1)
Xr = torch.zeros((p, ldx), dtype=torch.float64, device="cuda")
Xr[:, :L] = torch.randn((p, L), dtype=torch.float64, device="cuda")
run("randn again", Xr, p, 1.0, 1)
Xr[:, :L] = torch.round(torch.randn((p, L), dtype=torch.float64, device="cuda"))
run("rounded randn (few distinct values)", Xr, p, 1.0, 1)
