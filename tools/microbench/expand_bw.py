"""Expansion (R/getK.R:138,148,173) store bandwidth at C5 size against the write-only fill rate of the same buffer.
usage: python tools/microbench/expand_bw.py   (BGGE_EXPAND_SCALAR=1 selects the 8-byte-store kernel)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from bgge_b200 import _lib as lib  # noqa: E402

L, E = 10000, 4
n = L * E
dev = torch.device("cuda:0")
K0 = torch.randn((L, L), dtype=torch.float64, device=dev)
gid = torch.arange(L, dtype=torch.int32, device=dev).repeat(E)
env = torch.arange(E, dtype=torch.int32, device=dev).repeat_interleave(L)
out = torch.empty((n, n), dtype=torch.float64, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3)
    return best


gb = 8.0 * n * n / 1e9
for scalar in (0, 1):
    if scalar:
        os.environ["BGGE_EXPAND_SCALAR"] = "1"
    else:
        os.environ.pop("BGGE_EXPAND_SCALAR", None)
    for mode in (0, 1):
        fn = lambda: lib.call("bgge_dev_expand", vp(K0), L, L, vp(gid), vp(env), n, mode, 0, vp(out), n, st)
        fn()
        t = timed(fn)
        print(f"expand mode {mode} {'8B stores ' if scalar else '16B stcs  '} {t * 1e3:7.3f} ms  {gb / t:7.1f} GB/s", flush=True)
t = timed(out.zero_)
print(f"torch fill (zero_)           {t * 1e3:7.3f} ms  {gb / t:7.1f} GB/s")
rt = torch.cuda.cudart()
t = timed(lambda: rt.cudaMemset(out.data_ptr(), 0, n * n * 8) if hasattr(rt, "cudaMemset") else out.zero_())
print(f"cudaMemset                   {t * 1e3:7.3f} ms  {gb / t:7.1f} GB/s")
ref = K0[gid.long()][:, gid.long()] if False else None
# exactness of the new kernel against the old one on a ragged odd/even mix of sizes is covered by tests/test_gpu_getk.py
