"""Host-side plans for the two places the path shards across GPUs:

* independent chains / CV folds -> `shard_units` (no data-path collective), and
* the getK contraction          -> `panel_plan` / `assemble_panels`: 2*world equal tile-row chunks of
  the lower triangle, rank r owns chunks r and 2*world-1-r (equal work), one all-gather of the
  panels, local mirror.  The functions are backend-agnostic (the tests drive them over gloo on
  CPU with a numpy stand-in for the Gram kernel; bench.py drives them over NCCL).
"""
from __future__ import annotations

TILE = 128


def shard_units(n_units: int, world: int, rank: int):
    """Contiguous, balanced assignment of independent units (chains x folds) to ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def panel_plan(L: int, world: int):
    """Returns (chunk_tile_rows, panel_rows, owners) where owners[r] = the two chunk indices of rank r."""
    T = (L + TILE - 1) // TILE
    ch = (T + 2 * world - 1) // (2 * world)
    owners = [(r, 2 * world - 1 - r) for r in range(world)]
    return ch, ch * TILE, owners


def chunk_tile_rows(L: int, world: int, chunk: int):
    """Tile-row range [t0, t1) of one chunk, clamped to the matrix."""
    T = (L + TILE - 1) // TILE
    ch, _, _ = panel_plan(L, world)
    return min(T, chunk * ch), min(T, (chunk + 1) * ch)


def tiles_per_rank(L: int, world: int):
    """Lower-triangle tile count each rank computes (for balance checks / reporting)."""
    out = []
    for r in range(world):
        n = 0
        for c in (r, 2 * world - 1 - r):
            t0, t1 = chunk_tile_rows(L, world, c)
            n += sum(t + 1 for t in range(t0, t1))
        out.append(n)
    return out


def assemble_panels(S, gathered, L: int, world: int):
    """gathered[r][k] is rank r's k-th panel stored [column j][local row]; S is indexed S[j, i]
    (column-major L x L seen as a row-major (L, L) array).  Fills the rows every panel covers."""
    _, R, owners = panel_plan(L, world)
    for r in range(world):
        for k, c in enumerate(owners[r]):
            r0 = c * R
            r1 = min(L, r0 + R)
            if r1 > r0:
                S[:, r0:r1] = gathered[r][k][:, :r1 - r0]
    return S


# ----------------------------------------------------------------------------------------------------------
# column-sharded single chain: host loop around bgge_gibbs_shard_begin / _end
# ----------------------------------------------------------------------------------------------------------
def column_range(nr: int, world: int, rank: int, group: int = 1):
    """Eigen-columns [lo, hi) of a matrix with nr columns that rank sweeps; boundaries on multiples of `group`
    (the column-group size of the sweep kernel).  Mirrors the planner in csrc/sweep_fused.cu."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    lo = nr * rank // world // group * group if world > 1 else 0
    hi = nr * (rank + 1) // world // group * group if (world > 1 and rank + 1 < world) else nr
    return lo, hi


class DeviceVector:
    """A device buffer owned by libbgge_b200, exposed through __cuda_array_interface__ so that torch can wrap it
    without a copy (torch.as_tensor(DeviceVector(ptr, count), device="cuda"))."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def nccl_allreduce(group=None):
    """reduce(ptr, count): in-place sum over the ranks of `group` with torch.distributed (NCCL over NVLink).  The
    collective is ordered against torch's current stream, which knows nothing about the handle's own stream: the
    callback therefore returns only when the reduction has finished (run_sharded drains the handle's stream before it
    calls).  Correct but latency-bound -- the fast path is a library communicator (Comm below, BGGE(shard=(rank,
    world, comm))): the all-reduce is then issued on the handle's stream inside libbgge_b200, no host sync at all."""
    import torch
    import torch.distributed as dist
    cache = {}

    def reduce(ptr, count):
        t = cache.get((ptr, count))
        if t is None:
            t = cache[(ptr, count)] = torch.as_tensor(DeviceVector(ptr, count), device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        torch.cuda.current_stream().synchronize()

    return reduce


def run_sharded(h, nk: int, iters: int, reduce):
    """`iters` iterations of a column-sharded chain with a caller-supplied collective: per kernel step sweep the local
    columns, wait for them (the callback's stream is not the handle's), add the partial W b and sum b^2/s of all
    processes (reduce(ptr, count): in place, complete on return), finish the step."""
    import ctypes as C
    from . import _lib
    vec, cnt = C.c_void_p(), C.c_int64()
    for _ in range(int(iters)):
        for j in range(nk):
            _lib.call("bgge_gibbs_shard_begin", h, j, C.byref(vec), C.byref(cnt))
            _lib.call("bgge_gibbs_sync", h)
            reduce(vec.value, cnt.value)
            _lib.call("bgge_gibbs_shard_end", h, j)
        _lib.call("bgge_gibbs_shard_iteration_end", h)


class Comm:
    """A communicator owned by libbgge_b200 (NCCL inside the library, include/bgge_b200.h "multi-GPU inside the
    library").  Comm.local(ndev) -> one per GPU of this process; Comm.from_torch_distributed() -> this process' rank
    of the default torch.distributed group (torch only ships the 128-byte id; the collectives are the library's)."""

    def __init__(self, handle, rank, world, device):
        self.handle, self.rank, self.world, self.device = handle, rank, world, device

    @staticmethod
    def local(ndev, devices=None):
        import ctypes as C
        import numpy as np
        from . import _lib
        arr = (C.c_void_p * ndev)()
        dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        _lib.call("bgge_comm_init_all", int(ndev), _lib.ptr(dv), C.cast(arr, C.POINTER(C.c_void_p)))
        return [Comm(C.c_void_p(arr[d]), d, ndev, d if devices is None else int(devices[d])) for d in range(ndev)]

    @staticmethod
    def from_torch_distributed(device):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        rank, world = dist.get_rank(), dist.get_world_size()
        idbuf = (C.c_uint8 * 128)()
        if rank == 0:
            _lib.call("bgge_comm_unique_id", idbuf)
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=device)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        idbuf = (C.c_uint8 * 128).from_buffer_copy(raw)
        _lib.call("bgge_set_device", int(device.index if hasattr(device, "index") else device))
        h = C.c_void_p()
        _lib.call("bgge_comm_init_rank", idbuf, world, rank, C.byref(h))
        return Comm(h, rank, world, int(device.index if hasattr(device, "index") else device))

    def destroy(self):
        from . import _lib
        if self.handle:
            _lib.load().bgge_comm_destroy(self.handle)
            self.handle = None

