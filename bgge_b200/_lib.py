"""ctypes binding of libbgge_b200.so (include/bgge_b200.h).

The shared library is the product; this module only loads it and checks status codes.
There is no Python/NumPy fallback: if the library is missing or no sm_100 device is
present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BGGE_LIB selects another build of the same library (e.g. libbgge_b200_checks.so from `make CHECKS=1`)
LIB_PATH = os.environ.get("BGGE_LIB") or os.path.join(_HERE, "libbgge_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)
c_uint32_p = C.POINTER(C.c_uint32)
c_uint64_p = C.POINTER(C.c_uint64)
c_void_pp = C.POINTER(C.c_void_p)

GB, GK = 0, 1
EXPAND_G, EXPAND_GE, EXPAND_ENV, EXPAND_GI = 0, 1, 2, 3
TYPE_D, TYPE_BD = 0, 1
RNG_PHILOX, RNG_TAPE = 0, 1


KERNEL_DENSE = -1


class KernelDesc(C.Structure):
    """struct bgge_kernel_desc (include/bgge_b200.h)."""
    _fields_ = [("type", C.c_int), ("mode", C.c_int), ("env_k", C.c_int), ("reserved", C.c_int), ("L", C.c_int64),
                ("K0", C.c_void_p), ("dense", C.c_void_p)]


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_int64, C.c_double, C.c_void_p)


class BGGEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbgge_b200 error {code}: {msg}")
        self.code = code


# name -> (restype, argtypes); must list every symbol of include/bgge_b200.h
SIGNATURES = {
    "bgge_version": (C.c_int, []),
    "bgge_last_error": (C.c_char_p, []),
    "bgge_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bgge_set_device": (C.c_int, [C.c_int]),
    "bgge_release_cache": (C.c_int, []),
    "bgge_factor_matches": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_void_p]),
    "bgge_device_memory": (C.c_int, [c_uint64_p, c_uint64_p]),
    "bgge_getk_base": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                 C.c_void_p, c_double_p]),
    "bgge_getk_expand": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_void_p]),
    "bgge_getk": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p,
                            C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, c_void_pp, C.c_int, c_double_p]),
    "bgge_marker_ld": (C.c_int, [C.c_int64, c_int64_p]),
    "bgge_marker_cols": (C.c_int, [C.c_int64, c_int64_p]),
    "bgge_dev_gram": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_void_p]),
    "bgge_dev_symmetrize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "bgge_dev_gk_distance": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "bgge_dev_quantile7": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p, c_double_p,
                                     C.c_void_p]),
    "bgge_dev_gk_exp": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p]),
    "bgge_dev_expand": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                  C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "bgge_eig": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_void_p, c_int64_p]),
    "bgge_dev_eig": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_int64, c_int64_p, C.c_void_p]),
    "bgge_gibbs_create": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_void_p, c_void_pp]),
    "bgge_gibbs_destroy": (None, [C.c_void_p]),
    "bgge_gibbs_set_y": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgge_gibbs_set_xf": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bgge_gibbs_set_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double]),
    "bgge_gibbs_add_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_int]),
    "bgge_gibbs_add_kernel_matrix": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                               C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "bgge_gibbs_add_expanded_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                 C.c_double, C.c_int]),
    "bgge_gibbs_block_count": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "bgge_gibbs_block_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_void_pp, c_int64_p]),
    "bgge_gibbs_fetch_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int64_p, c_int64_p, c_int64_p, C.c_void_p,
                                         C.c_void_p, C.c_int64]),
    "bgge_gibbs_set_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "bgge_gibbs_resident_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "bgge_gibbs_set_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "bgge_gibbs_shard_begin": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "bgge_gibbs_shard_end": (C.c_int, [C.c_void_p, C.c_int]),
    "bgge_gibbs_shard_iteration_end": (C.c_int, [C.c_void_p]),
    "bgge_gibbs_shard_advance": (C.c_int, [C.c_void_p, C.c_int64]),
    "bgge_gibbs_merged_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bgge_gibbs_kernel_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bgge_gibbs_kernel_bytes": (C.c_int, [C.c_void_p, C.c_int, c_int64_p, C.POINTER(C.c_int)]),
    "bgge_gibbs_share_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "bgge_gibbs_set_tape": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "bgge_gibbs_init": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_uint64,
                                  C.c_uint32]),
    "bgge_gibbs_run": (C.c_int, [C.c_void_p, C.c_int64]),
    "bgge_gibbs_sync": (C.c_int, [C.c_void_p]),
    "bgge_gibbs_profile": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "bgge_gibbs_sizes": (C.c_int, [C.c_void_p, C.c_void_p, c_int64_p, c_int64_p, c_int64_p, c_int64_p]),
    "bgge_gibbs_state": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "bgge_gibbs_chain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgge_gibbs_summary": (C.c_int, [C.c_void_p, C.c_void_p, c_double_p, c_double_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgge_batch_create": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_void_p, c_void_pp]),
    "bgge_batch_destroy": (None, [C.c_void_p]),
    "bgge_batch_set_y": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgge_batch_set_xf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "bgge_batch_set_kernel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double]),
    "bgge_batch_add_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_int]),
    "bgge_batch_init": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_uint64, C.c_uint32]),
    "bgge_batch_run": (C.c_int, [C.c_void_p, C.c_int64]),
    "bgge_batch_sync": (C.c_int, [C.c_void_p]),
    "bgge_batch_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), c_int64_p, C.POINTER(C.c_int)]),
    "bgge_batch_chain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgge_batch_chain_bet": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bgge_batch_summary": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "bgge_fit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_void_pp, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                           C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_int,
                           C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, c_double_p,
                           c_double_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p]),
    "bgge_comm_init_all": (C.c_int, [C.c_int, C.c_void_p, c_void_pp]),
    "bgge_comm_unique_id": (C.c_int, [C.c_void_p]),
    "bgge_comm_init_rank": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_void_pp]),
    "bgge_comm_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bgge_comm_destroy": (None, [C.c_void_p]),
    "bgge_dev_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "bgge_dev_gram_panels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_double,
                                       C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "bgge_getk_base_multi": (C.c_int, [c_void_pp, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                                       C.c_double, C.c_void_p, c_double_p]),
    "bgge_gibbs_set_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bgge_gibbs_run_sharded": (C.c_int, [C.c_void_p, C.c_int64]),
    "bgge_fit_kernels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                   C.c_int, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, c_double_p, c_double_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_int)]),
    "bgge_test_philox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgge_test_draws": (C.c_int, [C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p, C.c_double, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C bgge_b200/csrc` (or __graft_entry__.build()). "
            "bgge_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_nccl_preloaded = False
_NCCL_ENTRY = ("bgge_comm_init_all", "bgge_comm_unique_id", "bgge_comm_init_rank")


def preload_nccl():
    """libbgge_b200 binds NCCL at run time (dlopen "libnccl.so.2"; a copy the process already holds is reused).  A
    Python process that will ALSO import torch must end up with torch's own (newer) NCCL under that soname, otherwise
    `import torch` fails on symbols the system copy lacks -- so before the first communicator is made, load the copy
    bundled with the Python environment if there is one (BGGE_NCCL_LIB overrides the path).  Without such a copy the
    library falls back to the system NCCL by itself."""
    global _nccl_preloaded
    if _nccl_preloaded:
        return
    _nccl_preloaded = True
    path = os.environ.get("BGGE_NCCL_LIB")
    if not path:
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for loc in (spec.submodule_search_locations if spec is not None else []):
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    path = cand
                    break
        except (ImportError, ValueError, AttributeError):
            path = None
    if path and os.path.exists(path):
        C.CDLL(path, mode=C.RTLD_GLOBAL)


def check(code):
    if code != 0:
        raise BGGEError(code, load().bgge_last_error().decode("utf-8", "replace"))


def call(name, *args):
    lib = load()
    if name in _NCCL_ENTRY:
        preload_nccl()
    check(getattr(lib, name)(*args))


def ptr(a):
    """Raw pointer of a numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def f64(a, order="F"):
    return np.require(a, dtype=np.float64, requirements=["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS", "ALIGNED"])


def device_count():
    c = C.c_int(0)
    call("bgge_device_count", C.byref(c))
    return c.value
