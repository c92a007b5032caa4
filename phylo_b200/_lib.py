"""ctypes binding of libvcsmc_b200.so (the C ABI declared in include/vcsmc_b200.h).

There is no fallback: if the shared library is missing or fails to load, importing a compute
entry raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C phylo_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvcsmc_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_POOL, ERR_DATA, ERR_STATE = 0, -1, -2, -3, -4, -5


class VcsmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libvcsmc_b200 error %d: %s" % (code, msg))
        self.code = code


class SweepConfig(C.Structure):
    _fields_ = [("n_taxa", C.c_int32), ("n_sites", C.c_int32), ("n_particles", C.c_int64), ("jc", C.c_int32),
                ("keep_for_backward", C.c_int32), ("workspace_bytes", C.c_int64), ("n_sub", C.c_int32), ("reserved", C.c_int32)]


class SweepSizes(C.Structure):
    _fields_ = [("min_bytes", C.c_int64), ("retain_bytes", C.c_int64)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
COMM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p)
COMM_ALLGATHER, COMM_BARRIER, COMM_ALLREDUCE = 1, 2, 3

_P = C.c_void_p
_SIGNATURES = {
    "vcsmc_abi_version": (C.c_int, []),
    "vcsmc_last_error": (C.c_char_p, []),
    "vcsmc_launch_count": (C.c_uint64, []),
    "vcsmc_pack_alignment": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "vcsmc_gather_sites": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "vcsmc_transition_fwd": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P]),
    "vcsmc_transition_host": (C.c_int, [_P, _P, C.c_int64, _P]),
    "vcsmc_transition_bwd": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, _P, _P, _P]),
    "vcsmc_merge_tiles": (C.c_int, [C.c_int]),
    "vcsmc_merge_fwd": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P]),
    "vcsmc_merge_bwd": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P]),
    "vcsmc_propose_pairs": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, _P]),
    "vcsmc_resample_work_doubles": (C.c_int64, [C.c_int64]),
    "vcsmc_resample": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "vcsmc_philox_step_uniforms": (C.c_int, [C.c_uint64, C.c_int, C.c_int64, C.c_int64, C.c_int, _P, _P, _P, _P, _P]),
    "vcsmc_sweep_query": (C.c_int, [C.POINTER(SweepConfig), C.POINTER(SweepSizes)]),
    "vcsmc_sweep_create": (C.c_int, [C.POINTER(SweepConfig), _P, C.POINTER(_P)]),
    "vcsmc_sweep_destroy": (None, [_P]),
    "vcsmc_sweep_set_allreduce": (C.c_int, [_P, ALLREDUCE_FN, _P]),
    "vcsmc_sweep_set_comm": (C.c_int, [_P, C.c_int, C.c_int, COMM_FN, _P, C.POINTER(C.c_void_p)]),
    "vcsmc_ipc_export": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "vcsmc_ipc_open": (C.c_int, [_P, C.POINTER(_P)]),
    "vcsmc_ipc_close": (C.c_int, [_P]),
    "vcsmc_sweep_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "vcsmc_sweep_set_uniforms": (C.c_int, [_P, _P, _P, _P, _P]),
    "vcsmc_sweep_set_seed": (C.c_int, [_P, C.c_uint64]),
    "vcsmc_sweep_set_uniforms_nested": (C.c_int, [_P, _P, _P, _P, _P]),
    "vcsmc_sweep_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "vcsmc_sweep_backward": (C.c_int, [_P, C.c_double, _P, _P, _P, _P, _P]),
    "vcsmc_sweep_output": (_P, [_P, C.c_char_p]),
    "vcsmc_sweep_profile": (C.c_int, [_P, C.POINTER(C.c_double)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: the CUDA library is not built and there is no fallback path. "
            "Run `make -C phylo_b200/csrc` (or __graft_entry__.build())." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.vcsmc_abi_version() != 4:
        raise ImportError("libvcsmc_b200 ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise VcsmcError(rc, load().vcsmc_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().vcsmc_launch_count())
