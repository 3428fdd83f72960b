"""ctypes binding of the C ABI declared in ``include/gxsv.h``.

There is no CPU fallback: if the CUDA library is missing, or no CUDA device is
visible when a state is created, the import / the constructor fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

GXSV_OK, GXSV_EINDEX, GXSV_EVALUE, GXSV_EZERONORM, GXSV_ECUDA, GXSV_ENOMEM, GXSV_EUNSUPPORTED = range(7)
OPT_FUSION, OPT_STRICT, OPT_PROFILE, OPT_DIST, OPT_BATCH, OPT_BATCH_AXES, OPT_CHUNK_LOG2, OPT_TMA, OPT_TILE_REGS, OPT_DIST_QUEUE = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
OPT_TILE_IO, OPT_DEFER_VCZ, OPT_TILE_WARPS = 11, 12, 13
KIND_NAMES = ("fill", "cz", "contract", "apply1q", "expect", "gather", "fused", "tensor", "misc")
K_COUNT = len(KIND_NAMES)


class Stats(C.Structure):
    _fields_ = [
        ("launches", C.c_uint64 * K_COUNT),
        ("algo_bytes", C.c_double * K_COUNT),
        ("device_ms", C.c_double * K_COUNT),
    ]


# every symbol include/gxsv.h declares: (name, restype, argtypes)
_H = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
SIGNATURES = {
    "gxsv_version": (C.c_int, []),
    "gxsv_create": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.POINTER(_H)]),
    "gxsv_create_external": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(_H)]),
    "gxsv_destroy": (C.c_int, [_H]),
    "gxsv_reset": (C.c_int, [_H]),
    "gxsv_last_error": (C.c_char_p, [_H]),
    "gxsv_set_option": (C.c_int, [_H, C.c_int, C.c_int]),
    "gxsv_get_option": (C.c_int, [_H, C.c_int, _I]),
    "gxsv_nqubit": (C.c_int, [_H]),
    "gxsv_max_qubits": (C.c_int, [_H]),
    "gxsv_ensure_capacity": (C.c_int, [_H, C.c_int]),
    "gxsv_add_plus": (C.c_int, [_H]),
    "gxsv_add_qubit": (C.c_int, [_H, _D]),
    "gxsv_tensor": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "gxsv_entangle": (C.c_int, [_H, C.c_int, C.c_int]),
    "gxsv_flush": (C.c_int, [_H]),
    "gxsv_evolve_single": (C.c_int, [_H, C.c_int, _D]),
    "gxsv_swap": (C.c_int, [_H, C.c_int, C.c_int]),
    "gxsv_evolve": (C.c_int, [_H, C.c_int, _I, C.c_void_p]),
    "gxsv_expectation_single": (C.c_int, [_H, C.c_int, _D, _D]),
    "gxsv_project_qubit": (C.c_int, [_H, C.c_int, _D, C.c_double, _D]),
    "gxsv_permute": (C.c_int, [_H, _I, C.c_int]),
    "gxsv_flatten": (C.c_int, [_H, C.c_void_p]),
    "gxsv_flatten_device": (C.c_int, [_H, C.c_void_p]),
    "gxsv_fidelity": (C.c_int, [_H, _H, _D]),
    "gxsv_inner": (C.c_int, [_H, _H, _D]),
    "gxsv_clone": (C.c_int, [_H, C.POINTER(_H)]),
    "gxsv_nonzero": (C.c_int, [_H, C.c_double, C.c_uint64, C.POINTER(C.c_uint64), _D, C.POINTER(C.c_uint64)]),
    "gxsv_norm2": (C.c_int, [_H, _D]),
    "gxsv_set_state": (C.c_int, [_H, C.c_int, C.c_void_p]),
    "gxsv_sync": (C.c_int, [_H]),
    "gxsv_stats": (C.c_int, [_H, C.POINTER(Stats)]),
    "gxsv_reset_stats": (C.c_int, [_H]),
    "gxsv_layout": (C.c_int, [_H, _I]),
    "gxsv_dm_trace_expect": (C.c_int, [_H, C.c_int, _I, _I, C.c_int, _D, _D]),
    "gxsv_nreal": (C.c_int, [_H]),
    "gxsv_is_virtual": (C.c_int, [_H, C.c_int]),
    "gxsv_settle": (C.c_int, [_H, C.c_int]),
    "gxsv_set_norm": (C.c_int, [_H, C.c_double]),
    "gxsv_take_norms": (C.c_int, [_H, _D, C.c_int, _I]),
    "gxsv_scale_host": (C.c_int, [_H, C.c_double, C.c_double]),
    "gxsv_get_hscale": (C.c_int, [_H, _D]),
    "gxsv_scale": (C.c_int, [_H, C.c_double, C.c_double]),
    "gxsv_ipc_export": (C.c_int, [_H, C.c_char_p]),
    "gxsv_ipc_open": (C.c_int, [_H, C.c_int, C.c_char_p]),
    "gxsv_ipc_close": (C.c_int, [_H]),
    "gxsv_swap_peer": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]),
    "gxsv_flush_batch": (C.c_int, [_H]),
    "gxsv_debug_stage": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _D, _D, _D, C.c_uint64, C.c_uint64, _D]),
    "gxsv_plan_tile_layouts": (C.c_int, [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, _I,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gxsv_pack_half": (C.c_int, [_H, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    "gxsv_unpack_half": (C.c_int, [_H, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_double, C.c_double]),
}

_lib = None


class LibraryMissingError(ImportError):
    pass


def load() -> C.CDLL:
    """Load ``lib/libgxsv.so`` (built in-tree by ``graphix_b200._build``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("GXSV_LIBRARY", _build.LIBPATH)
    if not os.path.exists(path):
        raise LibraryMissingError(
            f"{path} not found: build it with `python -m graphix_b200._build` "
            "(nvcc, sm_100a). graphix_b200 has no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def raise_for(lib: C.CDLL, handle, code: int) -> None:
    """Map a gxsv status to the exception type the reference raises."""
    if code == GXSV_OK:
        return
    msg = lib.gxsv_last_error(handle)
    text = msg.decode() if msg else f"gxsv error {code}"
    if code == GXSV_EINDEX:
        raise IndexError(text)
    if code == GXSV_EVALUE:
        raise ValueError(text)
    if code == GXSV_ENOMEM:
        raise MemoryError(text)
    if code == GXSV_EUNSUPPORTED:
        raise NotImplementedError(text)
    raise RuntimeError(text)
