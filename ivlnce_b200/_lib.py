"""ctypes binding of libivlnmap.so (include/ivln_map.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, this
module raises.  The library is built in-tree by `ivlnce_b200.build.build_library`
(nvcc, sm_100a) so that it travels with the repository snapshot.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libivlnmap.so")

IVM_OK = 0
_ERRORS = {1: "invalid argument", 2: "workspace too small or misaligned", 3: "CUDA runtime error",
           4: "step counter overflow (call rebase_stamps)"}

ERR_STORE_OVERFLOW = 1
ERR_EDGE_OVERFLOW = 2
ERR_KNOWN_OVERFLOW = 4
ERR_GRID_BARRIER = 8
ERR_CAND_OVERFLOW = 16


class IvmConfig(ctypes.Structure):
    _fields_ = [
        ("max_envs", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("map_rows", ctypes.c_int32), ("map_cols", ctypes.c_int32),
        ("res", ctypes.c_float), ("half_res", ctypes.c_float), ("half_h", ctypes.c_float), ("half_w", ctypes.c_float),
        ("store_rows", ctypes.c_int32), ("store_cols", ctypes.c_int32), ("mode", ctypes.c_int32),
        ("known_capacity", ctypes.c_int64),
        ("tile_rows", ctypes.c_int32), ("tile_cols", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 4),
    ]


class IvmStatus(ctypes.Structure):
    _fields_ = [("error_flags", ctypes.c_uint32), ("pad", ctypes.c_uint32), ("stats", ctypes.c_uint64 * 8)]


class MapLibraryError(RuntimeError):
    pass


_lib = None
_vp = ctypes.c_void_p


def load() -> ctypes.CDLL:
    """Load libivlnmap.so; raises loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MapLibraryError(
            f"{LIB_PATH} not found: the CUDA library must be built first "
            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)

    L = ctypes.CDLL(LIB_PATH)
    L.ivm_workspace_bytes.restype = ctypes.c_size_t
    L.ivm_workspace_bytes.argtypes = [ctypes.POINTER(IvmConfig)]
    L.ivm_create.argtypes = [ctypes.POINTER(IvmConfig), _vp, ctypes.c_size_t, ctypes.POINTER(_vp)]
    L.ivm_destroy.argtypes = [_vp]
    L.ivm_set_camera.argtypes = [_vp, _vp, _vp, _vp]
    L.ivm_step_iterative.argtypes = [_vp, ctypes.c_int32, _vp, _vp, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp,
                                     ctypes.c_int32, _vp, _vp, _vp, _vp]
    L.ivm_known_load.argtypes = [_vp, ctypes.c_int32, ctypes.c_int64, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp]
    L.ivm_known_clear.argtypes = [_vp, ctypes.c_int32, _vp]
    L.ivm_step_known.argtypes = [_vp, ctypes.c_int32, _vp, _vp, _vp, ctypes.c_int32, _vp, _vp, _vp]
    L.ivm_export_world.argtypes = [_vp, ctypes.c_int32, ctypes.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]
    L.ivm_read_status.argtypes = [_vp, ctypes.POINTER(IvmStatus), _vp]
    L.ivm_set_timing.argtypes = [_vp, ctypes.c_int32]
    L.ivm_stage_times.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]
    L.ivm_kernel_launches.restype = ctypes.c_int64
    L.ivm_kernel_launches.argtypes = [_vp]
    L.ivm_rebase_stamps.argtypes = [_vp, _vp]
    L.ivm_debug_set_step.argtypes = [_vp, ctypes.c_uint32]
    L.ivm_set_pipelined.argtypes = [_vp, ctypes.c_int32]
    L.ivm_read_phase_ns.argtypes = [_vp, ctypes.POINTER(ctypes.c_uint64), _vp]
    L.ivm_read_cta_trace.argtypes = [_vp, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32, _vp]
    L.ivm_copy_state.argtypes = [_vp, _vp, _vp]
    L.ivm_map_features.argtypes = [_vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp, _vp, _vp]
    L.ivm_rednet_preprocess.argtypes = [_vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp,
                                        ctypes.c_int32, ctypes.c_int32, _vp, _vp, _vp]
    L.ivm_copy_error_flags_async.argtypes = [_vp, _vp, _vp]
    L.ivm_clear_error_flags.argtypes = [_vp, _vp]
    L.ivm_last_cuda_error.restype = ctypes.c_char_p
    L.ivm_last_cuda_error.argtypes = [_vp]
    L.ivm_version.restype = ctypes.c_char_p
    for name in ("ivm_create", "ivm_destroy", "ivm_set_camera", "ivm_step_iterative", "ivm_known_load", "ivm_known_clear",
                 "ivm_step_known", "ivm_export_world", "ivm_read_status", "ivm_set_timing", "ivm_stage_times",
                 "ivm_rebase_stamps", "ivm_debug_set_step", "ivm_set_pipelined", "ivm_copy_state", "ivm_read_phase_ns", "ivm_read_cta_trace", "ivm_map_features", "ivm_rednet_preprocess", "ivm_copy_error_flags_async",
                 "ivm_clear_error_flags"):
        getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


def check(rc: int, ctx=None, what: str = "") -> None:
    if rc == IVM_OK:
        return
    msg = _ERRORS.get(rc, f"error {rc}")
    if rc == 3 and ctx is not None:
        msg += ": " + (load().ivm_last_cuda_error(ctx) or b"").decode()
    raise MapLibraryError(f"{what}: {msg}" if what else msg)
