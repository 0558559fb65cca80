"""ctypes binding of include/enero_b200.h (libenero_b200.so, built by enero_b200.build).

There is no CPU fallback: if the library is missing it is an ImportError-class
failure, and every compute entry point fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

HOST, DEVICE = 0, 1
ACTOR, CRITIC = 0, 1
MATH_FAST, MATH_ACCURATE = 0, 1
NPARAMS = 4201
GRAD_STRIDE = 4204
GRAD_FLOATS = 2 * GRAD_STRIDE + 4
PPO_HISTORY = 4096

_lib = None


class EneroError(RuntimeError):
    pass


def _p(t):
    return C.POINTER(t)


_i32p, _f32p, _f64p, _u8p, _i64p = _p(C.c_int32), _p(C.c_float), _p(C.c_double), _p(C.c_uint8), _p(C.c_int64)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/enero_b200.h one to one
SIGNATURES = {
    "enero_last_error": (C.c_char_p, []),
    "enero_version": (C.c_int, []),
    "enero_device_count": (C.c_int, []),
    "enero_ctx_create": (C.c_int, [C.c_int, _p(_vp)]),
    "enero_ctx_destroy": (C.c_int, [_vp]),
    "enero_ctx_sync": (C.c_int, [_vp]),
    "enero_ctx_launch_count": (C.c_int64, [_vp]),
    "enero_ctx_stream": (_vp, [_vp]),
    "enero_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "enero_ctx_set_math": (C.c_int, [_vp, C.c_int]),
    "enero_ctx_get_math": (C.c_int, [_vp]),
    "enero_ctx_set_engine": (C.c_int, [_vp, C.c_int]),
    "enero_ctx_get_engine": (C.c_int, [_vp]),
    "enero_ctx_profile_begin": (C.c_int, [_vp]),
    "enero_ctx_profile_end": (C.c_int, [_vp, _vp, _vp]),
    "enero_ctx_fp32_peak": (C.c_int, [_vp, _vp]),
    "enero_weights_upload": (C.c_int, [_vp, C.c_int, _vp]),
    "enero_weights_download": (C.c_int, [_vp, C.c_int, _vp]),
    "enero_topo_build": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _p(_vp)]),
    "enero_topo_destroy": (C.c_int, [_vp]),
    "enero_topo_sizes": (C.c_int, [_vp, _vp]),
    "enero_topo_get_pairs": (C.c_int, [_vp, _vp, _vp]),
    "enero_topo_get_sp_edges": (C.c_int, [_vp, _vp, _vp]),
    "enero_topo_get_middlepoints": (C.c_int, [_vp, _vp, _vp]),
    "enero_topo_get_capacity_feature": (C.c_int, [_vp, _vp]),
    "enero_topo_upload": (C.c_int, [_vp, _vp]),
    "enero_parse_demands": (C.c_int, [C.c_char_p, C.c_int, _vp]),
    "enero_parse_demands_batch": (C.c_int, [_p(C.c_char_p), C.c_int, C.c_int, _vp, C.c_int]),
    "enero_gnn_forward": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int, C.c_int, _vp, C.c_int]),
    "enero_decide_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "enero_decide_value_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "enero_value_batch": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "enero_batch_create": (C.c_int, [_vp, _vp, C.c_int, C.c_double, _p(_vp)]),
    "enero_batch_destroy": (C.c_int, [_vp]),
    "enero_batch_reset": (C.c_int, [_vp, _vp]),
    "enero_batch_policy": (C.c_int, [_vp, _vp, _vp]),
    "enero_batch_value": (C.c_int, [_vp, _vp]),
    "enero_batch_step": (C.c_int, [_vp, _vp, _vp, _vp]),
    "enero_batch_step_state": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "enero_batch_run_greedy": (C.c_int, [_vp]),
    "enero_batch_snapshot": (C.c_int, [_vp]),
    "enero_batch_restore": (C.c_int, [_vp]),
    "enero_batch_get_state": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "enero_batch_dmax": (C.c_int, [_vp]),
    "enero_batch_get_eligible": (C.c_int, [_vp, _vp, _vp]),
    "enero_batch_get_history": (C.c_int, [_vp, _vp, _vp, _vp]),
    "enero_batch_get_best": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "enero_batch_local_search": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "enero_batch_get_ls_eligible": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "enero_batch_get_ls_state": (C.c_int, [_vp, _vp, _vp]),
    "enero_batch_ls_route_add": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, _vp, _vp, _vp]),
    "enero_batch_rollout": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "enero_topo_build_ksp": (C.c_int, [_vp, C.c_int]),
    "enero_topo_ksp_sizes": (C.c_int, [_vp, _vp]),
    "enero_topo_get_ksp": (C.c_int, [_vp, _vp, _vp, _vp]),
    "enero_batch_sap": (C.c_int, [_vp, _vp, C.c_int, C.c_uint32, _vp, _vp]),
    "enero_ppo_grad_buffer": (_vp, [_vp]),
    "enero_ppo_zero_grad": (C.c_int, [_vp]),
    "enero_ppo_grad_ready": (C.c_int, [_vp]),
    "enero_ppo_accumulate": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float,
                                       C.c_float, C.c_int]),
    "enero_ppo_store": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "enero_ppo_accumulate_stored": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_float, C.c_float, C.c_float]),
    "enero_ppo_apply": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, _vp, C.c_int]),
    "enero_ppo_history": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "enero_ppo_grad_download": (C.c_int, [_vp, _vp, _vp, _vp]),
    "enero_adam_slots_upload": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "enero_adam_slots_download": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "enero_gae": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp]),
}


def library_path() -> str:
    return _build.LIB


def load():
    """dlopen libenero_b200.so (building it first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path) or _build.is_stale():
        try:
            _build.build_library()
        except Exception as exc:  # no nvcc on this machine and no prebuilt library
            if not os.path.isfile(path):
                raise EneroError("libenero_b200.so is missing and could not be built: %s" % exc)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the .so disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().enero_last_error()
        raise EneroError("enero_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def ptr(a):
    """void* of a numpy array (must be C-contiguous), a torch tensor, an int, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data_as(_vp)
    if isinstance(a, int):
        return _vp(a)
    if hasattr(a, "data_ptr"):
        return _vp(a.data_ptr())
    raise TypeError("cannot take a pointer of %r" % type(a))
