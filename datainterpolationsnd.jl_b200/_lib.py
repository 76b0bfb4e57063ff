"""ctypes binding of libdind.so (include/dind.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libdind.so")

OK = 0
ERR_INVALID_ARGUMENT, ERR_SIZE_MISMATCH, ERR_UNSUPPORTED, ERR_CUDA, ERR_STATE, ERR_NO_DEVICE = -1, -2, -3, -4, -5, -6
ERR_NCCL = -7
UNIQUE_ID_BYTES = 128
F32, F64 = 0, 1
LINEAR, CONSTANT, BSPLINE = 0, 1, 2
FLAG_ASYNC, FLAG_COPY, FLAG_CACHED_IDX, FLAG_GENERIC, FLAG_CELL_SORTED = 1, 2, 4, 8, 16
FLAG_PLAN_ORDER = 32
MAX_DIMS, MAX_DEGREE = 8, 7

# every symbol include/dind.h declares: (name, restype, argtypes)
_vp, _i, _i64, _u = C.c_void_p, C.c_int, C.c_int64, C.c_uint
_pi64, _pint = C.POINTER(C.c_int64), C.POINTER(C.c_int)
SYMBOLS = [
    ("dind_create", _i, [C.POINTER(_vp), _i, _i, _i]),
    ("dind_destroy", _i, [_vp]),
    ("dind_set_stream", _i, [_vp, _vp]),
    ("dind_synchronize", _i, [_vp]),
    ("dind_set_dim", _i, [_vp, _i, _i, _vp, _i64, _i, _pi64, _i]),
    ("dind_set_t_eval", _i, [_vp, _i, _vp, _i64, _i, _u]),
    ("dind_set_u", _i, [_vp, _vp, _pi64, _i, _u]),
    ("dind_set_weights", _i, [_vp, _vp, _pi64, _i, _u]),
    ("dind_eval_unstructured", _i, [_vp, _vp, _pint, _u]),
    ("dind_eval_grid", _i, [_vp, _vp, _pint, _u]),
    ("dind_eval_unstructured_gradient", _i, [_vp, _vp, _u]),
    ("dind_eval_points", _i, [_vp, _vp, C.POINTER(_vp), _i64, _pint, _u]),
    ("dind_build_caches", _i, [_vp, _i, _pint, _u]),
    ("dind_get_idx_eval", _i, [_vp, _i, _pi64]),
    ("dind_get_knots_all", _i, [_vp, _i, _vp, _pi64]),
    ("dind_get_basis_function_eval", _i, [_vp, _i, _vp]),
    ("dind_get_plan_permutation", _i, [_vp, _vp]),
    ("dind_last_kernel", C.c_char_p, [_vp]),
    ("dind_launch_count", _i64, [_vp]),
    ("dind_shard_range", _i, [_i64, _i, _i, _pi64, _pi64]),
    ("dind_comm_get_unique_id", _i, [_vp]),
    ("dind_comm_init", _i, [_vp, _vp, _i, _i]),
    ("dind_comm_destroy", _i, [_vp]),
    ("dind_comm_info", _i, [_vp, _pint, _pint, _pint]),
    ("dind_allgather_out", _i, [_vp, _vp, _vp, _i64, _i64, _i64, _u]),
    ("dind_eval_unstructured_allgather", _i, [_vp, _vp, _i64, _pint, _i, _u]),
    ("dind_host_register", _i, [_vp, _i64]),
    ("dind_host_unregister", _i, [_vp]),
    ("dind_last_error", C.c_char_p, []),
    ("dind_version", _i, []),
    ("dind_device_count", _i, []),
]

_lib = None


class DindError(AssertionError):
    """A negative status from libdind.  Subclasses AssertionError because the reference
    signals every validation failure with `@assert`."""

    def __init__(self, code, msg):
        super().__init__(f"[dind {code}] {msg}")
        self.code = code


def load():
    """dlopen libdind.so and bind every symbol of include/dind.h; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python datainterpolationsnd.jl_b200/build.py` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise DindError(rc, load().dind_last_error().decode())


def shard_range(n, rank, world):
    lo, hi = C.c_int64(), C.c_int64()
    check(load().dind_shard_range(n, rank, world, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value
