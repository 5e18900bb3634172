"""ctypes binding of libfeddb200.so (the C ABI declared in include/feddb200.h).

The library is built in-tree (feddlib_b200/libfeddb200.so, see feddlib_b200/build.py).  There is no
CPU fallback: a missing library or a missing CUDA device raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEDDB200_LIB") or os.path.join(_HERE, "libfeddb200.so")

OK, ELOGIC, ERUNTIME = 0, -1, -2
SCATTER_ATOMIC, SCATTER_COLOURED, SCATTER_GATHER = 0, 1, 2
BLOCK_SCALAR, BLOCK_DIAG, BLOCK_FULL = 0, 1, 2


class LogicError(Exception):
    """std::logic_error of the reference (TEUCHOS_TEST_FOR_EXCEPTION(..., std::logic_error, ...))."""


class EngineRuntimeError(RuntimeError):
    """std::runtime_error of the reference / CUDA failures."""


_vp = C.c_void_p
_i64 = C.c_int64
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/feddb200.h declares
SIGNATURES = {
    "feddb200_last_error": (C.c_char_p, []),
    "feddb200_device_count": (C.c_int, []),
    "feddb200_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "feddb200_destroy": (None, [_vp]),
    "feddb200_set_stream": (C.c_int, [_vp, _vp]),
    "feddb200_use_own_stream": (C.c_int, [_vp]),
    "feddb200_set_scatter_mode": (C.c_int, [_vp, C.c_int]),
    "feddb200_get_scatter_mode": (C.c_int, [_vp]),
    "feddb200_set_row_phase": (C.c_int, [_vp, C.c_int]),
    "feddb200_synchronize": (C.c_int, [_vp]),
    "feddb200_launch_count": (_i64, [_vp]),
    "feddb200_dev_alloc": (C.c_int, [_vp, C.POINTER(_vp), _i64]),
    "feddb200_dev_free": (C.c_int, [_vp, _vp]),
    "feddb200_copy_h2d": (C.c_int, [_vp, _vp, _vp, _i64]),
    "feddb200_copy_d2h": (C.c_int, [_vp, _vp, _vp, _i64]),
    "feddb200_bind_host_numa": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "feddb200_host_alloc": (C.c_int, [_vp, C.POINTER(_vp), _i64]),
    "feddb200_host_free": (C.c_int, [_vp, _vp]),
    "feddb200_mesh_upload": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int, _i64, _vp, _i64, _vp]),
    "feddb200_mesh_update_coords": (C.c_int, [_vp, _vp, _vp]),
    "feddb200_mesh_free": (None, [_vp]),
    "feddb200_pattern_build": (C.c_int, [_vp, C.POINTER(_vp), _vp, _vp, _i64, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
    "feddb200_pat_free": (None, [_vp]),
    "feddb200_pattern_info": (C.c_int, [_vp, _i64p, _i64p, _i64p, _i64p, _i64p, _i32p, _i32p]),
    "feddb200_pattern_get_nodes": (C.c_int, [_vp, _vp, _vp, _vp]),
    "feddb200_pattern_nnz": (_i64, [_vp, C.c_int, C.c_int, C.c_int]),
    "feddb200_pattern_nnz_owned": (_i64, [_vp, C.c_int, C.c_int, C.c_int]),
    "feddb200_pattern_expand": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "feddb200_assemble_laplace_d": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "feddb200_assemble_mass_d": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "feddb200_assemble_mass": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "feddb200_assemble_bdstab_d": (C.c_int, [_vp, _vp, _vp]),
    "feddb200_stress_quadrature": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), _vp, _vp]),
    "feddb200_assemble_stress_d": (C.c_int, [_vp, _vp, C.c_double, _vp, _vp]),
    "feddb200_assemble_stress": (C.c_int, [_vp, _vp, C.c_double, _vp, _i64, _vp]),
    "feddb200_assemble_bdstab": (C.c_int, [_vp, _vp, _vp]),
    "feddb200_assemble_linelas_d": (C.c_int, [_vp, _vp, C.c_double, C.c_double, _vp]),
    "feddb200_assemble_advection_d": (C.c_int, [_vp, _vp, _vp, _vp]),
    "feddb200_assemble_advection_in_u_d": (C.c_int, [_vp, _vp, _vp, _vp]),
    "feddb200_assemble_div_divT_d": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "feddb200_assemble_ns_jacobian_d": (C.c_int, [_vp, _vp, C.c_double, C.c_double, _vp, C.c_int, _vp]),
    "feddb200_assemble_laplace": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "feddb200_assemble_linelas": (C.c_int, [_vp, _vp, C.c_double, C.c_double, _vp]),
    "feddb200_assemble_advection": (C.c_int, [_vp, _vp, _vp, _vp]),
    "feddb200_assemble_advection_in_u": (C.c_int, [_vp, _vp, _vp, _vp]),
    "feddb200_assemble_div_divT": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "feddb200_assemble_ns_jacobian": (C.c_int, [_vp, _vp, C.c_double, C.c_double, _vp, C.c_int, _vp]),
    "feddb200_unpack_add_d": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "feddb200_ipc_alloc": (C.c_int, [_vp, _i64, C.POINTER(C.c_void_p), _vp]),
    "feddb200_ipc_open": (C.c_int, [_vp, _vp, C.POINTER(C.c_void_p)]),
    "feddb200_ipc_close": (C.c_int, [_vp, _vp]),
    "feddb200_set_ghost_targets": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "feddb200_assemble_rhs_d": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "feddb200_assemble_rhs": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "feddb200_set_dirichlet_rows_d": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "feddb200_set_dirichlet_rhs_d": (C.c_int, [_vp, _i64, C.c_int, _vp, _vp, _vp]),
    "feddb200_scale_d": (C.c_int, [_vp, _vp, _i64, C.c_double]),
    "feddb200_csr_add_symbolic_d": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64)]),
    "feddb200_csr_add_numeric_d": (C.c_int, [_vp, _i64, C.c_double, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "feddb200_block_merge_symbolic_d": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.POINTER(C.c_int64)]),
    "feddb200_block_merge_numeric_d": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load libfeddb200.so and type every entry point.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineRuntimeError(
                f"{LIB_PATH} not found: build it with `python -m feddlib_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc == OK:
        return
    msg = load().feddb200_last_error().decode(errors="replace")
    if rc == ELOGIC:
        raise LogicError(msg)
    raise EngineRuntimeError(msg)


def ptr(a):
    """Host pointer of a contiguous numpy array, device pointer of a CUDA torch tensor, or an int."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):             # torch tensor
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))
