"""ctypes binding of ``libfemder_b200.so`` (declared in ``include/femder_b200.h``).

The library is the product: there is no Python / CPU fallback.  If it is missing or no B200 is usable,
the first call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDB_LIB_PATH") or os.path.join(_HERE, "libfemder_b200.so")   # FDB_LIB_PATH: A/B timing of two builds

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class SweepParams(C.Structure):
    _fields_ = [("n_freq", C.c_int64), ("omega", C.c_void_p), ("mu", C.c_void_p), ("n_src", C.c_int64),
                ("src_nodes", C.c_void_p), ("src_q", C.c_void_p), ("n_src_sets", C.c_int64), ("src_set_ptr", C.c_void_p),
                ("tol", C.c_double), ("max_iter", C.c_int32),
                ("batch", C.c_int32), ("check_every", C.c_int32), ("warm_start", C.c_int32), ("precond", C.c_int32),
                ("arithmetic", C.c_int32), ("n_rhs_vec", C.c_int64), ("rhs_ptr", C.c_void_p), ("rhs_nodes", C.c_void_p),
                ("rhs_vals", C.c_void_p), ("rhs_coef", C.c_void_p), ("n_rcv", C.c_int64), ("rcv_nodes", C.c_void_p), ("rcv_w", C.c_void_p),
                ("modal_delta_hz", C.c_double), ("shared_next", C.c_void_p)]


class SweepResult(C.Structure):
    _fields_ = [("pR", C.c_void_p), ("pN", C.c_void_p), ("iters", C.c_void_p), ("relres", C.c_void_p),
                ("seconds", C.c_double), ("n_spmv", C.c_int64), ("n_kernels", C.c_int64),
                ("n_unconverged", C.c_int32), ("real_arithmetic", C.c_int32), ("n_resumed", C.c_int32), ("precond_used", C.c_int32),
                ("solved", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/femder_b200.h declares
SIGNATURES = {
    "fdb_last_error": (C.c_char_p, []),
    "fdb_abi_version": (C.c_int, []),
    "fdb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fdb_system_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "fdb_system_destroy": (C.c_int, [C.c_void_p]),
    "fdb_synchronize": (C.c_int, [C.c_void_p]),
    "fdb_set_volume_mesh": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "fdb_pattern_nnz": (C.c_int, [C.c_void_p, c_i64p]),
    "fdb_pattern_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_elem2nnz_get": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fdb_set_pattern": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "fdb_assemble_volume": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "fdb_get_HQ": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_set_HQ": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_get_element_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_set_surface_mesh": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "fdb_heron_areas": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fdb_assemble_surface": (C.c_int, [C.c_void_p]),
    "fdb_surface_nnz": (C.c_int, [C.c_void_p, c_i64p]),
    "fdb_surface_pattern_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_surface_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_set_surface": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_locate_points": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "fdb_sweep": (C.c_int, [C.c_void_p, C.POINTER(SweepParams), C.POINTER(SweepResult)]),
    "fdb_sizeof_sweep_params": (C.c_int, []),
    "fdb_sizeof_sweep_result": (C.c_int, []),
    "fdb_modal_compute": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int)]),
    "fdb_modal_estimate": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
    "fdb_modal_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), c_f64p, C.c_void_p]),
    "fdb_set_nodes": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "fdb_modal_compression": (C.c_int, [C.c_void_p, C.c_int]),
    "fdb_comm_unique_id": (C.c_int, [C.c_char_p]),
    "fdb_comm_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "fdb_comm_free": (C.c_int, [C.c_void_p]),
    "fdb_modal_set": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "fdb_modal_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "fdb_modal_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_modal_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "fdb_spmv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdb_spmv_bench": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_f64p, c_i64p]),
    "fdb_spmv_enqueue": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_i64p]),
    "fdb_iteration_enqueue": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_i64p]),
}

_lib = None


class FemderB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built: ``python -m femder_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise FemderB200Error(f"{LIB_PATH} is missing - build it with `python -m femder_b200.build`; "
                              "femder_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.fdb_sizeof_sweep_params() != C.sizeof(SweepParams) or lib.fdb_sizeof_sweep_result() != C.sizeof(SweepResult):
        raise FemderB200Error("libfemder_b200.so and femder_b200/_lib.py disagree on the sweep struct layout - rebuild the library")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().fdb_last_error()
        raise FemderB200Error(msg.decode("utf-8", "replace") if msg else f"femder_b200 call failed ({status})")


def ptr(a):
    """Address of a C-contiguous numpy array, an int device pointer, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor (host or device): PyTorch owns the buffer
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take the address of {type(a)}")


def device_count():
    n = C.c_int(0)
    check(load().fdb_device_count(C.byref(n)))
    return n.value
