"""ctypes binding of the C ABI declared in include/hpddm_b200.h (K = double) and
include/hpddm_b200z.h (K = complex double; same entry points with the hpddm_b200z_ prefix).

This is the *only* way Python code reaches the product: through the same
extern "C" entry points a C++/MPI host program binds.  There is no CPU
fallback: if libhpddm_b200.so is missing or no CUDA device is visible the
import / context creation raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhpddm_b200.so")

HOST, DEVICE = 0, 1
PRCNDTNR = dict(NO=0, SY=1, GE=2, OS=3, OG=4)
CORRECTION = {None: -1, "none": -1, "deflated": 0, "additive": 1, "balanced": 2}


class Stats(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz_a", C.c_int64), ("nnz_factor", C.c_int64), ("factor_bytes", C.c_int64),
                ("index_bytes", C.c_int64), ("fronts", C.c_int64), ("levels", C.c_int64), ("halo", C.c_int64),
                ("nu", C.c_int64), ("symmetric", C.c_int), ("numfact_seconds", C.c_double), ("symbolic_seconds", C.c_double)]


class HpddmB200Error(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_SIGS = {
    "hpddm_b200_last_error": (C.c_char_p, []),
    "hpddm_b200_version": (C.c_char_p, []),
    "hpddm_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "hpddm_b200_debug_coarse_layout": (C.c_int, [C.c_int, _P, _P, C.POINTER(C.c_int)]),
    "hpddm_b200_debug_halo_schedule": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, C.POINTER(C.c_int)]),
    "hpddm_b200_ctx_create": (C.c_int, [C.c_int, _PP]),
    "hpddm_b200_ctx_destroy": (C.c_int, [_P]),
    "hpddm_b200_nccl_unique_id": (C.c_int, [_P]),
    "hpddm_b200_ctx_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "hpddm_b200_ctx_comm_init_host": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "hpddm_b200_ctx_transport": (C.c_int, [_P]),
    "hpddm_b200_ctx_synchronize": (C.c_int, [_P]),
    "hpddm_b200_ctx_stream": (C.c_void_p, [_P]),
    "hpddm_b200_ctx_launch_count": (C.c_int64, [_P]),
    "hpddm_b200_ctx_hostreg_count": (C.c_int64, [_P]),
    "hpddm_b200_malloc": (C.c_int, [_P, C.c_size_t, _PP]),
    "hpddm_b200_free": (C.c_int, [_P, _P]),
    "hpddm_b200_memcpy": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int, C.c_int]),
    "hpddm_b200_sub_create": (C.c_int, [_P, C.c_int, _PP]),
    "hpddm_b200_sub_destroy": (C.c_int, [_P]),
    "hpddm_b200_sub_set_matrix": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_char]),
    "hpddm_b200_sub_set_neighbors": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "hpddm_b200_sub_set_scaling": (C.c_int, [_P, _P]),
    "hpddm_b200_sub_set_grid_hint": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "hpddm_b200_multiplicity_scaling": (C.c_int, [_P, _P]),
    "hpddm_b200_sub_numfact": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_char]),
    "hpddm_b200_sub_set_vectors": (C.c_int, [_P, _P, C.c_int]),
    "hpddm_b200_sub_solve_gevp": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_char, C.c_int, C.c_double, C.c_int, _P]),
    "hpddm_b200_sub_get_vectors": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "hpddm_b200_build_coarse": (C.c_int, [_P]),
    "hpddm_b200_set_coarse": (C.c_int, [_P, _P, C.c_int]),
    "hpddm_b200_get_coarse": (C.c_int, [_P, _P, C.POINTER(C.c_int)]),
    "hpddm_b200_start": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "hpddm_b200_end": (C.c_int, [_P]),
    "hpddm_b200_apply": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "hpddm_b200_deflation": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "hpddm_b200_exchange": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "hpddm_b200_gmv": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "hpddm_b200_sub_solve": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "hpddm_b200_coarse_solve": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "hpddm_b200_dot": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int]),
    "hpddm_b200_sub_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "hpddm_b200_sub_boundary_conditions": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int)]),
    "hpddm_b200_rhs_norm": (C.c_int, [_P, _P, C.c_int, _P, C.c_int]),
    "hpddm_b200_compute_residual": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "hpddm_b200_solve": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), _P]),
    "hpddm_b200_solve_bgmres": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), _P]),
    "hpddm_b200_solve_cg": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), _P]),
    "hpddm_b200_solve_gcrodr": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), _P]),
    "hpddm_b200_solve_bgcrodr": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), _P]),
    "hpddm_b200_recycle_dim": (C.c_int, [_P]),
    "hpddm_b200_recycle_destroy": (C.c_int, [_P]),
}
# the complex instantiation exports the same set under the hpddm_b200z_ prefix, with identical
# argument shapes (scalars travel behind pointers)
_SIGS.update({k.replace("hpddm_b200_", "hpddm_b200z_", 1): v for k, v in list(_SIGS.items())})
EXPORTS = sorted(_SIGS)


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)   # hpddm_b200_allgather_fn


def lib():
    """Load libhpddm_b200.so (built in-tree by __graft_entry__.build / make)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HpddmB200Error(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, prefix="hpddm_b200"):
    if rc < 0:
        raise HpddmB200Error(f"libhpddm_b200 error {rc}: {getattr(lib(), prefix + '_last_error')().decode()}")
    return rc


class Api:
    """The entry points of one scalar type: api.apply(...) -> hpddm_b200[z]_apply(...)."""

    def __init__(self, prefix, dtype):
        self.prefix = prefix
        self.dtype = np.dtype(dtype)

    def __getattr__(self, name):
        return getattr(lib(), f"{self.prefix}_{name}")

    def check(self, rc):
        return check(rc, self.prefix)


REAL = Api("hpddm_b200", np.float64)
COMPLEX = Api("hpddm_b200z", np.complex128)


def api(dtype=np.float64):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return REAL
    if dtype == np.complex128:
        return COMPLEX
    raise HpddmB200Error(f"unsupported scalar type {dtype}: the library is built for float64 and complex128")


def ptr(a):
    """void* of a numpy array (or an int device address, or None)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def ptr_array(items):
    """void*[len(items)] from numpy arrays / raw addresses."""
    arr = (C.c_void_p * len(items))()
    for i, a in enumerate(items):
        if isinstance(a, (int, np.integer)):
            arr[i] = int(a)
        elif hasattr(a, "data_ptr"):      # torch tensor (host pinned or device)
            arr[i] = a.data_ptr()
        else:
            arr[i] = a.ctypes.data
    return arr
