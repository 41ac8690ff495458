"""ctypes loader for libpacmann_host.so (the C++ mirror of the reference's Go host code)."""
import ctypes as C
import os

from . import cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpacmann_host.so")
_lib = None

vp, u64, i64, u64p, i64p, u32p, f32p, i32p = (C.c_void_p, C.c_uint64, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_int64),
                                               C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_int32))
SIG = {
    "pmh_last_error": (C.c_char_p, []),
    "pmh_mix64": (u64, [u64, u64]),
    "pmh_derive_key": (None, [u64, u64, u64, u64, vp]),
    "pmh_prf": (u64, [vp, u64, u64]),
    "pmh_pir_new": (vp, [u64, u64, vp, u64, C.c_int]),
    "pmh_pir_free": (None, [vp]),
    "pmh_pir_set_seeds": (None, [vp, C.c_int, u64, u64, u64]),
    "pmh_pir_preprocessing": (C.c_int, [vp, C.c_int]),
    "pmh_pir_dummy_preprocessing": (C.c_int, [vp, C.c_int]),
    "pmh_pir_query": (C.c_int, [vp, C.c_int, u64, C.c_int, vp]),
    "pmh_pir_private_query": (C.c_int, [vp, C.c_int, vp, vp]),
    "pmh_pir_nonprivate_query": (C.c_int, [vp, C.c_int, u64, vp]),
    "pmh_pir_get": (u64, [vp, C.c_int, C.c_int]),
    "pmh_pir_table": (u64p, [vp, C.c_int, C.c_int]),
    "pmh_pir_long_key": (u32p, [vp, C.c_int]),
    "pmh_pir_local_storage": (C.c_double, [vp, C.c_int]),
    "pmh_pir_comm_cost": (C.c_double, [vp, C.c_int]),
    "pmh_batch_new": (vp, [u64, u64, u64, vp, u64, u64, C.c_int]),
    "pmh_batch_free": (None, [vp]),
    "pmh_batch_set_seeds": (None, [vp, u64, u64]),
    "pmh_batch_preprocessing": (C.c_int, [vp]),
    "pmh_batch_dummy_preprocessing": (C.c_int, [vp]),
    "pmh_batch_query": (C.c_int, [vp, vp, u64, vp]),
    "pmh_batch_sub": (vp, [vp, u64]),
    "pmh_batch_get": (u64, [vp, C.c_int]),
    "pmh_batch_enable_resident": (C.c_int, [vp]),
    "pmh_batch_local_storage": (C.c_double, [vp]),
    "pmh_batch_prep_time": (C.c_double, [vp]),
    "pmh_batch_prep_total": (C.c_double, [vp]),
    "pmh_batch_prep_count": (C.c_uint64, [vp]),
    "pmh_l2dist": (C.c_float, [vp, vp, u64, C.c_int]),
    "pmh_frontend_basic": (vp, [i64, i64, i64, vp, vp]),
    "pmh_frontend_pir": (vp, [i64, i64, i64, vp, vp, C.c_int, C.c_int, u64, C.c_int, C.c_int]),
    "pmh_frontend_pir_shared": (vp, [vp, u64, C.c_int]),
    "pmh_frontend_set_group_lanes": (C.c_int, [vp, C.c_uint32]),
    "pmh_frontend_pir_lane": (vp, [vp, u64, C.c_uint32]),
    "pmh_search_knn_lockstep": (C.c_int, [vp, i64, vp, i64, i64, i64, i64, C.c_int, vp, vp]),
    "pmh_device_search_stats": (None, [vp]),
    "pmh_robust_prune_batch": (C.c_int, [vp, i64, i64, vp, i64, vp, i64, i64, C.c_float, C.c_int, vp, vp]),
    "pmh_selftest": (C.c_int, [C.c_int]),
    "pmh_frontend_free": (None, [vp]),
    "pmh_frontend_preprocess": (C.c_int, [vp]),
    "pmh_frontend_start_ids": (i64, [vp, vp, i64]),
    "pmh_frontend_set_start_ids": (None, [vp, vp, i64]),
    "pmh_frontend_set_rand_seed": (None, [vp, u64]),
    "pmh_frontend_search_knn": (C.c_int, [vp, vp, i64, i64, i64, i64, C.c_int, vp, vp]),
    "pmh_frontend_pir_handle": (vp, [vp]),
    "pmh_frontend_stat": (i64, [vp, C.c_int]),
}


class HostError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        cabi.lib()  # libpacmann_cuda.so first (the host mirror links against it)
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run make -C pacmann_b200/csrc (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIG.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc == -100:
        raise HostError(lib().pmh_last_error().decode(errors="replace"))
    return rc
