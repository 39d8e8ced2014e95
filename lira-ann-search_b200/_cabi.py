"""ctypes binding of liblira_b200.so (include/lira_b200.h). Thin: argument marshalling only.

There is no CPU fallback anywhere in this package: if the shared library is missing, or no CUDA
device is visible when a compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblira_b200.so")

METRIC_L2, METRIC_IP = 0, 1
SELECT_GT, SELECT_GE_ARGMAX, SELECT_TOPN = 0, 1, 2

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_vp = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int

# name -> (restype, argtypes); must list every symbol include/lira_b200.h declares
SIGNATURES = {
    "lira_last_error": (ctypes.c_char_p, []),
    "lira_version": (c_int, []),
    "lira_device_count": (c_int, []),
    "lira_index_create": (c_int, [c_f32p, c_i64, c_int, c_i64p, c_i32p, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "lira_index_create_from_assign": (c_int, [c_f32p, c_i64, c_int, c_i32p, c_int, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "lira_index_create_dev": (c_int, [c_vp, c_i64, c_int, c_i64p, c_vp, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "lira_index_free": (c_int, [c_vp]),
    "lira_index_ntotal": (c_i64, [c_vp, c_int]),
    "lira_index_nlist": (c_int, [c_vp]),
    "lira_index_dim": (c_int, [c_vp]),
    "lira_index_list_search": (c_int, [c_vp, c_int, c_f32p, c_i64, c_int, c_f32p, c_i64p]),
    "lira_scan_all_pairs": (c_int, [c_vp, c_f32p, c_i64, c_int, c_i64p, c_i64p]),
    "lira_search": (c_int, [c_vp, c_f32p, c_i64, c_i64p, c_i32p, c_int, c_int, c_f32p, c_i64p, c_i64p]),
    "lira_search_dev": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "lira_model_create": (c_int, [c_f32p, c_f32p, c_f32p, c_int, c_int, ctypes.POINTER(c_f32p), c_int, ctypes.POINTER(c_vp)]),
    "lira_model_free": (c_int, [c_vp]),
    "lira_model_set_use_tensor_cores": (c_int, [c_vp, c_int]),
    "lira_centroid_features": (c_int, [c_f32p, c_i64, c_f32p, c_int, c_int, c_f32p, c_f32p, c_int, c_f32p]),
    "lira_model_scores": (c_int, [c_vp, c_f32p, c_i64, c_f32p, c_f32p]),
    "lira_probe_search": (c_int, [c_vp, c_vp, c_f32p, c_i64, c_int, ctypes.c_double, c_int, c_int, c_f32p, c_i64p, c_i32p, c_i64p]),
    "lira_probe_search_dev": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_int, ctypes.c_double, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lira_probe_search_enqueue_dev": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_int, ctypes.c_double, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lira_index_finish": (c_int, [c_vp]),
    "lira_probe_search_submit": (c_int, [c_vp, c_vp, c_f32p, c_i64, c_int, ctypes.c_double, c_int, c_int, c_int]),
    "lira_probe_search_wait": (c_int, [c_vp, c_int, c_f32p, c_i64p, c_i32p, c_i64p]),
    "lira_select_search_dev": (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_int, ctypes.c_double, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "lira_knn": (c_int, [c_f32p, c_i64, c_f32p, c_i64, c_int, c_int, c_int, c_int, c_f32p, c_i64p]),
    "lira_knn_create": (c_int, [c_f32p, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "lira_knn_search": (c_int, [c_vp, c_f32p, c_i64, c_int, c_f32p, c_i64p]),
    "lira_knn_create_dev": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_int, ctypes.POINTER(c_vp)]),
    "lira_knn_search_dev": (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp]),
    "lira_knn_free": (c_int, [c_vp]),
    "lira_knn_ntotal": (c_i64, [c_vp]),
    "lira_knn_set_use_tensor_cores": (c_int, [c_vp, c_int]),
    "lira_knn_last_path": (c_int, [c_vp]),
    "lira_knn_last_redo": (c_int, [c_vp]),
    "lira_knn_last_scan_kind": (c_int, [c_vp]),
    "lira_knn_ivf": (c_int, [c_f32p, c_i64, c_int, c_int, c_int, c_int, ctypes.c_uint64, c_int, c_f32p, c_i64p]),
    "lira_kmeans_train": (c_int, [c_f32p, c_i64, c_int, c_int, c_int, ctypes.c_uint64, c_f32p, c_int, c_f32p]),
    "lira_kmeans_train_dev": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_int, c_vp]),
    "lira_centroid_features_dev": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_vp]),
    "lira_feature_stats_dev": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), c_int, c_vp]),
    "lira_mul_partition_dev": (c_int, [c_vp, c_i64, c_i64, c_int, ctypes.c_float, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_vp]),
    "lira_pack_keys_dev": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_int, c_vp]),
    "lira_merge_ranks_dev": (c_int, [c_vp, c_int, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    "lira_launch_count": (c_i64, []),
    "lira_index_last_timing": (c_int, [c_vp, c_f32p, c_f32p, c_i64p, c_i64p]),
    "lira_index_set_timing": (c_int, [c_vp, c_int]),
    "lira_index_last_scan_total_ms": (c_int, [c_vp, c_f32p]),
    "lira_index_set_use_tensor_cores": (c_int, [c_vp, c_int]),
    "lira_index_last_path": (c_int, [c_vp]),
    "lira_index_last_redo": (c_int, [c_vp]),
    "lira_index_byte_scan_eligible": (c_int, [c_vp]),
    "lira_index_last_scan_kind": (c_int, [c_vp]),
    "lira_index_tensor_core_eligible": (c_int, [c_vp]),
    "lira_index_tensor_core_mode": (c_int, [c_vp]),
}

_lib = None


class LiraError(RuntimeError):
    pass


def lib():
    """Load liblira_b200.so. Raises (never falls back) when the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LiraError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI and this table disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().lira_last_error()
        raise LiraError((msg or b"unknown error").decode("utf-8", "replace"))


def require_gpu():
    n = lib().lira_device_count()
    if n <= 0:
        raise LiraError("no CUDA device visible: liblira_b200 has no CPU fallback")
    return n


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)
