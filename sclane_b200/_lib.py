"""Loader of libsclane_b200.so.  Fails loudly: there is no CPU fallback in this package."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsclane_b200.so")

_lib = None


class SclaneError(RuntimeError):
    """A global failure reported by the C ABI (bad arguments, CUDA error, no device)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"sclane_b200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "sclane_b200 has no CPU fallback and will not run without its CUDA library.")
    L = C.CDLL(LIB_PATH)
    p = C.c_void_p
    i32, i64, dbl = C.c_int32, C.c_int64, C.c_double
    err = (C.c_char_p, C.c_size_t)
    L.sclane_abi_version.restype = i32
    L.sclane_device_count.restype = i32
    L.sclane_record_size.restype = C.c_size_t
    L.sclane_launch_count.restype = i64
    L.sclane_reset_launch_count.restype = None
    L.sclane_default_opts.argtypes = [C.POINTER(_abi.SclaneOpts)]
    L.sclane_default_opts.restype = None
    L.sclane_get_profile.argtypes = [p]
    L.sclane_get_profile.restype = None
    L.sclane_release.restype = None
    L.sclane_test_dynamic.argtypes = [p, i64, i64, p, i32, p, C.POINTER(_abi.SclaneOpts), p, p, *err]
    L.sclane_test_dynamic.restype = i32
    L.sclane_test_dynamic_gee.argtypes = [p, i64, i64, p, i32, p, p, C.POINTER(_abi.SclaneOpts), p, p, *err]
    L.sclane_test_dynamic_gee.restype = i32
    L.sclane_test_dynamic_csc.argtypes = [p, p, p, i64, i64, p, i32, p, C.POINTER(_abi.SclaneOpts), p, p, *err]
    L.sclane_test_dynamic_csc.restype = i32
    L.sclane_marge_basis.argtypes = [p, i64, i64, p, i32, C.POINTER(_abi.SclaneOpts), p, p, *err]
    L.sclane_marge_basis.restype = i32
    L.sclane_test_dynamic_device.argtypes = [p, i32, i64, i64, p, i32, p, C.POINTER(_abi.SclaneOpts), p, p, *err]
    L.sclane_test_dynamic_device.restype = i32
    L.sclane_candidate_knots.argtypes = [p, i64, C.POINTER(_abi.SclaneOpts), p, p]
    L.sclane_candidate_knots.restype = i32
    L.sclane_lineage_plan.argtypes = [p, i64, C.POINTER(_abi.SclaneOpts), i32, i32, p, p, p, p, p, p, p, p, p, p, p, *err]
    L.sclane_lineage_plan.restype = i32
    L.sclane_null_stats_glm.argtypes = [p, i64, i64, p, p, i32, p, p, i32, i32, p, p, p, *err]
    L.sclane_null_stats_glm.restype = i32
    L.sclane_score_glm.argtypes = [p, p, i64, i64, p, p, i32, p, p, i32, p, i32, p, *err]
    L.sclane_score_glm.restype = i32
    L.sclane_link_preds.argtypes = [p, p, i64, p, i32, p, p, p, p, *err]
    L.sclane_link_preds.restype = i32
    L.sclane_smoothed_counts.argtypes = [p, i64, i32, i32, p, i64, p, i32, i32, i32, p, *err]
    L.sclane_smoothed_counts.restype = i32
    L.sclane_matmul.argtypes = [p, p, p, i64, i64, i64, i32, *err]
    L.sclane_matmul.restype = i32
    L.sclane_llt_inverse.argtypes = [p, p, i64, i32, *err]
    L.sclane_llt_inverse.restype = i32
    L.sclane_pinv.argtypes = [p, p, i64, i64, dbl, i32, *err]
    L.sclane_pinv.restype = i32
    if L.sclane_abi_version() != _abi.SCLANE_ABI_VERSION:
        raise ImportError("libsclane_b200.so ABI version does not match sclane_b200/_abi.py")
    if L.sclane_record_size() != C.sizeof(_abi.SclaneRecord):
        raise ImportError(f"sclane_record layout mismatch: library {L.sclane_record_size()} vs ctypes {C.sizeof(_abi.SclaneRecord)}")
    _lib = L
    return L


def check(rc: int, errbuf) -> None:
    if rc != 0:
        raise SclaneError(rc, errbuf.value.decode("utf-8", "replace"))
