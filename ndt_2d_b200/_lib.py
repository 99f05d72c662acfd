"""ctypes binding of libndt2d_b200.so (the C ABI declared in include/ndt2d_b200.h).

There is no fallback: if the shared library is missing the import fails, and if
no CUDA device is present every create() returns NDT2D_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

# NDT2D_B200_LIB: an alternative build of the same library (kernel A/B experiments)
_LIB_PATH = Path(os.environ.get("NDT2D_B200_LIB") or
                 Path(__file__).resolve().parent / "lib" / "libndt2d_b200.so")

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_NO_MAP, ERR_SIZE, ERR_STATE = range(7)
STATUS_NAMES = {
    0: "NDT2D_OK", 1: "NDT2D_ERR_INVALID", 2: "NDT2D_ERR_NO_DEVICE", 3: "NDT2D_ERR_CUDA",
    4: "NDT2D_ERR_NO_MAP", 5: "NDT2D_ERR_SIZE", 6: "NDT2D_ERR_STATE",
}
PARTIAL_DOUBLES = 16


class Ndt2dError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = f"{where}: {STATUS_NAMES.get(status, status)}"
        if detail:
            msg += f" [{detail}]"
        super().__init__(msg)


class Params(C.Structure):
    _fields_ = [
        ("ndt_resolution", C.c_double),
        ("search_angular_resolution", C.c_double),
        ("search_angular_size", C.c_double),
        ("search_linear_resolution", C.c_double),
        ("search_linear_size", C.c_double),
        ("laser_max_beams", C.c_int),
        ("range_max", C.c_double),
        ("device", C.c_int),
        ("stream", C.c_void_p),
        ("kernel_variant", C.c_int),
        ("n_devices", C.c_int),
        ("devices", C.c_int * 16),
    ]


_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes).  Every symbol declared in include/ndt2d_b200.h.
SIGNATURES = {
    "ndt2d_version": (C.c_char_p, []),
    "ndt2d_last_error": (C.c_char_p, []),
    "ndt2d_device_count": (C.c_int, []),
    "ndt2d_default_params": (None, [C.POINTER(Params)]),
    "ndt2d_matcher_create": (C.c_int, [C.POINTER(Params), C.POINTER(_vp)]),
    "ndt2d_matcher_destroy": (C.c_int, [_vp]),
    "ndt2d_matcher_reset": (C.c_int, [_vp]),
    "ndt2d_matcher_add_scans": (C.c_int, [_vp, C.c_size_t, _dp, _u64p, _dp]),
    "ndt2d_matcher_match_scan": (C.c_int, [_vp, _dp, _dp, C.c_size_t, _dp, _ip, _dp, _dp]),
    "ndt2d_matcher_score_points": (C.c_int, [_vp, _dp, C.c_size_t, _dp, _dp]),
    "ndt2d_matcher_score_poses": (C.c_int, [_vp, _dp, C.c_size_t, _dp, C.c_size_t, _dp]),
    "ndt2d_matcher_likelihood_scan": (C.c_int, [_vp, _dp, _dp, C.c_size_t, _dp]),
    "ndt2d_matcher_match_scan_batch": (
        C.c_int, [_vp, C.c_size_t, _u64p, _dp, _u64p, _dp, _dp, _u64p, _dp, _dp, _ip, _dp, _dp]),
    "ndt2d_matcher_close_loop": (
        C.c_int, [_vp, C.c_size_t, _dp, _u64p, _dp, _u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double,
                  _dp, _dp, C.c_size_t, _u64p, _dp, _ip, _dp, _dp, C.POINTER(C.c_size_t),
                  C.POINTER(C.c_size_t)]),
    "ndt2d_matcher_search_shape": (C.c_int, [_vp, _u64p, _u64p]),
    "ndt2d_matcher_search_values": (C.c_int, [_vp, _dp, _dp]),
    "ndt2d_matcher_stage_scan": (C.c_int, [_vp, _dp, _dp, C.c_size_t]),
    "ndt2d_matcher_search_staged": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "ndt2d_matcher_search_staged_strided": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp]),
    "ndt2d_matcher_exchange_init": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_ubyte)]),
    "ndt2d_matcher_exchange_connect": (C.c_int, [_vp, C.POINTER(C.c_ubyte)]),
    "ndt2d_matcher_search_exchange": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "ndt2d_matcher_fetch_result": (C.c_int, [_vp, _dp, _ip, _dp, _dp]),
    "ndt2d_matcher_fetch_partial": (C.c_int, [_vp, _dp]),
    "ndt2d_matcher_group_info": (C.c_int, [_vp, _u64p]),
    "ndt2d_matcher_set_group_threshold": (C.c_int, [_vp, C.c_double]),
    "ndt2d_matcher_set_timing": (C.c_int, [_vp, C.c_int]),
    "ndt2d_matcher_set_tallies": (C.c_int, [_vp, C.c_int]),
    "ndt2d_matcher_group_search_stats": (C.c_int, [_vp, _dp, C.c_size_t, _u64p]),
    "ndt2d_combine_partials": (C.c_int, [_vp, _dp, C.c_size_t, _dp, _ip, _dp, _dp]),
    "ndt2d_combine_partials_host": (
        C.c_int, [_dp, C.c_size_t, _dp, C.c_size_t, _dp, C.c_size_t, _dp, _ip, _dp, _dp]),
    "ndt2d_search_lattice": (C.c_int, [C.c_double, C.c_double, _dp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ndt2d_matcher_combine_device": (C.c_int, [_vp, _vp, C.c_size_t, _dp, _ip, _dp, _dp]),
    "ndt2d_matcher_grid_info": (C.c_int, [_vp, _dp]),
    "ndt2d_matcher_dump_cells": (C.c_int, [_vp, _dp]),
    "ndt2d_matcher_dump_keys": (C.c_int, [_vp, _i32p, C.c_size_t]),
    "ndt2d_matcher_dump_scores": (C.c_int, [_vp, _dp, _dp, C.c_size_t, _dp, C.c_size_t]),
    "ndt2d_matcher_counters": (C.c_int, [_vp, _u64p]),
    "ndt2d_matcher_search_stats": (C.c_int, [_vp, _u64p]),
    "ndt2d_matcher_build_stats": (C.c_int, [_vp, _u64p]),
    "ndt2d_matcher_stream": (_vp, [_vp]),
    "ndt2d_occupancy_create": (C.c_int, [C.c_double, C.c_double, C.c_int, C.POINTER(_vp)]),
    "ndt2d_occupancy_destroy": (C.c_int, [_vp]),
    "ndt2d_occupancy_render": (C.c_int, [_vp, C.c_size_t, _dp, _u64p, _dp, _dp]),
    "ndt2d_occupancy_fetch": (C.c_int, [_vp, C.POINTER(C.c_int8), C.c_size_t]),
    "ndt2d_laser_to_points": (
        C.c_int, [C.c_int, C.POINTER(C.c_float), C.c_size_t, C.c_float, C.c_float, C.c_double, _dp, _dp,
                  C.c_int, _dp, C.POINTER(C.c_size_t)]),
    "ndt2d_find_nearest": (
        C.c_int, [C.c_int, _dp, C.c_size_t, C.c_int64, _dp, C.c_double, _u64p, _dp, C.c_size_t,
                  C.POINTER(C.c_size_t)]),
    "ndt2d_probe_gather": (C.c_int, [C.c_int, C.c_size_t, _dp]),
    "ndt2d_probe_copy": (C.c_int, [C.c_int, C.c_size_t, _dp]),
    "ndt2d_probe_div_by_count": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, _u64p]),
    "ndt2d_probe_call_latency": (C.c_int, [_vp, C.c_int, C.c_size_t, _dp, _u64p, _dp, _dp, _dp,
                                           C.c_size_t, C.c_size_t, _dp]),
    "ndt2d_probe_ex2": (C.c_int, [C.c_int, _dp]),
    "ndt2d_filter_create": (C.c_int, [C.c_size_t, C.c_size_t, C.c_int, _vp, C.POINTER(_vp)]),
    "ndt2d_filter_destroy": (C.c_int, [_vp]),
    "ndt2d_filter_set_particles": (C.c_int, [_vp, _dp, _dp, C.c_size_t]),
    "ndt2d_filter_size": (C.c_int, [_vp, C.POINTER(C.c_size_t)]),
    "ndt2d_filter_get_particles": (C.c_int, [_vp, _dp, _dp]),
    "ndt2d_filter_init": (
        C.c_int, [_vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                  C.c_uint64]),
    "ndt2d_filter_update": (C.c_int, [_vp, C.c_double, C.c_double, C.c_double, _dp, C.c_uint64]),
    "ndt2d_filter_measure": (C.c_int, [_vp, _vp, _dp, C.c_size_t]),
    "ndt2d_filter_resample": (C.c_int, [_vp, C.c_double, C.c_double, _dp, C.c_size_t, C.c_uint64]),
    "ndt2d_filter_stats": (C.c_int, [_vp, _dp, _dp]),
    "ndt2d_filter_set_cov": (C.c_int, [_vp, _dp]),
    "ndt2d_filter_last_draws": (C.c_int, [_vp, _u64p]),
}


def lib_path() -> Path:
    return _LIB_PATH


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m ndt_2d_b200.build` "
            "(ndt_2d_b200 has no CPU or pure-Python fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(status: int, where: str, allow=()):
    if status != OK and status not in allow:
        raise Ndt2dError(status, where, lib.ndt2d_last_error().decode(errors="replace"))
    return status


def dptr(a):
    """double* of a C-contiguous float64 numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
    return a.ctypes.data_as(_dp)


def u64ptr(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], "need contiguous uint64"
    return a.ctypes.data_as(_u64p)


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)
