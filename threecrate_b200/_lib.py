"""ctypes binding of libthreecrate_cuda.so — the same C ABI a Rust `threecrate-cuda` crate binds.

There is no CPU fallback: if the shared library is missing this module raises, and on a box
without a CUDA device `Context()` raises GpuError.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libthreecrate_cuda.so")

TC_OK, TC_INVALID_DATA, TC_ALGORITHM, TC_GPU = 0, 1, 2, 3
TC_NO_INDEX = 0xFFFFFFFF
TC_COMM_ID_BYTES = 128
TC_IPC_HANDLE_BYTES = 64


class ThreecrateError(Exception):
    """Base of the errors mirroring threecrate_core::Error (threecrate-core/src/error.rs:7-28)."""


class InvalidData(ThreecrateError, ValueError):
    """Error::InvalidData"""


class AlgorithmError(ThreecrateError, RuntimeError):
    """Error::Algorithm"""


class GpuError(ThreecrateError, RuntimeError):
    """Error::Gpu"""


class IcpResultC(C.Structure):
    _fields_ = [("transform", C.c_float * 7), ("mse", C.c_float), ("iterations", C.c_uint32),
                ("converged", C.c_int32), ("n_correspondences", C.c_uint64)]


class IndexInfoC(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("n_cells", C.c_uint64), ("dims", C.c_uint32 * 3),
                ("cell_size", C.c_float), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
                ("occupied_cells", C.c_uint32), ("max_cell_population", C.c_uint32),
                ("n_levels", C.c_uint32), ("reserved", C.c_uint32)]


TC_STATS_MAX_ITERS = 64


class StatsC(C.Structure):
    _fields_ = [("queries", C.c_uint64), ("chain_queries", C.c_uint64), ("rounds", C.c_uint64),
                ("box_splits", C.c_uint64), ("retries", C.c_uint64),
                ("candidates_staged", C.c_uint64), ("merges", C.c_uint64),
                ("icp_iterations", C.c_uint32), ("reserved", C.c_uint32),
                ("icp_mse", C.c_float * TC_STATS_MAX_ITERS),
                ("icp_valid", C.c_uint64 * TC_STATS_MAX_ITERS)]


_vp = C.c_void_p
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)

# every symbol include/threecrate_cuda.h declares: name -> (restype, argtypes)
class IcpScaleLevelC(C.Structure):
    _fields_ = [("voxel_size", C.c_float), ("max_iterations", C.c_uint32),
                ("max_correspondence_distance", C.c_float)]


SYMBOLS = {
    "tc_context_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "tc_context_destroy": (None, [_vp]),
    "tc_last_error": (C.c_char_p, [_vp]),
    "tc_context_stream": (_vp, [_vp]),
    "tc_context_synchronize": (C.c_int, [_vp]),
    "tc_launch_count": (C.c_uint64, [_vp]),
    "tc_timer_start": (C.c_int, [_vp]),
    "tc_timer_stop": (C.c_int, [_vp, _f32p]),
    "tc_version": (C.c_char_p, []),
    "tc_stats_enable": (C.c_int, [_vp, C.c_int]),
    "tc_last_stats": (C.c_int, [_vp, C.POINTER(StatsC)]),
    "tc_cloud_upload": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "tc_cloud_upload_strided": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.POINTER(_vp)]),
    "tc_cloud_from_device": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(_vp)]),
    "tc_cloud_free": (None, [_vp]),
    "tc_cloud_len": (C.c_uint64, [_vp]),
    "tc_cloud_download": (C.c_int, [_vp, _vp, _vp]),
    "tc_voxel_grid_filter": (C.c_int, [_vp, _vp, C.c_float, C.POINTER(_vp)]),
    "tc_radius_outlier_removal": (C.c_int, [_vp, _vp, C.c_float, C.c_uint32, C.POINTER(_vp)]),
    "tc_statistical_outlier_removal": (C.c_int, [_vp, _vp, C.c_uint32, C.c_float, C.c_int, _vp,
                                                 C.POINTER(_vp)]),
    "tc_index_build": (C.c_int, [_vp, _vp, C.c_uint32, C.c_float, C.POINTER(_vp)]),
    "tc_index_build_sharded": (C.c_int, [_vp, _vp, C.c_uint32, C.c_float, C.c_int, C.c_int,
                                         C.POINTER(_vp)]),
    "tc_index_free": (None, [_vp]),
    "tc_index_get_info": (C.c_int, [_vp, C.POINTER(IndexInfoC)]),
    "tc_knn": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "tc_knn_device": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "tc_radius_search": (C.c_int, [_vp, _vp, _f32p, C.c_float, _vp, _vp, C.c_uint64, _u64p]),
    "tc_estimate_normals": (C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_float, C.c_int, _vp, _vp]),
    "tc_estimate_normals_indexed": (C.c_int, [_vp, _vp, C.c_uint32, C.c_float, C.c_int, _vp, _vp]),
    "tc_estimate_normals_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_float, C.c_int, _vp,
                                             C.c_uint64, C.c_uint64, _vp]),
    "tc_icp_point_to_plane": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_uint64, _f32p,
                                        C.c_uint32, C.c_float, C.c_float, C.POINTER(IcpResultC), _vp]),
    "tc_icp_point_to_plane_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f32p, C.c_uint32, C.c_float,
                                               C.c_float, C.POINTER(IcpResultC), _vp]),
    "tc_icp_point_to_point": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _f32p, C.c_uint32,
                                        C.c_float, C.c_float, C.POINTER(IcpResultC), _vp]),
    "tc_icp_point_to_point_device": (C.c_int, [_vp, _vp, _vp, _vp, _f32p, C.c_uint32, C.c_float,
                                               C.c_float, C.POINTER(IcpResultC), _vp]),
    "tc_gicp": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _f32p, C.c_uint32, C.c_float,
                          C.c_float, C.c_uint32, C.POINTER(IcpResultC), _vp]),
    "tc_multiscale_icp_point_to_point": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _f32p,
                                                   C.POINTER(IcpScaleLevelC), C.c_uint32,
                                                   C.c_uint32, C.c_float, C.c_float,
                                                   C.POINTER(IcpResultC), _vp]),
    "tc_comm_get_unique_id": (C.c_int, [_vp, _vp]),
    "tc_comm_init_rank": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "tc_comm_destroy": (None, [_vp]),
    "tc_comm_allreduce_f64": (C.c_int, [_vp, _vp, C.c_uint64]),
    "tc_comm_peer_handle": (C.c_int, [_vp, _vp]),
    "tc_comm_peer_open": (C.c_int, [_vp, _vp]),
    "tc_dist_chunk": (None, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64),
                             C.POINTER(C.c_uint64)]),
    "tc_comm_window_handle": (C.c_int, [_vp, C.c_uint64, _vp]),
    "tc_comm_window_open": (C.c_int, [_vp, _vp]),
    "tc_estimate_normals_distributed": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_int,
                                                  _vp, _vp]),
    "tc_device_alloc": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "tc_device_free": (C.c_int, [_vp, _vp]),
    "tc_copy_to_device": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "tc_copy_to_host": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "tc_host_alloc_pinned": (C.c_int, [C.c_uint64, C.POINTER(_vp)]),
    "tc_host_free_pinned": (C.c_int, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library (fails loudly if it was not built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpuError(
            f"{LIB_PATH} is missing: build it with `python -m threecrate_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, ctx_handle=None):
    if status == TC_OK:
        return
    msg = ""
    if ctx_handle:
        raw = load().tc_last_error(ctx_handle)
        msg = raw.decode("utf-8", "replace") if raw else ""
    if status == TC_INVALID_DATA:
        raise InvalidData(msg or "invalid data")
    if status == TC_ALGORITHM:
        raise AlgorithmError(msg or "algorithm error")
    raise GpuError(msg or "GPU error (no CUDA device or CUDA failure; there is no CPU fallback)")
