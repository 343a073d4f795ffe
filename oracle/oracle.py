"""ctypes loader for oracle/_build/liboracle.so (TEST INFRASTRUCTURE ONLY — see __init__)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile oracle.cpp with the committed Makefile (g++ -O3 -fopenmp -ffp-contract=off)."""
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _IcpRes(C.Structure):
    _fields_ = [
        ("t", C.c_float * 3),
        ("q", C.c_float * 4),
        ("mse", C.c_float),
        ("iterations", C.c_uint64),
        ("converged", C.c_int32),
        ("n_corr", C.c_uint64),
    ]


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(_SO)
    lib.orc_max_threads.restype = C.c_int
    lib.orc_kdtree_new.restype = C.c_void_p
    lib.orc_kdtree_new.argtypes = [_f32p, C.c_uint64]
    lib.orc_kdtree_free.argtypes = [C.c_void_p]
    lib.orc_kdtree_knn.restype = C.c_uint64
    lib.orc_kdtree_knn.argtypes = [C.c_void_p, _f32p, C.c_uint64, _u64p, _f32p, _f32p]
    lib.orc_kdtree_knn_batch.argtypes = [C.c_void_p, _f32p, C.c_uint64, C.c_uint64, _u64p, _f32p,
                                         _u64p, C.c_int]
    lib.orc_kdtree_radius.restype = C.c_uint64
    lib.orc_kdtree_radius.argtypes = [C.c_void_p, _f32p, C.c_float, _u64p, _f32p, C.c_uint64]
    lib.orc_brute_knn_batch.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64, C.c_uint64, _u64p,
                                        _f32p, C.c_int]
    lib.orc_k_nearest_neighbors.argtypes = [_f32p, C.c_uint64, C.c_uint64, _u64p, _f32p, _u64p,
                                            C.c_int]
    lib.orc_estimate_normals.restype = C.c_int
    lib.orc_estimate_normals.argtypes = [_f32p, C.c_uint64, C.c_uint64, C.c_float, C.c_int, _f32p,
                                         _f32p, C.c_int]
    lib.orc_normals_f64.argtypes = [_f32p, C.c_uint64, C.c_uint64, _f64p, _f64p, C.c_int]
    lib.orc_symmetric_eigen3.argtypes = [_f32p, _f32p, _f32p]
    lib.orc_icp_point_to_plane.restype = C.c_int
    lib.orc_icp_point_to_plane.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_uint64,
                                           _f32p, C.c_uint64, C.c_float, C.c_float,
                                           C.POINTER(_IcpRes), _u64p, C.c_int]
    lib.orc_icp_point_to_point.restype = C.c_int
    lib.orc_icp_point_to_point.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_uint64,
                                           C.c_float, C.c_float, C.POINTER(_IcpRes), _u64p, C.c_int]
    lib.orc_iso_apply.argtypes = [_f32p, _f32p, _f32p]
    lib.orc_iso_mul.argtypes = [_f32p, _f32p, _f32p]
    lib.orc_solve6.restype = C.c_int
    lib.orc_solve6.argtypes = [_f32p, _f32p, _f32p]
    lib.orc_gicp.restype = C.c_int
    lib.orc_gicp.argtypes = [_f32p, C.c_uint64, _f32p, C.c_uint64, _f32p, C.c_uint64, C.c_float,
                             C.c_float, C.c_uint64, C.POINTER(_IcpRes), _u64p,
                             C.POINTER(C.c_int), C.c_int]
    lib.orc_gicp_covariances.argtypes = [_f32p, C.c_uint64, C.c_uint64, _f32p, C.c_int]
    _lib = lib
    return lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a, t):
    return a.ctypes.data_as(t)


def max_threads() -> int:
    return int(_load().orc_max_threads())


class OracleKdTree:
    """KdTree (nearest_neighbor.rs:29-299) incl. Rust BinaryHeap tie behaviour."""

    def __init__(self, points):
        self.points = _f32(points, (-1, 3))
        self.n = self.points.shape[0]
        self._h = _load().orc_kdtree_new(_p(self.points, _f32p), self.n)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_kdtree_free(self._h)
            self._h = None

    def find_k_nearest(self, query, k: int):
        """-> (idx[u64], dist[f32] = sqrt(d2), d2[f32]) ascending, len = min(k, n)."""
        q = _f32(query, (3,))
        k = int(k)
        idx = np.empty(max(k, 1), np.uint64)
        dist = np.empty(max(k, 1), np.float32)
        d2 = np.empty(max(k, 1), np.float32)
        c = _load().orc_kdtree_knn(self._h, _p(q, _f32p), k, _p(idx, _u64p), _p(dist, _f32p),
                                   _p(d2, _f32p))
        return idx[:c].copy(), dist[:c].copy(), d2[:c].copy()

    def knn_batch(self, queries, k: int, threads: int = 0):
        """-> (idx[nq,k] u64 (pad = 2^64-1), d2[nq,k] f32 (pad = inf), count[nq])."""
        q = _f32(queries, (-1, 3))
        nq = q.shape[0]
        idx = np.empty((nq, k), np.uint64)
        d2 = np.empty((nq, k), np.float32)
        cnt = np.empty(nq, np.uint64)
        _load().orc_kdtree_knn_batch(self._h, _p(q, _f32p), nq, k, _p(idx, _u64p), _p(d2, _f32p),
                                     _p(cnt, _u64p), threads)
        return idx, d2, cnt

    def find_radius_neighbors(self, query, radius: float):
        q = _f32(query, (3,))
        cap = max(self.n, 1)
        idx = np.empty(cap, np.uint64)
        dist = np.empty(cap, np.float32)
        c = _load().orc_kdtree_radius(self._h, _p(q, _f32p), float(radius), _p(idx, _u64p),
                                      _p(dist, _f32p), cap)
        return idx[:c].copy(), dist[:c].copy()


def brute_knn(points, queries, k: int, threads: int = 0):
    """Canonical kNN: the k smallest under ascending (d2, index). -> (idx u64, d2 f32)."""
    pts = _f32(points, (-1, 3))
    q = _f32(queries, (-1, 3))
    nq = q.shape[0]
    idx = np.empty((nq, k), np.uint64)
    d2 = np.empty((nq, k), np.float32)
    _load().orc_brute_knn_batch(_p(pts, _f32p), pts.shape[0], _p(q, _f32p), nq, k,
                                _p(idx, _u64p), _p(d2, _f32p), threads)
    return idx, d2


def k_nearest_neighbors(points, k: int, threads: int = 1):
    """PointCloudNeighbors::k_nearest_neighbors (point_cloud_ops.rs:80-105).
    -> (idx[n,k] u64, dist[n,k] f32, count[n]); empty arrays if n == 0 or k == 0."""
    pts = _f32(points, (-1, 3))
    n = pts.shape[0]
    if n == 0 or k == 0:
        return (np.empty((0, k), np.uint64), np.empty((0, k), np.float32),
                np.empty(0, np.uint64))
    idx = np.empty((n, k), np.uint64)
    dist = np.empty((n, k), np.float32)
    cnt = np.empty(n, np.uint64)
    _load().orc_k_nearest_neighbors(_p(pts, _f32p), n, k, _p(idx, _u64p), _p(dist, _f32p),
                                    _p(cnt, _u64p), threads)
    return idx, dist, cnt


class InvalidData(ValueError):
    pass


class AlgorithmError(RuntimeError):
    pass


def estimate_normals(points, k: int, radius=None, consistent_orientation: bool = True,
                     viewpoint=None, threads: int = 0):
    """estimate_normals_with_config (normals.rs:257-357). -> [n,6] f32 (position, normal)."""
    pts = _f32(points, (-1, 3))
    n = pts.shape[0]
    out = np.zeros((n, 6), np.float32)
    vp = None if viewpoint is None else _f32(viewpoint, (3,))
    st = _load().orc_estimate_normals(
        _p(pts, _f32p), n, int(k), -1.0 if radius is None else float(radius),
        int(bool(consistent_orientation)), None if vp is None else _p(vp, _f32p),
        _p(out, _f32p), threads)
    if st == 1:
        raise InvalidData("k_neighbors must be at least 3")
    return out


def normals_f64(points, k: int, threads: int = 0):
    """f64 covariance + f64 Jacobi on the reference's neighbourhoods.
    -> (normal[n,3] f64 unoriented, relgap[n] = (l1-l0)/l2)."""
    pts = _f32(points, (-1, 3))
    n = pts.shape[0]
    nrm = np.zeros((n, 3), np.float64)
    gap = np.zeros(n, np.float64)
    _load().orc_normals_f64(_p(pts, _f32p), n, int(k), _p(nrm, _f64p), _p(gap, _f64p), threads)
    return nrm, gap


def symmetric_eigen3(m):
    """nalgebra Matrix3::symmetric_eigen restatement. -> (vals[3], vecs[3,3] columns)."""
    m = _f32(m, (3, 3))
    val = np.empty(3, np.float32)
    vec = np.empty(9, np.float32)
    _load().orc_symmetric_eigen3(_p(m, _f32p), _p(val, _f32p), _p(vec, _f32p))
    return val, vec.reshape(3, 3).T.copy()  # columns = eigenvectors


@dataclass
class IcpResult:
    translation: np.ndarray  # [3] f32
    rotation: np.ndarray     # [4] f32 quaternion (i, j, k, w)
    mse: float
    iterations: int
    converged: bool
    correspondences: np.ndarray  # [m,2] u64 (src, tgt)


def icp_point_to_plane(source, target, target_normals, init=None, max_iters: int = 30,
                       max_dist=None, conv: float = 1e-6, threads: int = 0) -> IcpResult:
    """icp_point_to_plane_detailed (registration.rs:508-602).
    init = [tx,ty,tz, qi,qj,qk,qw] (identity if None)."""
    src = _f32(source, (-1, 3))
    tgt = _f32(target, (-1, 3))
    nrm = _f32(target_normals, (-1, 3))
    init7 = _f32([0, 0, 0, 0, 0, 0, 1] if init is None else init, (7,))
    res = _IcpRes()
    pairs = np.zeros((max(src.shape[0], 1), 2), np.uint64)
    st = _load().orc_icp_point_to_plane(
        _p(src, _f32p), src.shape[0], _p(tgt, _f32p), tgt.shape[0], _p(nrm, _f32p), nrm.shape[0],
        _p(init7, _f32p), int(max_iters), -1.0 if max_dist is None else float(max_dist),
        float(conv), C.byref(res), _p(pairs, _u64p), threads)
    if st == 1:
        raise InvalidData("invalid ICP arguments")
    if st == 2:
        raise AlgorithmError("ICP numerical failure")
    return IcpResult(np.array(res.t[:], np.float32), np.array(res.q[:], np.float32),
                     float(res.mse), int(res.iterations), bool(res.converged),
                     pairs[: res.n_corr].copy())


def last_icp_mse_f64() -> float:
    """Diagnostic: the last icp_point_to_plane call's mse with the SAME f32 squared residuals
    accumulated in f64 (how far the reference's sequential f32 sum is from the exact mean)."""
    lib = _load()
    lib.orc_last_icp_mse_f64.restype = C.c_double
    return float(lib.orc_last_icp_mse_f64())


def icp_point_to_point(source, target, init=None, max_iters: int = 50, conv: float = 1e-6,
                       max_dist=None, threads: int = 0, validate_conv: bool = True) -> IcpResult:
    """icp_point_to_point (registration.rs:644-680) -> icp_detailed (:258-370)."""
    src = _f32(source, (-1, 3))
    tgt = _f32(target, (-1, 3))
    if validate_conv and src.shape[0] and tgt.shape[0] and max_iters and conv <= 0.0:
        raise InvalidData("Convergence threshold must be positive")  # :665-669
    init7 = _f32([0, 0, 0, 0, 0, 0, 1] if init is None else init, (7,))
    res = _IcpRes()
    pairs = np.zeros((max(src.shape[0], 1), 2), np.uint64)
    st = _load().orc_icp_point_to_point(
        _p(src, _f32p), src.shape[0], _p(tgt, _f32p), tgt.shape[0], _p(init7, _f32p),
        int(max_iters), -1.0 if max_dist is None else float(max_dist), float(conv),
        C.byref(res), _p(pairs, _u64p), threads)
    if st == 1:
        raise InvalidData("invalid ICP arguments")
    if st == 2:
        raise AlgorithmError("Insufficient correspondences found")
    return IcpResult(np.array(res.t[:], np.float32), np.array(res.q[:], np.float32),
                     float(res.mse), int(res.iterations), bool(res.converged),
                     pairs[: res.n_corr].copy())


def iso_apply(iso7, p):
    out = np.empty(3, np.float32)
    _load().orc_iso_apply(_p(_f32(iso7, (7,)), _f32p), _p(_f32(p, (3,)), _f32p), _p(out, _f32p))
    return out


def iso_mul(a7, b7):
    out = np.empty(7, np.float32)
    _load().orc_iso_mul(_p(_f32(a7, (7,)), _f32p), _p(_f32(b7, (7,)), _f32p), _p(out, _f32p))
    return out


def solve6(ata, atb):
    """-> (x[6], path) path: 0 = cholesky, 1 = LU fallback, 2 = singular."""
    x = np.empty(6, np.float32)
    st = _load().orc_solve6(_p(_f32(ata, (36,)), _f32p), _p(_f32(atb, (6,)), _f32p), _p(x, _f32p))
    return x, st


def gicp_covariances(points, k: int = 20, threads: int = 0) -> np.ndarray:
    """compute_covariances (gicp.rs:58-95) -> [n, 3, 3] f32."""
    pts = _f32(points, (-1, 3))
    out = np.empty((pts.shape[0], 3, 3), np.float32)
    _load().orc_gicp_covariances(_p(pts, _f32p), pts.shape[0], int(k), _p(out, _f32p), threads)
    return out


_GICP_INVALID = {1: "GICP: source or target point cloud is empty",
                 2: "GICP: max_iterations must be > 0",
                 3: "GICP: clouds must have at least k points for reliable covariance estimation",
                 4: "GICP: source point cloud appears to be coplanar or collinear",
                 5: "GICP: target point cloud appears to be coplanar or collinear"}


def gicp(source, target, init=None, max_iterations: int = 50,
         max_correspondence_distance: float = 1.0, convergence_threshold: float = 1e-6,
         k_correspondences: int = 20, threads: int = 0) -> IcpResult:
    """gicp (gicp.rs:117-312) with GicpConfig's defaults (:37-46)."""
    src = _f32(source, (-1, 3))
    tgt = _f32(target, (-1, 3))
    init7 = _f32([0, 0, 0, 0, 0, 0, 1] if init is None else init, (7,))
    res = _IcpRes()
    why = C.c_int(0)
    pairs = np.zeros((max(src.shape[0], 1), 2), np.uint64)
    st = _load().orc_gicp(_p(src, _f32p), src.shape[0], _p(tgt, _f32p), tgt.shape[0],
                          _p(init7, _f32p), int(max_iterations), float(max_correspondence_distance),
                          float(convergence_threshold), int(k_correspondences), C.byref(res),
                          _p(pairs, _u64p), C.byref(why), threads)
    if st == 1:
        raise InvalidData(_GICP_INVALID.get(why.value, "GICP: invalid arguments"))
    if st == 2:
        raise AlgorithmError("GICP: insufficient correspondences (need >= 6)")
    if st == 3:
        raise AlgorithmError("GICP: Gauss-Newton system is ill-conditioned")
    return IcpResult(np.array(res.t[:], np.float32), np.array(res.q[:], np.float32),
                     float(res.mse), int(res.iterations), bool(res.converged),
                     pairs[: res.n_corr].copy())
