"""Host-side mirror of the reference's operator interface for the kNN -> normals -> ICP path.

Same names, argument meaning and error behaviour as the Rust functions they stand in for
(paths relative to the reference checkout):

  estimate_normals / estimate_normals_with_config / estimate_normals_radius
                                   threecrate-algorithms/src/normals.rs:238-380
  KdTree.new / find_k_nearest / find_radius_neighbors
                                   threecrate-algorithms/src/nearest_neighbor.rs:37,177,254
  k_nearest_neighbors              threecrate-algorithms/src/point_cloud_ops.rs:80-105
  icp_point_to_plane(_detailed)    threecrate-algorithms/src/registration.rs:488-602

Everything routes through the C ABI in include/threecrate_cuda.h (ctypes here, `extern "C"` in
the Rust shim); numpy arrays stand in for Vec<Point3f>.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import AlgorithmError, GpuError, InvalidData, ThreecrateError  # noqa: F401

_vp = C.c_void_p


def _pts(a, name="points") -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1 and a.size % 3 == 0:
        a = a.reshape(-1, 3)
    if a.ndim != 2 or a.shape[1] != 3:
        raise InvalidData(f"{name} must have shape (N, 3)")
    return a


# --------------------------------------------------------------------------------------------
# context / device-resident objects
# --------------------------------------------------------------------------------------------
class Context:
    """One CUDA stream on one device (tc_context)."""

    def __init__(self, device: Optional[int] = None):
        lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = _vp()
        st = lib.tc_context_create(int(device), C.byref(h))
        if st != _lib.TC_OK:
            raise GpuError(f"no usable CUDA device {device} (there is no CPU fallback)")
        self.h = h
        self.device = int(device)
        self.lib = lib

    def close(self):
        if getattr(self, "h", None):
            self.lib.tc_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st):
        _lib.check(st, self.h)

    @property
    def stream(self) -> int:
        return int(self.lib.tc_context_stream(self.h) or 0)

    def synchronize(self):
        self.check(self.lib.tc_context_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.tc_launch_count(self.h))

    def timer_start(self):
        self.check(self.lib.tc_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.check(self.lib.tc_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def enable_stats(self, on: bool = True):
        """Collect per-call statistics (tc_stats_enable); off by default."""
        self.check(self.lib.tc_stats_enable(self.h, 1 if on else 0))

    def last_stats(self) -> dict:
        """tc_last_stats: counters of the last kNN / normals launch and the per-iteration history
        of the last ICP call."""
        s = _lib.StatsC()
        self.check(self.lib.tc_last_stats(self.h, C.byref(s)))
        n = int(s.icp_iterations)
        return {"queries": int(s.queries), "chain_queries": int(s.chain_queries),
                "rounds": int(s.rounds), "box_splits": int(s.box_splits),
                "retries": int(s.retries), "candidates_staged": int(s.candidates_staged),
                "merges": int(s.merges), "icp_iterations": n,
                "icp_mse": [float(s.icp_mse[i]) for i in range(min(n, _lib.TC_STATS_MAX_ITERS))],
                "icp_valid": [int(s.icp_valid[i]) for i in range(min(n, _lib.TC_STATS_MAX_ITERS))]}

    def issue_rate(self) -> float:
        """Measured warp-instruction issue ceiling of the device (debug microbenchmark)."""
        self.lib.tc_debug_issue_rate.argtypes = [_vp, C.POINTER(C.c_double)]
        self.lib.tc_debug_issue_rate.restype = C.c_int
        v = C.c_double()
        self.check(self.lib.tc_debug_issue_rate(self.h, C.byref(v)))
        return float(v.value)

    # raw device memory
    def alloc(self, nbytes: int) -> int:
        p = _vp()
        self.check(self.lib.tc_device_alloc(self.h, int(nbytes), C.byref(p)))
        return int(p.value)

    def free(self, dptr: int):
        self.check(self.lib.tc_device_free(self.h, _vp(dptr)))

    def to_device(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self.check(self.lib.tc_copy_to_device(self.h, _vp(dptr), _vp(arr.ctypes.data), arr.nbytes))

    def to_host(self, arr: np.ndarray, dptr: int):
        self.check(self.lib.tc_copy_to_host(self.h, _vp(arr.ctypes.data), _vp(dptr), arr.nbytes))


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """Page-locked host array (freed with the process)."""
    lib = _lib.load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = _vp()
    if lib.tc_host_alloc_pinned(max(n, 1), C.byref(p)) != _lib.TC_OK:
        raise GpuError("pinned host allocation failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class DeviceCloud:
    """Device-resident PointCloud<Point3f> (tc_cloud)."""

    def __init__(self, points, ctx: Optional[Context] = None, stride_bytes: int = 12):
        self.ctx = ctx or default_context()
        h = _vp()
        if stride_bytes == 12:
            pts = _pts(points)
            self.n = pts.shape[0]
            st = self.ctx.lib.tc_cloud_upload(self.ctx.h, _vp(pts.ctypes.data), self.n, C.byref(h))
        else:  # raw strided records, e.g. KITTI .bin (x,y,z,intensity), lidar.rs:310-345
            raw = np.ascontiguousarray(points)
            self.n = raw.nbytes // stride_bytes
            st = self.ctx.lib.tc_cloud_upload_strided(self.ctx.h, _vp(raw.ctypes.data), self.n,
                                                      stride_bytes, C.byref(h))
        self.ctx.check(st)
        self.h = h

    @classmethod
    def from_device(cls, d_xyz: int, n: int, ctx: Optional["Context"] = None) -> "DeviceCloud":
        """Wrap (copy) n x 3 f32 AoS points that already live in device memory
        (tc_cloud_from_device) - e.g. a torch / cupy buffer's data pointer."""
        ctx = ctx or default_context()
        h = _vp()
        ctx.check(ctx.lib.tc_cloud_from_device(ctx.h, _vp(int(d_xyz)), int(n), C.byref(h)))
        c = cls.__new__(cls)
        c.ctx, c.h, c.n = ctx, h, int(n)
        return c

    @classmethod
    def _from_handle(cls, ctx: "Context", h) -> "DeviceCloud":
        c = cls.__new__(cls)
        c.ctx, c.h = ctx, h
        c.n = int(ctx.lib.tc_cloud_len(h))
        return c

    def __len__(self):
        return self.n

    def download(self) -> np.ndarray:
        """The points as a host (n, 3) float32 array (tc_cloud_download)."""
        out = np.empty((self.n, 3), np.float32)
        self.ctx.check(self.ctx.lib.tc_cloud_download(self.ctx.h, self.h, _vp(out.ctypes.data)))
        return out

    def free(self):
        if getattr(self, "h", None):
            self.ctx.lib.tc_cloud_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GridIndex:
    """Uniform-grid spatial index over a DeviceCloud (tc_index) — the KdTree::new replacement."""

    def __init__(self, cloud: DeviceCloud, k_hint: int = 16, cell_size: float = 0.0,
                 shard: Optional[tuple] = None):
        """shard=(rank, world): slab-sharded build for multi-GPU normals
        (tc_index_build_sharded) - estimate_normals_device then writes this rank's rows only."""
        self.cloud = cloud
        self.ctx = cloud.ctx
        self.shard = shard
        h = _vp()
        if shard is None:
            self.ctx.check(self.ctx.lib.tc_index_build(self.ctx.h, cloud.h, int(k_hint),
                                                       float(cell_size), C.byref(h)))
        else:
            self.ctx.check(self.ctx.lib.tc_index_build_sharded(
                self.ctx.h, cloud.h, int(k_hint), float(cell_size), int(shard[0]), int(shard[1]),
                C.byref(h)))
        self.h = h

    def info(self) -> dict:
        inf = _lib.IndexInfoC()
        self.ctx.check(self.ctx.lib.tc_index_get_info(self.h, C.byref(inf)))
        return {"n_points": inf.n_points, "n_cells": inf.n_cells, "dims": tuple(inf.dims),
                "cell_size": inf.cell_size, "bbox_min": tuple(inf.bbox_min),
                "bbox_max": tuple(inf.bbox_max), "occupied_cells": inf.occupied_cells,
                "max_cell_population": inf.max_cell_population, "n_levels": inf.n_levels}

    def free(self):
        if getattr(self, "h", None):
            self.ctx.lib.tc_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- kNN on the index -------------------------------------------------------------------
    def knn(self, queries, k: int, exclude_self: bool = False):
        """Batch kNN.  queries=None: the cloud queries itself.
        -> (idx[nq,k] u32 (TC_NO_INDEX pad), dist[nq,k] f32 (inf pad), count[nq] u32)."""
        ctx = self.ctx
        if queries is None:
            nq, qptr = self.cloud.n, None
        else:
            q = _pts(queries, "queries")
            nq, qptr = q.shape[0], _vp(q.ctypes.data)
        k = int(k)
        idx = np.full((nq, k), _lib.TC_NO_INDEX, np.uint32)
        dist = np.full((nq, k), np.inf, np.float32)
        cnt = np.zeros(nq, np.uint32)
        ctx.check(ctx.lib.tc_knn(ctx.h, self.h, qptr, nq, k, int(bool(exclude_self)),
                                 _vp(idx.ctypes.data), _vp(dist.ctypes.data),
                                 _vp(cnt.ctypes.data)))
        return idx, dist, cnt

    def estimate_normals(self, k: int, radius: Optional[float] = None,
                         consistent_orientation: bool = True, viewpoint=None) -> np.ndarray:
        ctx = self.ctx
        out = np.zeros((self.cloud.n, 6), np.float32)
        vp = None if viewpoint is None else np.ascontiguousarray(viewpoint, np.float32)
        ctx.check(ctx.lib.tc_estimate_normals_indexed(
            ctx.h, self.h, int(k), -1.0 if radius is None else float(radius),
            int(bool(consistent_orientation)), None if vp is None else _vp(vp.ctypes.data),
            _vp(out.ctypes.data)))
        return out

    def estimate_normals_device(self, d_out: int, k: int, consistent_orientation: bool = True,
                                viewpoint=None, shard=(0, None)):
        """Normals into a device buffer (n x 6 f32, original order); optional sorted-order shard."""
        ctx = self.ctx
        vp = None if viewpoint is None else np.ascontiguousarray(viewpoint, np.float32)
        end = self.cloud.n if shard[1] is None else int(shard[1])
        if self.shard is not None and shard == (0, None):
            end = 0xFFFFFFFFFFFFFFFF  # the rank's own rows of a tc_index_build_sharded index
        ctx.check(ctx.lib.tc_estimate_normals_device(
            ctx.h, self.h, int(k), -1.0, int(bool(consistent_orientation)),
            None if vp is None else _vp(vp.ctypes.data), int(shard[0]), end, _vp(d_out)))


# --------------------------------------------------------------------------------------------
# KdTree / PointCloudNeighbors mirror
# --------------------------------------------------------------------------------------------
class KdTree:
    """Mirror of threecrate_algorithms::KdTree (nearest_neighbor.rs:29-299) on the device grid."""

    def __init__(self, points, ctx: Optional[Context] = None, k_hint: int = 8):
        self.points = _pts(points)
        self.cloud = DeviceCloud(self.points, ctx)
        self.index = GridIndex(self.cloud, k_hint=k_hint)

    @classmethod
    def new(cls, points, **kw) -> "KdTree":
        return cls(points, **kw)

    def find_k_nearest(self, query, k: int):
        """-> (indices[<=k] u32, distances f32) ascending; k == 0 or empty tree -> empty."""
        if k == 0 or self.cloud.n == 0:
            return np.empty(0, np.uint32), np.empty(0, np.float32)
        q = np.ascontiguousarray(query, np.float32).reshape(1, 3)
        idx, dist, cnt = self.index.knn(q, k)
        c = int(cnt[0])
        return idx[0, :c].copy(), dist[0, :c].copy()

    def knn(self, queries, k: int):
        """Batched find_k_nearest (one row per query)."""
        return self.index.knn(_pts(queries, "queries"), k)

    def find_radius_neighbors(self, query, radius: float):
        if radius <= 0.0 or self.cloud.n == 0:
            return np.empty(0, np.uint32), np.empty(0, np.float32)
        ctx = self.cloud.ctx
        q = np.ascontiguousarray(query, np.float32).reshape(3)
        cap = self.cloud.n
        idx = np.empty(cap, np.uint32)
        dist = np.empty(cap, np.float32)
        found = C.c_uint64()
        ctx.check(ctx.lib.tc_radius_search(ctx.h, self.index.h, q.ctypes.data_as(C.POINTER(C.c_float)),
                                           float(radius), _vp(idx.ctypes.data),
                                           _vp(dist.ctypes.data), cap, C.byref(found)))
        return idx[: found.value].copy(), dist[: found.value].copy()


def k_nearest_neighbors(points, k: int, ctx: Optional[Context] = None):
    """PointCloudNeighbors::k_nearest_neighbors (point_cloud_ops.rs:80-105): every point's k
    nearest other points.  -> (idx[n,k], dist[n,k], count[n]); empty cloud or k == 0 -> empty."""
    pts = _pts(points)
    if pts.shape[0] == 0 or k == 0:
        return (np.empty((0, k), np.uint32), np.empty((0, k), np.float32), np.empty(0, np.uint32))
    cloud = DeviceCloud(pts, ctx)
    index = GridIndex(cloud, k_hint=k)
    return index.knn(None, k, exclude_self=True)


# --------------------------------------------------------------------------------------------
# normals
# --------------------------------------------------------------------------------------------
@dataclass
class NormalEstimationConfig:
    """normals.rs:17-37"""
    k_neighbors: int = 10
    radius: Optional[float] = None
    consistent_orientation: bool = True
    viewpoint: Optional[tuple] = None


def estimate_normals_with_config(points, config: NormalEstimationConfig,
                                 ctx: Optional[Context] = None) -> np.ndarray:
    """normals.rs:257-357 -> [n, 6] f32 rows of NormalPoint3f (position, normal)."""
    pts = _pts(points)
    n = pts.shape[0]
    if n == 0:  # Ok(empty) BEFORE the k check (normals.rs:261-269)
        return np.zeros((0, 6), np.float32)
    ctx = ctx or default_context()
    out = np.zeros((n, 6), np.float32)
    vp = None if config.viewpoint is None else np.ascontiguousarray(config.viewpoint, np.float32)
    ctx.check(ctx.lib.tc_estimate_normals(
        ctx.h, _vp(pts.ctypes.data), n, int(config.k_neighbors),
        -1.0 if config.radius is None else float(config.radius),
        int(bool(config.consistent_orientation)), None if vp is None else _vp(vp.ctypes.data),
        _vp(out.ctypes.data)))
    return out


def estimate_normals(points, k: int, ctx: Optional[Context] = None) -> np.ndarray:
    """normals.rs:238-247"""
    return estimate_normals_with_config(points, NormalEstimationConfig(k_neighbors=k), ctx)


def estimate_normals_radius(points, radius: float, consistent_orientation: bool,
                            ctx: Optional[Context] = None) -> np.ndarray:
    """normals.rs:368-380 (k_neighbors = 10 is the fallback value)"""
    return estimate_normals_with_config(
        points, NormalEstimationConfig(10, float(radius), consistent_orientation, None), ctx)


# --------------------------------------------------------------------------------------------
# ICP
# --------------------------------------------------------------------------------------------
IDENTITY = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0)
# default search-variant bits of the library (tc_search.cu g_tc_search_flags): per-lane two-pass
# kernels + Newton eigen solver; |32 selects the staged-tile kernels, |64 their TMA staging
DEFAULT_SEARCH_FLAGS = 159


@dataclass
class ICPResult:
    """registration.rs:13-24; transformation as [tx,ty,tz, qi,qj,qk,qw]."""
    transformation: np.ndarray
    mse: float
    iterations: int
    converged: bool
    correspondences: np.ndarray = field(default_factory=lambda: np.empty((0, 2), np.uint64))

    @property
    def translation(self):
        return self.transformation[:3]

    @property
    def rotation(self):
        return self.transformation[3:]

    def matrix(self) -> np.ndarray:
        i, j, k, w = [float(v) for v in self.transformation[3:]]
        m = np.eye(4)
        m[:3, :3] = [[1 - 2 * (j * j + k * k), 2 * (i * j - k * w), 2 * (i * k + j * w)],
                     [2 * (i * j + k * w), 1 - 2 * (i * i + k * k), 2 * (j * k - i * w)],
                     [2 * (i * k - j * w), 2 * (j * k + i * w), 1 - 2 * (i * i + j * j)]]
        m[:3, 3] = self.transformation[:3]
        return m


def icp_point_to_plane_detailed(source, target, target_normals, init=IDENTITY, max_iters: int = 30,
                                max_correspondence_distance: Optional[float] = None,
                                convergence_threshold: float = 1e-6,
                                ctx: Optional[Context] = None,
                                want_correspondences: bool = True) -> ICPResult:
    """registration.rs:508-602"""
    src, tgt = _pts(source, "source"), _pts(target, "target")
    nrm = _pts(target_normals, "target_normals")
    init7 = np.ascontiguousarray(init, np.float32).reshape(7)
    ctx = ctx or default_context()
    res = _lib.IcpResultC()
    pairs = np.zeros((max(src.shape[0], 1), 2), np.uint64) if want_correspondences else None
    ctx.check(ctx.lib.tc_icp_point_to_plane(
        ctx.h, _vp(src.ctypes.data), src.shape[0], _vp(tgt.ctypes.data), tgt.shape[0],
        _vp(nrm.ctypes.data), nrm.shape[0], init7.ctypes.data_as(C.POINTER(C.c_float)),
        int(max_iters), -1.0 if max_correspondence_distance is None else
        float(max_correspondence_distance), float(convergence_threshold), C.byref(res),
        None if pairs is None else _vp(pairs.ctypes.data)))
    corr = pairs[: res.n_correspondences].copy() if pairs is not None else np.empty((0, 2), np.uint64)
    return ICPResult(np.array(res.transform[:], np.float32), float(res.mse), int(res.iterations),
                     bool(res.converged), corr)


def icp_point_to_plane(source, target, target_normals, init=IDENTITY, max_iters: int = 30,
                       ctx: Optional[Context] = None) -> ICPResult:
    """registration.rs:488-496 (max distance None, convergence 1e-6)"""
    return icp_point_to_plane_detailed(source, target, target_normals, init, max_iters, None, 1e-6,
                                       ctx)


def icp_detailed(source, target, init=IDENTITY, max_iters: int = 50,
                 max_correspondence_distance: Optional[float] = None,
                 convergence_threshold: float = 1e-6, ctx: Optional[Context] = None,
                 want_correspondences: bool = True) -> ICPResult:
    """Point-to-point ICP, registration.rs:258-370."""
    src, tgt = _pts(source, "source"), _pts(target, "target")
    init7 = np.ascontiguousarray(init, np.float32).reshape(7)
    ctx = ctx or default_context()
    res = _lib.IcpResultC()
    pairs = np.zeros((max(src.shape[0], 1), 2), np.uint64) if want_correspondences else None
    ctx.check(ctx.lib.tc_icp_point_to_point(
        ctx.h, _vp(src.ctypes.data), src.shape[0], _vp(tgt.ctypes.data), tgt.shape[0],
        init7.ctypes.data_as(C.POINTER(C.c_float)), int(max_iters),
        -1.0 if max_correspondence_distance is None else float(max_correspondence_distance),
        float(convergence_threshold), C.byref(res),
        None if pairs is None else _vp(pairs.ctypes.data)))
    corr = pairs[: res.n_correspondences].copy() if pairs is not None else np.empty((0, 2), np.uint64)
    return ICPResult(np.array(res.transform[:], np.float32), float(res.mse), int(res.iterations),
                     bool(res.converged), corr)


def icp_point_to_point(source, target, init=IDENTITY, max_iterations: int = 50,
                       convergence_threshold: float = 1e-6,
                       max_correspondence_distance: Optional[float] = None,
                       ctx: Optional[Context] = None) -> ICPResult:
    """registration.rs:644-680 (argument order of the reference)."""
    src, tgt = _pts(source, "source"), _pts(target, "target")
    if src.shape[0] == 0 or tgt.shape[0] == 0:
        raise InvalidData("Source or target point cloud is empty")
    if max_iterations == 0:
        raise InvalidData("Max iterations must be positive")
    if convergence_threshold <= 0.0:
        raise InvalidData("Convergence threshold must be positive")
    return icp_detailed(src, tgt, init, max_iterations, max_correspondence_distance,
                        convergence_threshold, ctx)


def icp(source, target, init=IDENTITY, max_iters: int = 50, ctx: Optional[Context] = None) -> np.ndarray:
    """registration.rs:232-242: the final transformation, or `init` on ANY error."""
    try:
        return icp_detailed(source, target, init, max_iters, None, 1e-6, ctx,
                            want_correspondences=False).transformation
    except ThreecrateError:
        return np.ascontiguousarray(init, np.float32).reshape(7)


# --------------------------------------------------------------------------------------------
# filters (threecrate-algorithms/src/filtering.rs) — host arrays or DeviceCloud in, same kind out
# --------------------------------------------------------------------------------------------
def _as_cloud(points, ctx: Optional[Context]):
    if isinstance(points, DeviceCloud):
        return points, False
    return DeviceCloud(_pts(points), ctx), True


def _filter_result(ctx: Context, h, owned_input: Optional[DeviceCloud], host: bool):
    out = DeviceCloud._from_handle(ctx, h)
    if owned_input is not None:
        owned_input.free()
    if not host:
        return out
    pts = out.download()
    out.free()
    return pts


def voxel_grid_filter(points, voxel_size: float, ctx: Optional[Context] = None):
    """filtering.rs:38-133: one centroid per occupied voxel (ascending (z, y, x) voxel order)."""
    cloud, owned = _as_cloud(points, ctx)
    h = _vp()
    try:
        cloud.ctx.check(cloud.ctx.lib.tc_voxel_grid_filter(cloud.ctx.h, cloud.h, float(voxel_size),
                                                          C.byref(h)))
    except ThreecrateError:
        if owned:
            cloud.free()
        raise
    return _filter_result(cloud.ctx, h, cloud if owned else None, owned)


def radius_outlier_removal(points, radius: float, min_neighbors: int,
                           ctx: Optional[Context] = None):
    """filtering.rs:167-218: keep points with >= min_neighbors other points within radius."""
    if min_neighbors < 0:
        raise InvalidData("min_neighbors must be greater than 0")
    cloud, owned = _as_cloud(points, ctx)
    h = _vp()
    try:
        cloud.ctx.check(cloud.ctx.lib.tc_radius_outlier_removal(
            cloud.ctx.h, cloud.h, float(radius), int(min_neighbors), C.byref(h)))
    except ThreecrateError:
        if owned:
            cloud.free()
        raise
    return _filter_result(cloud.ctx, h, cloud if owned else None, owned)


def _sor(points, k_neighbors: int, value: float, mode: int, ctx, return_stats: bool):
    if k_neighbors < 0:
        raise InvalidData("k_neighbors must be greater than 0")
    cloud, owned = _as_cloud(points, ctx)
    h = _vp()
    stats = (C.c_float * 3)()
    try:
        cloud.ctx.check(cloud.ctx.lib.tc_statistical_outlier_removal(
            cloud.ctx.h, cloud.h, int(k_neighbors), float(value), int(mode), stats, C.byref(h)))
    except ThreecrateError:
        if owned:
            cloud.free()
        raise
    res = _filter_result(cloud.ctx, h, cloud if owned else None, owned)
    if return_stats:
        return res, {"mean": stats[0], "std_dev": stats[1], "threshold": stats[2]}
    return res


def statistical_outlier_removal(points, k_neighbors: int, std_dev_multiplier: float,
                                ctx: Optional[Context] = None, fast: bool = False,
                                return_stats: bool = False):
    """filtering.rs:253-321.  fast=False accumulates the global mean / variance like the
    reference (sequential f32, bit-exact); fast=True uses f64 tree sums."""
    return _sor(points, k_neighbors, std_dev_multiplier, 1 if fast else 0, ctx, return_stats)


def statistical_outlier_removal_with_threshold(points, k_neighbors: int, threshold: float,
                                               ctx: Optional[Context] = None):
    """filtering.rs:335-394"""
    return _sor(points, k_neighbors, threshold, 2, ctx, False)


# --------------------------------------------------------------------------------------------
# multiscale point-to-point ICP (registration.rs:26-71, 704-789)
# --------------------------------------------------------------------------------------------
@dataclass
class IcpScaleLevel:
    voxel_size: float
    max_iterations: int
    max_correspondence_distance: Optional[float] = None


@dataclass
class MultiScaleIcpConfig:
    """Defaults of registration.rs:46-71."""
    levels: list = field(default_factory=lambda: [IcpScaleLevel(0.20, 10, 0.50),
                                                  IcpScaleLevel(0.10, 10, 0.25),
                                                  IcpScaleLevel(0.05, 15, 0.15)])
    final_refinement_iterations: int = 10
    final_max_correspondence_distance: Optional[float] = 0.10
    convergence_threshold: float = 1e-5


def multiscale_icp_point_to_point(source, target, init=IDENTITY,
                                  config: Optional[MultiScaleIcpConfig] = None,
                                  ctx: Optional[Context] = None,
                                  want_correspondences: bool = True) -> ICPResult:
    """registration.rs:704-789"""
    config = config or MultiScaleIcpConfig()
    src, tgt = _pts(source, "source"), _pts(target, "target")
    init7 = np.ascontiguousarray(init, np.float32).reshape(7)
    ctx = ctx or default_context()
    nl = len(config.levels)
    levels = (_lib.IcpScaleLevelC * max(nl, 1))()
    for i, lv in enumerate(config.levels):
        levels[i].voxel_size = float(lv.voxel_size)
        levels[i].max_iterations = int(lv.max_iterations)
        levels[i].max_correspondence_distance = (
            -1.0 if lv.max_correspondence_distance is None else float(lv.max_correspondence_distance))
    res = _lib.IcpResultC()
    pairs = np.zeros((max(src.shape[0], 1), 2), np.uint64) if want_correspondences else None
    fm = config.final_max_correspondence_distance
    ctx.check(ctx.lib.tc_multiscale_icp_point_to_point(
        ctx.h, _vp(src.ctypes.data), src.shape[0], _vp(tgt.ctypes.data), tgt.shape[0],
        init7.ctypes.data_as(C.POINTER(C.c_float)), levels, nl,
        int(config.final_refinement_iterations), -1.0 if fm is None else float(fm),
        float(config.convergence_threshold), C.byref(res),
        None if pairs is None else _vp(pairs.ctypes.data)))
    corr = pairs[: res.n_correspondences].copy() if pairs is not None else np.empty((0, 2), np.uint64)
    return ICPResult(np.array(res.transform[:], np.float32), float(res.mse), int(res.iterations),
                     bool(res.converged), corr)


# --------------------------------------------------------------------------------------------
# Generalized ICP (threecrate-algorithms/src/gicp.rs)
# --------------------------------------------------------------------------------------------
@dataclass
class GicpConfig:
    """gicp.rs:24-46"""
    max_iterations: int = 50
    max_correspondence_distance: float = 1.0
    convergence_threshold: float = 1e-6
    k_correspondences: int = 20


def gicp(source, target, init=IDENTITY, config: Optional[GicpConfig] = None,
         ctx: Optional[Context] = None, want_correspondences: bool = True) -> ICPResult:
    """gicp.rs:117-312"""
    config = config or GicpConfig()
    src, tgt = _pts(source, "source"), _pts(target, "target")
    init7 = np.ascontiguousarray(init, np.float32).reshape(7)
    ctx = ctx or default_context()
    res = _lib.IcpResultC()
    pairs = np.zeros((max(src.shape[0], 1), 2), np.uint64) if want_correspondences else None
    ctx.check(ctx.lib.tc_gicp(
        ctx.h, _vp(src.ctypes.data), src.shape[0], _vp(tgt.ctypes.data), tgt.shape[0],
        init7.ctypes.data_as(C.POINTER(C.c_float)), int(config.max_iterations),
        float(config.max_correspondence_distance), float(config.convergence_threshold),
        int(config.k_correspondences), C.byref(res),
        None if pairs is None else _vp(pairs.ctypes.data)))
    corr = pairs[: res.n_correspondences].copy() if pairs is not None else np.empty((0, 2), np.uint64)
    return ICPResult(np.array(res.transform[:], np.float32), float(res.mse), int(res.iterations),
                     bool(res.converged), corr)


# --------------------------------------------------------------------------------------------
# multi-GPU (one process per GPU)
# --------------------------------------------------------------------------------------------
def dist_chunk(n_total: int, n_ranks: int, rank: int):
    """tc_dist_chunk: the contiguous row range of rank `rank` (equal lengths, multiple of 4)."""
    lo, hi = C.c_uint64(), C.c_uint64()
    _lib.load().tc_dist_chunk(int(n_total), int(n_ranks), int(rank), C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


class Comm:
    """NCCL communicator for the sharded ICP reduction (tc_comm).  The 128-byte unique id is
    produced on rank 0 and handed to the other ranks by the host (torch.distributed here)."""

    def __init__(self, ctx: Context, unique_id: bytes, n_ranks: int, rank: int):
        self.ctx = ctx
        h = _vp()
        buf = C.create_string_buffer(unique_id, _lib.TC_COMM_ID_BYTES)
        ctx.check(ctx.lib.tc_comm_init_rank(ctx.h, buf, int(n_ranks), int(rank), C.byref(h)))
        self.h = h
        self.n_ranks, self.rank = n_ranks, rank

    @staticmethod
    def unique_id(ctx: Context) -> bytes:
        buf = C.create_string_buffer(_lib.TC_COMM_ID_BYTES)
        ctx.check(ctx.lib.tc_comm_get_unique_id(ctx.h, buf))
        return buf.raw

    def peer_handle(self) -> bytes:
        """IPC handle of this rank's NVLink exchange buffer (gather from all ranks, then
        `open_peers`)."""
        buf = C.create_string_buffer(_lib.TC_IPC_HANDLE_BYTES)
        self.ctx.check(self.ctx.lib.tc_comm_peer_handle(self.h, buf))
        return buf.raw

    def open_peers(self, handles) -> None:
        """handles: the peer_handle() of every rank, in rank order.  Enables the fused
        (in-kernel, NVLink peer-memory) all-reduce of the ICP normal equations."""
        blob = b"".join(handles)
        assert len(blob) == self.n_ranks * _lib.TC_IPC_HANDLE_BYTES
        self.ctx.check(self.ctx.lib.tc_comm_peer_open(self.h, C.create_string_buffer(blob, len(blob))))

    # ---- distributed normals (tc_estimate_normals_distributed) ------------------------------
    def chunk(self, n_total: int):
        """Rows [lo, hi) of an n_total-row cloud this rank passes in and gets back."""
        return dist_chunk(n_total, self.n_ranks, self.rank)

    def window_handle(self, n_total: int) -> bytes:
        """Allocates this rank's NVLink window for clouds of n_total points and returns its IPC
        handle (gather from all ranks, then `open_window`)."""
        buf = C.create_string_buffer(_lib.TC_IPC_HANDLE_BYTES)
        self.ctx.check(self.ctx.lib.tc_comm_window_handle(self.h, int(n_total), buf))
        return buf.raw

    def open_window(self, handles) -> None:
        blob = b"".join(handles)
        assert len(blob) == self.n_ranks * _lib.TC_IPC_HANDLE_BYTES
        self.ctx.check(self.ctx.lib.tc_comm_window_open(self.h, C.create_string_buffer(blob, len(blob))))

    def estimate_normals(self, chunk_xyz: np.ndarray, n_total: int, k: int,
                         consistent_orientation: bool = True, viewpoint=None,
                         out: Optional[np.ndarray] = None) -> np.ndarray:
        """estimate_normals of ONE cloud over all ranks: pass this rank's rows (`chunk(n_total)`)
        and get the NormalPoint3f rows of the same range (bit-identical to the single-GPU
        result).  Collective: every rank must call it."""
        lo, hi = self.chunk(n_total)
        pts = np.ascontiguousarray(chunk_xyz, np.float32).reshape(-1, 3)
        if len(pts) != hi - lo:
            raise InvalidData(f"rank {self.rank} passes rows [{lo}, {hi}) of the cloud, got {len(pts)}")
        if out is None:
            out = np.empty((hi - lo, 6), np.float32)
        vp = None if viewpoint is None else np.ascontiguousarray(viewpoint, np.float32).reshape(3)
        self.ctx.check(self.ctx.lib.tc_estimate_normals_distributed(
            self.ctx.h, self.h, _vp(pts.ctypes.data), int(n_total), int(k),
            1 if consistent_orientation else 0, None if vp is None else _vp(vp.ctypes.data),
            _vp(out.ctypes.data)))
        return out

    def allreduce_f64(self, dptr: int, count: int):
        self.ctx.check(self.ctx.lib.tc_comm_allreduce_f64(self.h, _vp(dptr), int(count)))

    def destroy(self):
        if getattr(self, "h", None):
            self.ctx.lib.tc_comm_destroy(self.h)
            self.h = None


def icp_point_to_plane_device(src: DeviceCloud, tgt_index: GridIndex, d_tgt_normals: int,
                              init=IDENTITY, max_iters: int = 30,
                              max_correspondence_distance: Optional[float] = None,
                              convergence_threshold: float = 1e-6, comm: Optional[Comm] = None,
                              d_match_out: int = 0) -> ICPResult:
    """Device-resident ICP (tc_icp_point_to_plane_device); with `comm`, `src` is this rank's
    shard and the 29-scalar normal equations are all-reduced every iteration."""
    ctx = src.ctx
    init7 = np.ascontiguousarray(init, np.float32).reshape(7)
    res = _lib.IcpResultC()
    ctx.check(ctx.lib.tc_icp_point_to_plane_device(
        ctx.h, comm.h if comm else None, src.h, tgt_index.h, _vp(d_tgt_normals),
        init7.ctypes.data_as(C.POINTER(C.c_float)), int(max_iters),
        -1.0 if max_correspondence_distance is None else float(max_correspondence_distance),
        float(convergence_threshold), C.byref(res), _vp(d_match_out) if d_match_out else None))
    return ICPResult(np.array(res.transform[:], np.float32), float(res.mse), int(res.iterations),
                     bool(res.converged))


def icp_point_to_point_device(src: DeviceCloud, tgt_index: GridIndex, init=IDENTITY,
                              max_iters: int = 50,
                              max_correspondence_distance: Optional[float] = None,
                              convergence_threshold: float = 1e-6, comm: Optional[Comm] = None,
                              d_match_out: int = 0) -> ICPResult:
    """Device-resident point-to-point ICP (tc_icp_point_to_point_device)."""
    ctx = src.ctx
    init7 = np.ascontiguousarray(init, np.float32).reshape(7)
    res = _lib.IcpResultC()
    ctx.check(ctx.lib.tc_icp_point_to_point_device(
        ctx.h, comm.h if comm else None, src.h, tgt_index.h,
        init7.ctypes.data_as(C.POINTER(C.c_float)), int(max_iters),
        -1.0 if max_correspondence_distance is None else float(max_correspondence_distance),
        float(convergence_threshold), C.byref(res), _vp(d_match_out) if d_match_out else None))
    return ICPResult(np.array(res.transform[:], np.float32), float(res.mse), int(res.iterations),
                     bool(res.converged))
