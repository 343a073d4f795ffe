"""CPU restatement of the reference's filters and multiscale ICP (TEST INFRASTRUCTURE ONLY).

    voxel_grid_filter                           threecrate-algorithms/src/filtering.rs:38-133
    radius_outlier_removal                      filtering.rs:167-218
    statistical_outlier_removal                 filtering.rs:253-321
    statistical_outlier_removal_with_threshold  filtering.rs:335-394
    multiscale_icp_point_to_point               registration.rs:704-789

numpy float32 element-wise arithmetic is IEEE single without contraction, i.e. the reference's
f32 semantics; sequential f32 sums are taken with cumsum (numpy's sum() is pairwise, cumsum is
not).  PARITY UNPINNED like the rest of the oracle: pinned by the reference's own inline test
assertions (tests/test_oracle_filters.py), not by golden vectors.
"""
from __future__ import annotations

import numpy as np

from .oracle import (AlgorithmError, InvalidData, OracleKdTree, _f32, icp_point_to_point)

_F = np.float32


def _voxel_coords(pts: np.ndarray, voxel_size: float) -> np.ndarray:
    """floor((p - min) / voxel_size) as i32, per axis, in f32 (filtering.rs:95-100)."""
    mn = pts.min(axis=0)  # min_by(partial_cmp), filtering.rs:52-69
    q = np.floor((pts - mn) / _F(voxel_size))  # f32 subtract, f32 divide
    return np.clip(q, -2147483648.0, 2147483647.0).astype(np.int64)  # Rust `as i32` saturates


def voxel_grid_filter(points, voxel_size: float, return_coords: bool = False):
    """One centroid per voxel, f64 sums in cloud order (filtering.rs:108-130).  The reference
    emits the voxels in HashMap order (arbitrary); here they ascend by (z, y, x) coordinate."""
    pts = _f32(points, (-1, 3))
    if pts.shape[0] == 0:  # :42-44
        return (pts.copy(), np.empty((0, 3), np.int64)) if return_coords else pts.copy()
    if not voxel_size > 0.0:  # :46-50  (voxel_size <= 0.0 -> Err)
        if voxel_size <= 0.0:
            raise InvalidData("voxel_size must be positive")
    vc = _voxel_coords(pts, voxel_size)
    # unique voxels in ascending (z, y, x); `inv` maps every point to its voxel
    uniq, inv = np.unique(vc[:, ::-1], axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    sums = np.zeros((uniq.shape[0], 3), np.float64)
    np.add.at(sums, inv, pts.astype(np.float64))  # unbuffered: adds in cloud order
    cnt = np.bincount(inv, minlength=uniq.shape[0]).astype(np.float64)
    out = (sums * (1.0 / cnt)[:, None]).astype(_F)  # sum * (1 / count), then `as f32`
    return (out, uniq[:, ::-1].copy()) if return_coords else out


def _radius_counts(pts: np.ndarray, radius: float, chunk: int = 512) -> np.ndarray:
    """|find_radius_neighbors(p, radius)| per point (nearest_neighbor.rs:254-298): d2 <= r*r with
    d2 = dx*dx + dy*dy + dz*dz evaluated left to right in f32."""
    r2 = _F(radius) * _F(radius)
    n = pts.shape[0]
    out = np.empty(n, np.int64)
    for a in range(0, n, chunk):
        d = pts[a:a + chunk, None, :] - pts[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        out[a:a + chunk] = (d2 <= r2).sum(axis=1)
    return out


def radius_outlier_removal(points, radius: float, min_neighbors: int, return_mask: bool = False):
    pts = _f32(points, (-1, 3))
    if pts.shape[0] == 0:  # :172-174
        return (pts.copy(), np.zeros(0, bool)) if return_mask else pts.copy()
    if radius <= 0.0:
        raise InvalidData("radius must be positive")
    if min_neighbors == 0:
        raise InvalidData("min_neighbors must be greater than 0")
    counts = np.maximum(_radius_counts(pts, radius) - 1, 0)  # saturating_sub(1), :198
    keep = counts >= min_neighbors
    return (pts[keep], keep) if return_mask else pts[keep]


def sor_mean_distances(points, k_neighbors: int, threads: int = 0) -> np.ndarray:
    """Mean distance to the k+1 nearest neighbours, every neighbour with the point's own
    coordinates skipped; f32 sum in ascending-distance order (filtering.rs:279-301)."""
    pts = _f32(points, (-1, 3))
    n = pts.shape[0]
    k1 = k_neighbors + 1
    tree = OracleKdTree(pts)
    idx, d2, cnt = tree.knn_batch(pts, k1, threads)
    dist = np.sqrt(d2)  # f32 sqrt; find_k_nearest returns sqrt(d2)
    mean = np.zeros(n, _F)
    for i in range(n):
        c = int(cnt[i])
        nb = idx[i, :c].astype(np.int64)
        other = ~np.all(pts[nb] == pts[i], axis=1)  # cloud.points[idx] != *point
        dd = dist[i, :c][other]
        if dd.size:
            mean[i] = np.cumsum(dd, dtype=_F)[-1] / _F(dd.size)  # iter().sum::<f32>() / len
    return mean


def sor_threshold(mean: np.ndarray, std_dev_multiplier: float):
    """Global mean / std dev / threshold with the reference's sequential f32 sums (:304-312)."""
    n = _F(mean.shape[0])
    gm = np.cumsum(mean, dtype=_F)[-1] / n
    d = mean - gm
    var = np.cumsum(d * d, dtype=_F)[-1] / n
    sd = np.sqrt(var, dtype=_F)
    return gm, sd, gm + _F(std_dev_multiplier) * sd


def statistical_outlier_removal(points, k_neighbors: int, std_dev_multiplier: float,
                                return_details: bool = False, threads: int = 0):
    pts = _f32(points, (-1, 3))
    if pts.shape[0] == 0:
        return (pts.copy(), {}) if return_details else pts.copy()
    if k_neighbors == 0:
        raise InvalidData("k_neighbors must be greater than 0")
    if std_dev_multiplier <= 0.0:
        raise InvalidData("std_dev_multiplier must be positive")
    mean = sor_mean_distances(pts, k_neighbors, threads)
    gm, sd, thr = sor_threshold(mean, std_dev_multiplier)
    keep = mean <= thr  # :317
    if return_details:
        return pts[keep], {"mask": keep, "mean_distances": mean, "mean": float(gm),
                           "std_dev": float(sd), "threshold": float(thr)}
    return pts[keep]


def statistical_outlier_removal_with_threshold(points, k_neighbors: int, threshold: float,
                                               return_mask: bool = False, threads: int = 0):
    pts = _f32(points, (-1, 3))
    if pts.shape[0] == 0:
        return (pts.copy(), np.zeros(0, bool)) if return_mask else pts.copy()
    if k_neighbors == 0:
        raise InvalidData("k_neighbors must be greater than 0")
    if threshold <= 0.0:
        raise InvalidData("threshold must be positive")
    keep = sor_mean_distances(pts, k_neighbors, threads) <= _F(threshold)
    return (pts[keep], keep) if return_mask else pts[keep]


DEFAULT_LEVELS = ((0.20, 10, 0.50), (0.10, 10, 0.25), (0.05, 15, 0.15))  # registration.rs:46-71


def multiscale_icp_point_to_point(source, target, init=None, levels=DEFAULT_LEVELS,
                                  final_refinement_iterations: int = 10,
                                  final_max_correspondence_distance=0.10,
                                  convergence_threshold: float = 1e-5, threads: int = 0):
    """levels = [(voxel_size, max_iterations, max_correspondence_distance | None), ...]"""
    src = _f32(source, (-1, 3))
    tgt = _f32(target, (-1, 3))
    if src.shape[0] == 0 or tgt.shape[0] == 0:
        raise InvalidData("Source or target point cloud is empty")
    if len(levels) == 0:
        raise InvalidData("At least one ICP scale level is required")
    if convergence_threshold <= 0.0:
        raise InvalidData("Convergence threshold must be positive")
    if final_refinement_iterations == 0:
        raise InvalidData("Final refinement iterations must be positive")
    T = np.array([0, 0, 0, 0, 0, 0, 1], _F) if init is None else _f32(init, (7,))
    total = 0
    last = None
    for voxel_size, max_iterations, max_dist in levels:
        if voxel_size <= 0.0:
            raise InvalidData("Scale voxel_size must be positive")
        if max_iterations == 0:
            raise InvalidData("Scale max_iterations must be positive")
        sd = voxel_grid_filter(src, voxel_size)
        td = voxel_grid_filter(tgt, voxel_size)
        if sd.shape[0] < 3 or td.shape[0] < 3:
            continue
        r = icp_point_to_point(sd, td, T, max_iterations, convergence_threshold, max_dist, threads)
        T = np.concatenate([r.translation, r.rotation]).astype(_F)
        total += r.iterations
        last = r
    if last is None:
        raise AlgorithmError("No multiscale ICP level had enough downsampled points")
    fin = icp_point_to_point(src, tgt, T, final_refinement_iterations, convergence_threshold,
                             final_max_correspondence_distance, threads)
    fin.iterations += total
    return fin
