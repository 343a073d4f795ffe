"""Deterministic synthetic clouds for the BASELINE.json configs.

Pure numpy (no oracle, no CUDA): used by tests/, bench.py and __graft_entry__.smoke().
All clouds are float32 metres, generated from fixed seeds with numpy's PCG64.

  C1  bunny_standin()        35,947-pt bumpy sphere (assets/bunny.obj is absent from the mount)
  C2  kitti_frame()          64 beams x 1875 azimuth steps = 120,000 rays, KITTI-shaped
  C3  scan_pair(1_000_000)   two independently sampled terrain scans + known rigid offset
  C4  terrain(10_000_000)    terrain over [-100,100]^2
  C5  scan_pair(100_000_000) terrain over [-316,316]^2
Also the reference's own inline test fixtures (cube, planes, cylinder, Fibonacci sphere).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------------
# Reference test fixtures (lifted as inputs; citations are /root/reference paths)
# --------------------------------------------------------------------------------------------
def cube8() -> np.ndarray:
    """8-corner unit cube, threecrate-algorithms/src/nearest_neighbor.rs:395-406."""
    return np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1],
                     [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], F32)


def plane5() -> np.ndarray:
    """5-pt XY plane, normals.rs:401-406."""
    return np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0.5, 0.5, 0]], F32)


def plane4() -> np.ndarray:
    """4-pt XY plane, normals.rs:553-557."""
    return np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], F32)


def grid_plane(n: int = 20, spacing: float = 0.1) -> np.ndarray:
    """n x n grid on z = 0, normals.rs:443-451 (i outer, j inner; x = i*0.1, y = j*0.1)."""
    i, j = np.meshgrid(np.arange(n, dtype=F32), np.arange(n, dtype=F32), indexing="ij")
    pts = np.stack([i.ravel() * F32(spacing), j.ravel() * F32(spacing),
                    np.zeros(n * n, F32)], axis=1)
    return pts.astype(F32)


def cylinder(n_theta: int = 10, n_z: int = 10) -> np.ndarray:
    """10 x 10 unit cylinder, normals.rs:485-495 (theta = i/10 * 2pi, z = j/10 * 2 - 1)."""
    pts = []
    for i in range(n_theta):
        for j in range(n_z):
            theta = F32(i) / F32(n_theta) * F32(2.0) * F32(np.pi)
            z = F32(j) / F32(n_z) * F32(2.0) - F32(1.0)
            pts.append([np.cos(theta, dtype=F32), np.sin(theta, dtype=F32), z])
    return np.array(pts, F32)


def fibonacci_sphere(n: int, radius: float = 3.0):
    """Fibonacci sphere with analytic outward normals, registration.rs:1148-1165 (f32 ops)."""
    golden = F32(np.pi) * (F32(3.0) - np.sqrt(F32(5.0)))
    i = np.arange(n, dtype=F32)
    y = F32(1.0) - (i / max(F32(n) - F32(1.0), F32(1.0))) * F32(2.0)
    r = np.sqrt(np.maximum(F32(1.0) - y * y, F32(0.0))).astype(F32)
    theta = (golden * i).astype(F32)
    x = (np.cos(theta, dtype=F32) * r).astype(F32)
    z = (np.sin(theta, dtype=F32) * r).astype(F32)
    nrm = np.stack([x, y, z], axis=1).astype(F32)
    return (nrm * F32(radius)).astype(F32), nrm


# --------------------------------------------------------------------------------------------
# Rigid transforms (iso7 = [tx,ty,tz, qi,qj,qk,qw], nalgebra coordinate order)
# --------------------------------------------------------------------------------------------
def quat_from_euler(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """UnitQuaternion::from_euler_angles(roll, pitch, yaw) (f64 math, returned as [i,j,k,w])."""
    sr, cr = np.sin(roll / 2), np.cos(roll / 2)
    sp, cp = np.sin(pitch / 2), np.cos(pitch / 2)
    sy, cy = np.sin(yaw / 2), np.cos(yaw / 2)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                     cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy], np.float64)


def quat_to_matrix(q) -> np.ndarray:
    i, j, k, w = [float(v) for v in q]
    return np.array([[1 - 2 * (j * j + k * k), 2 * (i * j - k * w), 2 * (i * k + j * w)],
                     [2 * (i * j + k * w), 1 - 2 * (i * i + k * k), 2 * (j * k - i * w)],
                     [2 * (i * k - j * w), 2 * (j * k + i * w), 1 - 2 * (i * i + j * j)]])


def apply_iso(iso7, pts: np.ndarray) -> np.ndarray:
    """Apply iso7 in f64, return f32."""
    iso7 = np.asarray(iso7, np.float64)
    R = quat_to_matrix(iso7[3:7])
    return (pts.astype(np.float64) @ R.T + iso7[:3]).astype(F32)


def invert_iso(iso7) -> np.ndarray:
    iso7 = np.asarray(iso7, np.float64)
    qi = np.array([-iso7[3], -iso7[4], -iso7[5], iso7[6]])
    R = quat_to_matrix(qi)
    return np.concatenate([-(R @ iso7[:3]), qi])


# ground-truth offset of the reference's bench harness (examples/threecrate_dataset_bench.rs:281-287)
# t = (0.05, -0.02, 0.01), yaw 0.02 rad; SURVEY §8d adds roll 0.01 for the two-scan variant.
def bench_transform(roll: float = 0.0) -> np.ndarray:
    return np.concatenate([[0.05, -0.02, 0.01], quat_from_euler(roll, 0.0, 0.02)])


# --------------------------------------------------------------------------------------------
# C1: bunny stand-in
# --------------------------------------------------------------------------------------------
def bunny_standin(n: int = 35947) -> np.ndarray:
    """Bumpy sphere r = 0.1 (1 + 0.15 sin 3θ sin 4φ) on a Fibonacci lattice."""
    i = np.arange(n, dtype=np.float64)
    y = 1.0 - (i / (n - 1.0)) * 2.0
    rr = np.sqrt(np.maximum(1.0 - y * y, 0.0))
    phi = np.pi * (3.0 - np.sqrt(5.0)) * i
    theta = np.arccos(np.clip(y, -1, 1))
    r = 0.1 * (1.0 + 0.15 * np.sin(3 * theta) * np.sin(4 * phi))
    return np.stack([r * np.cos(phi) * rr, r * y, r * np.sin(phi) * rr], axis=1).astype(F32)


# --------------------------------------------------------------------------------------------
# C2: KITTI-shaped LiDAR frame
# --------------------------------------------------------------------------------------------
def kitti_frame(seed: int = 0x3C0FFEE, beams: int = 64, az_steps: int = 1875,
                with_intensity: bool = False) -> np.ndarray:
    """64 beams (elevation +2.0 .. -24.8 deg) x 1875 azimuth steps = 120,000 rays from a sensor
    at (0,0,1.73) into ground z=0 + 4 walls (x=±40, y=±12) + 12 seeded boxes; range noise
    N(0, 0.02); rays that miss or exceed 80 m are re-aimed at the ground ring.
    Returns [N,3] f32 (or [N,4] with an intensity column = KITTI .bin stride-16 records,
    threecrate-io/src/lidar.rs:310-345)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    elev = np.deg2rad(np.linspace(2.0, -24.8, beams))
    az = np.linspace(0.0, 2 * np.pi, az_steps, endpoint=False)
    E, A = np.meshgrid(elev, az, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    o = np.array([0.0, 0.0, 1.73])
    t_best = np.full(d.shape[0], np.inf)

    def hit_plane(axis, value, lo, hi):
        nonlocal t_best
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (value - o[axis]) / d[:, axis]
            p = o + t[:, None] * d
        ok = (t > 0.5) & np.isfinite(t)
        for a in range(3):
            if a != axis:
                ok &= (p[:, a] >= lo[a]) & (p[:, a] <= hi[a])
        t_best = np.where(ok & (t < t_best), t, t_best)

    big = 1e9
    hit_plane(2, 0.0, [-big, -big, 0], [big, big, 0])                 # ground
    for xw in (-40.0, 40.0):
        hit_plane(0, xw, [0, -12.0, 0.0], [0, 12.0, 6.0])             # walls x = ±40
    for yw in (-12.0, 12.0):
        hit_plane(1, yw, [-40.0, 0, 0.0], [40.0, 0, 6.0])             # walls y = ±12
    for _ in range(12):                                               # boxes (cars / poles)
        c = np.array([rng.uniform(-35, 35), rng.uniform(-10, 10), 0.0])
        if np.hypot(c[0], c[1]) < 4.0:
            c[0] += 8.0
        h = np.array([rng.uniform(0.8, 2.4), rng.uniform(0.8, 1.2), rng.uniform(1.2, 2.2)])
        lo, hi = c - [h[0], h[1], 0], c + [h[0], h[1], h[2]]
        for axis in range(3):
            for v in (lo[axis], hi[axis]):
                hit_plane(axis, v, lo, hi)
    miss = ~np.isfinite(t_best) | (t_best > 80.0)
    # re-aim misses at the ground between 5 m and 60 m along the same azimuth
    rr = rng.uniform(5.0, 60.0, size=d.shape[0])
    azm = np.arctan2(d[:, 1], d[:, 0])
    gp = np.stack([rr * np.cos(azm), rr * np.sin(azm), np.zeros_like(rr)], axis=1)
    dn = gp - o
    tn = np.linalg.norm(dn, axis=1)
    d = np.where(miss[:, None], dn / tn[:, None], d)
    t_best = np.where(miss, tn, t_best)
    t_best = t_best + rng.normal(0.0, 0.02, size=t_best.shape)
    pts = (o + t_best[:, None] * d).astype(F32)
    if with_intensity:
        inten = rng.uniform(0, 1, size=(pts.shape[0], 1)).astype(F32)
        return np.concatenate([pts, inten], axis=1)
    return pts


# --------------------------------------------------------------------------------------------
# C3/C4/C5: terrain
# --------------------------------------------------------------------------------------------
def terrain_height(x, y):
    return 0.5 * np.sin(0.3 * x) * np.cos(0.2 * y)


def terrain(n: int, half_extent: float = 100.0, seed: int = 4, noise: float = 0.0,
            wall_fraction: float = 0.05, return_normals: bool = False):
    """z = 0.5 sin(0.3x) cos(0.2y) over [-H,H]^2 plus four boundary walls (z in [0,3]).
    Returns points [n,3] f32 (and analytic unit normals when return_normals)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    H = float(half_extent)
    nw = int(n * wall_fraction)
    ng = n - nw
    x = rng.uniform(-H, H, ng)
    y = rng.uniform(-H, H, ng)
    z = terrain_height(x, y)
    nrm = None
    if return_normals:
        dzdx = 0.15 * np.cos(0.3 * x) * np.cos(0.2 * y)
        dzdy = -0.1 * np.sin(0.3 * x) * np.sin(0.2 * y)
        nrm = np.stack([-dzdx, -dzdy, np.ones_like(x)], axis=1)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    if noise > 0:
        z = z + rng.normal(0.0, noise, ng)
    g = np.stack([x, y, z], axis=1)
    if nw > 0:
        side = rng.integers(0, 4, nw)
        u = rng.uniform(-H, H, nw)
        zz = rng.uniform(0.0, 3.0, nw)
        wx = np.where(side == 0, -H, np.where(side == 1, H, u))
        wy = np.where(side == 2, -H, np.where(side == 3, H, u))
        if noise > 0:
            wx = wx + np.where(side < 2, rng.normal(0.0, noise, nw), 0.0)
            wy = wy + np.where(side >= 2, rng.normal(0.0, noise, nw), 0.0)
        w = np.stack([wx, wy, zz], axis=1)
        pts = np.concatenate([g, w], axis=0)
        if return_normals:
            wn = np.zeros((nw, 3))
            wn[:, 0] = np.where(side == 0, 1.0, np.where(side == 1, -1.0, 0.0))
            wn[:, 1] = np.where(side == 2, 1.0, np.where(side == 3, -1.0, 0.0))
            nrm = np.concatenate([nrm, wn], axis=0)
    else:
        pts = g
    perm = rng.permutation(n)  # scans are not spatially ordered
    pts = pts[perm].astype(F32)
    if return_normals:
        return pts, nrm[perm].astype(F32)
    return pts


def scan_pair(n: int, half_extent: float = 50.0, seed_target: int = 1, seed_source: int = 2,
              noise: float = 0.005, copy_variant: bool = False):
    """C3/C5: (source, target, target_analytic_normals, T_gt) with T_gt * source ~= target.
    Two independent samples of the same terrain (or, with copy_variant, source = exact inverse-
    transformed copy of target as in the reference's harness); source is pre-moved by T_gt^-1."""
    tgt, nrm = terrain(n, half_extent, seed_target, noise=noise, return_normals=True)
    T = bench_transform(roll=0.0 if copy_variant else 0.01)
    if copy_variant:
        base = tgt
    else:
        base = terrain(n, half_extent, seed_source, noise=noise)
    src = apply_iso(invert_iso(T), base)
    return src, tgt, nrm, T.astype(np.float64)
