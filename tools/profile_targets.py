#!/usr/bin/env python
"""Small driver for ncu captures: runs one workload a few times (no timing, no oracle).

    python tools/profile_targets.py c2|c4|c3|knn [--n N] [--reps R]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import threecrate_b200 as tc  # noqa: E402
from fixtures import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["c2", "c4", "c3", "c5far", "knn", "gicp", "filters"])
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--k", type=int, default=30)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
ctx = tc.default_context()

if a.what in ("c2", "knn"):
    pts = synth.kitti_frame()
    cloud = tc.DeviceCloud(pts, ctx)
    d_out = ctx.alloc(len(pts) * 24)
    for _ in range(a.reps):
        index = tc.GridIndex(cloud, k_hint=16)
        if a.what == "c2":
            index.estimate_normals_device(d_out, 16)
        else:
            index.knn(None, 16, exclude_self=True)
        ctx.synchronize()
        print(index.info())
        index.free()
elif a.what == "c4":
    n = a.n or 10_000_000
    pts = synth.terrain(n, 100.0 * (n / 1e7) ** 0.5, seed=4, noise=0.002)
    cloud = tc.DeviceCloud(pts, ctx)
    d_out = ctx.alloc(n * 24)
    for _ in range(a.reps):
        index = tc.GridIndex(cloud, k_hint=a.k)
        index.estimate_normals_device(d_out, a.k)
        ctx.synchronize()
        print(index.info())
        index.free()
elif a.what == "c3":
    n = a.n or 1_000_000
    src, tgt, nrm, T = synth.scan_pair(n, half_extent=50.0 * (n / 1e6) ** 0.5)
    tcloud, scloud = tc.DeviceCloud(tgt, ctx), tc.DeviceCloud(src, ctx)
    d_nrm = ctx.alloc(n * 12)
    ctx.to_device(d_nrm, nrm)
    for _ in range(a.reps):
        index = tc.GridIndex(tcloud, k_hint=1)
        r = tc.icp_point_to_plane_device(scloud, index, d_nrm, tc.IDENTITY, a.iters, None, -1.0)
        print(index.info(), r.translation)
        index.free()
elif a.what == "c5far":
    # the far regime of C5 at a size ncu can replay: C5's point density (250 / m^2), a source slab
    # lifted ~2.4-3 m off the surface at the cloud edge (roll about x), first iteration unseeded
    n = a.n or 4_000_000
    H = 316.0 * (n / 1e8) ** 0.5
    rng = np.random.default_rng(5)
    xy = rng.uniform(-H, H, (n, 2)).astype(np.float32)
    z = (0.5 * np.sin(0.3 * xy[:, 0]) * np.cos(0.2 * xy[:, 1]) + rng.normal(0, 0.005, n)).astype(np.float32)
    tgt = np.column_stack([xy, z]).astype(np.float32)
    gx = 0.15 * np.cos(0.3 * xy[:, 0]) * np.cos(0.2 * xy[:, 1])
    gy = -0.1 * np.sin(0.3 * xy[:, 0]) * np.sin(0.2 * xy[:, 1])
    nrm = np.column_stack([-gx, -gy, np.ones(n)])
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    m = n // 8
    sxy = np.column_stack([rng.uniform(-H, H, m), rng.uniform(0.75 * H, H, m)]).astype(np.float32)
    sz = (0.5 * np.sin(0.3 * sxy[:, 0]) * np.cos(0.2 * sxy[:, 1]) + 3.0 * sxy[:, 1] / H).astype(np.float32)
    src = np.column_stack([sxy, sz]).astype(np.float32)
    tcloud, scloud = tc.DeviceCloud(tgt, ctx), tc.DeviceCloud(src, ctx)
    d_nrm = ctx.alloc(n * 12)
    ctx.to_device(d_nrm, nrm)
    for _ in range(a.reps):
        index = tc.GridIndex(tcloud, k_hint=1)
        ctx.timer_start()
        r = tc.icp_point_to_plane_device(scloud, index, d_nrm, tc.IDENTITY, a.iters, None, -1.0)
        print(index.info()["dims"], index.info()["n_levels"], r.translation, f"{ctx.timer_stop():.3f} ms")
        index.free()
elif a.what == "gicp":
    n = a.n or 200_000
    src, tgt, _, _ = synth.scan_pair(n, half_extent=50.0 * (n / 1e6) ** 0.5)
    for _ in range(a.reps):
        r = tc.gicp(src, tgt, tc.IDENTITY, tc.GicpConfig(max_iterations=8), ctx, want_correspondences=False)
        print(r.iterations, r.translation)
elif a.what == "filters":
    pts = synth.kitti_frame()
    cloud = tc.DeviceCloud(pts, ctx)
    for _ in range(a.reps):
        for f in (lambda: tc.voxel_grid_filter(cloud, 0.2), lambda: tc.radius_outlier_removal(cloud, 0.5, 5),
                  lambda: tc.statistical_outlier_removal(cloud, 16, 1.0)):
            out = f()
            print(len(out))
            out.free()
