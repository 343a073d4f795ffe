#!/usr/bin/env python
"""Regenerates tests/golden/oracle_v1.npz: small fixed inputs and the ORACLE's outputs for them.

PARITY UNPINNED: the reference is Rust-only and cannot run in the build image, so these vectors
come from the C++/numpy restatement under oracle/ (itself pinned by the reference's inline test
assertions), not from the reference binary.  They serve two purposes: (1) freeze the checker -
tests/test_golden.py fails if the oracle's behaviour drifts; (2) let the -m gpu suite compare
the CUDA path with committed vectors without running the oracle.

    python tests/golden/make_golden.py        # from the repo root
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from fixtures import synth  # noqa: E402


def build():
    g = {}
    # ---- kNN: 600-point LiDAR-like subset, 64 external queries, k = 8; self-kNN k = 5
    pts = synth.kitti_frame(seed=21)[::200][:600].copy()
    rng = np.random.default_rng(7)
    q = (pts[rng.integers(0, len(pts), 64)] + rng.normal(0, 0.3, (64, 3))).astype(np.float32)
    idx, d2 = oracle.brute_knn(pts, q, 8)
    g["knn_points"], g["knn_queries"], g["knn_idx"], g["knn_d2"] = pts, q, idx.astype(np.uint32), d2
    sidx, sdist, scnt = oracle.k_nearest_neighbors(pts, 5)
    g["selfknn_idx"], g["selfknn_dist"] = sidx.astype(np.uint32), sdist
    # ---- normals: noisy terrain with a wall, k = 10 (oriented towards the default viewpoint)
    t = synth.terrain(1500, 3.0, seed=5, noise=0.004, wall_fraction=0.2)
    g["normals_points"] = t
    g["normals_k10"] = oracle.estimate_normals(t, 10)
    # neighbourhoods whose two smallest eigenvalues are separated (relative gap >= 1e-3): only
    # there is the eigenvector determined beyond f32 rounding (tests/test_gpu_normals.py)
    g["normals_k10_wellcond"] = oracle.normals_f64(t, 10)[1] >= 1e-3
    nr = g["normals_k10"]
    mn, mx = t.min(0), t.max(0)
    vp = (mn + mx) / 2 + np.array([0, 0, np.linalg.norm((mx - mn).astype(np.float32))])
    tv = vp - t
    tv /= np.linalg.norm(tv, axis=1, keepdims=True)
    g["normals_k10_sign_noise"] = np.abs((nr[:, 3:] * tv).sum(1)) < 1e-4
    # ---- point-to-plane / point-to-point ICP and GICP on a 3000-point pair
    src, tgt, nrm, T = synth.scan_pair(3000, half_extent=3.0, noise=0.003)
    g["icp_src"], g["icp_tgt"], g["icp_tgt_normals"], g["icp_T_gt"] = src, tgt, nrm, T
    for name, r in (("plane", oracle.icp_point_to_plane(src, tgt, nrm, None, 20)),
                    ("point", oracle.icp_point_to_point(src, tgt, None, 20, 1e-6, None)),
                    ("gicp", oracle.gicp(src, tgt, None, 15, 1.0, 1e-6, 20))):
        g[f"icp_{name}_T"] = np.concatenate([r.translation, r.rotation]).astype(np.float32)
        g[f"icp_{name}_meta"] = np.array([r.mse, r.iterations, float(r.converged),
                                          len(r.correspondences)], np.float64)
    # ---- filters on the kNN cloud + a few far outliers
    f = np.vstack([pts, [[300, 0, 0], [0, -300, 5], [40, 40, 80]]]).astype(np.float32)
    g["filter_points"] = f
    g["voxel_0p5"] = oracle.voxel_grid_filter(f, 0.5)
    g["radius_1p0_min3_mask"] = oracle.radius_outlier_removal(f, 1.0, 3, return_mask=True)[1]
    _, det = oracle.statistical_outlier_removal(f, 8, 1.0, return_details=True)
    g["sor_k8_mask"] = det["mask"]
    g["sor_k8_stats"] = np.array([det["mean"], det["std_dev"], det["threshold"]], np.float32)
    return g


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz")
    np.savez_compressed(out, **build())
    print(out, os.path.getsize(out), "bytes")
