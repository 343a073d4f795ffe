"""Committed vectors (tests/golden/oracle_v1.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them bit for bit (the checker is frozen).
GPU: the CUDA path matches them without the oracle in the loop.
PARITY UNPINNED: the vectors are oracle output, not reference-binary output (no Rust here)."""
import importlib.util
import os

import numpy as np
import pytest

from gpu_util import quat_angle as _quat_angle

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "oracle_v1.npz"))


def test_oracle_reproduces_golden():
    spec = importlib.util.spec_from_file_location("make_golden",
                                                  os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.build()
    assert sorted(fresh) == sorted(G.files)
    for k in G.files:
        a, b = G[k], np.asarray(fresh[k])
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.array_equal(a, b, equal_nan=True), k


@pytest.mark.gpu
def test_gpu_knn_matches_golden():
    import threecrate_b200 as tc
    tree = tc.KdTree(G["knn_points"], k_hint=8)
    idx, dist, cnt = tree.knn(G["knn_queries"], 8)
    assert np.array_equal(idx, G["knn_idx"])  # no ties in this fixture
    assert np.array_equal(dist, np.sqrt(G["knn_d2"]))
    sidx, sdist, _ = tc.k_nearest_neighbors(G["knn_points"], 5)
    assert np.array_equal(np.sort(sidx, 1), np.sort(G["selfknn_idx"], 1))
    assert np.array_equal(sdist, G["selfknn_dist"])


@pytest.mark.gpu
def test_gpu_normals_match_golden():
    import threecrate_b200 as tc
    got = tc.estimate_normals(G["normals_points"], 10)
    ref = G["normals_k10"]
    assert np.array_equal(got[:, :3], ref[:, :3])
    a, b = got[:, 3:].astype(np.float64), ref[:, 3:].astype(np.float64)
    ang = np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), (a * b).sum(1))
    # same bar as tests/test_gpu_normals.py: <= 1e-4 rad with matching sign wherever the
    # eigenvector is determined (relative eigengap >= 1e-3) and the orientation is not decided
    # by rounding noise (|n . view direction| >= 1e-4; there the unsigned angle is compared)
    ang = np.where(G["normals_k10_sign_noise"], np.minimum(ang, np.pi - ang), ang)
    well = G["normals_k10_wellcond"]
    assert well.mean() > 0.9
    assert ang[well].max() <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["plane", "point", "gicp"])
def test_gpu_icp_matches_golden(kind):
    import threecrate_b200 as tc
    src, tgt, nrm = G["icp_src"], G["icp_tgt"], G["icp_tgt_normals"]
    if kind == "plane":
        r = tc.icp_point_to_plane(src, tgt, nrm, tc.IDENTITY, 20)
    elif kind == "point":
        r = tc.icp_point_to_point(src, tgt, tc.IDENTITY, 20, 1e-6, None)
    else:
        r = tc.gicp(src, tgt, tc.IDENTITY, tc.GicpConfig(max_iterations=15))
    T, meta = G[f"icp_{kind}_T"], G[f"icp_{kind}_meta"]
    assert np.linalg.norm(r.translation.astype(np.float64) - T[:3]) <= 1e-5
    assert _quat_angle(r.rotation, T[3:]) <= 1e-5
    assert r.iterations == int(meta[1]) and r.converged == bool(meta[2])
    assert abs(r.mse - meta[0]) <= 1e-3 * abs(meta[0]) + 1e-10
    assert abs(len(r.correspondences) - int(meta[3])) <= 2


@pytest.mark.gpu
def test_gpu_filters_match_golden():
    import threecrate_b200 as tc
    f = G["filter_points"]
    assert np.array_equal(tc.voxel_grid_filter(f, 0.5).view(np.uint32), G["voxel_0p5"].view(np.uint32))
    assert np.array_equal(tc.radius_outlier_removal(f, 1.0, 3), f[G["radius_1p0_min3_mask"]])
    got, st = tc.statistical_outlier_removal(f, 8, 1.0, return_stats=True)
    assert np.array_equal(got, f[G["sor_k8_mask"]])
    assert np.array_equal(np.float32([st["mean"], st["std_dev"], st["threshold"]]), G["sor_k8_stats"])


# ------------------------------------------------------------------------------------------
# Reference-pinned vectors: produced by rust/golden_dumper/dump_golden.rs INSIDE the reference
# checkout (needs cargo, absent from the build image).  Until tests/golden/reference_v1.bin is
# committed these tests skip and parity stays "unpinned"; once it exists they hold BOTH the
# oracle restatement and the CUDA path to the reference's real output.
# ------------------------------------------------------------------------------------------
_REF = os.path.join(HERE, "golden", "reference_v1.bin")
_need_ref = pytest.mark.skipif(not os.path.exists(_REF),
                               reason="parity unpinned: tests/golden/reference_v1.bin not generated yet "
                                      "(run rust/golden_dumper/dump_golden.rs in the reference checkout)")


def _load_ref():
    spec = importlib.util.spec_from_file_location(
        "make_reference_inputs", os.path.join(HERE, "golden", "make_reference_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.read_reference(_REF), mod.CASES


def _check_against_reference(knn_fn, normals_fn, icp_fn):
    ref, cases = _load_ref()
    for r, c in zip(ref, cases):
        if r["kind"] == 1:
            pts, k = c[2], r["k"]
            idx, dist = knn_fn(pts, k + 1)
            # canonical (d2, index) order on both sides; exact wherever the distances are distinct
            distinct = np.all(np.diff(r["dist"], axis=1) > 0, axis=1)
            assert np.array_equal(idx[distinct].astype(np.uint64), r["idx"][distinct])
            assert np.array_equal(dist[distinct], r["dist"][distinct])
            got = normals_fn(pts, k)
            a, b = got[:, 3:].astype(np.float64), r["normals"][:, 3:].astype(np.float64)
            ang = np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), (a * b).sum(1))
            assert np.percentile(ang, 99) <= 1e-4
        else:
            got = icp_fn(c[2], c[3], c[4], c[5])
            assert np.linalg.norm(got.translation.astype(np.float64) - r["T"][:3]) <= 1e-5
            assert _quat_angle(got.rotation, r["T"][3:]) <= 1e-5
            assert got.iterations == r["iterations"] and got.converged == r["converged"]


@_need_ref
def test_oracle_matches_reference_pinned_vectors():
    import oracle

    def knn(pts, k):
        i, d2 = oracle.brute_knn(pts, pts, k)
        return i, np.sqrt(d2)
    _check_against_reference(knn, oracle.estimate_normals,
                             lambda s, t, n, it: oracle.icp_point_to_plane(s, t, n, max_iters=it, conv=-1.0))


@_need_ref
@pytest.mark.gpu
def test_gpu_matches_reference_pinned_vectors():
    import threecrate_b200 as tc

    def knn(pts, k):
        i, d, _ = tc.KdTree(pts, k_hint=k).knn(pts, k)
        return i, d
    _check_against_reference(knn, tc.estimate_normals,
                             lambda s, t, n, it: tc.icp_point_to_plane_detailed(s, t, n, tc.IDENTITY, it,
                                                                                None, -1.0))
