"""Oracle filters vs the reference's own inline test assertions
(threecrate-algorithms/src/filtering.rs:396-660) plus independent numpy cross-checks."""
import numpy as np
import pytest

import oracle
from fixtures import synth


def _cube(n, step=0.1):
    g = np.arange(n, dtype=np.float32) * np.float32(step)
    return np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)


def _has(points, p, tol=0.1):
    return bool(np.any(np.all(np.abs(points - np.asarray(p, np.float32)) < tol, axis=1)))


# ---- statistical outlier removal (filtering.rs:400-520) --------------------------------------
def test_sor_empty_and_single():
    assert len(oracle.statistical_outlier_removal(np.empty((0, 3), np.float32), 5, 1.0)) == 0
    assert len(oracle.statistical_outlier_removal(np.zeros((1, 3), np.float32), 1, 1.0)) == 1


def test_sor_with_outliers():  # filtering.rs:418-457
    pts = np.vstack([_cube(10), [[10, 10, 10], [-10, -10, -10], [5, 5, 5]]]).astype(np.float32)
    out = oracle.statistical_outlier_removal(pts, 5, 1.0)
    assert 0 < len(out) < len(pts)
    assert not _has(out, [10, 10, 10]) and not _has(out, [-10, -10, -10])


def test_sor_no_outliers():  # filtering.rs:460-480
    pts = _cube(5)
    assert len(oracle.statistical_outlier_removal(pts, 5, 1.0)) > len(pts) * 8 // 10


def test_sor_invalid():  # filtering.rs:483-497, 525-532
    one = np.zeros((1, 3), np.float32)
    for bad in ((0, 1.0), (5, 0.0), (5, -1.0)):
        with pytest.raises(oracle.InvalidData):
            oracle.statistical_outlier_removal(one, *bad)
    for bad in ((0, 1.0), (5, 0.0)):
        with pytest.raises(oracle.InvalidData):
            oracle.statistical_outlier_removal_with_threshold(one, *bad)


def test_sor_with_threshold():  # filtering.rs:500-522
    pts = np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1], [10, 10, 10]], np.float32)
    out = oracle.statistical_outlier_removal_with_threshold(pts, 3, 0.5)
    assert len(out) == 4 and not _has(out, [10, 10, 10])


def test_sor_mean_skips_duplicates_of_the_point():
    # a duplicated point is not its own neighbour: both copies see the other three at 1.0
    pts = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    m = oracle.sor_mean_distances(pts, 3)
    assert m[0] == m[1] == np.float32(1.0)


# ---- voxel grid (filtering.rs:534-574) -------------------------------------------------------
def test_voxel_empty_single_invalid():
    assert len(oracle.voxel_grid_filter(np.empty((0, 3), np.float32), 0.1)) == 0
    assert len(oracle.voxel_grid_filter(np.zeros((1, 3), np.float32), 0.1)) == 1
    for bad in (0.0, -1.0):
        with pytest.raises(oracle.InvalidData):
            oracle.voxel_grid_filter(np.zeros((1, 3), np.float32), bad)


def test_voxel_duplicates():  # filtering.rs:550-564
    pts = np.array([[0, 0, 0], [0, 0, 0], [0.1, 0, 0], [0.1, 0, 0], [0, 0.1, 0]], np.float32)
    assert len(oracle.voxel_grid_filter(pts, 0.05)) == 3


def test_voxel_centroids_against_pandas_groupby():
    import pandas as pd
    pts = synth.kitti_frame(seed=3)[:20000]
    vs = 0.5
    out, coords = oracle.voxel_grid_filter(pts, vs, return_coords=True)
    mn = pts.min(0)
    c = np.floor((pts - mn) / np.float32(vs)).astype(np.int64)
    df = pd.DataFrame({"x": c[:, 0], "y": c[:, 1], "z": c[:, 2], "px": pts[:, 0].astype(np.float64),
                       "py": pts[:, 1].astype(np.float64), "pz": pts[:, 2].astype(np.float64)})
    g = df.groupby(["z", "y", "x"], sort=True)[["px", "py", "pz"]].mean().to_numpy()
    assert out.shape == g.shape
    np.testing.assert_allclose(out, g, rtol=0, atol=1e-5)
    assert np.array_equal(coords[:, ::-1], np.array(sorted(map(tuple, coords[:, ::-1]))))


# ---- radius outlier removal (filtering.rs:577-660) -------------------------------------------
def test_radius_empty_single():
    assert len(oracle.radius_outlier_removal(np.empty((0, 3), np.float32), 0.5, 3)) == 0
    assert len(oracle.radius_outlier_removal(np.zeros((1, 3), np.float32), 0.5, 1)) == 0


def test_radius_with_outliers():  # filtering.rs:593-630
    g = np.arange(5, dtype=np.float32) * np.float32(0.1)
    plane = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    pts = np.vstack([np.c_[plane, np.zeros(25)], [[10, 10, 10], [-10, -10, -10]]]).astype(np.float32)
    out = oracle.radius_outlier_removal(pts, 0.5, 2)
    assert 0 < len(out) < len(pts)
    assert not _has(out, [10, 10, 10]) and not _has(out, [-10, -10, -10])


def test_radius_invalid():
    one = np.zeros((1, 3), np.float32)
    for bad in ((0.0, 3), (-1.0, 3), (0.5, 0)):
        with pytest.raises(oracle.InvalidData):
            oracle.radius_outlier_removal(one, *bad)


def test_radius_counts_against_kdtree_radius_query():
    pts = synth.terrain(1500, 4.0, seed=5, noise=0.01)
    _, keep = oracle.radius_outlier_removal(pts, 0.3, 4, return_mask=True)
    tree = oracle.OracleKdTree(pts)
    for i in range(0, 1500, 37):
        idx, _ = tree.find_radius_neighbors(pts[i], 0.3)
        assert keep[i] == (len(idx) - 1 >= 4)


# ---- multiscale ICP (registration.rs:704-789) ------------------------------------------------
def test_multiscale_validation():
    a = synth.terrain(200, 2.0, seed=1)
    with pytest.raises(oracle.InvalidData):
        oracle.multiscale_icp_point_to_point(np.empty((0, 3), np.float32), a)
    with pytest.raises(oracle.InvalidData):
        oracle.multiscale_icp_point_to_point(a, a, levels=())
    with pytest.raises(oracle.InvalidData):
        oracle.multiscale_icp_point_to_point(a, a, convergence_threshold=0.0)
    with pytest.raises(oracle.InvalidData):
        oracle.multiscale_icp_point_to_point(a, a, final_refinement_iterations=0)
    with pytest.raises(oracle.InvalidData):
        oracle.multiscale_icp_point_to_point(a, a, levels=((0.0, 5, None),))
    with pytest.raises(oracle.InvalidData):
        oracle.multiscale_icp_point_to_point(a, a, levels=((0.2, 0, None),))
    with pytest.raises(oracle.AlgorithmError):  # every level collapses to < 3 voxels
        oracle.multiscale_icp_point_to_point(a, a, levels=((100.0, 5, None),))


def test_multiscale_recovers_known_transform():
    src, tgt, _, T = synth.scan_pair(6000, half_extent=6.0, copy_variant=True)
    r = oracle.multiscale_icp_point_to_point(src, tgt, levels=((0.8, 10, 2.0), (0.4, 10, 1.0)),
                                             final_refinement_iterations=15,
                                             final_max_correspondence_distance=0.5)
    assert np.linalg.norm(r.translation - T[:3]) < 2e-2
    assert r.iterations > 15 or r.converged


def test_sor_against_scipy_ckdtree():
    """Independent f64 restatement with scipy: same mean distances (to f32 rounding) and the same
    keep / drop decision everywhere except within rounding of the threshold."""
    from scipy.spatial import cKDTree
    pts = synth.terrain(3000, 4.0, seed=8, noise=0.01)
    rng = np.random.default_rng(0)
    pts = np.vstack([pts, rng.uniform(-4, 4, (30, 3)) + [0, 0, 3]]).astype(np.float32)
    k = 10
    _, det = oracle.statistical_outlier_removal(pts, k, 1.5, return_details=True)
    d, _ = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k + 1)
    mean64 = d[:, 1:].mean(axis=1)  # no duplicate points in this cloud: column 0 is the point itself
    np.testing.assert_allclose(det["mean_distances"], mean64, rtol=2e-6)
    thr64 = mean64.mean() + 1.5 * mean64.std()
    assert abs(det["threshold"] - thr64) <= 1e-5 * thr64
    clear = np.abs(mean64 - thr64) > 1e-4 * thr64
    assert np.array_equal(det["mask"][clear], (mean64 <= thr64)[clear])


def test_radius_outlier_against_scipy_ckdtree():
    from scipy.spatial import cKDTree
    pts = synth.kitti_frame(seed=2)[::40][:3000].copy()
    r, m = 0.6, 4
    _, keep = oracle.radius_outlier_removal(pts, r, m, return_mask=True)
    p64 = pts.astype(np.float64)
    tree = cKDTree(p64)
    cnt = np.array([len(x) for x in tree.query_ball_point(p64, r)]) - 1
    # points whose m-th neighbour sits within f32 rounding of the radius may differ
    d, _ = tree.query(p64, m + 1)
    clear = np.abs(d[:, m] - r) > 1e-5
    assert np.array_equal(keep[clear], (cnt >= m)[clear])
