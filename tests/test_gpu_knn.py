"""GPU kNN parity (bit-exact indices) against the oracle, through the C ABI."""
import numpy as np
import pytest

import threecrate_b200 as tc
from fixtures import synth
from gpu_util import knn_parity

pytestmark = pytest.mark.gpu
NO = 0xFFFFFFFF


def _brute(orc, pts, q, k):
    return orc.brute_knn(pts, q, k)


def test_cube_ties_follow_the_canonical_rule(orc):
    # nearest_neighbor.rs:429-436 fixture: every corner is at d2 = 0.75 from the centre
    t = tc.KdTree(synth.cube8())
    idx, dist = t.find_k_nearest([0.5, 0.5, 0.5], 3)
    assert idx.tolist() == [0, 1, 2]  # ascending (d2, index)
    assert np.all(dist == np.sqrt(np.float32(0.75)))


def test_edge_cases():
    t = tc.KdTree(synth.cube8())
    assert len(t.find_k_nearest([0, 0, 0], 0)[0]) == 0          # k == 0
    idx, dist = t.find_k_nearest([0, 0, 0], 20)                  # k > n -> n results
    assert len(idx) == 8 and np.all(np.diff(dist) >= 0) and idx[0] == 0
    e = tc.KdTree(np.zeros((0, 3), np.float32))
    assert len(e.find_k_nearest([0, 0, 0], 3)[0]) == 0          # empty tree
    one = tc.KdTree(np.array([[1, 2, 3]], np.float32))
    idx, dist = one.find_k_nearest([1, 2, 4], 5)
    assert idx.tolist() == [0] and dist[0] == 1.0
    with pytest.raises(tc.InvalidData):
        tc.KdTree(np.array([[np.nan, 0, 0]], np.float32))


@pytest.mark.parametrize("n,k", [(1000, 1), (5000, 5), (20000, 16), (20000, 30)])
def test_external_queries_bit_exact_vs_bruteforce(orc, n, k):
    rng = np.random.default_rng(n + k)
    pts = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    q = rng.uniform(-12, 12, (700, 3)).astype(np.float32)  # some queries outside the bbox
    q[:5] = [[100, 100, 100], [-50, 0, 0], [0, 0, 37], [10, 10, 10], [-10, -10, -10]]
    idx, dist, cnt = tc.KdTree(pts).knn(q, k)
    bi, bd2 = _brute(orc, pts, q, k)
    assert np.all(cnt == k)
    assert np.array_equal(idx.astype(np.uint64), bi)
    assert np.array_equal(dist, np.sqrt(bd2))  # bitwise: sqrt of identical d2


def test_self_knn_excludes_self_bit_exact(orc):
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(15000, 3)).astype(np.float32)
    k = 16
    idx, dist, cnt = tc.k_nearest_neighbors(pts, k)
    bi, bd2 = _brute(orc, pts, pts, k + 1)
    assert np.all(cnt == k)
    assert np.all(bi[:, 0] == np.arange(len(pts)))  # self first (d2 = 0, no duplicates)
    assert np.array_equal(idx.astype(np.uint64), bi[:, 1:])
    assert np.array_equal(dist, np.sqrt(bd2[:, 1:]))


def test_tie_dense_grid_matches_canonical_rule_exactly(orc):
    pts = synth.grid_plane(20, 0.1)
    idx, dist, cnt = tc.KdTree(pts).knn(pts, 9)
    bi, bd2 = _brute(orc, pts, pts, 9)
    assert np.array_equal(idx.astype(np.uint64), bi)
    # and agrees with the reference kd-tree modulo ties
    ki, kd2, _ = orc.OracleKdTree(pts).knn_batch(pts, 9)
    exact, modulo, mismatch = knn_parity(idx, dist, ki, kd2, pts, pts)
    assert not mismatch and modulo > 0


def test_duplicates_and_small_clouds(orc):
    rng = np.random.default_rng(9)
    base = rng.integers(0, 3, (400, 3)).astype(np.float32)  # many exact duplicates
    idx, dist, cnt = tc.k_nearest_neighbors(base, 7)
    bi, bd2 = _brute(orc, base, base, 8)
    for i in range(len(base)):  # k+1 search, drop own index, keep 7
        exp = [j for j in bi[i].tolist() if j != i][:7]
        assert idx[i].tolist() == exp
    # fewer points than k: n-1 neighbours, padded
    small = rng.normal(size=(5, 3)).astype(np.float32)
    idx, dist, cnt = tc.k_nearest_neighbors(small, 10)
    assert np.all(cnt == 4) and np.all(idx[:, 4:] == NO) and np.all(np.isinf(dist[:, 4:]))


def test_kitti_frame_k16_vs_reference_kdtree(orc):
    """BASELINE config 2: indices bit-exact vs the kd-tree restatement (modulo documented ties)."""
    pts = synth.kitti_frame()
    k = 16
    idx, dist, cnt = tc.k_nearest_neighbors(pts, k)
    ri, rdist, rcnt = orc.k_nearest_neighbors(pts, k, threads=0)
    assert np.all(cnt == k)
    rd2 = (rdist * rdist).astype(np.float32)
    exact, modulo, mismatch = knn_parity(idx, dist, ri, rd2, pts, pts)
    print(f"kNN C2: exact={exact} modulo_ties={modulo} mismatch={len(mismatch)}")
    assert not mismatch
    assert exact > 0.999 * len(pts)
    assert np.array_equal(dist[idx == ri.astype(np.uint32)], rdist[idx == ri.astype(np.uint32)])


def test_clustered_and_skewed_density(orc):
    """Ring expansion: dense blob + isolated far points + a line (degenerate extents)."""
    rng = np.random.default_rng(21)
    blob = rng.normal(scale=0.01, size=(4000, 3))
    far = rng.uniform(-50, 50, (60, 3))
    line = np.stack([np.linspace(-5, 5, 500), np.zeros(500), np.zeros(500)], 1)
    pts = np.concatenate([blob, far, line]).astype(np.float32)
    q = np.concatenate([pts[::7], rng.uniform(-60, 60, (100, 3)).astype(np.float32)])
    idx, dist, cnt = tc.KdTree(pts).knn(q, 11)
    bi, bd2 = _brute(orc, pts, q, 11)
    assert np.array_equal(idx.astype(np.uint64), bi)
    # all points on a plane / line / single point
    for cloud in (line.astype(np.float32), np.repeat(np.array([[1, 1, 1]], np.float32), 50, 0)):
        idx, dist, cnt = tc.KdTree(cloud).knn(cloud[:20], 5)
        bi, bd2 = _brute(orc, cloud, cloud[:20], 5)
        assert np.array_equal(idx.astype(np.uint64), bi)


def test_full_size_properties_1m():
    """Size-independent properties at 1M points: ascending rows, no self, symmetric sanity."""
    pts = synth.terrain(1_000_000, 50.0, seed=1)
    cloud = tc.DeviceCloud(pts)
    index = tc.GridIndex(cloud, k_hint=16)
    idx, dist, cnt = index.knn(None, 16, exclude_self=True)
    assert np.all(cnt == 16)
    assert np.all(np.diff(dist, axis=1) >= 0)
    assert not np.any(idx == np.arange(len(pts), dtype=np.uint32)[:, None])
    # distances recompute bit-exactly from the returned indices
    sel = np.random.default_rng(0).integers(0, len(pts), 2000)
    d = pts[idx[sel]] - pts[sel][:, None, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    assert np.array_equal(np.sqrt(d2.astype(np.float32)), dist[sel])
    info = index.info()
    assert info["n_points"] == len(pts) and info["occupied_cells"] > 0


def test_radius_search_matches_reference(orc):
    """KdTree::find_radius_neighbors (nearest_neighbor.rs:254-298): inclusive d2 <= r2, ascending."""
    rng = np.random.default_rng(31)
    pts = rng.uniform(-5, 5, (30000, 3)).astype(np.float32)
    t = tc.KdTree(pts)
    o = orc.OracleKdTree(pts)
    for q, r in [(pts[17], 0.6), ([0.1, 0.2, 0.3], 1.1), ([7.0, 7.0, 7.0], 4.0), ([0, 0, 0], 0.05),
                 ([100, 0, 0], 1.0)]:
        gi, gd = t.find_radius_neighbors(q, r)
        oi, od = o.find_radius_neighbors(q, r)
        assert len(gi) == len(oi)
        assert np.array_equal(np.sort(gi), np.sort(oi.astype(np.uint32)))
        assert np.array_equal(gd, np.sort(od))  # same distances, ascending
        assert np.all(np.diff(gd) >= 0) and np.all(gd <= np.float32(r) + 1e-7)
    # radius <= 0 and the reference's cube fixture (point_cloud_ops.rs:186-203 shape)
    assert len(t.find_radius_neighbors([0, 0, 0], 0.0)[0]) == 0
    assert len(t.find_radius_neighbors([0, 0, 0], -1.0)[0]) == 0
    cube = tc.KdTree(synth.cube8())
    idx, dist = cube.find_radius_neighbors([0, 0, 0], 1.0)
    assert sorted(idx.tolist()) == [0, 1, 2, 3]  # boundary points included


def test_kitti_strided_upload_matches_packed():
    """KITTI .bin records (x,y,z,intensity; stride 16, threecrate-io/src/lidar.rs:310-345) uploaded
    raw and de-interleaved on the device give the same cloud as the packed path."""
    raw = synth.kitti_frame(with_intensity=True)
    a = tc.GridIndex(tc.DeviceCloud(raw, stride_bytes=16), k_hint=16).estimate_normals(16)
    b = tc.estimate_normals(np.ascontiguousarray(raw[:, :3]), 16)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("k", [65, 100, 257])
def test_large_k_heap_path(orc, k):
    """k beyond the 64-entry register lists: global-memory heap kernels, same exact result."""
    pts = synth.kitti_frame(seed=9)[::40][:3000].copy()
    rng = np.random.default_rng(1)
    q = (pts[rng.integers(0, len(pts), 200)] + rng.normal(0, 0.2, (200, 3))).astype(np.float32)
    tree = tc.KdTree(pts, k_hint=k)
    idx, dist, cnt = tree.knn(q, k)
    ref_idx, ref_d2 = orc.brute_knn(pts, q, k)
    exact, modulo, mismatch = knn_parity(idx, dist, ref_idx, ref_d2, pts, q)
    assert not mismatch and exact + modulo == len(q)
    assert np.array_equal(dist, np.sqrt(ref_d2))
    assert np.all(cnt == k)
    # self-query with exclusion and k > N - 1: every other point, padded
    small = pts[:70]
    sidx, sdist, scnt = tc.k_nearest_neighbors(small, k)
    m = min(k, 69)
    assert np.all(scnt == m) and np.all(sidx[:, m:] == 0xFFFFFFFF)
    ridx, rdist, rcnt = orc.k_nearest_neighbors(small, k)
    assert np.array_equal(sdist[:, :m], rdist[:, :m])


@pytest.mark.parametrize("case", ["identical", "collinear", "far_offset", "tiny_extent", "two_clusters"])
def test_degenerate_geometry(orc, case):
    """Clouds that stress the grid sizing: zero / tiny extents, huge coordinates, empty space."""
    rng = np.random.default_rng(3)
    if case == "identical":
        pts = np.tile(np.float32([[1.5, -2.0, 0.25]]), (300, 1))
    elif case == "collinear":
        pts = np.zeros((500, 3), np.float32)
        pts[:, 0] = np.linspace(0, 10, 500, dtype=np.float32)
    elif case == "far_offset":
        pts = (rng.uniform(0, 5, (2000, 3)) + [4.0e5, -3.0e5, 1.0e4]).astype(np.float32)
    elif case == "tiny_extent":
        pts = (rng.uniform(0, 1e-5, (1500, 3)) + [1.0, 1.0, 1.0]).astype(np.float32)
    else:
        a = rng.normal(0, 0.05, (800, 3))
        b = rng.normal(0, 0.05, (800, 3)) + [5000.0, 0, 0]
        pts = np.vstack([a, b]).astype(np.float32)
    k = 9
    idx, dist, cnt = tc.KdTree(pts, k_hint=k).knn(pts[::7], k)
    ref_idx, ref_d2 = orc.brute_knn(pts, pts[::7], k)
    exact, modulo, mismatch = knn_parity(idx, dist, ref_idx, ref_d2, pts, pts[::7])
    assert not mismatch and exact + modulo == len(pts[::7])
    assert np.array_equal(dist, np.sqrt(ref_d2))
    # the pipeline on top survives the same clouds (normals are unit or the +z default)
    nrm = tc.estimate_normals(pts, 8)
    assert np.all(np.isfinite(nrm)) and np.allclose(np.linalg.norm(nrm[:, 3:], axis=1), 1.0, atol=1e-5)
    vox = tc.voxel_grid_filter(pts, 0.5)
    assert np.array_equal(vox.view(np.uint32), orc.voxel_grid_filter(pts, 0.5).view(np.uint32))


@pytest.mark.parametrize("k", [15, 16, 17, 31, 32])
@pytest.mark.parametrize("cloud", ["lattice", "duplicates", "lattice3d"])
def test_ties_wider_than_the_member_table_are_resolved_in_place(orc, cloud, k):
    """k + 1 == table width (k = 16 / 32 and their neighbours): ANY bit-equal d2 straddling rank k
    overflows the member table of the two-pass selection, which then collects the members in place
    (resolve_ties: strictly-below first, then the points at the k-th distance in ascending index).
    Lattices tie at every rank, duplicated points tie at distance zero; the rows must equal the
    canonical (d2, index) order of the brute force bit for bit."""
    rng = np.random.default_rng(12)
    pts = {"lattice": synth.grid_plane(40, 0.125),
           "duplicates": np.repeat(rng.uniform(-1, 1, (600, 3)).astype(np.float32), 5, axis=0),
           "lattice3d": np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32) * 0.25] * 3,
                                             indexing="ij"), -1).reshape(-1, 3)}[cloud]
    pts = np.ascontiguousarray(pts[rng.permutation(len(pts))])
    idx, dist, cnt = tc.KdTree(pts, k_hint=k).knn(pts, k)
    bi, bd2 = _brute(orc, pts, pts, k)
    assert np.array_equal(idx.astype(np.uint64), bi)
    assert np.array_equal(dist, np.sqrt(bd2))
    sidx, sdist, scnt = tc.k_nearest_neighbors(pts, k)      # kNN(k+1), own index dropped
    bi1, _ = _brute(orc, pts, pts, k + 1)
    for i in range(0, len(pts), 37):
        assert sidx[i].tolist() == [j for j in bi1[i].tolist() if j != i][:k]
    # the normals kernel runs the same selection: it must terminate and produce unit normals
    nrm = tc.estimate_normals(pts, k)[:, 3:]
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
