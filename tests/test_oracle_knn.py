"""Oracle kNN: the reference's own inline tests (nearest_neighbor.rs:408-727,
point_cloud_ops.rs:146-239) re-run against the C++ restatement, plus scipy cross-checks."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from fixtures import synth


def _canon(idx, d2):
    o = np.lexsort((idx, d2))
    return idx[o], d2[o]


def test_cube_k3_all_ties(orc):
    # nearest_neighbor.rs:429-436 + :461-483: distances equal, indices not required to match
    t = orc.OracleKdTree(synth.cube8())
    idx, dist, d2 = t.find_k_nearest([0.5, 0.5, 0.5], 3)
    assert len(idx) == 3
    assert np.all(d2 == np.float32(0.75))
    assert np.allclose(dist, np.sqrt(0.75))
    assert len(set(idx.tolist())) == 3


def test_k_zero_k_gt_n_empty(orc):
    # nearest_neighbor.rs:540-563 edge cases
    t = orc.OracleKdTree(synth.cube8())
    assert len(t.find_k_nearest([0, 0, 0], 0)[0]) == 0
    idx, dist, _ = t.find_k_nearest([0, 0, 0], 20)
    assert len(idx) == 8 and np.all(np.diff(dist) >= 0)
    e = orc.OracleKdTree(np.zeros((0, 3), np.float32))
    assert len(e.find_k_nearest([0, 0, 0], 3)[0]) == 0


def test_radius_edge_cases(orc):
    t = orc.OracleKdTree(synth.cube8())
    assert len(t.find_radius_neighbors([0, 0, 0], 0.0)[0]) == 0
    assert len(t.find_radius_neighbors([0, 0, 0], -1.0)[0]) == 0
    idx, dist = t.find_radius_neighbors([0.5, 0.5, 0.0], 1.0)  # point_cloud_ops.rs:186-203 shape
    assert np.all(dist <= 1.0) and np.all(np.diff(dist) >= 0)
    idx, _ = t.find_radius_neighbors([0, 0, 0], 1.0)
    assert sorted(idx.tolist()) == [0, 1, 2, 3]  # d2 <= r2 is inclusive (nearest_neighbor.rs:270)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_kd_equals_bruteforce_random(orc, seed):
    # nearest_neighbor.rs:566-641: 100 random points x 10 random queries, distances within 1e-6
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-10, 10, (100, 3)).astype(np.float32)
    q = rng.uniform(-10, 10, (10, 3)).astype(np.float32)
    t = orc.OracleKdTree(pts)
    for k in (1, 3, 5, 10):
        ki, kd2, cnt = t.knn_batch(q, k)
        bi, bd2 = orc.brute_knn(pts, q, k)
        assert np.all(cnt == k)
        assert np.all(np.abs(np.sqrt(kd2) - np.sqrt(bd2)) < 1e-6)
        assert np.array_equal(kd2, bd2)
        assert np.array_equal(ki, bi)  # random floats: no ties, so indices agree too


def test_kd_vs_scipy_sets(orc):
    rng = np.random.default_rng(7)
    pts = rng.normal(size=(20000, 3)).astype(np.float32)
    q = rng.normal(size=(500, 3)).astype(np.float32)
    ki, kd2, _ = orc.OracleKdTree(pts).knn_batch(q, 17)
    _, si = cKDTree(pts.astype(np.float64)).query(q.astype(np.float64), 17)
    # f64 ordering may legitimately differ from f32-rounded d2 ordering at near-ties: compare
    # rows whose f32 d2 values (including the 18th) are well separated
    ki18, kd218, _ = orc.OracleKdTree(pts).knn_batch(q, 18)
    gaps = np.min(np.diff(kd218, axis=1) / kd218[:, 1:], axis=1)
    ok = gaps > 1e-5
    assert ok.sum() > 400
    assert np.array_equal(ki[ok].astype(np.int64), si[ok])


def test_kd_ties_canonicalise_to_bruteforce(orc):
    # regular grid => many exact d2 ties. Membership may differ only at the rank-k boundary.
    pts = synth.grid_plane(20, 0.1)
    t = orc.OracleKdTree(pts)
    k = 9
    ki, kd2, _ = t.knn_batch(pts, k)
    bi, bd2 = orc.brute_knn(pts, pts, k)
    assert np.array_equal(kd2, bd2)  # the multiset of distances is unique
    n_boundary = 0
    for r in range(pts.shape[0]):
        a, _ = _canon(ki[r], kd2[r])
        if not np.array_equal(a, bi[r]):
            # differing members must all sit at the k-th distance (tie straddling rank k)
            diff = set(a.tolist()) ^ set(bi[r].tolist())
            for j in diff:
                d = np.float32(((pts[j] - pts[r]) ** 2).sum())
                assert abs(d - bd2[r, -1]) <= 1e-6
            n_boundary += 1
    assert n_boundary > 0  # the grid really does exercise the tie rule


def test_k_nearest_neighbors_excludes_self(orc):
    # point_cloud_ops.rs:153-167 (ignored upstream): 4 points, k=2 -> 2 neighbours each, no self
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], np.float32)
    idx, dist, cnt = orc.k_nearest_neighbors(pts, 2)
    assert idx.shape == (4, 2) and np.all(cnt == 2)
    for i in range(4):
        assert i not in idx[i].tolist()
        assert dist[i, 0] <= dist[i, 1]
    # k >= n: n-1 neighbours
    idx, dist, cnt = orc.k_nearest_neighbors(pts, 10)
    assert np.all(cnt == 3)
    assert orc.k_nearest_neighbors(np.zeros((0, 3), np.float32), 3)[0].shape[0] == 0
    assert orc.k_nearest_neighbors(pts, 0)[0].shape[0] == 0


def test_binaryheap_emulation_sorted_output(orc):
    rng = np.random.default_rng(3)
    pts = rng.integers(0, 4, (300, 3)).astype(np.float32)  # heavy ties
    t = orc.OracleKdTree(pts)
    for q in pts[:50]:
        idx, dist, d2 = t.find_k_nearest(q, 12)
        assert np.all(np.diff(d2) >= 0)
        b, bd2 = orc.brute_knn(pts, q[None], 12)
        assert np.array_equal(d2, bd2[0])


def test_fuzz_kdtree_distances_equal_bruteforce(orc):
    """Property test (hypothesis): on small clouds full of duplicates and lattice ties the
    kd-tree restatement returns the same multiset of distances as the canonical brute force, for
    every k, and radius queries return exactly the points within the radius."""
    from hypothesis import given, settings, strategies as st

    coord = st.integers(min_value=-3, max_value=3).map(lambda v: np.float32(v) * np.float32(0.25))
    point = st.tuples(coord, coord, coord)

    @settings(max_examples=60, deadline=None)
    @given(st.lists(point, min_size=1, max_size=40), point, st.integers(min_value=1, max_value=45),
           st.integers(min_value=0, max_value=6))
    def check(pts, q, k, rq):
        p = np.array(pts, np.float32)
        tree = orc.OracleKdTree(p)
        idx, dist, d2 = tree.find_k_nearest(np.array(q, np.float32), k)
        bi, bd2 = orc.brute_knn(p, np.array([q], np.float32), min(k, len(p)))
        assert len(idx) == min(k, len(p))
        assert np.array_equal(np.sort(d2), np.sort(bd2[0]))
        assert np.all(np.diff(dist) >= 0)
        radius = np.float32(rq) * np.float32(0.25)
        ridx, rdist = tree.find_radius_neighbors(np.array(q, np.float32), float(radius))
        dd = p - np.array(q, np.float32)
        ref = np.nonzero((dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]
                         <= radius * radius)[0]
        if radius <= 0:  # nearest_neighbor.rs:255-257: a non-positive radius finds nothing
            assert len(ridx) == 0
        else:
            assert sorted(ridx.tolist()) == ref.tolist()

    check()
