"""Parity at the sizes bench.py reports (BASELINE configs at FULL size), against the CPU oracle:

  C3  point-to-plane ICP, 1M <-> 1M, 30 iterations: conv = -1 (all 30 run) and the API default 1e-6
  C4  normals k = 30 on the 10M-point cloud, and the headline k = 16 on the same cloud:
      <= 1e-4 rad + sign on every well-conditioned point, kNN rows bit-exact on a 200k-query sample

The north-star bars are asserted as written: kNN indices bit-exact (modulo the documented rank-k
ties), normals <= 1e-4 rad with sign, ICP <= 1e-5 in rotation and translation with equal
iterations / converged and agreeing correspondences.  The oracle needs ~1 min per ICP run and
~20 s per 10M normals pass on 16 host threads."""
import numpy as np
import pytest

import threecrate_b200 as tc
from fixtures import synth
from gpu_util import angle, knn_parity, quat_angle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    return synth.scan_pair(1_000_000, half_extent=50.0)


@pytest.mark.parametrize("conv", [-1.0, 1e-6])
def test_c3_full_size_icp_parity(orc, c3, conv):
    src, tgt, nrm, T = c3
    r = tc.icp_point_to_plane_detailed(src, tgt, nrm, tc.IDENTITY, 30, None, conv)
    ref = orc.icp_point_to_plane(src, tgt, nrm, max_iters=30, conv=conv)
    rot = quat_angle(r.rotation, ref.rotation)
    tr = float(np.linalg.norm(r.translation.astype(np.float64) - ref.translation))
    same = (r.correspondences == ref.correspondences).all(axis=1).mean() \
        if len(r.correspondences) == len(ref.correspondences) else 0.0
    print(f"C3 full size conv={conv}: rot_err={rot:.3e} rad trans_err={tr:.3e} iterations "
          f"{r.iterations}/{ref.iterations} converged {r.converged}/{ref.converged} "
          f"mse {r.mse:.9e}/{ref.mse:.9e} pairs equal {same:.7f}")
    assert rot <= 1e-5 and tr <= 1e-5
    assert r.iterations == ref.iterations and r.converged == ref.converged
    # mse: the reference adds 1M squared residuals sequentially in f32 (registration.rs:461-469),
    # which alone is off the exact mean by ~5e-4 relative; the device reduces in f64.  Bar: 1e-4
    # relative against the f64 mean of the SAME residuals, and no farther from the reference's
    # value than that value's own accumulation error (+ 1e-4).
    exact = orc.last_icp_mse_f64()
    ref_err = abs(ref.mse - exact)
    print(f"    mse: device {r.mse:.9e} reference f32 {ref.mse:.9e} same residuals in f64 "
          f"{exact:.9e} (reference's own accumulation error {ref_err / exact:.2e} relative)")
    assert abs(r.mse - exact) <= 1e-4 * exact
    assert abs(r.mse - ref.mse) <= 1e-4 * abs(ref.mse) + 1.01 * ref_err
    assert len(r.correspondences) == len(ref.correspondences)
    assert same > 0.9999
    assert np.linalg.norm(r.translation - T[:3]) < 1e-3


@pytest.fixture(scope="module")
def c4():
    return synth.terrain(10_000_000, 100.0, seed=4, noise=0.002)


@pytest.mark.parametrize("k", [30, 16])
def test_c4_full_size_normals_parity(orc, c4, k):
    pts = c4
    got = tc.estimate_normals(pts, k)
    ref = orc.estimate_normals(pts, k)
    _, relgap = orc.normals_f64(pts, k)
    assert np.array_equal(got[:, :3], pts)
    ang = angle(got[:, 3:], ref[:, 3:])
    mn, mx = pts.min(0), pts.max(0)
    vp = (mn + mx) / 2 + np.array([0, 0, np.linalg.norm((mx - mn).astype(np.float32))])
    tv = vp - pts
    tv /= np.linalg.norm(tv, axis=1, keepdims=True)
    sign_noise = np.abs((ref[:, 3:] * tv).sum(1)) < 1e-4
    ang_eff = np.where(sign_noise, np.minimum(ang, np.pi - ang), ang)
    well = relgap >= 1e-3
    # stratified report: angular error by relative eigengap decade
    for lo, hi in ((1e-3, 1e-2), (1e-2, 1e-1), (1e-1, 10.0)):
        m = (relgap >= lo) & (relgap < hi)
        if m.any():
            print(f"C4 k={k} gap [{lo:g},{hi:g}): n={int(m.sum())} max={ang_eff[m].max():.3e} "
                  f"p99.9={np.percentile(ang_eff[m], 99.9):.3e}")
    bad = (ang_eff > 1e-4) & well
    # A point may legitimately differ when two candidates are BIT-EQUAL in d2 across the
    # neighbourhood boundary (rank k of kNN(k+1)): the reference keeps whichever its traversal met
    # first, the device the smaller index (documented tie rule, SURVEY a-3).  Every point over the
    # bar must be such a tie; they are counted and reported.
    tie_excused = 0
    if bad.any():
        tree = orc.OracleKdTree(pts)
        bi = np.nonzero(bad)[0]
        _, d2, _ = tree.knn_batch(pts[bi], k + 2)
        straddle = d2[:, k] == d2[:, k + 1]
        tie_excused = int(straddle.sum())
        for i, s_ in zip(bi, straddle):
            print(f"    point {i}: {ang_eff[i]:.3e} rad, gap {relgap[i]:.3f}, d2[k]={d2[list(bi).index(i), k]!r} "
                  f"d2[k+1]={d2[list(bi).index(i), k + 1]!r} tie straddling rank k: {bool(s_)}")
        assert straddle.all(), "a normal beyond 1e-4 rad that is not explained by a rank-k tie"
    print(f"C4 k={k}: n={len(pts)} max={ang_eff[well & ~bad].max():.3e} over_tol={int(bad.sum())} "
          f"(all rank-k ties: {tie_excused}) ill_conditioned={int((~well).sum())} "
          f"sign_noise={int(sign_noise.sum())}")
    assert well.mean() > 0.99
    assert np.allclose(np.linalg.norm(got[:, 3:].astype(np.float64), axis=1), 1.0, atol=1e-5)


@pytest.mark.parametrize("k", [31, 17])
def test_c4_full_size_knn_sample_bit_exact(orc, c4, k):
    """200k sample queries (points of the cloud, so the query itself is neighbour 0) against the
    kd-tree restatement on all 10M points: rows bit-exact, rank-k ties reported."""
    pts = c4
    rng = np.random.default_rng(k)
    q = pts[rng.choice(len(pts), 200_000, replace=False)]
    idx, dist, cnt = tc.KdTree(pts, k_hint=k - 1).knn(q, k)
    ri, rd2, _ = orc.OracleKdTree(pts).knn_batch(q, k)
    assert np.all(cnt == k)
    exact, modulo, mismatch = knn_parity(idx, dist, ri, rd2, pts, q)
    print(f"C4 kNN k={k}: exact={exact} modulo_ties={modulo} mismatch={len(mismatch)}")
    assert not mismatch
    assert exact > 0.9999 * len(q)
    same = idx == ri.astype(np.uint32)
    assert np.array_equal(dist[same], np.sqrt(rd2)[same])  # bitwise: sqrt of identical d2
