"""GPU point-to-point ICP parity against the oracle (1e-5 in rotation and translation)."""
import numpy as np
import pytest

import threecrate_b200 as tc
from fixtures import synth
from gpu_util import quat_angle

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _compare(got, ref):
    rot = quat_angle(got.rotation, ref.rotation)
    tr = float(np.linalg.norm(got.translation.astype(np.float64) - ref.translation))
    print(f"p2p ICP parity: rot_err={rot:.3e} trans_err={tr:.3e} iters={got.iterations}/{ref.iterations} "
          f"mse={got.mse:.6e}/{ref.mse:.6e}")
    assert rot <= TOL and tr <= TOL
    assert got.iterations == ref.iterations and got.converged == ref.converged
    assert abs(got.mse - ref.mse) <= 1e-3 * max(abs(ref.mse), 1e-12) + 1e-10
    # the oracle (like the reference) sums 1e5 f32 terms sequentially, so its pose differs from
    # the f64-reduced one at the 1e-6 level: a handful of 1-NN / max-distance decisions may flip
    assert abs(len(got.correspondences) - len(ref.correspondences)) <= 1e-4 * len(ref.correspondences) + 1


def test_identity_and_validation(orc):
    src = synth.terrain(2000, 4.0, seed=3)
    r = tc.icp_point_to_point(src, src, tc.IDENTITY, 10)
    assert r.converged and r.iterations <= 3 and r.mse < 1e-6
    e = np.zeros((0, 3), np.float32)
    with pytest.raises(tc.InvalidData):
        tc.icp_point_to_point(e, src)
    with pytest.raises(tc.InvalidData):
        tc.icp_point_to_point(src, src, tc.IDENTITY, 0)
    with pytest.raises(tc.InvalidData):
        tc.icp_point_to_point(src, src, tc.IDENTITY, 10, 0.0)
    with pytest.raises(tc.AlgorithmError):
        tc.icp_point_to_point(src, src + np.float32(100.0), tc.IDENTITY, 5, 1e-6, 0.5)
    # icp() swallows errors and returns init (registration.rs:238-241)
    init = np.array([1, 2, 3, 0, 0, 0, 1], np.float32)
    assert np.array_equal(tc.icp(e, src, init, 5), init)


def test_parity_bench_workload(orc):
    """The reference's benchmark workload: target = source moved by the bench transform."""
    src = synth.terrain(100_000, 16.0, seed=3, noise=0.002)
    tgt = synth.apply_iso(synth.bench_transform(), src)
    got = tc.icp_point_to_point(src, tgt, tc.IDENTITY, 10, 1e-5)
    ref = orc.icp_point_to_point(src, tgt, max_iters=10, conv=1e-5)
    _compare(got, ref)
    same = (got.correspondences == ref.correspondences).all(axis=1).mean()
    assert same > 0.999


def test_parity_two_scans_not_converged_and_max_distance(orc):
    src, tgt, _, T = synth.scan_pair(60_000, half_extent=12.0)
    got = tc.icp_detailed(src, tgt, tc.IDENTITY, 6, None, -1.0)   # never converges: final-mse path
    ref = orc.icp_point_to_point(src, tgt, max_iters=6, conv=-1.0, validate_conv=False)
    assert not got.converged and got.iterations == 6
    _compare(got, ref)
    got = tc.icp_detailed(src, tgt, tc.IDENTITY, 8, 0.05, 1e-7)
    ref = orc.icp_point_to_point(src, tgt, max_iters=8, conv=1e-7, max_dist=0.05)
    _compare(got, ref)


def test_not_converged_mse_counts_accepted_pairs_only(orc):
    """registration.rs:342-361: when the loop ends without converging the reported mse is over the
    ACCEPTED correspondences of the last iteration; pairs the max-distance test rejected keep
    seeding the next search but must not enter that mean."""
    src, tgt, _, T = synth.scan_pair(40_000, half_extent=10.0)
    # a third of the source is pushed 0.3 m off the surface: rejected at max distance 0.08
    src = src.copy()
    src[::3, 2] += np.float32(0.3)
    got = tc.icp_detailed(src, tgt, tc.IDENTITY, 5, 0.08, -1.0)
    ref = orc.icp_point_to_point(src, tgt, max_iters=5, conv=-1.0, max_dist=0.08, validate_conv=False)
    assert not got.converged and got.iterations == 5
    assert len(got.correspondences) < 0.8 * len(src)          # pairs really were rejected
    _compare(got, ref)


def test_gicp_negative_max_distance_rejects_every_pair():
    """GicpConfig.max_correspondence_distance is a plain f32: negative means no pair passes
    `dist > max` (gicp.rs:207-213) -> insufficient correspondences."""
    src, tgt, _, _ = synth.scan_pair(3000, half_extent=3.0)
    with pytest.raises(tc.AlgorithmError):
        tc.gicp(src, tgt, tc.IDENTITY, tc.GicpConfig(max_correspondence_distance=-1.0))
