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
