"""Oracle point-to-point ICP: the reference's inline tests (registration.rs:797-1140), restated."""
import numpy as np
import pytest

from fixtures import synth


def test_identity_converges_quickly(orc):
    # registration.rs:797-830: identical clouds converge within 3 iterations, mse ~ 0
    src = synth.terrain(2000, 4.0, seed=3)
    r = orc.icp_point_to_point(src, src, max_iters=10)
    assert r.converged and r.iterations <= 3 and r.mse < 1e-6


def test_recovers_bench_transform(orc):
    # the published benchmark's target (examples/threecrate_dataset_bench.rs:281-287)
    src = synth.terrain(5000, 6.0, seed=3)
    T = synth.bench_transform()
    r = orc.icp_point_to_point(src, synth.apply_iso(T, src), max_iters=50, conv=1e-9)
    assert np.linalg.norm(r.translation - T[:3]) < 1e-5
    assert min(np.linalg.norm(r.rotation - T[3:]), np.linalg.norm(r.rotation + T[3:])) < 1e-5


def test_validation(orc):
    # registration.rs:266-276, 653-669
    p = synth.terrain(100, 2.0, seed=1)
    e = np.zeros((0, 3), np.float32)
    with pytest.raises(ValueError):
        orc.icp_point_to_point(e, p)
    with pytest.raises(ValueError):
        orc.icp_point_to_point(p, p, max_iters=0)
    with pytest.raises(ValueError):
        orc.icp_point_to_point(p, p, conv=0.0)


def test_insufficient_correspondences(orc):
    p = synth.terrain(100, 2.0, seed=1)
    with pytest.raises(RuntimeError):
        orc.icp_point_to_point(p, p + np.float32(100.0), max_iters=5, max_dist=0.5)


def test_not_converged_reports_mse_under_final_transform(orc):
    # registration.rs:342-369
    src = synth.terrain(3000, 5.0, seed=5)
    tgt = synth.apply_iso(synth.bench_transform(), src)
    r = orc.icp_point_to_point(src, tgt, max_iters=2, conv=1e-12)
    assert not r.converged and r.iterations == 2 and len(r.correspondences) == 3000
    moved = synth.apply_iso(np.concatenate([r.translation, r.rotation]), src)
    mse = np.mean(((moved[r.correspondences[:, 0]] - tgt[r.correspondences[:, 1]]) ** 2).sum(1))
    assert abs(mse - r.mse) <= 1e-3 * max(mse, 1e-12) + 1e-10
