"""Oracle point-to-plane ICP: the reference's inline tests (registration.rs:1144-1267)."""
import numpy as np
import pytest

from fixtures import synth


def test_identity(orc):
    # registration.rs:1167-1177
    p, n = synth.fibonacci_sphere(50)
    r = orc.icp_point_to_plane(p, p, n, max_iters=20)
    assert r.converged and r.mse < 1e-6


def test_translation(orc):
    # registration.rs:1179-1196
    p, n = synth.fibonacci_sphere(100)
    shift = np.array([0.15, 0, 0], np.float32)
    r = orc.icp_point_to_plane(p, p + shift, n, max_iters=50)
    assert np.linalg.norm(r.translation - shift) < 0.3
    assert r.mse < 0.1
    assert len(r.correspondences) == 100


def test_validation(orc):
    # registration.rs:1198-1216
    p, n = synth.fibonacci_sphere(20)
    with pytest.raises(ValueError):
        orc.icp_point_to_plane(p, p, np.array([[0, 0, 1]], np.float32), max_iters=10)
    with pytest.raises(ValueError):
        orc.icp_point_to_plane(np.zeros((0, 3), np.float32), p, n, max_iters=10)
    with pytest.raises(ValueError):
        orc.icp_point_to_plane(p, p, n, max_iters=0)


def test_convergence_vs_shift(orc):
    # registration.rs:1218-1251 (p2plane half)
    p, n = synth.fibonacci_sphere(80)
    shift = np.array([0.1, 0.05, 0], np.float32)
    r = orc.icp_point_to_plane(p, p + shift, n, max_iters=50)
    assert np.linalg.norm(r.translation) > 0.05
    assert r.converged or r.mse < 0.1


def test_max_distance(orc):
    # registration.rs:1253-1267
    p, n = synth.fibonacci_sphere(50)
    r = orc.icp_point_to_plane(p, p + np.array([0.1, 0, 0], np.float32), n, max_iters=30,
                               max_dist=5.0)
    assert r.mse < 0.5


def test_insufficient_correspondences(orc):
    # registration.rs:568-572: < 6 valid pairs -> Algorithm error
    p, n = synth.fibonacci_sphere(50)
    with pytest.raises(RuntimeError):
        orc.icp_point_to_plane(p, p + np.float32(100.0), n, max_iters=5, max_dist=0.5)


def test_not_converged_returns_prev_mse_and_max_iters(orc):
    # registration.rs:595-601 with conv <= 0 (never converges)
    p, n = synth.fibonacci_sphere(100)
    r = orc.icp_point_to_plane(p, p + np.array([0.1, 0, 0], np.float32), n, max_iters=7, conv=-1.0)
    assert not r.converged and r.iterations == 7 and len(r.correspondences) == 100


def test_recovers_bench_transform(orc):
    src, tgt, nrm, T = synth.scan_pair(20000, half_extent=8.0, copy_variant=True, noise=0.0)
    r = orc.icp_point_to_plane(src, tgt, nrm, max_iters=30, conv=-1.0)
    assert np.linalg.norm(r.translation - T[:3]) < 2e-3
    dq = min(np.linalg.norm(r.rotation - T[3:]), np.linalg.norm(r.rotation + T[3:]))
    assert dq < 1e-3


def test_nalgebra_restatements(orc):
    rng = np.random.default_rng(5)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        iso = np.concatenate([rng.normal(size=3), q]).astype(np.float32)
        p = rng.normal(size=3).astype(np.float32)
        ref = synth.apply_iso(iso, p[None])[0]
        assert np.allclose(orc.iso_apply(iso, p), ref, atol=2e-6)
        q2 = rng.normal(size=4)
        q2 /= np.linalg.norm(q2)
        iso2 = np.concatenate([rng.normal(size=3), q2]).astype(np.float32)
        comp = orc.iso_mul(iso, iso2)
        assert np.allclose(orc.iso_apply(comp, p), orc.iso_apply(iso, orc.iso_apply(iso2, p)),
                           atol=1e-5)
        a = rng.normal(size=(20, 6)).astype(np.float32)
        ata = (a.T @ a).astype(np.float32)
        b = rng.normal(size=6).astype(np.float32)
        x, path = orc.solve6(ata, b)
        assert path == 0
        assert np.allclose(x, np.linalg.solve(ata.astype(np.float64), b), rtol=2e-3, atol=2e-4)
    x, path = orc.solve6(np.zeros((6, 6), np.float32), np.ones(6, np.float32))
    assert path == 2
