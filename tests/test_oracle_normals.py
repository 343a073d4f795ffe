"""Oracle normals: the reference's inline tests (normals.rs:398-624) + numpy cross-checks."""
import numpy as np
import pytest

from fixtures import synth


def test_simple_plane(orc):
    # normals.rs:399-422: XY plane, k=3 -> |n.z| > 0.8
    out = orc.estimate_normals(synth.plane5(), 3)
    assert out.shape == (5, 6)
    assert np.all(np.abs(out[:, 5]) > 0.8)
    assert np.array_equal(out[:, :3], synth.plane5())


def test_empty_cloud(orc):
    # normals.rs:424-429
    assert orc.estimate_normals(np.zeros((0, 3), np.float32), 5).shape == (0, 6)


def test_insufficient_k(orc):
    # normals.rs:431-438
    with pytest.raises(ValueError):
        orc.estimate_normals(synth.plane5(), 2)
    # empty is checked BEFORE k (normals.rs:261-269)
    assert orc.estimate_normals(np.zeros((0, 3), np.float32), 2).shape == (0, 6)


def test_radius_plane(orc):
    # normals.rs:440-480: 20x20 grid, radius 0.2: unit length +-0.1, >80% |n.z|>0.8
    out = orc.estimate_normals(synth.grid_plane(), 10, radius=0.2)
    n = out[:, 3:]
    assert np.all(np.abs(np.linalg.norm(n, axis=1) - 1.0) < 0.1)
    assert (np.abs(n[:, 2]) > 0.8).mean() > 0.8


def test_cylinder_viewpoint(orc):
    # normals.rs:482-548: k=8, viewpoint (0,0,2): >60% of normals perpendicular to z
    out = orc.estimate_normals(synth.cylinder(), 8, viewpoint=[0, 0, 2])
    assert (np.abs(out[:, 5]) < 0.5).mean() > 0.6


def test_orientation_consistency(orc):
    # normals.rs:550-592: all n.z same sign with viewpoint (0,0,1)
    out = orc.estimate_normals(synth.plane4(), 3, viewpoint=[0, 0, 1])
    assert np.all(out[:, 5] > 0)


def test_default_viewpoint_and_unit_length(orc):
    pts = synth.bunny_standin(4000)
    out = orc.estimate_normals(pts, 10)
    n = out[:, 3:].astype(np.float64)
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    mn, mx = pts.min(0), pts.max(0)
    vp = (mn + mx) / 2 + np.array([0, 0, np.linalg.norm(mx - mn)])
    tv = vp - pts
    assert ((n * tv).sum(1) >= -1e-6).all()


def test_f32_solver_vs_f64_jacobi(orc):
    """The f32 nalgebra restatement agrees with an independent f64 Jacobi on well-conditioned
    neighbourhoods (relative eigengap > 1e-2) to far better than the 1e-4 rad parity budget."""
    pts = synth.terrain(20000, 10.0, seed=11, noise=0.002)
    out = orc.estimate_normals(pts, 16, consistent_orientation=False)
    n64, gap = orc.normals_f64(pts, 16)
    n32 = out[:, 3:].astype(np.float64)
    ang = np.arctan2(np.linalg.norm(np.cross(n32, n64), axis=1), np.abs((n32 * n64).sum(1)))
    good = gap > 1e-2
    assert good.mean() > 0.9
    assert ang[good].max() < 2e-5


def test_symmetric_eigen_vs_numpy(orc):
    rng = np.random.default_rng(0)
    for t in range(500):
        a = rng.normal(size=(3, 3)).astype(np.float32)
        m = (a @ a.T).astype(np.float32)
        m = ((m + m.T) / 2).astype(np.float32)
        val, vec = orc.symmetric_eigen3(m)
        M, V = m.astype(np.float64), vec.astype(np.float64)
        assert np.abs(M @ V - V * val[None, :]).max() / np.abs(m).max() < 5e-6
        assert np.abs(V.T @ V - np.eye(3)).max() < 5e-6
        assert np.allclose(np.sort(val), np.linalg.eigvalsh(M), atol=5e-6 * np.abs(m).max())
    val, vec = orc.symmetric_eigen3(np.zeros((3, 3), np.float32))
    assert np.all(val == 0)
    val, vec = orc.symmetric_eigen3(np.diag([3.0, 1.0, 2.0]).astype(np.float32))
    assert np.argmin(val) == 1 and abs(abs(vec[1, 1]) - 1) < 1e-6
