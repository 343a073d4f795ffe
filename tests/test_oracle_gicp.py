"""Oracle GICP vs the reference's own inline tests (threecrate-algorithms/src/gicp.rs:314-583)."""
import numpy as np
import pytest

import oracle
from fixtures import synth

F = np.float32


def make_sphere(n, radius):  # gicp.rs:318-333, f32 arithmetic
    i = np.arange(n, dtype=F)
    golden = F(np.pi) * (F(3.0) - np.sqrt(F(5.0)))
    y = F(1.0) - (i / max(F(n) - F(1.0), F(1.0))) * F(2.0)
    r = np.sqrt(np.maximum(F(1.0) - y * y, F(0.0)))
    th = golden * i
    return np.stack([np.cos(th) * r * F(radius), y * F(radius), np.sin(th) * r * F(radius)],
                    1).astype(F)


def quat_axis(axis, angle):
    q = np.zeros(4, F)
    q[axis] = np.sin(F(angle) / 2)
    q[3] = np.cos(F(angle) / 2)
    return q


def rotate(q, p):
    return synth.apply_iso(np.concatenate([[0, 0, 0], q]).astype(F), p)


def angle_to(qa, qb):
    d = abs(float(np.dot(qa.astype(np.float64), qb.astype(np.float64))))
    return 2.0 * np.arccos(min(1.0, d / (np.linalg.norm(qa) * np.linalg.norm(qb))))


def test_identity_converges():  # :335-345
    c = make_sphere(100, 3.0)
    r = oracle.gicp(c, c, max_iterations=30)
    assert r.converged and r.mse < 1e-4


def test_recovers_small_translation():  # :347-364
    s = make_sphere(150, 3.0)
    t = s + F([0.1, 0, 0])
    r = oracle.gicp(s, t, max_iterations=60, max_correspondence_distance=2.0)
    assert np.linalg.norm(r.translation - [0.1, 0, 0]) < 0.05 and r.mse < 0.1


def test_errors():  # :366-380, 556-582
    c = make_sphere(30, 1.0)
    with pytest.raises(oracle.InvalidData, match="empty"):
        oracle.gicp(np.empty((0, 3), F), c)
    with pytest.raises(oracle.InvalidData, match="max_iterations"):
        oracle.gicp(c, c, max_iterations=0)
    few = make_sphere(10, 1.0)
    with pytest.raises(oracle.InvalidData, match="at least"):
        oracle.gicp(few, few)
    g = np.arange(50, dtype=F) * F(0.1)
    plane = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    plane = np.c_[plane, np.zeros(len(plane))].astype(F)
    with pytest.raises(oracle.InvalidData, match="coplanar"):
        oracle.gicp(plane, plane)


def test_result_fields_populated():  # :382-392
    c = make_sphere(60, 2.0)
    r = oracle.gicp(c, c, max_iterations=10)
    assert r.iterations > 0 and len(r.correspondences) > 0


def test_recovers_tiny_rotation_from_identity():  # :404-428
    s = make_sphere(300, 3.0)
    q = quat_axis(2, np.deg2rad(2.0))
    r = oracle.gicp(s, rotate(q, s), max_iterations=60, max_correspondence_distance=0.8)
    assert angle_to(r.rotation, q) < np.deg2rad(0.5) and r.mse < 0.01


def test_refines_rotation_from_near_correct_init():  # :430-455
    s = make_sphere(200, 3.0)
    q = quat_axis(2, np.deg2rad(8.0))
    init = np.concatenate([[0, 0, 0], quat_axis(2, np.deg2rad(6.0))]).astype(F)
    r = oracle.gicp(s, rotate(q, s), init, max_iterations=60, max_correspondence_distance=0.8)
    assert angle_to(r.rotation, q) < np.deg2rad(0.5)


def test_refines_combined_rotation_and_translation():  # :457-497
    s = make_sphere(200, 3.0)
    q = quat_axis(1, np.deg2rad(6.0))
    iso = np.concatenate([[0.3, 0, 0], q]).astype(F)
    t = synth.apply_iso(iso, s)
    init = np.concatenate([[0.24, 0, 0], quat_axis(1, np.deg2rad(4.8))]).astype(F)
    r = oracle.gicp(s, t, init, max_iterations=80, max_correspondence_distance=0.8)
    assert np.linalg.norm(r.translation - [0.3, 0, 0]) < 0.05
    assert angle_to(r.rotation, q) < np.deg2rad(0.5)


def test_robust_to_noise():  # :503-536
    s = make_sphere(200, 3.0)
    i = np.arange(200, dtype=F)
    noise = np.stack([np.sin(i * F(1.6180339887)), np.cos(i * F(2.7182818284)),
                      np.sin(i * F(3.1415926535))], 1).astype(F) * F(0.05)
    r = oracle.gicp(s, s + noise, max_iterations=50, max_correspondence_distance=1.0)
    assert r.mse < 0.05 and np.linalg.norm(r.translation) < 0.1


def test_robust_to_outliers():  # :538-560
    s = make_sphere(200, 3.0)
    t = np.arange(20, dtype=F)
    out = np.stack([t * F(7.3) - 50, t * F(3.1) - 30, t * F(5.7) - 40], 1).astype(F)
    r = oracle.gicp(s, np.vstack([s, out]), max_iterations=40, max_correspondence_distance=0.5)
    assert np.linalg.norm(r.translation) < 0.05 and r.mse < 0.01


def test_covariances_against_numpy():
    pts = synth.terrain(800, 3.0, seed=2, noise=0.02)
    cov = oracle.gicp_covariances(pts, 20)
    idx, _ = oracle.brute_knn(pts, pts, 20)
    for i in range(0, 800, 53):
        nb = pts[idx[i].astype(np.int64)].astype(np.float64)
        ref = np.cov(nb.T, ddof=1) + 1e-4 * np.eye(3)
        np.testing.assert_allclose(cov[i], ref, rtol=2e-3, atol=1e-6)


def test_one_gauss_newton_step_against_independent_numpy():
    """One GICP iteration re-derived in f64 numpy straight from the formulas (M = C_t + R C_s R^T,
    J = [-skew(Ts) | I], H d = g, delta = Rz Ry Rx) must reproduce the oracle's first step."""
    src, tgt, _, _ = synth.scan_pair(1500, half_extent=2.5, noise=0.004)
    r = oracle.gicp(src, tgt, None, 1, 0.6, 1e-6, 12)
    cs = oracle.gicp_covariances(src, 12).astype(np.float64)
    ct = oracle.gicp_covariances(tgt, 12).astype(np.float64)
    nn_idx, nn_d2 = oracle.brute_knn(tgt, src, 1)
    H = np.zeros((6, 6))
    g = np.zeros(6)
    n_corr, mse = 0, 0.0
    for i in range(len(src)):
        d = float(np.sqrt(nn_d2[i, 0]))
        if d > 0.6:
            continue
        j = int(nn_idx[i, 0])
        M = ct[j] + cs[i]  # R = identity on the first iteration
        Mi = np.linalg.inv(M)
        s = src[i].astype(np.float64)
        A = -np.array([[0, -s[2], s[1]], [s[2], 0, -s[0]], [-s[1], s[0], 0]])
        J = np.hstack([A, np.eye(3)])
        res = tgt[j].astype(np.float64) - s
        H += J.T @ Mi @ J
        g += J.T @ Mi @ res
        n_corr += 1
        mse += d * d
    x = np.linalg.solve(H, g)

    def q_axis(a, ang):
        q = np.zeros(4)
        q[a] = np.sin(ang / 2)
        q[3] = np.cos(ang / 2)
        return q

    def q_mul(a, b):  # [i, j, k, w]
        ai, aj, ak, aw = a
        bi, bj, bk, bw = b
        return np.array([aw * bi + ai * bw + aj * bk - ak * bj, aw * bj - ai * bk + aj * bw + ak * bi,
                         aw * bk + ai * bj - aj * bi + ak * bw, aw * bw - ai * bi - aj * bj - ak * bk])

    q = q_mul(q_mul(q_axis(2, x[2]), q_axis(1, x[1])), q_axis(0, x[0]))
    assert len(r.correspondences) == n_corr
    assert abs(r.mse - mse / n_corr) <= 1e-5 * (mse / n_corr)
    np.testing.assert_allclose(r.translation, x[3:], rtol=0, atol=2e-5)
    assert angle_to(r.rotation, q.astype(F)) <= 2e-5
