"""The staged-tile kNN / normals kernels (tc_tile.cu: warp-owned boxes staged into shared memory
by 1-D TMA bulk copies, uniform scan) are an A/B variant of the per-lane kernels.  They must give
the same bits: kNN rows equal to the oracle's canonical brute force, normals equal to the per-lane
kernels' (same neighbour order, same f32 sums; the eigen solver is switched separately)."""
import ctypes as C

import numpy as np
import pytest

import threecrate_b200 as tc
from threecrate_b200 import _lib
from fixtures import synth

pytestmark = pytest.mark.gpu

PER_LANE, TILE_LDG, TILE_TMA = 31, 31 | 32, 31 | 32 | 64
NEWTON = 128


@pytest.fixture
def flags():
    lib = _lib.load()
    lib.tc_debug_set_search_flags.argtypes = [C.c_int]

    def set_flags(f):
        lib.tc_debug_set_search_flags(int(f))
    yield set_flags
    lib.tc_debug_set_search_flags(tc.DEFAULT_SEARCH_FLAGS)


def _clouds():
    rng = np.random.default_rng(11)
    yield "terrain", synth.terrain(60_000, 10.0, seed=3, noise=0.002)
    yield "kitti", synth.kitti_frame()[::2]
    yield "blob", rng.normal(size=(20_000, 3)).astype(np.float32)
    g = synth.grid_plane(40, 0.1)                      # lattice: every query is tie-dense
    yield "lattice", g
    yield "dups", np.repeat(rng.uniform(-1, 1, (3000, 3)).astype(np.float32), 3, axis=0)
    yield "tiny", rng.uniform(-1, 1, (9, 3)).astype(np.float32)


@pytest.mark.parametrize("k", [3, 8, 16, 30, 32])
def test_tile_knn_rows_bit_exact(orc, flags, k):
    for name, pts in _clouds():
        want_i, want_d2 = orc.brute_knn(pts, pts, min(k + 1, len(pts)))
        for f in (TILE_LDG, TILE_TMA):
            flags(f)
            idx, dist, cnt = tc.k_nearest_neighbors(pts, k)
            kk = min(k, len(pts) - 1)
            assert np.all(cnt == kk), (name, f)
            # self is dropped by index; duplicates at distance 0 keep the (d2, index) order
            for r in range(0, len(pts), max(1, len(pts) // 4000)):
                row = [int(j) for j in want_i[r] if j != r][:kk]
                assert idx[r, :kk].tolist() == row, (name, f, r)


@pytest.mark.parametrize("k", [10, 16, 30])
def test_tile_normals_equal_the_per_lane_kernels(flags, k):
    for name, pts in _clouds():
        if len(pts) <= k:
            continue
        for solver in (0, NEWTON):
            flags(PER_LANE | solver)
            ref = tc.estimate_normals(pts, k)
            for f in (TILE_LDG, TILE_TMA):
                flags(f | solver)
                got = tc.estimate_normals(pts, k)
                assert np.array_equal(ref, got), (name, k, f, solver)


def test_newton_solver_matches_jacobi_within_parity_bar(orc, flags):
    pts = synth.terrain(80_000, 12.0, seed=9, noise=0.002)
    _, gap = orc.normals_f64(pts, 16)
    flags(PER_LANE)
    a = tc.estimate_normals(pts, 16)[:, 3:].astype(np.float64)
    flags(PER_LANE | NEWTON)
    b = tc.estimate_normals(pts, 16)[:, 3:].astype(np.float64)
    ang = np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), (a * b).sum(1))
    assert ang[gap > 1e-3].max() <= 1e-6   # both are f64 solves of the same f32 covariance
    assert np.all((a * b).sum(1)[gap > 1e-3] > 0)


def test_tile_stats_are_reported(flags):
    pts = synth.terrain(50_000, 9.0, seed=2, noise=0.002)
    ctx = tc.default_context()
    flags(TILE_TMA)
    ctx.enable_stats(True)
    try:
        tc.estimate_normals(pts, 16)
        s = ctx.last_stats()
    finally:
        ctx.enable_stats(False)
    assert s["queries"] == len(pts)
    assert s["rounds"] >= (len(pts) + 31) // 32
    assert s["candidates_staged"] > 16 * s["rounds"]
    assert s["chain_queries"] < 0.05 * len(pts)   # cloud-edge queries need a wider search
