"""tc_estimate_normals_distributed on ONE GPU (a 1-rank communicator): the whole protocol runs -
window, upload of the chunk, peer barriers, slab/complete index build, rows routed through the
window, download - and must reproduce estimate_normals bit for bit.  The N-rank run is
tests/multi_worker.py (gpurun --gpus 2) and the parity block of bench.py --gpus N."""
import numpy as np
import pytest

import threecrate_b200 as tc
from fixtures import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    ctx = tc.default_context()
    c = tc.Comm(ctx, tc.Comm.unique_id(ctx), 1, 0)
    yield c
    c.destroy()


@pytest.mark.parametrize("name,k", [("terrain", 16), ("kitti", 10), ("tiny", 5), ("big_k", 70)])
def test_one_rank_distributed_equals_estimate_normals(comm, name, k):
    rng = np.random.default_rng(1)
    pts = {"terrain": synth.terrain(150_001, 12.0, seed=2, noise=0.002),
           "kitti": synth.kitti_frame(seed=3)[:50_000],
           "tiny": rng.uniform(-1, 1, (7, 3)).astype(np.float32),
           "big_k": synth.terrain(20_000, 4.0, seed=5, noise=0.002)}[name]
    n = len(pts)
    assert comm.chunk(n) == (0, n)
    comm.open_window([comm.window_handle(n)])
    for _ in range(2):  # back-to-back calls reuse the window (barrier epochs keep advancing)
        got = comm.estimate_normals(pts, n, k)
        ref = tc.estimate_normals(pts, k)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    vp = np.array([0.0, 0.0, 50.0], np.float32)
    got = comm.estimate_normals(pts, n, k, viewpoint=vp)
    ref = tc.estimate_normals_with_config(
        pts, tc.NormalEstimationConfig(k_neighbors=k, viewpoint=vp))
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_validation(comm):
    pts = synth.terrain(1000, 2.0, seed=1)
    comm.open_window([comm.window_handle(1000)])
    with pytest.raises(tc.InvalidData):
        comm.estimate_normals(pts, 1000, 2)  # k < 3 (normals.rs:264-266)
    with pytest.raises(tc.InvalidData):
        comm.estimate_normals(pts[:10], 1000, 8)  # not this rank's chunk
    with pytest.raises(tc.InvalidData):
        comm.estimate_normals(pts[:500], 500, 8)  # no window of that size
