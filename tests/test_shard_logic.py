"""N > 1 host logic on CPU: world_size-2 gloo processes shard the work, all-reduce the 29 ICP
scalars and must reproduce the unsharded solve; shard ranges must tile."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import threecrate_b200 as tc
from threecrate_b200 import sharding
from fixtures import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile():
    for n in (0, 1, 7, 120000, 10_000_001):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(i, w, n) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(2, 2, 10)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # terrain + walls: all six degrees of freedom are observable (a sphere is not)
        p, nrm = synth.terrain(4000, 10.0, seed=3, return_normals=True, wall_fraction=0.2)
        tgt = p + np.array([0.02, -0.01, 0.015], np.float32)
        lo, hi = sharding.shard_range(rank, world, len(p))
        part = sharding.normal_equation_sums(p[lo:hi], tgt[lo:hi], nrm[lo:hi])
        t = torch.from_numpy(part.copy())
        dist.all_reduce(t)  # the per-iteration collective: 29 f64, sum
        x = sharding.solve_normal_equations(t.numpy())
        # every rank holds identical sums, hence identical solutions
        xs = [torch.zeros(6, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(xs, torch.from_numpy(x.copy()))
        same = all(torch.equal(xs[0], v) for v in xs)
        full = sharding.solve_normal_equations(sharding.normal_equation_sums(p, tgt, nrm))
        q.put((rank, same, float(np.abs(x - full).max()), float(t[28])))
    finally:
        dist.destroy_process_group()


def test_gloo_allreduce_of_normal_equations_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, err, nvalid in res:
        assert same, "ranks disagree"
        assert err < 1e-9
        assert nvalid == 4000


def test_distributed_normals_chunks_tile_and_route():
    """tc_dist_chunk (the row split of tc_estimate_normals_distributed): contiguous, equal length,
    a multiple of 4 rows (16-byte aligned chunk boundaries in the 12-byte point array), and the
    kernel's routing rule `owner = i // chunk` sends every row to the rank whose range holds it."""
    for n in (0, 1, 5, 1023, 120000, 10_000_001):
        for w in (1, 2, 3, 8):
            r = [tc.dist_chunk(n, w, i) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            chunk = max(4, -(-n // w) + 3 & ~3)
            assert chunk % 4 == 0 and all(b - a in (chunk, max(0, n - i * chunk), 0)
                                          for i, (a, b) in enumerate(r))
            for i in (0, n // 3, n - 1):
                if 0 <= i < n:
                    a, b = r[i // chunk]
                    assert a <= i < b
                    # the device form: __umulhi(i, 2^32 // chunk), corrected upwards once
                    q = (i * ((1 << 32) // chunk)) >> 32
                    q += (q + 1) * chunk <= i
                    assert q == i // chunk


def _chunk_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts = synth.terrain(1001, 5.0, seed=9)
        n = len(pts)
        lo, hi = tc.dist_chunk(n, world, rank)
        chunk = tc.dist_chunk(n, world, 0)[1]
        # all-gather of the chunks rebuilds the cloud (what k_gather_chunks does over NVLink)
        mine = np.zeros((chunk, 3), np.float32)
        mine[: hi - lo] = pts[lo:hi]
        parts = [torch.zeros(chunk, 3) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(mine))
        full = torch.cat(parts)[:n].numpy()
        # every rank computes the rows of "its" points (here: a split by y) and routes row i to
        # rank i // chunk: the owners end up with exactly their range, each row once
        order = np.argsort(full[:, 1], kind="stable")
        own = order[rank * n // world:(rank + 1) * n // world]
        sent = torch.zeros(world * chunk, dtype=torch.int32)
        sent[torch.from_numpy(own)] = 1
        dist.all_reduce(sent)
        q.put((rank, bool(np.array_equal(full, pts)), int(sent[:n].min()), int(sent[:n].max()),
               int(sent[lo:hi].sum()) == hi - lo))
    finally:
        dist.destroy_process_group()


def test_gloo_chunk_gather_and_row_routing_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30000 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_chunk_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rebuilt, lo_cnt, hi_cnt, mine_complete in res:
        assert rebuilt and lo_cnt == 1 and hi_cnt == 1 and mine_complete


def test_point_balanced_slab_boundaries():
    """Host logic of tc_index_build_sharded: from the per-plane point counts (identical on every
    rank) every rank derives the same slab boundaries - monotone, covering all planes, each within
    1/8 slab of its equal-planes position, and closer to equal point counts than equal planes."""
    import ctypes as C

    from threecrate_b200 import _lib
    lib = _lib.load()
    fn = lib.tc_debug_balanced_boundaries
    fn.argtypes = [C.POINTER(C.c_uint32), C.c_int, C.c_int, C.POINTER(C.c_int)]
    fn.restype = None
    rng = np.random.default_rng(4)

    def boundaries(counts, world):
        c = np.ascontiguousarray(counts, np.uint32)
        out = np.zeros(world + 1, np.int32)
        fn(c.ctypes.data_as(C.POINTER(C.c_uint32)), len(c), world, out.ctypes.data_as(C.POINTER(C.c_int)))
        return out

    for n_planes, world in ((1017, 8), (1017, 2), (64, 8), (33, 8), (4000, 3)):
        profiles = {
            "uniform": np.full(n_planes, 9800),
            "walls": np.r_[150_000, np.full(n_planes - 2, 9600), 150_000],   # the bench cloud's shape
            "ramp": np.linspace(100, 20_000, n_planes).astype(np.int64),
            "random": rng.integers(0, 30_000, n_planes),
            "empty_half": np.r_[np.zeros(n_planes // 2, np.int64), np.full(n_planes - n_planes // 2, 500)],
        }
        for name, counts in profiles.items():
            b = boundaries(counts, world)
            assert b[0] == 0 and b[-1] == n_planes
            assert np.all(np.diff(b) > 0), (name, n_planes, world, b)
            margin = n_planes // world // 8
            eq = np.array([r * n_planes // world for r in range(world + 1)])
            assert np.all(np.abs(b - eq) <= margin)
            if margin == 0:
                assert np.array_equal(b, eq)
                continue
            per = np.add.reduceat(counts, b[:-1])
            per_eq = np.add.reduceat(counts, eq[:-1])
            assert per.max() <= per_eq.max() + counts.max(), (name, per, per_eq)
    # the bench cloud's profile: equal planes leave the edge slabs 10 % heavier; balanced: < 1.5 %
    counts = np.r_[135_000, np.full(1015, 9600), 135_000]
    b = boundaries(counts, 8)
    per = np.add.reduceat(counts, b[:-1])
    assert per.max() / per.mean() < 1.015
