"""N > 1 host logic on CPU: world_size-2 gloo processes shard the work, all-reduce the 29 ICP
scalars and must reproduce the unsharded solve; shard ranges must tile."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

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
