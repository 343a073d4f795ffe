"""Slab-sharded index build (tc_index_build_sharded): every rank sorts only its slab of cell planes
plus a halo, and tc_estimate_normals_device writes exactly the rank's rows.  The ranks are played
one after the other on ONE GPU here (the driver's test box has one), which checks the same
invariants the multi-GPU bench asserts: the rows tile the cloud exactly once and are bit-identical
to the complete-index result, including when a search has to look beyond the halo."""
import ctypes as C

import numpy as np
import pytest

import threecrate_b200 as tc
from threecrate_b200 import _lib
from fixtures import synth

pytestmark = pytest.mark.gpu


def _sharded_rows(pts, k, world):
    ctx = tc.default_context()
    n = len(pts)
    cloud = tc.DeviceCloud(pts, ctx)
    full = np.zeros((n, 6), np.float32)
    d_out = ctx.alloc(n * 24)
    idx = tc.GridIndex(cloud, k_hint=k)
    idx.estimate_normals_device(d_out, k)
    ctx.to_host(full, d_out)
    idx.free()
    written = np.zeros(n, np.int32)
    got = np.zeros((n, 6), np.float32)
    for r in range(world):
        ctx.to_device(d_out, np.zeros((n, 6), np.float32))
        ix = tc.GridIndex(cloud, k_hint=k, shard=(r, world))
        ix.estimate_normals_device(d_out, k)
        part = np.zeros((n, 6), np.float32)
        ctx.to_host(part, d_out)
        ix.free()
        rows = np.abs(part[:, 3:]).sum(1) > 0
        written += rows
        got[rows] = part[rows]
    ctx.free(d_out)
    cloud.free()
    return full, got, written


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("k", [16, 30])
def test_slabs_tile_the_cloud_and_match_the_complete_index(world, k):
    pts = synth.terrain(400_000, 20.0, seed=21, noise=0.002)
    full, got, written = _sharded_rows(pts, k, world)
    assert np.all(written == 1)
    assert np.array_equal(full, got)


@pytest.mark.parametrize("name", ["kitti_multilevel", "few_planes", "tiny"])
def test_clouds_that_fall_back_to_a_complete_index(name):
    rng = np.random.default_rng(3)
    pts = {"kitti_multilevel": synth.kitti_frame(),
           "few_planes": synth.terrain(3000, 1.0, seed=5, noise=0.002),
           "tiny": rng.uniform(-1, 1, (40, 3)).astype(np.float32)}[name]
    full, got, written = _sharded_rows(pts, 10, 4)
    assert np.all(written == 1)
    assert np.array_equal(full, got)


def test_search_beyond_the_halo_is_redone_on_a_complete_index():
    lib = _lib.load()
    lib.tc_debug_set_shard_halo.argtypes = [C.c_int]
    # uneven density along the slab axis: sparse stripes need rings of several cells
    rng = np.random.default_rng(8)
    dense = synth.terrain(200_000, 15.0, seed=4, noise=0.002)
    sparse = rng.uniform([-15, -15, 0.8], [15, 15, 2.5], (1500, 3)).astype(np.float32)
    pts = np.concatenate([dense, sparse]).astype(np.float32)
    lib.tc_debug_set_shard_halo(1)
    try:
        full, got, written = _sharded_rows(pts, 16, 4)
    finally:
        lib.tc_debug_set_shard_halo(4)
    assert np.all(written == 1)
    assert np.array_equal(full, got)
