#!/usr/bin/env python
"""Kernel A/B bench: times the search kernels for each variant-flag setting and checks the outputs
are bit-identical across variants.  python tools/kbench.py [--flags 0,1,2,3] [--cells a,b,c]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import threecrate_b200 as tc  # noqa: E402
from threecrate_b200 import _lib  # noqa: E402
from fixtures import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--flags", default="0,1,2,3")
ap.add_argument("--what", default="c2,c4,c3")
ap.add_argument("--c4n", type=int, default=2_000_000)
ap.add_argument("--c4k", default="30")
ap.add_argument("--cellscale", default="1.0")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--finecap", type=int, default=0)
ap.add_argument("--iters", default="30", help="c3: comma list of iteration counts to time")
ap.add_argument("--maxlevels", type=int, default=0, help="cap the index resolutions (1: no coarse/fine level)")
a = ap.parse_args()
ctx = tc.default_context()
lib = _lib.load()
setflags = lib.tc_debug_set_search_flags
setflags.argtypes = [C.c_int]
if a.finecap:
    lib.tc_debug_set_fine_cap.argtypes = [C.c_int]
    lib.tc_debug_set_fine_cap(a.finecap)
flags = [int(f) for f in a.flags.split(",")]
scales = [float(s) for s in a.cellscale.split(",")]


def timed(fn, reps):
    fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop())
    return float(np.median(ts))


def normals_case(name, pts, k):
    cloud = tc.DeviceCloud(pts, ctx)
    n = len(pts)
    d_out = ctx.alloc(n * 24)
    base_index = tc.GridIndex(cloud, k_hint=k)
    auto = base_index.info()["cell_size"]
    import time
    bt = []
    for _ in range(a.reps):
        ctx.synchronize(); t0 = time.perf_counter()
        tmp = tc.GridIndex(cloud, k_hint=k); ctx.synchronize()
        bt.append((time.perf_counter() - t0) * 1e3); del tmp
    print(f"{name:4s} index build {np.median(bt):.3f} ms (host clock), levels={base_index.info()['n_levels']}")
    ref = None
    for sc in scales:
        index = base_index if sc == 1.0 else tc.GridIndex(cloud, k_hint=k, cell_size=auto * sc)
        info = index.info()
        for f in flags:
            setflags(f)
            ms = timed(lambda: index.estimate_normals_device(d_out, k), a.reps)
            out = np.empty((n, 6), np.float32)
            ctx.to_host(out, d_out)
            if ref is None:
                ref = out
            same = np.array_equal(ref, out)
            a_, b_ = ref[:, 3:].astype(np.float64), out[:, 3:].astype(np.float64)
            ang = np.arctan2(np.linalg.norm(np.cross(a_, b_), axis=1), (a_ * b_).sum(1))
            ctx.enable_stats(True)
            index.estimate_normals_device(d_out, k)
            st = ctx.last_stats()
            ctx.enable_stats(False)
            print(f"{name:4s} k={k:2d} cell={info['cell_size']:.4f} (x{sc}) occ={info['occupied_cells']} "
                  f"maxpop={info['max_cell_population']} flags={f} kernel={ms:8.3f} ms "
                  f"{n / ms / 1e3:8.1f} Mpts/s same={same} pos_same={np.array_equal(ref[:, :3], out[:, :3])} "
                  f"ang max={ang.max():.2e} n>1e-5={(ang > 1e-5).sum()} "
                  f"chain={st['chain_queries']} rounds={st['rounds']} splits={st['box_splits']} "
                  f"retries={st['retries']} cand/round={st['candidates_staged'] / max(st['rounds'], 1):.0f} "
                  f"merges/round={st['merges'] / max(st['rounds'], 1):.1f}", flush=True)
        if index is not base_index:
            index.free()
    ctx.free(d_out)


if a.maxlevels:
    lib.tc_debug_set_max_levels.argtypes = [C.c_int]
    lib.tc_debug_set_max_levels(a.maxlevels)
what = a.what.split(",")
if "c2" in what:
    normals_case("c2", synth.kitti_frame(), 16)
if "c4" in what:
    n = a.c4n
    c4pts = synth.terrain(n, 100.0 * (n / 1e7) ** 0.5, seed=4, noise=0.002)
    for kk in a.c4k.split(","):
        normals_case("c4", c4pts, int(kk))
if "c3" in what:
    n = 1_000_000
    src, tgt, nrm, T = synth.scan_pair(n, half_extent=50.0)
    tcloud, scloud = tc.DeviceCloud(tgt, ctx), tc.DeviceCloud(src, ctx)
    d_nrm = ctx.alloc(n * 12)
    ctx.to_device(d_nrm, nrm)
    base_index = tc.GridIndex(tcloud, k_hint=1)
    auto = base_index.info()["cell_size"]
    ref = None
    lib.tc_debug_set_icp_keep.argtypes = [C.c_int]
    for sc in scales:
        index = base_index if sc == 1.0 else tc.GridIndex(tcloud, k_hint=1, cell_size=auto * sc)
        info = index.info()
        for f in flags:
            setflags(f & 0xFFFF)
            lib.tc_debug_set_icp_keep(0 if (f >> 16) & 1 else 1)   # flag bit 16: search every time
            res = []
            for it in [int(x) for x in a.iters.split(",")][:-1]:
                ms = timed(lambda: res.append(tc.icp_point_to_plane_device(
                    scloud, index, d_nrm, tc.IDENTITY, it, None, -1.0)), 3)
                print(f"   iters={it}: {ms:8.3f} ms", flush=True)
            res = []
            ms = timed(lambda: res.append(tc.icp_point_to_plane_device(
                scloud, index, d_nrm, tc.IDENTITY, int(a.iters.split(",")[-1]), None, -1.0)), 3)
            r = res[-1]
            if ref is None:
                ref = r.transformation
            print(f"c3 icp cell={info['cell_size']:.4f} (x{sc}) occ={info['occupied_cells']} "
                  f"maxpop={info['max_cell_population']} flags={f} total={ms:8.3f} ms "
                  f"per_iter={ms / 30 * 1e3:7.1f} us same={np.array_equal(ref, r.transformation)}",
                  flush=True)
