#!/usr/bin/env python
"""Per-query cycle histogram of the normals kernel (debug): which queries form the tail?"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import threecrate_b200 as tc
from threecrate_b200 import _lib, synth
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ctx = tc.default_context(); lib = _lib.load()
lib.tc_debug_set_search_flags.argtypes = [C.c_int]
lib.tc_debug_set_query_clock_buffer.argtypes = [C.c_void_p]
pts = synth.kitti_frame(); n = len(pts)
cloud = tc.DeviceCloud(pts, ctx)
index = tc.GridIndex(cloud, k_hint=16)
if scale != 1.0:
    index = tc.GridIndex(cloud, k_hint=16, cell_size=index.info()['cell_size'] * scale)
d_out = ctx.alloc(n * 24); d_dbg = ctx.alloc(n * 8)
lib.tc_debug_set_search_flags(flags)
index.estimate_normals_device(d_out, 16); ctx.synchronize()
lib.tc_debug_set_query_clock_buffer(C.c_void_p(d_dbg))
ctx.timer_start(); index.estimate_normals_device(d_out, 16); ms = ctx.timer_stop()
dbg = np.zeros((n, 2), np.uint32); ctx.to_host(dbg, d_dbg)
cyc, aux, lvl = dbg[:, 0].astype(np.float64), dbg[:, 1] & 0xFFFF, dbg[:, 1] >> 16
print(f"flags={flags} cell={index.info()['cell_size']:.3f} kernel {ms:.3f} ms; per-query cycles: mean {cyc.mean():.0f} "
      f"p50 {np.percentile(cyc,50):.0f} p90 {np.percentile(cyc,90):.0f} p99 {np.percentile(cyc,99):.0f} "
      f"p99.9 {np.percentile(cyc,99.9):.0f} max {cyc.max():.0f}")
top = np.argsort(-cyc)[:8]
for i in top:
    print(f"  q{i}: cycles {cyc[i]:.0f} aux(R or n)={aux[i]} pos={pts[i]} range={np.linalg.norm(pts[i,:2]):.1f}")
print("info:", index.info()); print("level hist:", np.bincount(lvl), "aux hist:", np.bincount(aux)[:20])
for l in range(int(lvl.max())+1):
    m = lvl == l
    print(f"  level {l}: n={m.sum()} mean cycles {cyc[m].mean():.0f} max {cyc[m].max():.0f}")
for r in range(1, 17):
    m = aux == r
    if m.any(): print(f"  aux={r}: n={m.sum()} mean cycles {cyc[m].mean():.0f} max {cyc[m].max():.0f}")
