#!/usr/bin/env python
"""Per-query cycle breakdown of the two-pass normals kernel (debug): which queries / warps form
the tail?  python tools/qclock.py [flags] [cell scale] [k]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import threecrate_b200 as tc
from threecrate_b200 import _lib
from fixtures import synth
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 15
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
k = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ctx = tc.default_context(); lib = _lib.load()
lib.tc_debug_set_search_flags.argtypes = [C.c_int]
lib.tc_debug_set_query_clock_buffer.argtypes = [C.c_void_p]
pts = synth.kitti_frame(); n = len(pts)
cloud = tc.DeviceCloud(pts, ctx)
index = tc.GridIndex(cloud, k_hint=k)
if scale != 1.0:
    index = tc.GridIndex(cloud, k_hint=k, cell_size=index.info()['cell_size'] * scale)
d_out = ctx.alloc(n * 24); d_dbg = ctx.alloc(n * 32)
lib.tc_debug_set_search_flags(flags)
index.estimate_normals_device(d_out, k); ctx.synchronize()
lib.tc_debug_set_query_clock_buffer(C.c_void_p(d_dbg))
ctx.timer_start(); index.estimate_normals_device(d_out, k); ms = ctx.timer_stop()
dbg = np.zeros((n, 8), np.uint32); ctx.to_host(dbg, d_dbg)
cyc = dbg[:, 0].astype(np.float64)
nmem, lvl = dbg[:, 1] & 0xFFFF, dbg[:, 1] >> 16
qi, t_ns, R = dbg[:, 2], dbg[:, 3], dbg[:, 6]
p1 = (dbg[:, 4] - dbg[:, 7]).astype(np.float64)
p2 = (dbg[:, 5] - dbg[:, 4]).astype(np.float64)
emit = cyc - p1 - p2
print(f"flags={flags} k={k} kernel {ms:.3f} ms (with clocks); info {index.info()}")
pc = lambda a: " ".join(f"{np.percentile(a, p):.0f}" for p in (50, 90, 99, 99.9, 100))
print(f"per-query cycles mean {cyc.mean():.0f}  p50/90/99/99.9/max {pc(cyc)}")
print(f"  pass1 mean {p1.mean():.0f} [{pc(p1)}]  pass2+rank mean {p2.mean():.0f} [{pc(p2)}]  emit mean {emit.mean():.0f} [{pc(emit)}]")
# warps: 32 consecutive sorted positions
order = np.argsort(qi)
w_cyc = cyc[order][: n // 32 * 32].reshape(-1, 32)
w_max, w_mean = w_cyc.max(1), w_cyc.mean(1)
print(f"per-warp max: mean {w_max.mean():.0f} [{pc(w_max)}]; lane mean/max ratio {w_mean.sum() / w_max.sum():.2f}")
t0 = t_ns.min()
start = (t_ns[order][: n // 32 * 32].reshape(-1, 32)[:, 0] - t0).astype(np.float64) / 1e3
print(f"warp start times us: [{pc(start)}]; kernel span ~{(start + w_max / 1.965e3).max():.0f} us")
worst = np.argsort(-w_max)[:6]
for w in worst:
    ids = order[w * 32:(w + 1) * 32]
    print(f"  warp {w}: start {start[w]:.0f} us, max {w_max[w]:.0f} mean {w_mean[w]:.0f}; levels {np.bincount(lvl[ids], minlength=3)} "
          f"R {np.bincount(R[ids])[:8]} range {np.linalg.norm(pts[ids, :2], axis=1).mean():.1f} m "
          f"p1 {p1[ids].max():.0f} p2 {p2[ids].max():.0f} emit {emit[ids].max():.0f}")
for l in range(int(lvl.max()) + 1):
    m = lvl == l
    print(f"  level {l}: n={m.sum()} mean cycles {cyc[m].mean():.0f} max {cyc[m].max():.0f}; R hist {np.bincount(R[m])[:10]}")
