#!/usr/bin/env python
"""Index-build phase timing: python tools/buildtrace.py [n] [k] [world]  (TC_TRACE=1 for host phases)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import threecrate_b200 as tc
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pts = bench.head_cloud(n)
ctx = tc.default_context()
cloud = tc.DeviceCloud(pts, ctx)
for rep in range(3):
    for r in ([None] if world == 1 else [0, world // 2]):
        ctx.synchronize(); t0 = time.perf_counter()
        ctx.timer_start()
        ix = tc.GridIndex(cloud, k_hint=k, shard=None if r is None else (r, world))
        ms = ctx.timer_stop()
        print(f"rep {rep} rank {r}: build {ms:.3f} ms (events) {1e3*(time.perf_counter()-t0):.3f} ms (host)  {ix.info()['dims']}", flush=True)
        ix.free()
os._exit(0)
