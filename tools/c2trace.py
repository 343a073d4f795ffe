import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np
import threecrate_b200 as tc
from fixtures import synth
pts = synth.kitti_frame()
ctx = tc.default_context()
cloud = tc.DeviceCloud(pts, ctx)
for rep in range(4):
    ctx.synchronize(); t0 = time.perf_counter()
    ctx.timer_start()
    ix = tc.GridIndex(cloud, k_hint=16)
    ms = ctx.timer_stop()
    print(f"rep {rep}: build {ms:.3f} ms (events) {1e3*(time.perf_counter()-t0):.3f} ms (host) {ix.info()['n_levels']} levels", flush=True)
    ix.free()
os._exit(0)
