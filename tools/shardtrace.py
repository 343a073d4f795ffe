#!/usr/bin/env python
"""Per-rank cost of the slab-sharded headline step, emulated on ONE GPU:
python tools/shardtrace.py [n] [k] [world]   (TC_TRACE=1 adds the build's phase timings)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import threecrate_b200 as tc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
world = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reps = int(os.environ.get("REPS", "8"))
pts = bench.head_cloud(n)
ctx = tc.default_context()
cloud = tc.DeviceCloud(pts, ctx)
d_out = ctx.alloc(n * 24)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step(shard):
    tb = tk = 0.0
    info = None
    for rep in range(reps + 2):
        flush.fill_(rep & 255)
        torch.cuda.synchronize()
        ctx.timer_start()
        ix = tc.GridIndex(cloud, k_hint=k, shard=shard)
        b = ctx.timer_stop()
        ctx.timer_start()
        ix.estimate_normals_device(d_out, k)
        kk = ctx.timer_stop()
        info = ix.info()
        ix.free()
        if rep >= 2:
            tb += b
            tk += kk
    return tb / reps, tk / reps, info


b, kk, info = step(None)
print(f"complete: build {b:.3f} kernel {kk:.3f} ms  dims {info['dims']}", flush=True)
tot = []
for r in range(world):
    b, kk, info = step((r, world))
    tot.append(b + kk)
    print(f"rank {r}/{world}: build {b:.3f} kernel {kk:.3f} step {b + kk:.3f} ms", flush=True)
print(f"max step {max(tot):.3f}  mean {np.mean(tot):.3f}")
os._exit(0)
