import sys, numpy as np
sys.path.insert(0, '/root/repo')
import threecrate_b200 as tc
from fixtures import synth
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
pts = bench.head_cloud(n)
ctx = tc.default_context()
cloud = tc.DeviceCloud(pts, ctx)
d_out = ctx.alloc(n * 24)
def full(k):
    out = np.zeros((n, 6), np.float32)
    ix = tc.GridIndex(cloud, k_hint=k); ix.estimate_normals_device(d_out, k); ctx.to_host(out, d_out); ix.free()
    return out
for k in (16, 30):
    a = full(k); b = full(k)
    d = np.any(a != b, axis=1)
    print(f"k={k}: complete vs complete: {d.sum()} rows differ")
    for world in (2,):
        got = np.zeros((n, 6), np.float32); written = np.zeros(n, np.int32)
        for r in range(world):
            ctx.to_device(d_out, np.zeros((n, 6), np.float32))
            ix = tc.GridIndex(cloud, k_hint=k, shard=(r, world)); ix.estimate_normals_device(d_out, k)
            part = np.zeros((n, 6), np.float32); ctx.to_host(part, d_out); ix.free()
            rows = np.abs(part[:, 3:]).sum(1) > 0
            written += rows; got[rows] = part[rows]
        d = np.any(a != got, axis=1)
        print(f"k={k} world={world}: written once {np.all(written==1)} missing {(written==0).sum()} dup {(written>1).sum()}; rows differing from complete: {d.sum()}")
        if d.sum():
            idx = np.nonzero(d)[0][:10]
            for i in idx: print("  row", i, "pos", a[i,:3], "full", a[i,3:], "shard", got[i,3:])
            print("  y range of differing rows", a[d,1].min(), a[d,1].max())
