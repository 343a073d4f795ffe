import sys, os, numpy as np, torch, time
sys.path.insert(0, '/root/repo')
import threecrate_b200 as tc
import bench
n = 10_000_000
pts = bench.head_cloud(n)
ctx = tc.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
cloud = tc.DeviceCloud(pts, ctx)
out_t = torch.zeros((n, 6), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
def ev(): return torch.cuda.Event(enable_timing=True)
for mode in ("created_in_loop",):
    shard = (0, 2)
    pool = [(ev(), ev(), ev()) for _ in range(8)]
    for a, b, c in pool:  # torch creates the CUDA event lazily at first record
        a.record(ext); b.record(ext); c.record(ext)
    torch.cuda.synchronize()
    evs = []; host = []
    for it in range(8):
        t0 = time.perf_counter()
        e0, e1, e2 = pool[it] if mode.startswith("precreated") else (ev(), ev(), ev())
        e0.record(ext)
        ix = tc.GridIndex(cloud, k_hint=16, shard=shard)
        t1 = time.perf_counter()
        if mode != "precreated_e0_only":
            e1.record(ext)
        ix.estimate_normals_device(out_t.data_ptr(), 16)
        t2 = time.perf_counter()
        e2.record(ext)
        evs.append((e0, e1, e2))
        ix.free()
        host.append((1e3*(t1-t0), 1e3*(t2-t1)))
    ctx.synchronize(); torch.cuda.synchronize()
    print(mode, "e0->e2", np.round([a.elapsed_time(c) for a, b, c in evs[3:]], 3), "host build/normals ms", np.round(host[3:], 2).tolist(), flush=True)
os._exit(0)
