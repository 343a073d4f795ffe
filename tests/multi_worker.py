"""Worker for tests/test_gpu_multi.py (launched by torchrun, one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import threecrate_b200 as tc  # noqa: E402
from fixtures import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = tc.Context(local)

# ---- ICP: source sharded, target replicated, all-reduce of the normal equations
n = 200_000
src, tgt, nrm, T = synth.scan_pair(n, half_extent=22.0)
ids = [tc.Comm.unique_id(ctx) if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = tc.Comm(ctx, ids[0], world, rank)
tcloud = tc.DeviceCloud(tgt, ctx)
index = tc.GridIndex(tcloud, k_hint=1)
d_nrm = ctx.alloc(n * 12)
ctx.to_device(d_nrm, nrm)
lo, hi = rank * n // world, (rank + 1) * n // world
shard = tc.DeviceCloud(src[lo:hi], ctx)
r_nccl = tc.icp_point_to_plane_device(shard, index, d_nrm, tc.IDENTITY, 20, None, -1.0, comm)
# fused path: IPC-map every rank's exchange buffer, all-reduce inside the correspondence kernel
handles = [None] * world
dist.all_gather_object(handles, comm.peer_handle())
comm.open_peers(handles)
r = tc.icp_point_to_plane_device(shard, index, d_nrm, tc.IDENTITY, 20, None, -1.0, comm)
assert np.array_equal(r.transformation, r_nccl.transformation), "fused != NCCL all-reduce"
assert r.iterations == r_nccl.iterations and r.mse == r_nccl.mse
rp = tc.icp_point_to_point_device(shard, index, tc.IDENTITY, 8, None, 1e-9, comm)
assert rp.iterations >= 1 and np.isfinite(rp.transformation).all()
# all ranks hold the identical transform
t = torch.tensor(r.transformation, device="cuda")
gathered = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(gathered, t)
for g in gathered:
    assert torch.equal(g, gathered[0]), "ranks disagree on the transform"
# and it matches the single-GPU run to f32 rounding
if rank == 0:
    full = tc.icp_point_to_plane_device(tc.DeviceCloud(src, ctx), index, d_nrm, tc.IDENTITY, 20,
                                        None, -1.0, None)
    dt = np.abs(full.transformation - r.transformation).max()
    print(f"sharded vs single ICP: max |dT| = {dt:.3e}; t_err vs truth "
          f"{np.linalg.norm(r.translation - T[:3]):.3e}")
    assert dt < 2e-6 and r.iterations == full.iterations == 20
    assert abs(r.mse - full.mse) <= 1e-5 * abs(full.mse) + 1e-12

# ---- normals: queries sharded by sorted position over a replicated grid (whole-cell ownership)
pts = synth.kitti_frame()
m = len(pts)
cloud = tc.DeviceCloud(pts, ctx)
ix = tc.GridIndex(cloud, k_hint=16)
d_out = ctx.alloc(m * 24)
torch.cuda.synchronize()
# poison, then each rank writes only its shard
ctx.to_device(d_out, np.full((m, 6), np.nan, np.float32))
ix.estimate_normals_device(d_out, 16, shard=(rank * m // world, (rank + 1) * m // world))
mine = np.empty((m, 6), np.float32)
ctx.to_host(mine, d_out)
written = ~np.isnan(mine[:, 0])
cnt = torch.tensor(written.astype(np.int32), device="cuda")
dist.all_reduce(cnt)
assert int(cnt.min()) == 1 and int(cnt.max()) == 1, "shards must tile the cloud exactly once"
ref = ix.estimate_normals(16)
assert np.array_equal(mine[written], ref[written])
# ---- distributed estimate_normals: chunks in, chunks out, rows routed over NVLink windows
for cloud_pts, kk in ((synth.terrain(300_001, 18.0, seed=11, noise=0.002), 16), (pts, 10)):
    nn = len(cloud_pts)
    wh = [None] * world
    dist.all_gather_object(wh, comm.window_handle(nn))
    comm.open_window(wh)
    lo, hi = comm.chunk(nn)
    full_ix = tc.GridIndex(tc.DeviceCloud(cloud_pts, ctx), k_hint=kk)
    ref = full_ix.estimate_normals(kk)
    for _ in range(3):
        got = comm.estimate_normals(cloud_pts[lo:hi], nn, kk)
        assert np.array_equal(got.view(np.uint32), ref[lo:hi].view(np.uint32)), \
            "distributed normals differ from the single-GPU rows"
dist.barrier()
if rank == 0:
    print("MULTI_OK")
comm.destroy()
dist.destroy_process_group()
