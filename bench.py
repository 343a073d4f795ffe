#!/usr/bin/env python
"""bench.py — headline benchmark of the kNN -> normals -> point-to-plane ICP hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, via the C ABI)
    python bench.py --impl reference ...                      # reference arm: CPU oracle port

Headline (`metric`/`value`): normals points/s at k=16 on BASELINE config 2 (a synthetic
120,000-point KITTI-shaped LiDAR frame), one step = one full `estimate_normals` pass
(index build + fused kNN/covariance/eigen/orientation kernel) with the input cloud already
resident in HBM and the output left in HBM.  At N > 1 every rank processes its own frame per step
(frames are the shard unit; weak scaling, no data-path collective).  `e2e` is the same metric
through the host-buffer C-ABI call `tc_estimate_normals` (pinned host in, host out; H2D + D2H
inside the timed region).  `extra` carries the other BASELINE configs: C4 (normals k=30 on a
10M-point cloud, queries sharded over the ranks on a replicated grid) and C3 (30 point-to-plane
ICP iterations on two 1M-point scans, source sharded, 29-scalar NCCL all-reduce per iteration).

Timing: CUDA events on the library's stream; >= 3 warm-ups; L2 is flushed (256 MiB write)
before every timed step; max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "normals_points_per_s_k16"
UNIT = "points/s"
K_C2 = 16
BYTES_NORMALS_PER_PT = 36.0      # 12 B read + 24 B NormalPoint3f written   
BYTES_ICP_PER_PT_ITER = 36.0     # 12 B read + 24 B gathered                
BYTES_INDEX_PER_PT = 116.0       # bbox + keys + radix passes + gather + ranges


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "20", "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
            time.sleep(0.25)  # let the sampler come up before the timed region starts
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nm in enumerate(names):
                    if any(r[5 + i].strip().lower() == "active" for r in rows if len(r) >= 9):
                        out["reasons"].append(nm)
                out["samples"] = len(sm)
        except Exception:
            pass
        return out


def _bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU (what numactl / NCCL's own
    affinity do in a deployment): the LiDAR-frame step is bounded by launch latency and by the
    GPU's writes into pinned host memory, both of which cross the socket interconnect otherwise."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:  # NVML enumerates in PCI order, CUDA may not: go through the bus id
            p = torch.cuda.get_device_properties(gpu_index)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(
                f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


# --------------------------------------------------------------------------------- reference arm
def run_reference(args, rank: int, world: int):
    """The reference's own CPU path cannot run here (pure Rust, no cargo/rustc): this arm times
    the C++ oracle port with all host threads on the same workload (cpu_baseline.kind = port)."""
    if rank != 0:
        return
    import oracle
    from fixtures import synth

    oracle.build()
    pts = synth.kitti_frame()
    n = pts.shape[0]
    # every host core this process may use (torchrun pins OMP_NUM_THREADS=1; the thread count is
    # passed explicitly, and under torchrun rank 0 alone runs this arm)
    threads = max(len(os.sched_getaffinity(0)), oracle.max_threads())
    for _ in range(max(args.warmup, 1)):
        oracle.estimate_normals(pts, K_C2, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.estimate_normals(pts, K_C2, threads=threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C2: estimate_normals k=16, 120000-pt KITTI-shaped frame "
                               "(kd-tree build + PCA normals), CPU oracle port, all host threads"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full C2 frames ({n} pts, k=16)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C3/C4 extra workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--c4-points", type=int, default=10_000_000)
    ap.add_argument("--c3-points", type=int, default=1_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import threecrate_b200 as tc
    from fixtures import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    affinity = _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = tc.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak_gbs, peak_kind = _peaks()

    def flush_l2():
        with torch.cuda.stream(ext):
            flush_buf.zero_()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ----------------------------------------------------------------------------- C2 headline
    # every rank takes the SAME frame: weak scaling with identical per-GPU work (frames drawn with
    # different seeds differ by +-15 % in kernel time, which max-over-ranks would book as a
    # scaling loss)
    pts = synth.kitti_frame(seed=0x3C0FFEE)
    n = pts.shape[0]
    h_in = tc.pinned_empty((n, 3))
    h_in[:] = pts
    h_out = tc.pinned_empty((n, 6))
    cloud = tc.DeviceCloud(h_in, ctx)
    d_out = ctx.alloc(n * 24)

    def step_resident(record=None):
        """index build + fused normals kernel; input and output resident in HBM."""
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(ext)
        index = tc.GridIndex(cloud, k_hint=K_C2)
        e1.record(ext)
        index.estimate_normals_device(d_out, K_C2)
        e2.record(ext)
        if record is not None:
            record.append((e0, e1, e2))
        return index

    # one nvidia-smi poller per JOB (rank 0's GPU): eight of them polling the driver every 20 ms
    # measurably slow every rank's launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        flush_l2()
        step_resident().free()
    ctx.synchronize()
    barrier()
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        step_resident(evs).free()
    ctx.synchronize()
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - l0
    ms_index = [a.elapsed_time(b) for a, b, _ in evs]
    ms_kernel = [b.elapsed_time(c) for _, b, c in evs]
    ms_step = [a.elapsed_time(c) for a, _, c in evs]
    total_ms = max_over_ranks(float(np.sum(ms_step)))
    value = n * world * args.steps / (total_ms * 1e-3)
    per_rank_ms = [float(np.sum(ms_step)) / args.steps]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank_ms[0])
        per_rank_ms = [float(x) for x in gathered]

    # e2e: host buffers through the drop-in C-ABI call, copies inside the timed region
    lib = ctx.lib
    import ctypes as C

    def step_e2e():
        ctx.check(lib.tc_estimate_normals(ctx.h, C.c_void_p(h_in.ctypes.data), n, K_C2, -1.0, 1, None,
                                          C.c_void_p(h_out.ctypes.data)))

    for _ in range(args.warmup):
        flush_l2()
        step_e2e()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush_l2()
        ctx.synchronize()
        t0 = time.perf_counter()
        step_e2e()          # returns after the D2H copy completed (synchronous host API)
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
    e2e_total = max_over_ranks(float(np.sum(e2e_ms)))
    e2e_value = n * world * args.steps / (e2e_total * 1e-3)
    clocks = sampler.stop()

    kern_s = float(np.mean(ms_kernel)) * 1e-3
    achieved = BYTES_NORMALS_PER_PT * n / kern_s / 1e9
    roofline = {"bound": "hbm", "kernel": "k_normals2<16,+1> (fused two-pass kNN + covariance + eigen + orientation; incl. the tie-list launch)",
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "peak_kind": f"of {peak_kind}", "traffic": 5.61e6,
                "traffic_source": "profiles/r01c_c2_raw.csv (ncu --set full, dram read+write per launch)",
                "algorithmic_bytes_per_launch": BYTES_NORMALS_PER_PT * n,
                "kernel_ms": 1e3 * kern_s, "index_build_ms": float(np.mean(ms_index)),
                "note": "C2 (1.4 MB) is L2-resident and issue-bound, not HBM-bound; see DESIGN.md"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: estimate_normals k=16 on a 120000-pt KITTI-shaped LiDAR frame "
                               "(index build + fused normals kernel), one frame per rank per step (the same "
                               "synthetic frame on every rank)",
                   "k": K_C2, "points_per_rank": n, "l2": "flushed (256 MiB write) before every step",
                   "timed": "CUDA events on the library stream, summed over steps, max over ranks",
                   "ms_per_step_by_rank": [round(x, 5) for x in per_rank_ms],
                   "cpu_affinity": (f"NVML-local CPUs of the rank's GPU ({len(affinity)} cores)"
                                    if affinity else "unchanged")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 12 * world),
                "d2h_bytes_per_step": int(n * 24 * world), "ms_per_step": e2e_total / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "wall_s_timed_region": t_wall,
    }
    ctx.free(d_out)
    cloud.free()

    # -------------------------------------------------------------------------- cpu_baseline
    if rank == 0 and world == 1 and not args.no_cpu:  # the CPU baseline is an N=1 figure
        try:
            import oracle

            oracle.build()
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core back
            threads = max(len(os.sched_getaffinity(0)), oracle.max_threads())
            oracle.estimate_normals(pts, K_C2, threads=threads)
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 200):
                oracle.estimate_normals(pts, K_C2, threads=threads)
                reps += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n * reps / dt, "unit": UNIT, "cores": threads,
                                    "kind": "port",
                                    "sample": f"{reps} full C2 frames ({n} pts, k=16), "
                                              "kd-tree build serial + OpenMP over points"}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": f"failed: {e}"}

    # ------------------------------------------------------------------------------- extras
    extra = {}
    if not args.no_extra:
        try:
            extra.update(bench_c4(args, tc, synth, ctx, ext, rank, world, barrier, max_over_ranks,
                                  peak_gbs, flush_l2, ev))
        except Exception as e:
            extra["c4_error"] = repr(e)
        try:
            extra.update(bench_c3(args, tc, synth, ctx, ext, rank, world, barrier, max_over_ranks,
                                  peak_gbs, flush_l2, ev, dist if world > 1 else None))
        except Exception as e:
            extra["c3_error"] = repr(e)
        if rank == 0:  # single-GPU rows of SURVEY §8(f): filters, multiscale ICP, GICP
            try:
                extra.update(bench_next_rows(tc, synth, ctx, ext, flush_l2, ev))
            except Exception as e:
                extra["next_rows_error"] = repr(e)
    line["extra"] = extra

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_c4(args, tc, synth, ctx, ext, rank, world, barrier, max_over_ranks, peak_gbs, flush_l2, ev):
    """C4: normals k=30 on a 10M-point cloud; grid replicated, queries sharded (strong scaling)."""
    n = args.c4_points
    pts = synth.terrain(n, 100.0 * (n / 10_000_000) ** 0.5, seed=4, noise=0.002)
    cloud = tc.DeviceCloud(pts, ctx)
    d_out = ctx.alloc(n * 24)
    from threecrate_b200.sharding import shard_range
    lo, hi = shard_range(rank, world, n)
    steps, warm = 3, 2
    res = []
    for it in range(warm + steps):
        flush_l2()
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(ext)
        index = tc.GridIndex(cloud, k_hint=30)     # redundant build on every rank
        e1.record(ext)
        index.estimate_normals_device(d_out, 30, shard=(lo, hi))
        e2.record(ext)
        ctx.synchronize()
        if it >= warm:
            res.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        info = index.info()
        index.free()
    barrier()
    ms_index = float(np.mean([r[0] for r in res]))
    ms_kernel = float(np.mean([r[1] for r in res]))
    ms_total = max_over_ranks(ms_index + ms_kernel)
    ms_kernel_max = max_over_ranks(ms_kernel)
    ctx.free(d_out)
    cloud.free()
    ach = 36.0 * (hi - lo) / (ms_kernel * 1e-3) / 1e9
    ach_build = BYTES_INDEX_PER_PT * n / (ms_index * 1e-3) / 1e9
    return {"c4_normals_k30": {
        "points": n, "n_gpus": world, "points_per_s": n / (ms_total * 1e-3),
        "points_per_s_kernel_only": n / (ms_kernel_max * 1e-3), "ms_index_build": ms_index,
        "ms_normals_kernel": ms_kernel_max, "scaling": "strong (queries sharded, grid replicated)",
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                     "frac": ach / peak_gbs, "kernel": "k_normals2<32> (16-candidate batches)",
                     "traffic": 1.124e9, "traffic_source": "profiles/r01c_c4_raw.csv"},
        "index_build_roofline": {"achieved": ach_build, "peak": peak_gbs, "unit": "GB/s",
                                 "frac": ach_build / peak_gbs, "bytes_per_point": BYTES_INDEX_PER_PT},
        "grid": {"cell_size": info["cell_size"], "dims": info["dims"],
                 "occupied_cells": info["occupied_cells"],
                 "max_cell_population": info["max_cell_population"]}}}


def bench_c3(args, tc, synth, ctx, ext, rank, world, barrier, max_over_ranks, peak_gbs, flush_l2, ev,
             dist):
    """C3: 30 point-to-plane ICP iterations, 1M <-> 1M; target replicated, source sharded,
    29-scalar all-reduce per iteration at N > 1."""
    n = args.c3_points
    src, tgt, nrm, T = synth.scan_pair(n, half_extent=50.0 * (n / 1_000_000) ** 0.5)
    from threecrate_b200.sharding import shard_range
    lo, hi = shard_range(rank, world, n)
    comm = None
    if world > 1:
        ids = [tc.Comm.unique_id(ctx) if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = tc.Comm(ctx, ids[0], world, rank)
        handles = [None] * world          # NVLink peer buffers for the fused all-reduce
        dist.all_gather_object(handles, comm.peer_handle())
        comm.open_peers(handles)
    tcloud = tc.DeviceCloud(tgt, ctx)
    scloud = tc.DeviceCloud(src[lo:hi], ctx)
    d_nrm = ctx.alloc(n * 12)
    ctx.to_device(d_nrm, nrm)
    iters = 30
    steps, warm = 3, 2
    res = []
    r = None
    for it in range(warm + steps):
        flush_l2()
        barrier()
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(ext)
        index = tc.GridIndex(tcloud, k_hint=1)
        e1.record(ext)
        r = tc.icp_point_to_plane_device(scloud, index, d_nrm, tc.IDENTITY, iters, None, -1.0, comm)
        e2.record(ext)
        ctx.synchronize()
        if it >= warm:
            res.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        index.free()
    ms_index = float(np.mean([x[0] for x in res]))
    ms_icp = max_over_ranks(float(np.mean([x[1] for x in res])))
    ctx.free(d_nrm)
    scloud.free()
    tcloud.free()
    if comm:
        comm.destroy()
    t_err = float(np.linalg.norm(r.translation.astype(np.float64) - T[:3]))
    ach = BYTES_ICP_PER_PT_ITER * (hi - lo) * iters / (ms_icp * 1e-3) / 1e9
    return {"c3_icp_point_to_plane": {
        "source_points": n, "target_points": n, "iterations": iters, "n_gpus": world,
        "iters_per_s": iters / (ms_icp * 1e-3),
        "iters_per_s_incl_index_build": iters / ((ms_icp + ms_index) * 1e-3),
        "ms_per_iter": ms_icp / iters, "ms_index_build": ms_index,
        "translation_error_vs_ground_truth": t_err,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                     "frac": ach / peak_gbs, "kernel": "k_icp_correspond (+ k_icp_solve)",
                     "note": "36 MB/iter working set is L2-resident; issue/latency-bound"}}}


def bench_next_rows(tc, synth, ctx, ext, flush_l2, ev):
    """Filters on the C2 frame (device-resident in and out) and the two other registrations on a
    200k terrain pair (host arrays through the C ABI, uploads included)."""
    def timed(fn, reps=5, warm=2):
        ts = []
        for it in range(warm + reps):
            flush_l2()
            e0, e1 = ev(), ev()
            e0.record(ext)
            r = fn()
            e1.record(ext)
            ctx.synchronize()
            if it >= warm:
                ts.append(e0.elapsed_time(e1))
            if isinstance(r, tc.DeviceCloud):
                r.free()
        return float(np.median(ts))

    pts = synth.kitti_frame()
    n = len(pts)
    cloud = tc.DeviceCloud(pts, ctx)
    out = {}
    ms = timed(lambda: tc.voxel_grid_filter(cloud, 0.2))
    out["voxel_grid_filter_0.2m"] = {"points": n, "ms": ms, "points_per_s": n / (ms * 1e-3)}
    ms = timed(lambda: tc.radius_outlier_removal(cloud, 0.5, 5))
    out["radius_outlier_removal_r0.5_min5"] = {"points": n, "ms": ms, "points_per_s": n / (ms * 1e-3)}
    ms = timed(lambda: tc.statistical_outlier_removal(cloud, 16, 1.0))
    out["statistical_outlier_removal_k16_exact"] = {"points": n, "ms": ms, "points_per_s": n / (ms * 1e-3)}
    ms = timed(lambda: tc.statistical_outlier_removal(cloud, 16, 1.0, fast=True))
    out["statistical_outlier_removal_k16_fast"] = {"points": n, "ms": ms, "points_per_s": n / (ms * 1e-3)}
    cloud.free()
    src, tgt, _, _ = synth.scan_pair(200_000, half_extent=22.0)
    r = [None]

    def run_gicp():
        r[0] = tc.gicp(src, tgt, tc.IDENTITY, tc.GicpConfig(max_iterations=20), ctx,
                       want_correspondences=False)
    ms = timed(run_gicp, reps=3, warm=1)
    out["gicp_200k_k20"] = {"ms": ms, "iterations": r[0].iterations,
                            "ms_per_iteration_incl_covariances": ms / max(r[0].iterations, 1)}

    def run_ms():
        r[0] = tc.multiscale_icp_point_to_point(
            src, tgt, tc.IDENTITY, tc.MultiScaleIcpConfig(
                levels=[tc.IcpScaleLevel(1.0, 10, 2.0), tc.IcpScaleLevel(0.5, 10, 1.0),
                        tc.IcpScaleLevel(0.25, 15, 0.6)], final_max_correspondence_distance=0.4),
            ctx, want_correspondences=False)
    ms = timed(run_ms, reps=3, warm=1)
    out["multiscale_icp_200k_3_levels"] = {"ms": ms, "iterations": r[0].iterations}
    return {"next_rows": out}


if __name__ == "__main__":
    main()
