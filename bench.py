#!/usr/bin/env python
"""bench.py — headline benchmark of the kNN -> normals -> point-to-plane ICP hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, via the C ABI)
    python bench.py --impl reference ...                      # reference arm: CPU oracle port

Headline (`metric`/`value`, every N): normals points/s at k=16 on a synthetic 10M-point cloud - the
workload BASELINE.json's north_star quotes its target on ("normals (k=16) on a 10M-point cloud ...
scaling >= 6x at 8 GPUs").  One step = one full `estimate_normals` pass: index build + fused
kNN / covariance / eigen / orientation kernel, input cloud resident in HBM, output left in HBM.
At N > 1 the QUERIES are sharded over the ranks (contiguous ranges of the cell-sorted order) on a
grid every rank builds itself: strong scaling, no data-path collective; the timed region is
bracketed by barriers and the value is the 10M points over the slowest rank's time.
`e2e` is the same metric through the host-buffer C-ABI call `tc_estimate_normals` (pinned host in,
host out; H2D + D2H inside the timed region); `e2e_pageable` repeats it from ordinary pageable
memory, which is what a Rust `Vec<Point3f>` is.

`extra` carries the other BASELINE configs, each with its own roofline and (N = 1) CPU leg:
  c2  k=16 on the 120k-point KITTI-shaped frame (configs[1]; N = 1 only - one frame does not shard)
  c4  k=30 on the 10M cloud, queries sharded                       (configs[3])
  c3  30 point-to-plane ICP iterations on two 1M-point scans, source sharded, 29-scalar
      all-reduce per iteration fused into the correspondence kernel (configs[2])
  c5  ICP on a 100M-point target with 12.5M source points per rank (= configs[4] at N = 8)
At N > 1 the line also carries `parity`: the sharded results compared with the single-GPU call
(normals rows tile the cloud exactly once and are bit-identical; ICP transforms agree; the fused
peer all-reduce equals the NCCL path bit for bit).  A mismatch exits non-zero.

Timing: CUDA events on the library's stream; >= 3 warm-ups; L2 is flushed (256 MiB write) before
every timed step (the 10M working set is larger than L2 anyway); max over ranks.
"""
from __future__ import annotations

import argparse
import csv
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
K_HEAD = 16
N_HEAD = 10_000_000
BYTES_NORMALS_PER_PT = 36.0      # 12 B read + 24 B NormalPoint3f written   
BYTES_ICP_PER_PT_ITER = 36.0     # 12 B read + 24 B gathered                
# bytes the BUILT index pipeline must move per point and per cell: bbox 12 r,
# histogram 12 r, scatter 12 r + 16 w = 52 B/pt; histogram table 4 w (atomics) + 4 r (statistics)
# + scan 4 r + 4 w + scatter cursor 4 r + 4 w = 24 B/cell
BYTES_INDEX_PER_PT = 52.0
BYTES_INDEX_PER_CELL = 24.0
WORKLOAD = ("estimate_normals k=16 on a 10,000,000-point synthetic terrain cloud (index build + "
            "fused normals kernel); at N > 1 every rank builds the index of its slab of cell planes "
            "(+ halo) and computes the rows of the points in its slab")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def _profile_numbers(tag: str, kernel_substr: str):
    """(dram bytes, warp instructions, source string, points) of one launch of a kernel, read from
    the committed ncu raw export profiles/<tag>_raw.csv (its command and commit are recorded in
    profiles/<tag>.meta.json).  Missing pieces come back as None with the reason in the string."""
    path = os.path.join(ROOT, "profiles", f"{tag}_raw.csv")
    meta = os.path.join(ROOT, "profiles", f"{tag}.meta.json")
    if not os.path.exists(path):
        return None, None, f"profiles/{tag}_raw.csv not found", None
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if kernel_substr in r[col["Kernel Name"]]:
                tr = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tr += float(r[col[m]].replace(",", "")) * scale.get(units[col[m]], 1.0)
                inst = float(r[col["smsp__inst_executed.sum"]].replace(",", ""))
                src, pts = f"profiles/{tag}_raw.csv", None
                if os.path.exists(meta):
                    mj = json.load(open(meta))
                    pts = mj.get("points")
                    src += f" (ncu --set full at commit {mj.get('commit', '?')}, {pts} points)"
                return tr, inst, src, pts
    except Exception as e:  # the roofline block must never break the bench
        return None, None, f"profiles/{tag}_raw.csv unreadable: {e}", None
    return None, None, f"kernel {kernel_substr} not in profiles/{tag}_raw.csv", None


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
    affinity do in a deployment)."""
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


def head_cloud(n: int):
    """The headline / C4 cloud: terrain over [-100,100]^2 at 10M points (same density when n is
    scaled down), seed 4, 2 mm noise."""
    from fixtures import synth
    return synth.terrain(n, 100.0 * (n / 10_000_000) ** 0.5, seed=4, noise=0.002)


# --------------------------------------------------------------------------------- reference arm
def run_reference(args, rank: int, world: int):
    """The reference's own CPU path cannot run here (pure Rust, no cargo/rustc): this arm times
    the C++ oracle port (kd-tree build serial, OpenMP over the loops the reference gives to rayon)
    with every host thread, on a bounded sample of the headline workload: a terrain cloud of the
    same density and k, sized so one step takes a few seconds."""
    if rank != 0:
        return
    import oracle

    oracle.build()
    threads = max(len(os.sched_getaffinity(0)), oracle.max_threads())
    n = args.ref_points
    pts = head_cloud(n)
    if args.warmup > 0:  # one warm-up pass: a pass is seconds of CPU work
        oracle.estimate_normals(pts, K_HEAD, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.estimate_normals(pts, K_HEAD, threads=threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = (f"{args.steps} passes over a {n}-point terrain cloud of the headline density "
              f"(1/{max(N_HEAD // n, 1)} of the 10M workload), k=16: kd-tree build serial + "
              "OpenMP over points")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "k": K_HEAD, "points": N_HEAD,
                   "reference_sample_points": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------- our arm
class Env:
    """Everything the workload functions share."""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C2/C3/C4/C5 extra workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-c5", action="store_true", help="skip the 100M-point ICP extra")
    ap.add_argument("--only", default="", help="comma list of extras to run (c4,c3,c2,c5,next_rows)")
    ap.add_argument("--points", type=int, default=N_HEAD, help="headline cloud size")
    ap.add_argument("--c3-points", type=int, default=1_000_000)
    ap.add_argument("--c5-target", type=int, default=100_000_000)
    ap.add_argument("--c5-source-per-rank", type=int, default=12_500_000)
    ap.add_argument("--ref-points", type=int, default=1_000_000,
                    help="reference arm / cpu_baseline sample size")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    E = Env()
    E.args, E.rank, E.world, E.local, E.tc, E.torch, E.dist = args, rank, world, local, tc, torch, dist
    E.all_cpus = os.sched_getaffinity(0)
    E.affinity = _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    E.barrier = barrier
    E.max_over_ranks = lambda x: reduce_ranks(x, dist.ReduceOp.MAX)
    E.ctx = ctx = tc.Context(local)
    E.ext = ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    E.peak_gbs, E.peak_kind = _peaks()
    E.parity_failures = []

    def flush_l2():
        with torch.cuda.stream(ext):
            flush_buf.zero_()

    E.flush_l2 = flush_l2
    E.ev = lambda: torch.cuda.Event(enable_timing=True)

    # ---------------------------------------------------------------------------- headline
    n = args.points
    pts = head_cloud(n)
    from threecrate_b200.sharding import shard_range
    lo, hi = shard_range(rank, world, n)   # nominal share; the slab build owns whole cell planes
    shard = (rank, world) if world > 1 else None
    E.shard = shard
    h_in = tc.pinned_empty((n, 3))
    h_in[:] = pts
    h_out = tc.pinned_empty((n, 6))
    cloud = tc.DeviceCloud(h_in, ctx)
    out_t = torch.zeros((n, 6), dtype=torch.float32, device="cuda")  # rows by original index
    d_out = out_t.data_ptr()
    torch.cuda.synchronize()

    def step_resident(record=None):
        """index build (every rank) + fused normals kernel on this rank's shard."""
        e0, e1, e2 = E.ev(), E.ev(), E.ev()
        e0.record(ext)
        index = tc.GridIndex(cloud, k_hint=K_HEAD, shard=shard)
        e1.record(ext)
        index.estimate_normals_device(d_out, K_HEAD)
        e2.record(ext)
        if record is not None:
            record.append((e0, e1, e2))
        return index

    sampler = ClockSampler(local)  # one nvidia-smi poller per JOB (rank 0's GPU)
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
    total_ms = E.max_over_ranks(float(np.sum(ms_step)))
    value = n * args.steps / (total_ms * 1e-3)
    per_rank_ms = [float(np.sum(ms_step)) / args.steps]
    per_rank_split = [(float(np.mean(ms_index)), float(np.mean(ms_kernel)))]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (per_rank_ms[0],) + per_rank_split[0])
        per_rank_ms = [float(x[0]) for x in gathered]
        per_rank_split = [(float(x[1]), float(x[2])) for x in gathered]
    info = tc.GridIndex(cloud, k_hint=K_HEAD)
    grid_info = info.info()
    info.free()

    # ---- e2e: host buffers through the C-ABI call, copies inside the timed region.
    # N = 1: the drop-in tc_estimate_normals.  N > 1: tc_estimate_normals_distributed - every rank
    # passes ITS contiguous chunk of the rows (1/N of the upload), the chunks are exchanged over
    # NVLink peer memory, and every rank gets the normal rows of its chunk back (1/N of the
    # download); window set-up (once per cloud size) is outside the timed region like the
    # communicator's.
    lib = ctx.lib
    import ctypes as C
    h_page_in = np.array(pts, copy=True)          # ordinary pageable memory (a Rust Vec<Point3f>)
    h_page_out = np.empty((n, 6), np.float32)
    dcomm = None
    if world > 1:
        dcomm = _make_comm(E, peers=False)
        wh = [None] * world
        dist.all_gather_object(wh, dcomm.window_handle(n))
        dcomm.open_window(wh)
        c_lo, c_hi = dcomm.chunk(n)

    def e2e_pass(src, dst):
        if world == 1:
            ctx.check(lib.tc_estimate_normals(ctx.h, C.c_void_p(src.ctypes.data), n, K_HEAD, -1.0, 1,
                                              None, C.c_void_p(dst.ctypes.data)))
            return
        dcomm.estimate_normals(src[c_lo:c_hi], n, K_HEAD, out=dst[c_lo:c_hi])

    def time_e2e(src, dst, steps):
        for _ in range(2):
            flush_l2()
            e2e_pass(src, dst)
        barrier()
        ms = []
        for _ in range(steps):
            flush_l2()
            ctx.synchronize()
            barrier()               # (the call is collective: start the ranks together)
            t0 = time.perf_counter()
            e2e_pass(src, dst)      # returns after the D2H copy completed (synchronous host API)
            ms.append(1e3 * (time.perf_counter() - t0))
        tot = E.max_over_ranks(float(np.sum(ms)))
        return n * steps / (tot * 1e-3), tot / steps

    e2e_value, e2e_ms = time_e2e(h_in, h_out, args.steps)
    e2e_pg_value, e2e_pg_ms = time_e2e(h_page_in, h_page_out, max(3, args.steps // 4))
    clocks = sampler.stop()
    del h_page_in, h_page_out
    dist_parity = None
    if world > 1:
        # the rows the distributed call returned vs the single-GPU call, bit for bit
        h_out[c_lo:c_hi] = 0
        e2e_pass(h_in, h_out)
        ix1 = tc.GridIndex(cloud, k_hint=K_HEAD)
        ix1.estimate_normals_device(d_out, K_HEAD)
        ref_rows = np.empty((n, 6), np.float32)
        ctx.to_host(ref_rows, d_out)
        ix1.free()
        bad = int(np.any(h_out[c_lo:c_hi].view(np.uint32) != ref_rows[c_lo:c_hi].view(np.uint32),
                         axis=1).sum())
        bad_all = int(E.max_over_ranks(float(bad)))
        dist_parity = {"rows_differing_max_over_ranks": bad_all,
                       "bitwise_equal_to_single_gpu": bad_all == 0}
        if bad_all:
            E.parity_failures.append(f"distributed normals: {bad_all} rows differ")
        del ref_rows
        dcomm.destroy()

    # ---- roofline of the dominant kernel (this rank's launch), live CUDA-event time
    kern_s = float(np.mean(ms_kernel)) * 1e-3
    q_launch = hi - lo
    achieved = BYTES_NORMALS_PER_PT * q_launch / kern_s / 1e9
    tr, inst, tr_src, prof_pts = _profile_numbers("r02_head", "k_normals2")
    scaled = bool(tr and prof_pts)
    roofline = {
        "bound": "hbm",
        "kernel": "k_normals2<16,+1> (fused two-pass kNN + covariance + eigen + orientation; ties at "
                  "rank k resolved inside the kernel)",
        "achieved": achieved, "peak": E.peak_gbs, "unit": "GB/s", "frac": achieved / E.peak_gbs,
        "peak_kind": f"of {E.peak_kind}",
        "traffic": (tr * q_launch / prof_pts) if scaled else None,
        "traffic_source": tr_src + ("; scaled to this launch's query count" if scaled else ""),
        "algorithmic_bytes_per_launch": BYTES_NORMALS_PER_PT * q_launch,
        "kernel_ms": 1e3 * kern_s, "index_build_ms": float(np.mean(ms_index)),
        "queries_per_launch": int(q_launch),
        "note": "exact kNN is bounded by instruction issue, not by HBM (see roofline_issue)"}
    cells = float(np.prod(grid_info["dims"]))
    build_bytes = BYTES_INDEX_PER_PT * n + BYTES_INDEX_PER_CELL * cells
    ach_b = build_bytes / (float(np.mean(ms_index)) * 1e-3) / 1e9
    index_roofline = {"bound": "hbm", "achieved": ach_b, "peak": E.peak_gbs, "unit": "GB/s",
                      "frac": ach_b / E.peak_gbs, "bytes": build_bytes,
                      "bytes_model": "52 B/point + 24 B/cell of the built counting-sort pipeline",
                      "cells": cells, "ms": float(np.mean(ms_index))}
    try:
        peak_issue = ctx.issue_rate()
        if inst and prof_pts:
            ipq = inst / prof_pts
            ach_i = ipq * q_launch / kern_s
            issue = {"bound": "issue", "achieved": ach_i, "peak": peak_issue, "unit": "warp-inst/s",
                     "frac": ach_i / peak_issue, "warp_inst_per_query": ipq,
                     "warp_inst_source": tr_src,
                     "peak_source": "k_issue_rate microbenchmark (16 independent FMNMX/FMUL chains "
                                    "per thread), measured in this run"}
        else:
            issue = {"bound": "issue", "achieved": None, "peak": peak_issue, "unit": "warp-inst/s",
                     "frac": None, "note": tr_src}
    except Exception as e:
        issue = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "k": K_HEAD, "points": n,
                   "queries_per_rank": int(hi - lo),
                   "l2": "flushed (256 MiB write) before every step; working set > L2",
                   "timed": "CUDA events on the library stream, summed over steps, max over ranks",
                   "ms_per_step_by_rank": [round(x, 5) for x in per_rank_ms],
                   "ms_build_and_kernel_by_rank": [[round(b, 4), round(k, 4)]
                                                   for b, k in per_rank_split],
                   "grid": {"cell_size": grid_info["cell_size"], "dims": list(grid_info["dims"]),
                            "levels": grid_info["n_levels"]},
                   "cpu_affinity": (f"NVML-local CPUs of the rank's GPU ({len(E.affinity)} cores)"
                                    if E.affinity else "unchanged")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 12),
                "d2h_bytes_per_step": int(n * 24), "ms_per_step": e2e_ms,
                "host_memory": "pinned",
                "call": ("tc_estimate_normals" if world == 1 else
                         "tc_estimate_normals_distributed (each rank moves its 1/N chunk of the "
                         "rows over PCIe; chunks and result rows cross NVLink peer windows)")},
        "e2e_pageable": {"value": e2e_pg_value, "unit": UNIT, "ms_per_step": e2e_pg_ms,
                         "host_memory": "pageable (what a Rust Vec<Point3f> is)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_issue": issue,
        "index_build_roofline": index_roofline,
        "wall_s_timed_region": t_wall,
    }

    # ---- N > 1: the sharded rows tile the cloud exactly once and equal the single-GPU result
    if world > 1:
        line["parity"] = {"normals_k16": parity_normals(E, cloud, out_t, lo, hi, K_HEAD),
                          "normals_k16_distributed_e2e": dist_parity}

    # -------------------------------------------------------------------------- cpu_baseline
    if rank == 0 and world == 1 and not args.no_cpu:  # the CPU baseline is an N=1 figure
        line["cpu_baseline"] = cpu_normals(E, K_HEAD, args.ref_points)

    # ------------------------------------------------------------------------------- extras
    extra = {}
    if not args.no_extra:
        for name, fn in (("c4", lambda: bench_c4(E, cloud, out_t, lo, hi)),
                         ("c3", lambda: bench_c3(E)),
                         ("c2", lambda: bench_c2(E) if world == 1 else {}),
                         ("c5", lambda: bench_c5(E) if not args.no_c5 else {}),
                         ("next_rows", lambda: bench_next_rows(E) if rank == 0 and world == 1 else {})):
            if args.only and name not in args.only.split(","):
                continue
            try:
                extra.update(fn())
            except Exception as e:
                extra[f"{name}_error"] = repr(e)
    line["extra"] = extra
    if world > 1:
        line["parity"]["failures"] = E.parity_failures

    fail = len(E.parity_failures) > 0
    if world > 1:
        fail = E.max_over_ranks(1.0 if fail else 0.0) > 0.0
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if fail:
        sys.stderr.write("bench.py: multi-GPU parity FAILED: " + "; ".join(E.parity_failures) + "\n")
        sys.exit(3)


# ------------------------------------------------------------------------------- parity helpers
def parity_normals(E, cloud, out_t, lo, hi, k):
    """Sharded rows vs the single-GPU call.  Every rank's row buffer holds only ITS rows (zero
    elsewhere): the all-reduced sum is the assembled result, and a per-row ownership count must
    be exactly one everywhere."""
    tc, torch, dist, ctx = E.tc, E.torch, E.dist, E.ctx
    n = cloud.n
    out_t.zero_()
    torch.cuda.synchronize()   # torch's streams and the library's stream are not ordered
    index = tc.GridIndex(cloud, k_hint=k, shard=E.shard)
    index.estimate_normals_device(out_t.data_ptr(), k)
    ctx.synchronize()
    torch.cuda.synchronize()
    index.free()
    owned = (out_t[:, 3:].abs().sum(dim=1) > 0).to(torch.int32)  # a written row has a unit normal
    dist.all_reduce(owned)
    assembled = out_t.clone()
    dist.all_reduce(assembled)
    single = torch.zeros_like(out_t)
    torch.cuda.synchronize()
    index = tc.GridIndex(cloud, k_hint=k)                # complete index, whole cloud on this GPU
    index.estimate_normals_device(single.data_ptr(), k)
    ctx.synchronize()
    torch.cuda.synchronize()
    index.free()
    rows_once = bool((owned == 1).all().item())
    differing = int((assembled != single).any(dim=1).sum().item())
    identical = differing == 0
    res = {"rows_written_exactly_once": rows_once, "rows": int(n), "rows_differing": differing,
           "rows_missing": int((owned == 0).sum().item()),
           "rows_duplicated": int((owned > 1).sum().item()),
           "assembled_equals_single_gpu_bitwise": identical}
    if not rows_once:
        E.parity_failures.append(f"normals k={k}: shards do not tile the cloud exactly once")
    if not identical:
        E.parity_failures.append(f"normals k={k}: sharded rows differ from the single-GPU rows")
    del owned, assembled, single
    return res


def cpu_normals(E, k, n):
    """CPU leg: the C++ oracle port on a terrain cloud of the workload's density."""
    try:
        import oracle
        oracle.build()
        os.sched_setaffinity(0, E.all_cpus)  # the CPU baseline gets every host core back
        threads = max(len(os.sched_getaffinity(0)), oracle.max_threads())
        pts = head_cloud(n)
        t0 = time.perf_counter()
        reps = 0
        while reps < 1 or (time.perf_counter() - t0 < 12.0 and reps < 50):
            oracle.estimate_normals(pts, k, threads=threads)
            reps += 1
        dt = time.perf_counter() - t0
        return {"value": n * reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{reps} passes over a {n}-point terrain cloud of the workload's density "
                          f"(1/{max(N_HEAD // n, 1)} of the 10M cloud), k={k}: kd-tree build serial "
                          "+ OpenMP over points"}
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}


# ------------------------------------------------------------------------------------- extras
def bench_c4(E, cloud, out_t, lo, hi):
    """C4: normals k=30 on the 10M cloud; grid replicated, queries sharded (strong scaling)."""
    tc, ctx, ext = E.tc, E.ctx, E.ext
    n, k = cloud.n, 30
    d_out = out_t.data_ptr()
    steps, warm = 10, 3
    res = []
    info = None
    for it in range(warm + steps):
        E.flush_l2()
        e0, e1, e2 = E.ev(), E.ev(), E.ev()
        e0.record(ext)
        index = tc.GridIndex(cloud, k_hint=k, shard=E.shard)   # slab build on every rank
        e1.record(ext)
        index.estimate_normals_device(d_out, k)
        e2.record(ext)
        ctx.synchronize()
        if it >= warm:
            res.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        info = index.info()
        index.free()
    E.barrier()
    ms_index = float(np.mean([r[0] for r in res]))
    ms_kernel = float(np.mean([r[1] for r in res]))
    ms_total = E.max_over_ranks(ms_index + ms_kernel)
    ms_kernel_max = E.max_over_ranks(ms_kernel)
    ach = BYTES_NORMALS_PER_PT * (hi - lo) / (ms_kernel * 1e-3) / 1e9
    tr, inst, tr_src, prof_pts = _profile_numbers("r02_c4", "k_normals2")
    cells = float(np.prod(info["dims"]))
    build_bytes = BYTES_INDEX_PER_PT * n + BYTES_INDEX_PER_CELL * cells
    ach_build = build_bytes / (ms_index * 1e-3) / 1e9
    out = {
        "points": n, "k": k, "n_gpus": E.world, "timed_reps": steps,
        "points_per_s": n / (ms_total * 1e-3),
        "points_per_s_kernel_only": n / (ms_kernel_max * 1e-3), "ms_index_build": ms_index,
        "ms_normals_kernel": ms_kernel_max, "scaling": "strong (queries sharded, grid replicated)",
        "roofline": {"bound": "hbm", "achieved": ach, "peak": E.peak_gbs, "unit": "GB/s",
                     "frac": ach / E.peak_gbs, "kernel": "k_normals2<32> (16-candidate batches)",
                     "traffic": (tr * (hi - lo) / prof_pts) if (tr and prof_pts) else None,
                     "traffic_source": tr_src},
        "index_build_roofline": {"achieved": ach_build, "peak": E.peak_gbs, "unit": "GB/s",
                                 "frac": ach_build / E.peak_gbs, "bytes": build_bytes,
                                 "bytes_model": "52 B/point + 24 B/cell (built pipeline)"},
        "grid": {"cell_size": info["cell_size"], "dims": list(info["dims"]),
                 "occupied_cells": info["occupied_cells"],
                 "max_cell_population": info["max_cell_population"]}}
    if E.world > 1:
        out["parity"] = parity_normals(E, cloud, out_t, lo, hi, k)
    if E.rank == 0 and E.world == 1 and not E.args.no_cpu:
        out["cpu_baseline"] = cpu_normals(E, k, E.args.ref_points)
    return {"c4_normals_k30": out}


def bench_c2(E):
    """C2 (BASELINE configs[1]): k=16 on the 120k-point KITTI-shaped frame, one GPU."""
    from fixtures import synth
    import ctypes as C
    tc, ctx, ext = E.tc, E.ctx, E.ext
    pts = synth.kitti_frame(seed=0x3C0FFEE)
    n, k = pts.shape[0], 16
    h_in = tc.pinned_empty((n, 3))
    h_in[:] = pts
    h_out = tc.pinned_empty((n, 6))
    cloud = tc.DeviceCloud(h_in, ctx)
    d_out = ctx.alloc(n * 24)
    steps, warm = 100, 5
    res = []
    for it in range(warm + steps):
        E.flush_l2()
        e0, e1, e2 = E.ev(), E.ev(), E.ev()
        e0.record(ext)
        index = tc.GridIndex(cloud, k_hint=k)
        e1.record(ext)
        index.estimate_normals_device(d_out, k)
        e2.record(ext)
        if it >= warm:
            res.append((e0, e1, e2))
        index.free()
    ctx.synchronize()
    ms_index = float(np.mean([a.elapsed_time(b) for a, b, _ in res]))
    ms_kernel = float(np.mean([b.elapsed_time(c) for _, b, c in res]))
    e2e = []
    for it in range(5 + 50):
        E.flush_l2()
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.check(ctx.lib.tc_estimate_normals(ctx.h, C.c_void_p(h_in.ctypes.data), n, k, -1.0, 1,
                                              None, C.c_void_p(h_out.ctypes.data)))
        if it >= 5:
            e2e.append(1e3 * (time.perf_counter() - t0))
    ctx.free(d_out)
    cloud.free()
    ach = BYTES_NORMALS_PER_PT * n / (ms_kernel * 1e-3) / 1e9
    tr, inst, tr_src, prof_pts = _profile_numbers("r02_c2", "k_normals2")
    out = {"points": n, "k": k, "timed_reps": steps,
           "points_per_s": n / ((ms_index + ms_kernel) * 1e-3),
           "ms_index_build": ms_index, "ms_normals_kernel": ms_kernel,
           "e2e_points_per_s": n / (float(np.mean(e2e)) * 1e-3), "e2e_ms": float(np.mean(e2e)),
           "roofline": {"bound": "hbm", "achieved": ach, "peak": E.peak_gbs, "unit": "GB/s",
                        "frac": ach / E.peak_gbs, "kernel": "k_normals2<16,+1>", "traffic": tr,
                        "traffic_source": tr_src,
                        "note": "1.4 MB working set: L2-resident, one wave of 3750 warps, bounded "
                                "by the slowest warp"}}
    if not E.args.no_cpu:
        try:
            import oracle
            threads = max(len(os.sched_getaffinity(0)), oracle.max_threads())
            oracle.estimate_normals(pts, k, threads=threads)
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t0 < 5.0 and reps < 200):
                oracle.estimate_normals(pts, k, threads=threads)
                reps += 1
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": n * reps / dt, "unit": UNIT, "cores": threads,
                                   "kind": "port", "sample": f"{reps} full C2 frames ({n} pts, k=16)"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "sample": f"failed: {e}"}
    return {"c2_normals_k16_kitti_frame": out}


def _make_comm(E, peers=True):
    tc, ctx, dist = E.tc, E.ctx, E.dist
    if E.world == 1:
        return None
    ids = [tc.Comm.unique_id(ctx) if E.rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = tc.Comm(ctx, ids[0], E.world, E.rank)
    if not peers:
        return comm
    handles = [None] * E.world          # NVLink peer buffers for the fused all-reduce
    dist.all_gather_object(handles, comm.peer_handle())
    comm.open_peers(handles)
    return comm


def _set_icp_fuse(E, on: bool):
    import ctypes as C
    E.ctx.lib.tc_debug_set_icp_fuse.argtypes = [C.c_int]
    E.ctx.lib.tc_debug_set_icp_fuse(1 if on else 0)


def bench_c3(E):
    """C3: 30 point-to-plane ICP iterations, 1M <-> 1M; target replicated, source sharded,
    29-scalar all-reduce per iteration (fused into the kernel over NVLink peer memory) at N > 1."""
    from fixtures import synth
    from threecrate_b200.sharding import shard_range
    tc, ctx, ext = E.tc, E.ctx, E.ext
    n = E.args.c3_points
    src, tgt, nrm, T = synth.scan_pair(n, half_extent=50.0 * (n / 1_000_000) ** 0.5)
    lo, hi = shard_range(E.rank, E.world, n)
    comm = _make_comm(E)
    tcloud = tc.DeviceCloud(tgt, ctx)
    scloud = tc.DeviceCloud(src[lo:hi], ctx)
    d_nrm = ctx.alloc(n * 12)
    ctx.to_device(d_nrm, nrm)
    iters = 30
    steps, warm = 10, 3
    res = []
    r = None
    for it in range(warm + steps):
        E.flush_l2()
        E.barrier()
        e0, e1, e2 = E.ev(), E.ev(), E.ev()
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
    ms_icp = E.max_over_ranks(float(np.mean([x[1] for x in res])))
    t_err = float(np.linalg.norm(r.translation.astype(np.float64) - T[:3]))
    ach = BYTES_ICP_PER_PT_ITER * (hi - lo) * iters / (ms_icp * 1e-3) / 1e9
    tr, inst, tr_src, _ = _profile_numbers("r02_c3", "k_icp_correspond")
    out = {
        "source_points": n, "target_points": n, "iterations": iters, "n_gpus": E.world,
        "timed_reps": steps, "iters_per_s": iters / (ms_icp * 1e-3),
        "iters_per_s_incl_index_build": iters / ((ms_icp + ms_index) * 1e-3),
        "ms_per_iter": ms_icp / iters, "ms_index_build": ms_index,
        "translation_error_vs_ground_truth": t_err,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": E.peak_gbs, "unit": "GB/s",
                     "frac": ach / E.peak_gbs, "kernel": "k_icp_correspond (solve fused)",
                     "traffic": tr, "traffic_source": tr_src,
                     "note": "36 MB/iter working set is L2-resident; latency-bound"}}
    if E.world > 1:
        # sharded vs single GPU (rank 0 holds the whole source once), fused vs NCCL all-reduce
        index = tc.GridIndex(tcloud, k_hint=1)
        _set_icp_fuse(E, False)
        r_nccl = tc.icp_point_to_plane_device(scloud, index, d_nrm, tc.IDENTITY, iters, None, -1.0, comm)
        _set_icp_fuse(E, True)
        r_fused = tc.icp_point_to_plane_device(scloud, index, d_nrm, tc.IDENTITY, iters, None, -1.0, comm)
        E.barrier()
        par = {"fused_equals_nccl_bitwise": bool(np.array_equal(r_nccl.transformation,
                                                                r_fused.transformation)),
               "iterations": [int(r_nccl.iterations), int(r_fused.iterations)]}
        if E.rank == 0:
            full = tc.DeviceCloud(src, ctx)
            r1 = tc.icp_point_to_plane_device(full, index, d_nrm, tc.IDENTITY, iters, None, -1.0, None)
            full.free()
            dT = np.abs(r1.transformation.astype(np.float64) - r_fused.transformation.astype(np.float64))
            par["sharded_vs_single_max_abs_dT"] = float(dT.max())
            par["sharded_vs_single_mse_rel"] = float(abs(r1.mse - r_fused.mse) / max(abs(r1.mse), 1e-30))
            if dT.max() > 1e-5:
                E.parity_failures.append(
                    f"C3 ICP: sharded transform differs from single GPU by {dT.max():.3g}")
        # every rank must hold the same transform bits
        t = E.torch.tensor(r_fused.transformation.view(np.int32).astype(np.int64), device="cuda")
        tmax, tmin = t.clone(), t.clone()
        E.dist.all_reduce(tmax, op=E.dist.ReduceOp.MAX)
        E.dist.all_reduce(tmin, op=E.dist.ReduceOp.MIN)
        par["transform_identical_on_all_ranks"] = bool(E.torch.equal(tmax, tmin))
        if not par["fused_equals_nccl_bitwise"]:
            E.parity_failures.append("C3 ICP: fused peer all-reduce differs from the NCCL path")
        if not par["transform_identical_on_all_ranks"]:
            E.parity_failures.append("C3 ICP: ranks hold different transforms")
        index.free()
        out["parity"] = par
    if E.rank == 0 and E.world == 1 and not E.args.no_cpu:
        try:
            import oracle
            threads = max(len(os.sched_getaffinity(0)), oracle.max_threads())
            cpu_it = 5
            t0 = time.perf_counter()
            oracle.icp_point_to_plane(src, tgt, nrm, max_iters=cpu_it, conv=-1.0)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {
                "value": cpu_it / dt, "unit": "iterations/s (incl. the serial kd-tree build)",
                "cores": threads, "kind": "port",
                "sample": f"{cpu_it} of the 30 iterations on the full 1M <-> 1M pair "
                          "(correspondences OpenMP-parallel, everything else serial as in the reference)"}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "sample": f"failed: {e}"}
    ctx.free(d_nrm)
    scloud.free()
    tcloud.free()
    if comm:
        comm.destroy()
    return {"c3_icp_point_to_plane": out}


def bench_c5(E):
    """C5 (BASELINE configs[4]): point-to-plane ICP against a 100M-point target, replicated on
    every rank (cloud + grid + normals), with 12.5M source points per rank (= the 100M <-> 100M pair
    at N = 8; at N < 8 the first N/8 of the source).  The clouds are generated on the device
    (torch Philox; the same seed on every rank gives the same target everywhere)."""
    tc, ctx, ext, torch = E.tc, E.ctx, E.ext, E.torch
    from fixtures import synth
    nt, ns = E.args.c5_target, E.args.c5_source_per_rank
    H = 316.0 * (nt / 100_000_000) ** 0.5
    T = synth.bench_transform(roll=0.01)
    Tinv = synth.invert_iso(T)
    dev = torch.device("cuda", E.local)

    def terrain_dev(n, seed, noise, want_normals, y_range=(-1.0, 1.0)):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        xy = torch.rand((n, 2), generator=g, device=dev, dtype=torch.float32)
        x = (xy[:, 0] * 2.0 - 1.0) * H
        y = (y_range[0] + xy[:, 1] * (y_range[1] - y_range[0])) * H
        z = 0.5 * torch.sin(0.3 * x) * torch.cos(0.2 * y)
        nrm = None
        if want_normals:
            dzdx = 0.15 * torch.cos(0.3 * x) * torch.cos(0.2 * y)
            dzdy = -0.1 * torch.sin(0.3 * x) * torch.sin(0.2 * y)
            nrm = torch.stack([-dzdx, -dzdy, torch.ones_like(x)], dim=1)
            nrm = (nrm / nrm.norm(dim=1, keepdim=True)).contiguous()
        z = z + torch.randn(n, generator=g, device=dev, dtype=torch.float32) * noise
        return torch.stack([x, y, z], dim=1).contiguous(), nrm

    tgt, nrm = terrain_dev(nt, 5, 0.005, True)
    # the source is sharded SPATIALLY (contiguous spatial ranges keep a rank's working
    # set compact): rank r holds the source points of the r-th of 8 slabs along y, so it touches
    # one eighth of the replicated target; N < 8 runs use the first N slabs
    slabs = max(8, E.world)
    srank = int(os.environ.get("TC_C5_SLAB", E.rank))  # (debug: which slab a 1-GPU run takes)
    yr = (-1.0 + 2.0 * srank / slabs, -1.0 + 2.0 * (srank + 1) / slabs)
    src, _ = terrain_dev(ns, 600 + srank, 0.005, False, yr)  # this rank's source shard
    R = torch.tensor(synth.quat_to_matrix(Tinv[3:7]), dtype=torch.float32, device=dev)
    src = (src @ R.T + torch.tensor(Tinv[:3], dtype=torch.float32, device=dev)).contiguous()
    torch.cuda.synchronize()
    tcloud = tc.DeviceCloud.from_device(tgt.data_ptr(), nt, ctx)
    scloud = tc.DeviceCloud.from_device(src.data_ptr(), ns, ctx)
    ctx.synchronize()
    del tgt, src
    comm = _make_comm(E)
    iters = 30
    steps, warm = 3, 1
    res = []
    r, info = None, None
    for it in range(warm + steps):
        E.barrier()
        e0, e1, e2 = E.ev(), E.ev(), E.ev()
        e0.record(ext)
        index = tc.GridIndex(tcloud, k_hint=1)
        e1.record(ext)
        r = tc.icp_point_to_plane_device(scloud, index, nrm.data_ptr(), tc.IDENTITY, iters, None,
                                         -1.0, comm)
        e2.record(ext)
        ctx.synchronize()
        if it >= warm:
            res.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        info = index.info()
        index.free()
    ms_index = float(np.mean([x[0] for x in res]))
    ms_icp = E.max_over_ranks(float(np.mean([x[1] for x in res])))
    # where the time goes: the same call stopped after 1, 2, 3, 5 and 10 iterations
    cumulative = {}
    index = tc.GridIndex(tcloud, k_hint=1)
    for it in (1, 2, 3, 5, 10):
        E.barrier()
        e1, e2 = E.ev(), E.ev()
        e1.record(ext)
        tc.icp_point_to_plane_device(scloud, index, nrm.data_ptr(), tc.IDENTITY, it, None, -1.0, comm)
        e2.record(ext)
        ctx.synchronize()
        cumulative[str(it)] = round(E.max_over_ranks(e1.elapsed_time(e2)), 3)
    cumulative[str(iters)] = round(ms_icp, 3)
    index.free()
    scloud.free()
    tcloud.free()
    if comm:
        comm.destroy()
    t_err = float(np.linalg.norm(r.translation.astype(np.float64) - T[:3]))
    ach = BYTES_ICP_PER_PT_ITER * ns * iters / (ms_icp * 1e-3) / 1e9
    return {"c5_icp_100m": {
        "target_points": nt, "source_points_per_rank": ns, "source_points_total": ns * E.world,
        "iterations": iters, "n_gpus": E.world, "timed_reps": steps,
        "iters_per_s": iters / (ms_icp * 1e-3), "ms_per_iter": ms_icp / iters,
        "source_points_per_s": ns * E.world * iters / (ms_icp * 1e-3),
        "ms_index_build_100m": ms_index, "cumulative_ms_after_iterations": cumulative,
        "scaling": "weak in the source (12.5M points per rank = one of 8 spatial slabs of the "
                   "100M-point source), target + grid replicated; N = 8 is BASELINE config 5 "
                   "(100M <-> 100M)",
        "translation_error_vs_ground_truth": t_err,
        "transform": [float(v) for v in r.transformation],
        "grid": {"cell_size": info["cell_size"], "dims": list(info["dims"])},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": E.peak_gbs, "unit": "GB/s",
                     "frac": ach / E.peak_gbs, "kernel": "k_icp_correspond (per rank)"}}}


def bench_next_rows(E):
    """Filters on the C2 frame (device-resident in and out) and the two other registrations on a
    200k terrain pair (host arrays through the C ABI, uploads included)."""
    from fixtures import synth
    tc, ctx, ext = E.tc, E.ctx, E.ext

    def timed(fn, reps=5, warm=2):
        ts = []
        for it in range(warm + reps):
            E.flush_l2()
            e0, e1 = E.ev(), E.ev()
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
