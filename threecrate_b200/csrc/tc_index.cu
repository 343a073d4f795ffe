// tc_index.cu — device-resident multi-resolution uniform-grid index.  Replaces KdTree::new
// (threecrate-algorithms/src/nearest_neighbor.rs:37-159), which is a serial O(N log N) quickselect.
//
// Build: bbox -> one measured histogram trial (+ statistics) -> histogram of every level in one
// launch -> cell-range scan of every level in one launch (decoupled look-back) -> counting-sort
// scatter of every level in one launch into float4 {x, y, z, bits(original index)} arrays.
// A hand-written stable LSD radix sort (tci_radix_sort_pairs) orders external query sets, ICP
// sources and voxel keys, where no cell table exists.
//
// The kernels are streaming passes (coalesced loads, grids capped at a multiple of the SM count);
// algorithmic bytes per point: bbox 12, histogram 12r + 4w per level, scatter
// 12r + 16w per level, cell-range scan 8 B/cell.  For a LiDAR-frame-sized cloud the build is
// bounded by host API latency, hence the fused launches, the polled host flags and the cached,
// self-cleaning workspaces below.
#include <atomic>
#include <chrono>
#include <cstdlib>

#include <vector>

#include "tc_internal.cuh"

namespace {

// TC_TRACE=1: print host-side phase timings of tc_index_build (adds stream syncs; debug only)
struct PhaseTrace {
  bool on, sync;  // TC_TRACE=1: phases with stream syncs; TC_TRACE=2: host time only, no syncs
  tc_context* ctx;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTrace(tc_context* c) : on(std::getenv("TC_TRACE") != nullptr), sync(true), ctx(c) {
    if (on) {
      sync = std::atoi(std::getenv("TC_TRACE")) != 2;
      if (sync) cudaStreamSynchronize(ctx->stream);
      t0 = std::chrono::steady_clock::now();
    }
  }
  void mark(const char* what) {
    if (!on) return;
    if (sync) cudaStreamSynchronize(ctx->stream);
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[tc_index_build] %-22s %8.1f us\n", what,
            std::chrono::duration<double, std::micro>(t1 - t0).count());
    t0 = t1;
  }
};

constexpr int kThreads = 256;

inline int grid_for(tc_context* ctx, uint64_t n, int per_block, int waves = 8) {
  uint64_t blocks = (n + per_block - 1) / per_block;
  uint64_t cap = (uint64_t)ctx->sm_count * waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------- bbox
// Scratch layout (64 words): [0..6] bbox accumulators (min x3, max x3, non-finite flag), [7] block
// ticket, [8..13] cell statistics accumulators, [14] block ticket, [16..22] published bbox,
// [24..29] published statistics.  The accumulators are armed once at context creation; the last
// block of k_bbox / k_cell_stats publishes the result and re-arms them, so neither needs an
// init launch or a memset in front (every API call on this path costs ~2-3 us of host time).
__global__ void k_scratch_init(uint32_t* s) {
  const int t = threadIdx.x;
  if (t < 3) s[t] = 0xFFFFFFFFu;       // min (ordered encoding)
  else if (t < 64) s[t] = 0u;          // max, flags, tickets, statistics
}

__global__ void __launch_bounds__(kThreads) k_bbox(const float* __restrict__ xyz, uint64_t n,
                                                   uint32_t* __restrict__ s,
                                                   volatile uint32_t* host, uint32_t seq) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  uint32_t bad = 0;
  auto upd = [&](int a, float v) {
    if (!isfinite(v)) bad = 1;
    mn[a] = fminf(mn[a], v);
    mx[a] = fmaxf(mx[a], v);
  };
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t done = 0;
  if ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) {
    // four points = 48 bytes = three aligned 128-bit loads, all in flight together
    const float4* v4 = reinterpret_cast<const float4*>(xyz);
    const uint64_t groups = n / 4;
    auto four = [&](const float4& a, const float4& b, const float4& c) {
      upd(0, a.x); upd(1, a.y); upd(2, a.z);
      upd(0, a.w); upd(1, b.x); upd(2, b.y);
      upd(0, b.z); upd(1, b.w); upd(2, c.x);
      upd(0, c.y); upd(1, c.z); upd(2, c.w);
    };
    uint64_t gi = tid;
    for (; gi + stride < groups; gi += 2 * stride) {  // six 128-bit loads in flight
      const uint64_t gj = gi + stride;
      const float4 a = v4[3 * gi], b = v4[3 * gi + 1], c = v4[3 * gi + 2];
      const float4 d = v4[3 * gj], e = v4[3 * gj + 1], f = v4[3 * gj + 2];
      four(a, b, c);
      four(d, e, f);
    }
    for (; gi < groups; gi += stride) four(v4[3 * gi], v4[3 * gi + 1], v4[3 * gi + 2]);
    done = 4 * groups;
  }
  for (uint64_t i = done + tid; i < n; i += stride) {
#pragma unroll
    for (int a = 0; a < 3; ++a) upd(a, xyz[3 * i + a]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  bad = __any_sync(0xffffffffu, bad);
  // block-level reduction first: six atomics per BLOCK on six addresses (per warp they were the
  // kernel's bottleneck: ~57k same-address atomics serialise in L2)
  __shared__ float smn[kThreads / 32][3], smx[kThreads / 32][3];
  __shared__ uint32_t sbad[kThreads / 32];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      smn[wid][a] = mn[a];
      smx[wid][a] = mx[a];
    }
    sbad[wid] = bad;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = fminf(mn[a], smn[w][a]);
        mx[a] = fmaxf(mx[a], smx[w][a]);
      }
      bad |= sbad[w];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (mn[a] <= mx[a]) {
        atomicMin(&s[a], f2ord(mn[a]));
        atomicMax(&s[3 + a], f2ord(mx[a]));
      }
    }
    if (bad) atomicOr(&s[6], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&s[7], 1u) == gridDim.x - 1) {  // last block: publish and re-arm
      __threadfence();
      // results go straight into the context's pinned host words (zero-copy), then the
      // sequence number the host is spinning on: no memcpy call, no stream synchronisation
      for (int j = 0; j < 7; ++j) host[j] = atomicExch(&s[j], j < 3 ? 0xFFFFFFFFu : 0u);
      s[7] = 0u;
      __threadfence_system();
      host[7] = seq;
    }
  }
}

// ------------------------------------------------------------------- cell keys + histogram
// Keys only (spatial sort of queries / ICP sources).
__global__ void __launch_bounds__(kThreads) k_cell_keys(const float* __restrict__ xyz, uint32_t n,
                                                        GridParams g, uint32_t* __restrict__ keys) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float u;
    const int cx = cell_coord(xyz[3 * (uint64_t)i + 0], g.ox, g.inv, g.nx, u);
    const int cy = cell_coord(xyz[3 * (uint64_t)i + 1], g.oy, g.inv, g.ny, u);
    const int cz = cell_coord(xyz[3 * (uint64_t)i + 2], g.oz, g.inv, g.nz, u);
    keys[i] = cell_id(g, cx, cy, cz);
  }
}

// Up to kMaxLevels resolutions are histogrammed / scattered by ONE launch each: the point is read
// once and its cell id recomputed per level (a dozen ALU ops; no key arrays, fewer launches).
struct LevelJob {
  GridParams g;
  int aggregate;               // warp-aggregate the atomics (tables with few cells per point)
  uint32_t* counts;            // n_cells + 1 (histogram, then scatter cursor)
  const uint32_t* cell_start;  // n_cells + 1 (scatter only)
  float4* out;                 // n (scatter only)
  uint32_t cell_lo, cell_hi;   // only cells in [cell_lo, cell_hi) take part (slab-sharded build)
  // histogram only (slab build): the planes p of the slowest axis with p % sample == 0 (sample a
  // power of two, 0: none) are ALSO counted, compactly, in sample_counts[(p / sample) * plane +
  // offset in plane] - the statistics sample every rank takes alike
  int sample;
  uint32_t plane;              // cells per plane
  uint32_t* sample_counts;
  // ... and EVERY point is counted by its plane (block-private shared-memory histogram, one
  // global atomic per plane and block): the ranks cut the cloud into slabs of equal point
  // counts from these, every rank alike.  nullptr: off (more planes than the table holds).
  uint32_t* plane_counts;
  int n_planes;
};
struct LevelJobs {
  int n;
  LevelJob l[kMaxLevels];
};

__device__ __forceinline__ uint32_t point_cell(const GridParams& g, float x, float y, float z,
                                               int* plane = nullptr) {
  float u;
  const int cx = cell_coord(x, g.ox, g.inv, g.nx, u);
  const int cy = cell_coord(y, g.oy, g.inv, g.ny, u);
  const int cz = cell_coord(z, g.oz, g.inv, g.nz, u);
  if (plane) *plane = g.ymajor ? cy : cz;  // coordinate along the slowest axis
  return cell_id(g, cx, cy, cz);
}

// f(i, x, y, z) for every point, grid-stride.  A 16-byte aligned array is read four points (three
// 128-bit loads) per thread and trip: a third of the load instructions, 48 bytes in flight.
// (clouds below kVecPoints keep one point per thread: a LiDAR frame would otherwise fill fewer
//  blocks than there are SMs; the launch sites size their grids with points_per_thread())
constexpr uint32_t kVecPoints = 1u << 20;
inline int points_per_thread(uint64_t n) { return n >= kVecPoints ? 4 : 1; }
template <class F>
__device__ __forceinline__ void for_each_point(const float* __restrict__ xyz, uint32_t n, F&& f) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  uint32_t done = 0;
  if (n >= kVecPoints && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0) {
    const float4* v4 = reinterpret_cast<const float4*>(xyz);
    const uint32_t groups = n / 4;
    for (uint32_t gi = tid; gi < groups; gi += stride) {
      const float4 a = __ldg(&v4[3 * (uint64_t)gi]), b = __ldg(&v4[3 * (uint64_t)gi + 1]),
                   c = __ldg(&v4[3 * (uint64_t)gi + 2]);
      f(4 * gi + 0, a.x, a.y, a.z);
      f(4 * gi + 1, a.w, b.x, b.y);
      f(4 * gi + 2, b.z, b.w, c.x);
      f(4 * gi + 3, c.y, c.z, c.w);
    }
    done = 4 * groups;
  }
  for (uint32_t i = done + tid; i < n; i += stride)
    f(i, xyz[3 * (uint64_t)i + 0], xyz[3 * (uint64_t)i + 1], xyz[3 * (uint64_t)i + 2]);
}

__global__ void __launch_bounds__(kThreads) k_hist_levels(const float* __restrict__ xyz, uint32_t n,
                                                          const __grid_constant__ LevelJobs jobs) {
  __shared__ uint32_t s_planes[tc_context::kPlaneWords];
  uint32_t* const plane_counts = jobs.l[0].plane_counts;  // (slab trial: one level)
  if (plane_counts) {
    for (int p = threadIdx.x; p < jobs.l[0].n_planes; p += kThreads) s_planes[p] = 0u;
    __syncthreads();
  }
  for_each_point(xyz, n, [&](uint32_t, float x, float y, float z) {
    const uint32_t active = __activemask();
    for (int l = 0; l < jobs.n; ++l) {
      const LevelJob& jb = jobs.l[l];
      int plane;
      const uint32_t c = point_cell(jb.g, x, y, z, &plane);
      const bool in = c >= jb.cell_lo && c < jb.cell_hi;
      if (jb.plane_counts) atomicAdd(&s_planes[plane], 1u);
      if (jb.sample && (plane & (jb.sample - 1)) == 0)
        atomicAdd(&jb.sample_counts[(uint64_t)(plane / jb.sample) * jb.plane +
                                    (c - (uint32_t)plane * jb.plane)], 1u);
      if (!jb.aggregate) {  // more cells than points: plain reductions, no matching
        if (in) atomicAdd(&jb.counts[c], 1u);
        continue;
      }
      // warp-aggregated: consecutive points of a scan usually share a cell, and coarse levels
      // funnel thousands of points into one counter
      const uint32_t peers = __match_any_sync(active, c);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1 && in)
        atomicAdd(&jb.counts[c], (uint32_t)__popc(peers));
    }
  });
  if (plane_counts) {
    __syncthreads();
    for (int p = threadIdx.x; p < jobs.l[0].n_planes; p += kThreads)
      if (s_planes[p]) atomicAdd(&plane_counts[p * tc_context::kPlaneStride], s_planes[p]);
  }
}

// occupied cells, max population, points living in cells with population <= low_thr x {1,2,4,8}
// (how much of the cloud is too sparse for this cell size), points counted -> host[8..14]
__global__ void __launch_bounds__(kThreads) k_cell_stats(const uint32_t* __restrict__ counts,
                                                         uint64_t n_cells, uint32_t low_thr,
                                                         uint32_t* __restrict__ s,
                                                         volatile uint32_t* host, uint32_t seq,
                                                         uint32_t* __restrict__ planes,
                                                         int n_planes /* 0: none to publish */) {
  uint32_t occ = 0, mx = 0, low[4] = {0, 0, 0, 0}, tot = 0;
  auto upd = [&](uint32_t c) {
    occ += (c != 0);
    mx = max(mx, c);
    tot += c;
#pragma unroll
    for (int j = 0; j < 4; ++j) low[j] += (c <= (low_thr << j)) ? c : 0u;  // thresholds x1,2,4,8
  };
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t done = 0;
  if ((reinterpret_cast<uintptr_t>(counts) & 15) == 0) {
    const uint4* c4 = reinterpret_cast<const uint4*>(counts);
    const uint64_t groups = n_cells / 4;
    uint64_t gi = tid;
    for (; gi + 3 * stride < groups; gi += 4 * stride) {  // four 128-bit loads in flight
      const uint4 a = c4[gi], b = c4[gi + stride], c = c4[gi + 2 * stride], d = c4[gi + 3 * stride];
      upd(a.x); upd(a.y); upd(a.z); upd(a.w);
      upd(b.x); upd(b.y); upd(b.z); upd(b.w);
      upd(c.x); upd(c.y); upd(c.z); upd(c.w);
      upd(d.x); upd(d.y); upd(d.z); upd(d.w);
    }
    for (; gi < groups; gi += stride) {
      const uint4 c = c4[gi];
      upd(c.x); upd(c.y); upd(c.z); upd(c.w);
    }
    done = 4 * groups;
  }
  for (uint64_t i = done + tid; i < n_cells; i += stride) upd(counts[i]);
  for (int o = 16; o > 0; o >>= 1) {
    occ += __shfl_xor_sync(0xffffffffu, occ, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) low[j] += __shfl_xor_sync(0xffffffffu, low[j], o);
  }
  __shared__ uint32_t sred[kThreads / 32][7];  // per-warp partials -> seven atomics per block
  if ((threadIdx.x & 31) == 0) {
    uint32_t* r = sred[threadIdx.x >> 5];
    r[0] = occ;
    r[1] = mx;
#pragma unroll
    for (int j = 0; j < 4; ++j) r[2 + j] = low[j];
    r[6] = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kThreads / 32; ++w) {
      occ += sred[w][0];
      mx = max(mx, sred[w][1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) low[j] += sred[w][2 + j];
      tot += sred[w][6];
    }
    if (occ) atomicAdd(&s[8], occ);
    if (mx) atomicMax(&s[9], mx);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (low[j]) atomicAdd(&s[10 + j], low[j]);
    if (tot) atomicAdd(&s[15], tot);
  }
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&s[14], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // last block: publish and re-arm (the per-plane counts of a slab build: all threads)
  __threadfence();
  for (int p = threadIdx.x; p < n_planes; p += kThreads)
    host[64 + p] = atomicExch(&planes[p * tc_context::kPlaneStride], 0u);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 0; j < 6; ++j) host[8 + j] = atomicExch(&s[8 + j], 0u);
    host[14] = atomicExch(&s[15], 0u);
    s[14] = 0u;
    __threadfence_system();
    host[15] = seq;
  }
}

// Counting-sort scatter: point i goes to cursor[cell(i)]++ as float4 (x, y, z, bits(i)), for
// every level of `jobs`.  The order INSIDE a cell is arrival order (not deterministic) — nothing
// downstream depends on it: every selection is keyed by (d2, original index) and multi-GPU
// shards own whole cells.
// (`zero` lists up to three word ranges nothing in this launch reads - the scan's tile states, an
//  abandoned trial histogram, the statistics sample of a slab build - which are cleared here so
//  the cached workspaces are left all zero)
struct ZeroJobs {
  uint32_t* p[5];
  uint64_t n[5];
};
__global__ void __launch_bounds__(kThreads) k_scatter_levels(const float* __restrict__ xyz,
                                                             uint32_t n,
                                                             const __grid_constant__ LevelJobs jobs,
                                                             const ZeroJobs zero) {
  for (int z = 0; z < 5; ++z)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero.n[z];
         i += (uint64_t)gridDim.x * blockDim.x)
      zero.p[z][i] = 0u;
  for_each_point(xyz, n, [&](uint32_t i, float x, float y, float z) {
    const float4 v = make_float4(x, y, z, __uint_as_float(i));
    const uint32_t active = __activemask();
    const int lane = threadIdx.x & 31;
    for (int l = 0; l < jobs.n; ++l) {
      const LevelJob& jb = jobs.l[l];
      const uint32_t c = point_cell(jb.g, x, y, z);
      // the histogram itself is the cursor: it counts down to zero while the cell fills up
      if (!jb.aggregate) {  // more cells than points: one atomic per point, no matching
        if (c >= jb.cell_lo && c < jb.cell_hi)
          jb.out[__ldg(&jb.cell_start[c]) + atomicSub(&jb.counts[c], 1u) - 1u] = v;
        continue;
      }
      // (warp-aggregated: one atomic per distinct cell per warp)
      const uint32_t peers = __match_any_sync(active, c);
      const int leader = __ffs(peers) - 1;
      const bool in = c >= jb.cell_lo && c < jb.cell_hi;  // (uniform across the peers)
      uint32_t old = 0;
      if (lane == leader && in) old = atomicSub(&jb.counts[c], (uint32_t)__popc(peers));
      old = __shfl_sync(active, old, leader);
      if (in) {
        const uint32_t pos =
            __ldg(&jb.cell_start[c]) + old - 1u - (uint32_t)__popc(peers & ((1u << lane) - 1u));
        jb.out[pos] = v;
      }
    }
  });
}

// ---------------------------------------------------------------------- exclusive scan (u32)
// Single-pass scan with decoupled look-back: one launch, every element read once and written
// once (8 B/cell).  Tiles are handed out through an atomic ticket so that a tile's predecessors
// are always already running; each tile publishes its aggregate, then walks back over earlier
// tiles' (status, value) words until it meets an inclusive prefix.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp,
                                                         uint32_t& total) {
  // returns exclusive prefix of v across the block (kScanThreads threads), total = block sum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (kScanThreads / 32) ? smem_warp[lane] : 0;
    uint32_t wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < (kScanThreads / 32)) smem_warp[lane] = wi - w;  // exclusive warp offsets
    if (lane == (kScanThreads / 32) - 1) smem_warp[kScanThreads / 32] = wi;
  }
  __syncthreads();
  total = smem_warp[kScanThreads / 32];
  const uint32_t res = smem_warp[warp] + incl - v;
  __syncthreads();
  return res;
}

constexpr unsigned long long kTileAggregate = 1ull << 32, kTileInclusive = 2ull << 32;

// Up to kMaxLevels independent tables scanned by one launch: tickets [tile_begin[j],
// tile_begin[j+1]) belong to table j, whose look-back stops at its own first tile.
struct ScanJobs {
  int n;
  const uint32_t* in[kMaxLevels];
  uint32_t* out[kMaxLevels];  // n + 1 entries: out[n] = grand total.  in == out is allowed.
  uint64_t len[kMaxLevels];
  uint32_t tile_begin[kMaxLevels + 1];
};

// 16 consecutive items of one thread (vector loads when aligned and fully inside the table)
__device__ __forceinline__ uint32_t scan_load16(const uint32_t* in, uint64_t base, uint64_t n,
                                                uint32_t (&v)[kScanItems]) {
  if (base + kScanItems <= n && (((uintptr_t)(in + base)) & 15) == 0) {
#pragma unroll
    for (int it = 0; it < kScanItems; it += 4) {
      const uint4 q = *reinterpret_cast<const uint4*>(in + base + it);
      v[it] = q.x;
      v[it + 1] = q.y;
      v[it + 2] = q.z;
      v[it + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int it = 0; it < kScanItems; ++it) v[it] = (base + it < n) ? in[base + it] : 0u;
  }
  uint32_t s = 0;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) s += v[it];
  return s;
}

// state[0] = ticket counter (as u64), state[1 + t] = (status << 32 | value) of ticket t; zeroed.
// A tile is `sub` consecutive chunks of kScanTile items: inclusive prefixes travel at most 32
// tiles per look-back step (~1 us each), so a multi-million-cell table is cut into a few hundred
// large tiles rather than a thousand small ones.  Phase A sums the tile and publishes, the
// look-back resolves the prefix, phase B re-reads the tile (L2-hot) and writes the scan.
__global__ void __launch_bounds__(kScanThreads)
k_scan_lookback(const __grid_constant__ ScanJobs jobs, unsigned long long* state, int sub) {
  __shared__ uint32_t sw[kScanThreads / 32 + 1];
  __shared__ uint32_t s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd(&state[0], 1ull);
  __syncthreads();
  const uint32_t ticket = s_tile;
  int j = 0;
  while (j + 1 < jobs.n && ticket >= jobs.tile_begin[j + 1]) ++j;
  const uint32_t first = jobs.tile_begin[j];  // first ticket of this table
  const uint32_t* in = jobs.in[j];
  uint32_t* out = jobs.out[j];
  const uint64_t n = jobs.len[j];
  const uint32_t tile = ticket - first;
  const uint64_t tile_base = (uint64_t)tile * kScanTile * sub;
  uint32_t v[kScanItems];
  // ---- phase A: tile total
  uint32_t s = 0;
  for (int u = 0; u < sub; ++u)
    s += scan_load16(in, tile_base + (uint64_t)u * kScanTile + (uint64_t)threadIdx.x * kScanItems, n, v);
  uint32_t total;
  block_exclusive_scan(s, sw, total);
  if (threadIdx.x < 32) {
    // warp-wide look-back: lane j inspects predecessor (p - j); the nearest tile that already
    // published an inclusive prefix ends the walk, the aggregates in front of it are summed
    volatile unsigned long long* st = state + 1;
    const int lane = threadIdx.x;
    uint32_t prefix = 0;
    if (tile == 0) {
      if (lane == 0) st[ticket] = kTileInclusive | total;
    } else {
      if (lane == 0) {
        st[ticket] = kTileAggregate | total;
        __threadfence();
      }
      for (int64_t p = (int64_t)ticket - 1; p >= (int64_t)first; p -= 32) {
        const int64_t idx = p - lane;
        unsigned long long w = kTileInclusive;  // before the first tile: an inclusive prefix of 0
        if (idx >= (int64_t)first) {
          do {
            w = st[idx];
          } while ((w >> 32) == 0);  // predecessor is running (ticket order): short spin
        }
        const unsigned incl = __ballot_sync(0xffffffffu, (w >> 32) == 2);
        const int stop = incl ? (__ffs(incl) - 1) : 31;
        uint32_t val = lane <= stop ? (uint32_t)w : 0u;
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        prefix += val;
        if (incl) break;
      }
      if (lane == 0) {
        __threadfence();
        st[ticket] = kTileInclusive | (unsigned long long)(prefix + total);
      }
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  // ---- phase B: scan chunk by chunk with a running prefix
  uint32_t running = s_prefix;
  for (int u = 0; u < sub; ++u) {
    const uint64_t base = tile_base + (uint64_t)u * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    const uint32_t cs = scan_load16(in, base, n, v);
    uint32_t chunk_total;
    uint32_t ex = block_exclusive_scan(cs, sw, chunk_total) + running;
#pragma unroll
    for (int it = 0; it < kScanItems; ++it) {
      const uint64_t i = base + it;
      if (i < n) out[i] = ex;
      ex += v[it];
      if (i == n - 1) out[n] = ex;
    }
    running += chunk_total;
  }
}

// ------------------------------------------------------------------------- LSD radix sort
// 8-bit digits; per pass: tile histogram -> exclusive scan of hist[digit][tile] -> stable scatter.
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per block
constexpr int kRsWarps = kRsThreads / 32;

__global__ void __launch_bounds__(kRsThreads) k_radix_hist(const uint32_t* __restrict__ keys,
                                                           uint32_t n, int shift,
                                                           uint32_t* __restrict__ hist,
                                                           uint32_t n_tiles) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll
  for (int it = 0; it < kRsItems; ++it) {
    const uint64_t i = base + (uint64_t)it * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// vals_in == nullptr means the identity permutation (first pass).
__global__ void __launch_bounds__(kRsThreads)
k_radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                int shift, const uint32_t* __restrict__ offs, uint32_t n_tiles) {
  __shared__ uint32_t whist[kRsWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) ((uint32_t*)whist)[i] = 0;
  __syncthreads();
  // tile element order: warp-major, then round, then lane  (== global order)
  const uint64_t base = (uint64_t)blockIdx.x * kRsTile + (uint64_t)warp * (32 * kRsItems);
  uint32_t key[kRsItems], rank[kRsItems];
#pragma unroll
  for (int it = 0; it < kRsItems; ++it) {
    const uint64_t i = base + it * 32 + lane;
    const bool valid = i < n;
    key[it] = valid ? keys_in[i] : 0xFFFFFFFFu;
    const uint32_t digit = (key[it] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (0x100u | lane));
    const int leader = __ffs(peers) - 1;
    uint32_t pre = 0;
    if (lane == leader && valid) {
      pre = whist[warp][digit];
      whist[warp][digit] = pre + __popc(peers);
    }
    pre = __shfl_sync(0xffffffffu, pre, leader);
    rank[it] = pre + __popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;
    uint32_t run = offs[(uint64_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      const uint32_t c = whist[w][d];
      whist[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kRsItems; ++it) {
    const uint64_t i = base + it * 32 + lane;
    if (i < n) {
      const uint32_t digit = (key[it] >> shift) & 255u;
      const uint32_t pos = whist[warp][digit] + rank[it];
      keys_out[pos] = key[it];
      vals_out[pos] = vals_in ? vals_in[i] : (uint32_t)i;
    }
  }
}

// -------------------------------------------------------------------------------- gather
__global__ void __launch_bounds__(kThreads) k_gather(const float* __restrict__ xyz,
                                                     const uint32_t* __restrict__ perm, uint32_t n,
                                                     float4* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t j = perm[i];
    const float* p = xyz + 3 * (uint64_t)j;
    out[i] = make_float4(p[0], p[1], p[2], __uint_as_float(j));
  }
}

int key_bits_for(uint64_t n_cells) {
  int b = 1;
  while (b < 32 && (1ull << b) < n_cells) ++b;
  return b;
}

}  // namespace

// ==========================================================================================
// `d_state` (tiles + 1 u64 words, zeroed) may be supplied by the caller; otherwise it is
// allocated and cleared here.
// chunks per tile such that the largest table is cut into at most ~1024 tiles (every SM holds
// several tiles in both phases; the look-back walks 32 tiles per step)
static int scan_sub(const ScanJobs& jobs) {
  uint64_t longest = 0;
  for (int j = 0; j < jobs.n; ++j) longest = std::max(longest, jobs.len[j]);
  const uint64_t chunks = (longest + kScanTile - 1) / kScanTile;
  return (int)std::min<uint64_t>(64, std::max<uint64_t>(1, (chunks + 1023) / 1024));
}
static uint32_t scan_tiles(ScanJobs& jobs, int sub) {
  uint32_t tiles = 0;
  const uint64_t per_tile = (uint64_t)kScanTile * sub;
  for (int j = 0; j < jobs.n; ++j) {
    jobs.tile_begin[j] = tiles;
    tiles += (uint32_t)((jobs.len[j] + per_tile - 1) / per_tile);
  }
  jobs.tile_begin[jobs.n] = tiles;
  return tiles;
}

static int scan_tables(tc_context* ctx, ScanJobs& jobs, unsigned long long* d_state) {
  const int sub = scan_sub(jobs);
  const uint32_t tiles = scan_tiles(jobs, sub);
  if (tiles == 0) return TC_OK;
  unsigned long long* own = nullptr;
  if (!d_state) {
    TC_TRY(tc_alloc(ctx, &own, (uint64_t)tiles + 1));
    TC_CUDA(ctx, cudaMemsetAsync(own, 0, ((uint64_t)tiles + 1) * sizeof(unsigned long long),
                                 ctx->stream));
    d_state = own;
  }
  k_scan_lookback<<<tiles, kScanThreads, 0, ctx->stream>>>(jobs, d_state, sub);
  TC_LAUNCHED(ctx);
  tc_free(ctx, own);
  return TC_OK;
}

int tci_exclusive_scan_u32(tc_context* ctx, const uint32_t* d_in, uint32_t* d_out, uint64_t n) {
  if (n == 0) {
    TC_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(uint32_t), ctx->stream));
    return TC_OK;
  }
  ScanJobs jobs{};
  jobs.n = 1;
  jobs.in[0] = d_in;
  jobs.out[0] = d_out;
  jobs.len[0] = n;
  return scan_tables(ctx, jobs, nullptr);
}

int tci_radix_sort_pairs(tc_context* ctx, uint32_t* d_keys, uint32_t* d_vals, uint32_t* d_keys_alt,
                         uint32_t* d_vals_alt, uint32_t n, int key_bits, bool vals_given,
                         uint32_t** keys_out, uint32_t** vals_out) {
  // Stable.  Values start as the identity permutation unless `vals_given` (then d_vals holds
  // them); the caller always provides both value buffers, the result lands in one of each pair.
  const uint32_t n_tiles = (n + kRsTile - 1) / kRsTile;
  uint32_t* d_hist = nullptr;
  TC_TRY(tc_alloc(ctx, &d_hist, (uint64_t)256 * n_tiles + 1));
  uint32_t *kin = d_keys, *kout = d_keys_alt, *vin = vals_given ? d_vals : nullptr,
           *vout = d_vals_alt;
  uint32_t* vother = d_vals;
  const int passes = (key_bits + 7) / 8;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    k_radix_hist<<<n_tiles, kRsThreads, 0, ctx->stream>>>(kin, n, shift, d_hist, n_tiles);
    TC_LAUNCHED(ctx);
    TC_TRY(tci_exclusive_scan_u32(ctx, d_hist, d_hist, (uint64_t)256 * n_tiles));
    k_radix_scatter<<<n_tiles, kRsThreads, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, d_hist,
                                                             n_tiles);
    TC_LAUNCHED(ctx);
    // ping-pong
    uint32_t* t = kin;
    kin = kout;
    kout = t;
    vin = vout;
    vout = vother;
    vother = vin;
  }
  tc_free(ctx, d_hist);
  *keys_out = kin;
  *vals_out = vin;
  return TC_OK;
}

int tci_scratch_arm(tc_context* ctx) {
  k_scratch_init<<<1, 64, 0, ctx->stream>>>(ctx->d_scratch);
  TC_LAUNCHED(ctx);
  TC_CUDA(ctx, cudaMemsetAsync(ctx->d_planes, 0,
                               (size_t)tc_context::kPlaneWords * tc_context::kPlaneStride *
                                   sizeof(uint32_t),
                               ctx->stream));
  return TC_OK;
}

// Wait until a kernel's last block has stored `seq` into the pinned host word `flag` (its results
// were stored before it, fenced system-wide).  Polling pinned memory costs ~2 us after the kernel
// ends; a memcpy + cudaStreamSynchronize pair costs 10-15 us, and the index build is host-bound.
static int wait_host_flag(tc_context* ctx, volatile uint32_t* flag, uint32_t seq) {
  for (uint64_t spins = 0;; ++spins) {
    if (*flag == seq) {
      std::atomic_thread_fence(std::memory_order_acquire);  // results are read after the flag
      return TC_OK;
    }
    if ((spins & 0xFFF) == 0xFFF) {  // every few thousand polls: did the stream die or drain?
      const cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e == cudaSuccess) {  // nothing left in flight: the flag must be there now
        if (*flag == seq) {
          std::atomic_thread_fence(std::memory_order_acquire);
          return TC_OK;
        }
        return tc_fail(ctx, TC_GPU, "device result flag never arrived");
      }
      if (e != cudaErrorNotReady)
        return tc_fail(ctx, TC_GPU, std::string("kernel failed: ") + cudaGetErrorString(e));
    }
  }
}

int tci_bbox(tc_context* ctx, const float* d_xyz, uint64_t n, float mn[3], float mx[3]) {
  TcRange nvtx_range("tc:bbox");
  const uint32_t seq = ++ctx->seq;
  k_bbox<<<grid_for(ctx, n, kThreads * 4), kThreads, 0, ctx->stream>>>(d_xyz, n, ctx->d_scratch,
                                                                      ctx->h_scratch, seq);
  TC_LAUNCHED(ctx);
  TC_TRY(wait_host_flag(ctx, ctx->h_scratch + 7, seq));
  if (ctx->h_scratch[6] != 0)
    return tc_fail(ctx, TC_INVALID_DATA, "point coordinates must be finite (NaN/Inf found)");
  for (int a = 0; a < 3; ++a) {
    mn[a] = ord2f(ctx->h_scratch[a]);
    mx[a] = ord2f(ctx->h_scratch[3 + a]);
  }
  return TC_OK;
}

namespace {

constexpr uint64_t kMaxCells = 1ull << 27;

// dims for a given cell size, shrinking resolution if the table would be too large
bool g_make_grid_ymajor = false;  // set by tci_index_build for the build in progress
GridParams make_grid(const float mn[3], const float mx[3], float cell, uint64_t n,
                     uint64_t max_cells) {
  GridParams g{};
  g.ox = mn[0];
  g.oy = mn[1];
  g.oz = mn[2];
  g.ex = mx[0] - mn[0];
  g.ey = mx[1] - mn[1];
  g.ez = mx[2] - mn[2];
  g.n = (uint32_t)n;
  for (int attempt = 0; attempt < 64; ++attempt) {
    const double dx = std::floor((double)g.ex / cell) + 1, dy = std::floor((double)g.ey / cell) + 1,
                 dz = std::floor((double)g.ez / cell) + 1;
    if (dx * dy * dz <= (double)max_cells && dx < 2e9 && dy < 2e9 && dz < 2e9) {
      g.nx = (int)dx;
      g.ny = (int)dy;
      g.nz = (int)dz;
      break;
    }
    cell *= 1.26f;  // ~2x fewer cells per step
  }
  if (g.nx < 1) g.nx = g.ny = g.nz = 1;
  // Slab-sharded builds (world > 1) make the axis with more cells the slowest one, so the shard
  // unit is a run of whole planes along a long axis.  Single-GPU builds keep z slowest: the order
  // makes no difference on the 10M terrain cloud, but the LiDAR frame's slowest warp is 12 % faster.
  g.ymajor = (g_make_grid_ymajor && g.ny >= g.nz) ? 1 : 0;
  g.cell = cell;
  g.inv = 1.0f / cell;
  return g;
}

float target_population(uint32_t k_hint) {
  // points per occupied cell aimed for: ring-1 (3x3x3) search is exact when the k-th neighbour
  // lies within one cell edge; ~0.55 (k+1) per cell keeps that true for surface-like data.
  float t = 0.55f * (float)(k_hint + 1);
  if (t < 4.0f) t = 4.0f;  // 1-NN (ICP): ~4 per cell measured best (fewer rows, still few candidates)
  if (t > 48.0f) t = 48.0f;
  return t;
}

}  // namespace

namespace {

inline uint64_t round_up4(uint64_t v) { return (v + 3) & ~(uint64_t)3; }
inline uint64_t cells_of(const GridParams& g) { return (uint64_t)g.nx * g.ny * g.nz; }

// Slab of a grid owned by `rank`: whole planes along the slowest axis, built with `halo` planes
// on either side.
struct Slab {
  bool on = false;
  uint64_t plane = 0;  // cells per plane
  int n_planes = 0;
  uint64_t cell_lo = 0, cell_hi = 0, own_cell_lo = 0, own_cell_hi = 0;
};
constexpr int kSamplePlanes = 8;  // the statistics sample of a slab build: every 8th plane
// own planes [p_lo, p_hi) of the slowest axis, built with `halo` planes on either side
Slab slab_planes(const GridParams& g, int world, int p_lo, int p_hi, int halo) {
  Slab s;
  s.plane = g.ymajor ? (uint64_t)g.nz * g.nx : (uint64_t)g.ny * g.nx;
  s.n_planes = g.ymajor ? g.ny : g.nz;
  s.cell_hi = s.own_cell_hi = cells_of(g);
  s.on = world > 1 && s.n_planes >= 4 * world;
  if (s.on) {
    s.own_cell_lo = (uint64_t)p_lo * s.plane;
    s.own_cell_hi = (uint64_t)p_hi * s.plane;
    s.cell_lo = (uint64_t)std::max(p_lo - halo, 0) * s.plane;
    s.cell_hi = (uint64_t)std::min(p_hi + halo, s.n_planes) * s.plane;
  }
  return s;
}
inline int equal_boundary(int n_planes, int world, int r) { return (int)((int64_t)r * n_planes / world); }
// slabs of equal plane counts
Slab slab_of(const GridParams& g, int rank, int world, int halo) {
  const int n_planes = g.ymajor ? g.ny : g.nz;
  return slab_planes(g, world, equal_boundary(n_planes, world, rank),
                     equal_boundary(n_planes, world, rank + 1), halo);
}
// how far a slab boundary may move away from the equal-planes position to balance the ranks'
// point counts (the trial histogram counts that many planes more on either side)
inline int balance_margin(int n_planes, int world) { return n_planes / world / 8; }
// Boundaries that give every rank the same number of POINTS (per-plane counts of the whole cloud,
// identical on every rank), each kept within `margin` planes of its equal-planes position.
void balanced_boundaries(const uint32_t* plane_counts, int n_planes, uint64_t n, int world,
                         int margin, int* bnd /* world + 1 */) {
  bnd[0] = 0;
  bnd[world] = n_planes;
  uint64_t cum = 0;
  int r = 1;
  for (int p = 0; p < n_planes && r < world; ++p) {
    cum += plane_counts[p];
    while (r < world && cum * (uint64_t)world >= (uint64_t)r * n) bnd[r++] = p + 1;
  }
  for (; r < world; ++r) bnd[r] = n_planes;
  for (r = 1; r < world; ++r) {
    const int e = equal_boundary(n_planes, world, r);
    bnd[r] = std::max(e - margin, std::min(e + margin, bnd[r]));
  }
}

// the trial workspace: the histogram (n_cells + 1 counters), then the compact sample table
inline uint64_t sample_off_of(const GridParams& g) { return round_up4(cells_of(g) + 1); }
inline uint64_t sample_words_of(const Slab& s) {
  return s.on ? (uint64_t)((s.n_planes + kSamplePlanes - 1) / kSamplePlanes) * s.plane : 0;
}
inline uint64_t trial_words_of(const GridParams& g, const Slab& s) {
  return sample_off_of(g) + sample_words_of(s);
}

// Trial histogram of `cloud` on grid g plus its statistics (one host sync):
// stats = occupied / max_pop / points in cells with population <= low_thr x {1,2,4,8} / points
// counted.  A slab build counts only the rank's slab, plus - in a compact table behind the
// histogram - the sample planes (every rank the same ones, so every rank takes the same
// decisions), and takes the statistics over the sample; the scatter launch clears the sample.
int trial_histogram(tc_context* ctx, const tc_cloud* cloud, const GridParams& g, uint32_t low_thr,
                    const Slab& slab, uint32_t** d_counts_out, uint32_t stats[7]) {
  const uint64_t n = cloud->n, n_cells = cells_of(g);
  uint32_t* d_counts = nullptr;
  TC_TRY(tc_ws_get_zeroed(ctx, 0, &d_counts, trial_words_of(g, slab)));
  *d_counts_out = d_counts;
  LevelJobs jobs{};
  jobs.n = 1;
  jobs.l[0].g = g;
  jobs.l[0].counts = d_counts;
  jobs.l[0].cell_lo = slab.on ? (uint32_t)slab.cell_lo : 0u;
  jobs.l[0].cell_hi = slab.on ? (uint32_t)slab.cell_hi : 0xFFFFFFFFu;
  jobs.l[0].sample = slab.on ? kSamplePlanes : 0;
  jobs.l[0].plane = (uint32_t)slab.plane;
  jobs.l[0].sample_counts = d_counts + sample_off_of(g);
  const bool planes = slab.on && slab.n_planes <= tc_context::kPlaneWords;
  jobs.l[0].plane_counts = planes ? ctx->d_planes : nullptr;
  jobs.l[0].n_planes = slab.n_planes;
  jobs.l[0].aggregate = n_cells < n ? 1 : 0;
  k_hist_levels<<<grid_for(ctx, n, kThreads * points_per_thread(n)), kThreads, 0, ctx->stream>>>(
      cloud->d_xyz, (uint32_t)n, jobs);
  TC_LAUNCHED(ctx);
  const uint32_t seq = ++ctx->seq;
  const uint32_t* d_stat = slab.on ? d_counts + sample_off_of(g) : d_counts;
  const uint64_t n_stat = slab.on ? sample_words_of(slab) : n_cells;
  k_cell_stats<<<grid_for(ctx, n_stat, kThreads * 4), kThreads, 0, ctx->stream>>>(
      d_stat, n_stat, low_thr, ctx->d_scratch, ctx->h_scratch, seq, ctx->d_planes,
      planes ? slab.n_planes : 0);
  TC_LAUNCHED(ctx);
  TC_TRY(wait_host_flag(ctx, ctx->h_scratch + 15, seq));
  for (int j = 0; j < 7; ++j) stats[j] = ctx->h_scratch[8 + j];
  return TC_OK;
}

}  // namespace

int g_tc_max_levels = kMaxLevels;
uint64_t g_tc_fine_cap = 64;  // fine-level table cap, cells per point
extern "C" void tc_debug_set_fine_cap(int cells_per_point) { g_tc_fine_cap = (uint64_t)cells_per_point; }
extern "C" void tc_debug_set_max_levels(int n) {
  g_tc_max_levels = n < 1 ? 1 : (n > kMaxLevels ? kMaxLevels : n);
}

int g_tc_shard_halo = 4;  // planes built beyond the rank's own slab on either side
extern "C" void tc_debug_set_shard_halo(int planes) { g_tc_shard_halo = planes < 1 ? 1 : planes; }
// (host logic of the slab build, exported for the CPU tests: the boundaries every rank derives
//  from the per-plane point counts; bnd has world + 1 entries)
extern "C" void tc_debug_balanced_boundaries(const uint32_t* plane_counts, int n_planes, int world,
                                             int* bnd) {
  uint64_t n = 0;
  for (int p = 0; p < n_planes; ++p) n += plane_counts[p];
  balanced_boundaries(plane_counts, n_planes, n, world, balance_margin(n_planes, world), bnd);
}

extern "C" int tc_index_build(tc_context* ctx, const tc_cloud* cloud, uint32_t k_hint,
                              float cell_size, tc_index** out) {
  return tci_index_build(ctx, cloud, k_hint, cell_size, 0, 1, out);
}
extern "C" int tc_index_build_sharded(tc_context* ctx, const tc_cloud* cloud, uint32_t k_hint,
                                      float cell_size, int rank, int world, tc_index** out) {
  if (world < 1 || rank < 0 || rank >= world) return tc_fail(ctx, TC_INVALID_DATA, "rank out of range");
  return tci_index_build(ctx, cloud, k_hint, cell_size, rank, world, out);
}

int tci_index_build(tc_context* ctx, const tc_cloud* cloud, uint32_t k_hint, float cell_size,
                    int rank, int world, tc_index** out, const GridParams* like) {
  if (like) cell_size = like->cell;  // same cell edge and cell order as an existing index
  TcRange nvtx_range("tc_index_build");
  if (!ctx || !cloud || !out) return TC_INVALID_DATA;
  *out = nullptr;
  const uint64_t n = cloud->n;
  if (n >= 0xFFFFFFFFull) return tc_fail(ctx, TC_INVALID_DATA, "cloud too large (N must be < 2^32-1)");
  tc_index* ix = new tc_index();
  ix->ctx = ctx;
  ix->cloud = cloud;
  ix->n = n;
  ix->n_local = n;
  ix->k_hint = k_hint;
  ix->cell_size_arg = cell_size;
  ix->shard_rank = rank;
  ix->shard_world = world;
  ix->own_lo = rank * n / (uint64_t)world;  // complete index: the rank's share of the sorted order
  ix->own_hi = (rank + 1) * n / (uint64_t)world;
  if (n == 0) {
    ix->n_levels = 1;
    ix->lv[0].g.nx = ix->lv[0].g.ny = ix->lv[0].g.nz = 1;
    ix->lv[0].g.cell = 1.0f;
    ix->lv[0].g.inv = 1.0f;
    ix->lv[0].n_cells = 1;
    *out = ix;
    return TC_OK;
  }
  PhaseTrace trace(ctx);
  g_make_grid_ymajor = like ? like->ymajor != 0 : world > 1;
  const int halo = std::min(g_tc_shard_halo, 255);
  int st = tci_bbox(ctx, cloud->d_xyz, n, ix->bbox_min, ix->bbox_max);
  trace.mark("bbox");
  if (st != TC_OK) {
    delete ix;
    return st;
  }
  const float* mn = ix->bbox_min;
  const float* mx = ix->bbox_max;
  const double ex = (double)mx[0] - mn[0], ey = (double)mx[1] - mn[1], ez = (double)mx[2] - mn[2];
  const double emax = std::max(ex, std::max(ey, ez));

  uint32_t* d_trial = nullptr;  // histogram of the measured trial
  const uint64_t table_cap = std::min<uint64_t>(kMaxCells, std::max<uint64_t>(64 * n, 1u << 20));
  const float target = target_population(k_hint);
  // a 3x3x3 block of surface-like data spans ~9 occupied cells: it cannot even hold k+1 points
  // when the cells hold fewer than (k+1)/9 each
  const uint32_t low_thr = (k_hint + 1) / 9;

  float cell = cell_size;
  const bool auto_cell = !(cell_size > 0.0f);
  if (auto_cell) {
    // first guess: surface-like data, area proxy = sum of the three bbox face areas
    double area = ex * ey + ey * ez + ex * ez;
    if (area <= 0) area = emax * emax;
    if (area <= 0) area = 1.0;
    cell = (float)std::sqrt(target * area / (double)n);
    if (!(cell > 0) || !std::isfinite(cell)) cell = 1.0f;
    if (emax > 0 && cell > emax) cell = (float)emax;
    if (emax > 0 && cell < emax * 1e-6) cell = (float)(emax * 1e-6);
  }
  uint32_t stats[7] = {0, 0, 0, 0, 0, 0, 0};
  // One measured trial (histogram + stats + one host sync).  If the mean population of occupied
  // cells is off target the cell is rescaled ONCE (surface-like scaling: occupied ~ cell^-2) and
  // re-histogrammed together with the extra resolutions, without waiting for new statistics;
  // the level decisions below use the measured trial's skew, which is scale-free.
  float stat_scale = 1.0f;  // final cell / measured cell
  GridParams g = make_grid(mn, mx, cell, n, table_cap);
  cell = g.cell;
  // (a slab build samples its statistics, and counts `margin` planes more on either side so the
  //  boundaries can still move to balance the ranks' point counts)
  const int margin = balance_margin(g.ymajor ? g.ny : g.nz, world);
  const Slab tslab = slab_of(g, rank, world, halo + margin);
  st = trial_histogram(ctx, cloud, g, low_thr, tslab, &d_trial, stats);
  trace.mark("trial histogram+stats");
  std::vector<uint32_t> plane_counts;
  if (st == TC_OK && tslab.on && tslab.n_planes <= tc_context::kPlaneWords)
    plane_counts.assign(ctx->h_scratch + 64, ctx->h_scratch + 64 + tslab.n_planes);
  if (st != TC_OK) {
    tc_ws_release_zeroed(ctx, 0, d_trial, false);
    delete ix;
    return st;
  }
  const uint64_t trial_words = cells_of(g) + 1, trial_sample_off = sample_off_of(g);
  bool primary_counted = true;  // d_trial holds the primary histogram
  // points the statistics were taken over: all of them, or the sample planes of a slab build
  const double samp_n = std::max<double>(1.0, (double)stats[6]);
  const double samp_ratio = (double)n / samp_n;
  const float pop1 = stats[0] ? (float)(samp_n / (double)stats[0]) : target;
  if (auto_cell && !(pop1 > target * 0.7f && pop1 < target * 1.4f)) {
    float scale = std::sqrt(target / pop1);
    scale = std::min(4.0f, std::max(0.25f, scale));
    float next = cell * scale;
    if (emax > 0 && next > emax) next = (float)emax;
    const GridParams g2 = make_grid(mn, mx, next, n, table_cap);
    if (std::fabs(g2.cell - cell) > 1e-3f * cell) {
      primary_counted = false;
      stat_scale = g2.cell / cell;
      g = g2;
      cell = g2.cell;
    }
  }
  // statistics of the final grid, extrapolated from the measured one when it was rescaled
  const float s2 = stat_scale * stat_scale;
  const uint32_t occ_est = (uint32_t)std::max(1.0, (double)stats[0] * samp_ratio / s2);
  const uint32_t maxpop_est = (uint32_t)std::min<double>(
      (double)n, std::ceil((double)stats[1] * std::max(1.0f, s2 * stat_scale)));
  // skew of the measured trial: densest cell vs mean, and share of points in near-empty cells
  const float skew = (float)stats[1] / std::max(1.0f, pop1);
  const float target_ratio = 6.0f;
  // "near-empty" is judged at the measured cell size: a cell pop1/target times more populated
  // than intended needs a proportionally higher threshold (counters exist for x1, x2, x4, x8)
  int lj = 0;
  for (float r = pop1 / target; r >= 1.5f && lj < 3; r *= 0.5f) ++lj;
  const uint32_t low_pts = stats[2 + lj];
  // Density skew -> extra resolutions: a 4x finer grid when some cells are far
  // over target (dense LiDAR near field), a 4x coarser one when a visible share of the points
  // sits in nearly empty cells (far field: ring growth would otherwise walk thousands of rows).
  bool want_fine = auto_cell && g_tc_max_levels > 1 && skew > target_ratio && n > 4096;
  // ... and always for a 1-NN index (k_hint <= 1: an ICP target): the first iterations of a badly
  // aligned pair query it from several cells away, where the fine grid is thousands of empty
  // rows and the coarse one a handful (tc_icp.cu picks the level by the distance in hand)
  const bool want_coarse = auto_cell && g_tc_max_levels > (want_fine ? 2 : 1) && n > 4096 &&
                           ((low_thr > 0 && (double)low_pts > 0.01 * samp_n) || k_hint <= 1);

  // Level geometry, finest first.
  GridParams lg[kMaxLevels];
  int nl = 0, primary = 0;
  if (want_fine) {
    // (a sparse cloud in a big bbox would make a 4x finer dense table mostly empty cells: cap it)
    const GridParams gf = make_grid(
        mn, mx, g.cell * 0.25f, n,
        std::min<uint64_t>(table_cap, std::max<uint64_t>(g_tc_fine_cap * n, 1u << 18)));
    if (gf.cell < g.cell * 0.9f) lg[nl++] = gf;  // the table cap may refuse to refine
    else want_fine = false;
  }
  primary = nl;
  lg[nl++] = g;
  if (want_coarse) {
    // 4x for sparse regions of a kNN index (ring growth there walks thousands of empty rows);
    // 3x for the far queries of a 1-NN index, where the coarse cells' points are all candidates
    // (measured on the 100M-point ICP target: 2x 300, 3x 294, 4x 233, 6x 200 iterations/s, and
    // the build grows with the table: 27.0 / 24.1 / 22.5 / 21.4 ms)
    float f = (low_thr > 0 && (double)low_pts > 0.01 * samp_n) ? 4.0f : 3.0f;
    if (const char* e = std::getenv("TC_COARSE_FACTOR")) f = std::max(1.5f, (float)atof(e));  // (debug)
    lg[nl++] = make_grid(mn, mx, g.cell * f, n, table_cap);
  }

  // One persistent arena (cell_start tables + sorted points of every level) and one temporary
  // arena (histograms still to be counted + scan state), cleared by a single memset: every API
  // call costs the host 2-3 us, which is what bounds the build of a LiDAR-frame-sized cloud.
  uint64_t words = 0, cs_off[kMaxLevels], pts_off[kMaxLevels];
  for (int l = 0; l < nl; ++l) {
    cs_off[l] = words;
    words += round_up4(cells_of(lg[l]) + 1);
  }
  for (int l = 0; l < nl; ++l) {
    pts_off[l] = words;
    words += 4 * n;
  }
  // Slab-sharded build (one resolution, slabs of whole planes along the slowest axis): this
  // rank scans / scatters only the cells of its planes +- halo.  Nothing outside the slab is
  // written or read: the search's ring cap (GridParams.flags bits 8..15) keeps it inside.
  Slab slab_off;
  slab_off.cell_hi = slab_off.own_cell_hi = cells_of(lg[0]);
  Slab sl = nl == 1 ? slab_of(lg[0], rank, world, halo) : slab_off;
  if (sl.on && primary_counted && !plane_counts.empty() && margin > 0) {
    // point-balanced slabs (the grid is the trial's: its histogram covers them)
    std::vector<int> bnd((size_t)world + 1);
    balanced_boundaries(plane_counts.data(), sl.n_planes, n, world, margin, bnd.data());
    sl = slab_planes(lg[0], world, bnd[rank], bnd[rank + 1], halo);
  }
  const bool slab = sl.on;
  const uint64_t cell_lo = sl.cell_lo, cell_hi = sl.cell_hi, own_cell_lo = sl.own_cell_lo,
                 own_cell_hi = sl.own_cell_hi;
  // a sampled trial histogram is complete only inside the slab it was taken for
  if (tslab.on && primary_counted && !slab) primary_counted = false;
  uint64_t twords = 0, cnt_off[kMaxLevels], tiles = 0;
  for (int l = 0; l < nl; ++l) {
    tiles += (cells_of(lg[l]) + kScanTile - 1) / kScanTile;
    if (l == primary && primary_counted) continue;
    cnt_off[l] = twords;
    twords += round_up4(cells_of(lg[l]) + 1);
  }
  const uint64_t state_off = twords;  // u64 words, 8-byte aligned since twords % 4 == 0
  twords += 2 * (tiles + 1);
  uint32_t* d_tmp = nullptr;
  trace.mark("host: level decisions");
  // reuse an arena a freed index left behind (all use is ordered on ctx->stream)
  for (int a = 0; a < tc_context::kArenaSlots && !ix->d_arena; ++a)
    if (ctx->arena_cache[a] && ctx->arena_words[a] >= words && ctx->arena_words[a] <= 2 * words + 1024) {
      ix->d_arena = (uint32_t*)ctx->arena_cache[a];
      ix->arena_alloc_words = ctx->arena_words[a];
      ctx->arena_cache[a] = nullptr;
      ctx->arena_words[a] = 0;
    }
  st = TC_OK;
  if (!ix->d_arena) {
    st = tc_alloc(ctx, &ix->d_arena, words);
    ix->arena_alloc_words = words;
  }
  trace.mark("host: arena alloc");
  if (st == TC_OK) st = tc_ws_get_zeroed(ctx, 1, &d_tmp, twords);
  trace.mark("host: workspace");
  bool restored = false;
  LevelJobs all{}, todo{};
  ScanJobs scan{};
  if (st == TC_OK) {
    all.n = scan.n = nl;
    for (int l = 0; l < nl; ++l) {
      LevelJob& jb = all.l[l];
      jb.g = lg[l];
      jb.counts = (l == primary && primary_counted) ? d_trial : d_tmp + cnt_off[l];
      jb.cell_start = ix->d_arena + cs_off[l];
      jb.out = reinterpret_cast<float4*>(ix->d_arena + pts_off[l]);
      jb.aggregate = cells_of(lg[l]) < n ? 1 : 0;
      jb.cell_lo = slab ? (uint32_t)cell_lo : 0u;
      jb.cell_hi = slab ? (uint32_t)cell_hi : 0xFFFFFFFFu;
      if (!(l == primary && primary_counted)) todo.l[todo.n++] = jb;
      scan.in[l] = jb.counts + (slab ? cell_lo : 0);
      scan.out[l] = ix->d_arena + cs_off[l] + (slab ? cell_lo : 0);
      scan.len[l] = slab ? cell_hi - cell_lo : cells_of(lg[l]);
    }
    const int grid = grid_for(ctx, n, kThreads * points_per_thread(n));
    if (todo.n > 0) {
      k_hist_levels<<<grid, kThreads, 0, ctx->stream>>>(cloud->d_xyz, (uint32_t)n, todo);
      ctx->launches++;
    }
    st = scan_tables(ctx, scan, reinterpret_cast<unsigned long long*>(d_tmp + state_off));
    if (st == TC_OK) {
      // every histogram counts back down to zero in the scatter; the scan states and (when the
      // cell was rescaled) the abandoned trial histogram are cleared by the same launch
      ZeroJobs zero{};
      zero.p[0] = d_tmp + state_off;
      zero.n[0] = twords - state_off;
      zero.p[1] = d_trial + (tslab.on ? tslab.cell_lo : 0);
      zero.n[1] = primary_counted ? 0 : tslab.on ? tslab.cell_hi - tslab.cell_lo + 1 : trial_words;
      zero.p[2] = d_trial + trial_sample_off;  // the statistics sample of a slab build
      zero.n[2] = sample_words_of(tslab);
      if (primary_counted && tslab.on && slab) {  // counted for the margin, outside the final slab
        zero.p[3] = d_trial + tslab.cell_lo;
        zero.n[3] = cell_lo - tslab.cell_lo;
        zero.p[4] = d_trial + cell_hi;
        zero.n[4] = tslab.cell_hi - cell_hi;
      }
      k_scatter_levels<<<grid, kThreads, 0, ctx->stream>>>(cloud->d_xyz, (uint32_t)n, all, zero);
      ctx->launches++;
      trace.mark("launches: hist/scan/scatter");
      if (slab) {
        // the rank's own query range and point count, in local sorted positions
        const uint32_t* cs = ix->d_arena + cs_off[0];
        uint32_t* h = ctx->h_scratch + 44;
        cudaMemcpyAsync(h + 0, cs + own_cell_lo, 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(h + 1, cs + own_cell_hi, 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(h + 2, cs + cell_hi, 4, cudaMemcpyDeviceToHost, ctx->stream);
        trace.mark("slab: readback enqueued");
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
          st = tc_fail(ctx, TC_GPU, "sharded index build failed");
        trace.mark("slab: stream sync");
        ix->sharded = true;
        ix->shard_halo = halo;
        ix->cell_lo = cell_lo;
        ix->cell_hi = cell_hi;
        ix->own_cell_lo = own_cell_lo;
        ix->own_cell_hi = own_cell_hi;
        ix->own_lo = h[0];
        ix->own_hi = h[1];
        ix->n_local = h[2];
      }
      if (st == TC_OK && cudaGetLastError() != cudaSuccess)
        st = tc_fail(ctx, TC_GPU, "index build launch failed");
      else if (st == TC_OK) restored = true;
    }
  }
  tc_ws_release_zeroed(ctx, 1, d_tmp, restored);
  tc_ws_release_zeroed(ctx, 0, d_trial, restored);
  trace.mark("levels: hist+scan+scatter");
  if (st != TC_OK) {
    tc_index_free(ix);
    return st;
  }
  for (int l = 0; l < nl; ++l) {
    GridLevel& lv = ix->lv[l];
    lv.g = lg[l];
    lv.n_cells = cells_of(lg[l]);
    lv.d_cell_start = ix->d_arena + cs_off[l];
    lv.d_pts = reinterpret_cast<float4*>(ix->d_arena + pts_off[l]);
    // finer cells nest inside primary cells (same origin, edge / 4): their population is bounded
    // by the primary maximum; the slack covers points that f32 rounding puts across a cell face
    lv.max_pop_bound = 2 * maxpop_est + 32;
  }
  ix->lv[primary].occupied = occ_est;
  ix->lv[primary].max_pop = maxpop_est;
  ix->primary = primary;
  ix->n_levels = nl;
  *out = ix;
  return TC_OK;
}

// Largest population of a level-0 cell, exact (max over adjacent differences of cell_start),
// computed on first use and cached: the sharded normals launch must cover every position of a
// cell that starts inside its range, and the build-time figures are only estimates.
__global__ void __launch_bounds__(kThreads) k_max_population(const uint32_t* __restrict__ cs,
                                                             uint64_t n_cells,
                                                             uint32_t* __restrict__ out) {
  uint32_t m = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells;
       i += (uint64_t)gridDim.x * blockDim.x)
    m = max(m, cs[i + 1] - cs[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
int tci_level0_max_population(tc_context* ctx, const tc_index* ix, uint32_t* out) {
  if (ix->level0_max_pop == 0 && ix->n > 0) {
    uint32_t* d = ctx->d_scratch + 40;  // a free scratch word
    TC_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(uint32_t), ctx->stream));
    const GridLevel& l0 = ix->lv[0];
    k_max_population<<<grid_for(ctx, l0.n_cells, kThreads * 4), kThreads, 0, ctx->stream>>>(
        l0.d_cell_start, l0.n_cells, d);
    TC_LAUNCHED(ctx);
    TC_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 40, d, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ix->level0_max_pop = ctx->h_scratch[40];
  }
  *out = ix->level0_max_pop;
  return TC_OK;
}

extern "C" void tc_index_free(tc_index* ix) {
  if (!ix) return;
  // the levels are views into the arena; keep it for the next build when a slot is free (or
  // holds a smaller one), else return it to the pool
  tc_context* ctx = ix->ctx;
  void* p = ix->d_arena;
  uint64_t w = ix->arena_alloc_words;
  if (p && w > 0) {
    int slot = -1;
    for (int a = 0; a < tc_context::kArenaSlots; ++a)
      if (!ctx->arena_cache[a]) slot = a;
    if (slot < 0)
      for (int a = 0; a < tc_context::kArenaSlots; ++a)
        if (ctx->arena_words[a] < w) slot = a;
    if (slot >= 0) {
      std::swap(p, ctx->arena_cache[slot]);
      std::swap(w, ctx->arena_words[slot]);
    }
  }
  tc_free(ctx, p);
  delete ix;
}

extern "C" int tc_index_get_info(const tc_index* ix, tc_index_info* out) {
  if (!ix || !out) return TC_INVALID_DATA;
  const GridLevel& p = ix->lv[ix->primary];
  out->n_points = ix->n;
  out->n_cells = p.n_cells;
  out->dims[0] = p.g.nx;
  out->dims[1] = p.g.ny;
  out->dims[2] = p.g.nz;
  out->cell_size = p.g.cell;
  for (int a = 0; a < 3; ++a) {
    out->bbox_min[a] = ix->bbox_min[a];
    out->bbox_max[a] = ix->bbox_max[a];
  }
  out->occupied_cells = p.occupied;
  out->max_cell_population = p.max_pop;
  out->n_levels = (uint32_t)ix->n_levels;
  return TC_OK;
}

// Sort arbitrary points (queries / ICP source) by the cells of grid g so that neighbouring
// threads walk neighbouring cells.  Key = clamped cell id in g.
int tci_sort_by_grid(tc_context* ctx, const float* d_xyz, uint64_t n, const GridParams& g,
                     float4** d_sorted) {
  TcRange nvtx_range("tc:sort_by_grid (radix)");
  *d_sorted = nullptr;
  if (n == 0) return TC_OK;
  uint32_t *d_keys = nullptr, *d_keys_alt = nullptr, *d_vals = nullptr, *d_vals_alt = nullptr;
  uint32_t *ks = nullptr, *vs = nullptr;
  int st = tc_alloc(ctx, &d_keys, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_keys_alt, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_vals, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_vals_alt, n);
  if (st == TC_OK) {
    k_cell_keys<<<grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(d_xyz, (uint32_t)n, g,
                                                                         d_keys);
    ctx->launches++;
    const uint64_t n_cells = (uint64_t)g.nx * g.ny * g.nz;
    st = tci_radix_sort_pairs(ctx, d_keys, d_vals, d_keys_alt, d_vals_alt, (uint32_t)n,
                              key_bits_for(n_cells), false, &ks, &vs);
  }
  if (st == TC_OK) st = tc_alloc(ctx, d_sorted, n);
  if (st == TC_OK) {
    k_gather<<<grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(d_xyz, vs, (uint32_t)n,
                                                                       *d_sorted);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "gather launch failed");
  }
  tc_free(ctx, d_keys);
  tc_free(ctx, d_keys_alt);
  tc_free(ctx, d_vals);
  tc_free(ctx, d_vals_alt);
  return st;
}
