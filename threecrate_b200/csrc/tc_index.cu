// tc_index.cu — device-resident uniform-grid index: bbox -> cell keys -> LSD radix sort ->
// cell-range scan -> sorted float4 points.  Replaces KdTree::new
// (threecrate-algorithms/src/nearest_neighbor.rs:37-159), which is a serial O(N log N) quickselect.
//
// All kernels are HBM-bound streaming passes (coalesced loads, grid sized in multiples of the SM
// count).  Algorithmic bytes per point: bbox 12, keys+histogram 12r+4w, radix passes
// (8r+8w) x P with P = ceil(key_bits/8), gather 4r+12r+16w, cell-range scan 8 B/cell.
#include <chrono>
#include <cstdlib>

#include "tc_internal.cuh"

namespace {

// TC_TRACE=1: print host-side phase timings of tc_index_build (adds stream syncs; debug only)
struct PhaseTrace {
  bool on;
  tc_context* ctx;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTrace(tc_context* c) : on(std::getenv("TC_TRACE") != nullptr), ctx(c) {
    if (on) {
      cudaStreamSynchronize(ctx->stream);
      t0 = std::chrono::steady_clock::now();
    }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
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
__global__ void k_scratch_init(uint32_t* s) {
  const int t = threadIdx.x;
  if (t < 3) s[t] = 0xFFFFFFFFu;       // min (ordered encoding)
  else if (t < 6) s[t] = 0u;           // max
  else if (t < 64) s[t] = 0u;
}

__global__ void __launch_bounds__(kThreads) k_bbox(const float* __restrict__ xyz, uint64_t n,
                                                   uint32_t* __restrict__ s) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  uint32_t bad = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = xyz[3 * i + a];
      if (!isfinite(v)) bad = 1;
      mn[a] = fminf(mn[a], v);
      mx[a] = fmaxf(mx[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  bad = __any_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (mn[a] <= mx[a]) {
        atomicMin(&s[a], f2ord(mn[a]));
        atomicMax(&s[3 + a], f2ord(mx[a]));
      }
    }
    if (bad) atomicOr(&s[6], 1u);
  }
}

// ------------------------------------------------------------------- cell keys + histogram
// MODE 0: keys + histogram (index build).  MODE 1: keys only (spatial sort of queries).
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_cell_keys(const float* __restrict__ xyz, uint32_t n,
                                                        GridParams g, uint32_t* __restrict__ keys,
                                                        uint32_t* __restrict__ counts) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float u;
    const int cx = cell_coord(xyz[3 * (uint64_t)i + 0], g.ox, g.inv, g.nx, u);
    const int cy = cell_coord(xyz[3 * (uint64_t)i + 1], g.oy, g.inv, g.ny, u);
    const int cz = cell_coord(xyz[3 * (uint64_t)i + 2], g.oz, g.inv, g.nz, u);
    const uint32_t c = cell_id(g, cx, cy, cz);
    keys[i] = c;
    if (MODE == 0) {
      // warp-aggregated: consecutive points of a scan usually share a cell, and coarse levels
      // funnel thousands of points into one counter
      const uint32_t peers = __match_any_sync(__activemask(), c);
      if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&counts[c], (uint32_t)__popc(peers));
    }
  }
}

// occupied cells -> s[8], max population -> s[9], points living in cells with population
// <= low_thr -> s[10] (how much of the cloud is too sparse for this cell size)
__global__ void __launch_bounds__(kThreads) k_cell_stats(const uint32_t* __restrict__ counts,
                                                         uint64_t n_cells, uint32_t low_thr,
                                                         uint32_t* __restrict__ s) {
  uint32_t occ = 0, mx = 0, low[4] = {0, 0, 0, 0};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = counts[i];
    occ += (c != 0);
    mx = max(mx, c);
#pragma unroll
    for (int j = 0; j < 4; ++j) low[j] += (c <= (low_thr << j)) ? c : 0u;  // thresholds x1,2,4,8
  }
  for (int o = 16; o > 0; o >>= 1) {
    occ += __shfl_xor_sync(0xffffffffu, occ, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll
    for (int j = 0; j < 4; ++j) low[j] += __shfl_xor_sync(0xffffffffu, low[j], o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (occ) atomicAdd(&s[8], occ);
    if (mx) atomicMax(&s[9], mx);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (low[j]) atomicAdd(&s[10 + j], low[j]);
  }
}

// Counting-sort scatter: point i goes to cursor[cell(i)]++ as float4 (x, y, z, bits(i)).
// The order INSIDE a cell is arrival order (not
// deterministic) — nothing downstream depends on it: every selection is keyed by (d2, original
// index) and multi-GPU shards own whole cells.
__global__ void __launch_bounds__(kThreads) k_scatter_cells(const float* __restrict__ xyz,
                                                            const uint32_t* __restrict__ keys,
                                                            uint32_t n,
                                                            const uint32_t* __restrict__ cell_start,
                                                            uint32_t* __restrict__ counts,
                                                            float4* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = xyz + 3 * (uint64_t)i;
    const uint32_t c = keys[i];
    // the histogram itself is the cursor: it counts down to zero while the cell fills up
    // (warp-aggregated: one atomic per distinct cell per warp)
    const uint32_t active = __activemask();
    const uint32_t peers = __match_any_sync(active, c);
    const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    uint32_t old = 0;
    if (lane == leader) old = atomicSub(&counts[c], (uint32_t)__popc(peers));
    old = __shfl_sync(active, old, leader);
    const uint32_t pos = __ldg(&cell_start[c]) + old - 1u - (uint32_t)__popc(peers & ((1u << lane) - 1u));
    out[pos] = make_float4(p[0], p[1], p[2], __uint_as_float(i));
  }
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

// state[0] = ticket counter (as u64), state[1 + t] = (status << 32 | value) of tile t; zeroed.
// out has n + 1 entries: out[n] = grand total.  in == out is allowed.
__global__ void __launch_bounds__(kScanThreads)
k_scan_lookback(const uint32_t* in, uint64_t n, uint32_t* out, unsigned long long* state) {
  __shared__ uint32_t sw[kScanThreads / 32 + 1];
  __shared__ uint32_t s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = (uint32_t)atomicAdd(&state[0], 1ull);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
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
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) s += v[it];
  uint32_t total;
  uint32_t ex = block_exclusive_scan(s, sw, total);
  if (threadIdx.x < 32) {
    // warp-wide look-back: lane j inspects predecessor (p - j); the nearest tile that already
    // published an inclusive prefix ends the walk, the aggregates in front of it are summed
    volatile unsigned long long* st = state + 1;
    const int lane = threadIdx.x;
    uint32_t prefix = 0;
    if (tile == 0) {
      if (lane == 0) st[0] = kTileInclusive | total;
    } else {
      if (lane == 0) {
        st[tile] = kTileAggregate | total;
        __threadfence();
      }
      for (int64_t p = (int64_t)tile - 1; p >= 0; p -= 32) {
        const int64_t idx = p - lane;
        unsigned long long w = kTileInclusive;  // before tile 0: an inclusive prefix of 0
        if (idx >= 0) {
          do {
            w = st[idx];
          } while ((w >> 32) == 0);  // predecessor is running (ticket order): short spin
        }
        const unsigned incl = __ballot_sync(0xffffffffu, (w >> 32) == 2);
        const int first = incl ? (__ffs(incl) - 1) : 31;
        uint32_t val = lane <= first ? (uint32_t)w : 0u;
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        prefix += val;
        if (incl) break;
      }
      if (lane == 0) {
        __threadfence();
        st[tile] = kTileInclusive | (unsigned long long)(prefix + total);
      }
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int it = 0; it < kScanItems; ++it) {
    const uint64_t i = base + it;
    if (i < n) out[i] = ex;
    ex += v[it];
    if (i == n - 1) out[n] = ex;
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
int tci_exclusive_scan_u32(tc_context* ctx, const uint32_t* d_in, uint32_t* d_out, uint64_t n) {
  if (n == 0) {
    TC_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(uint32_t), ctx->stream));
    return TC_OK;
  }
  const uint32_t n_tiles = (uint32_t)((n + kScanTile - 1) / kScanTile);
  unsigned long long* d_state = nullptr;
  TC_TRY(tc_alloc(ctx, &d_state, (uint64_t)n_tiles + 1));
  TC_CUDA(ctx, cudaMemsetAsync(d_state, 0, ((uint64_t)n_tiles + 1) * sizeof(unsigned long long),
                               ctx->stream));
  k_scan_lookback<<<n_tiles, kScanThreads, 0, ctx->stream>>>(d_in, n, d_out, d_state);
  TC_LAUNCHED(ctx);
  tc_free(ctx, d_state);
  return TC_OK;
}

int tci_radix_sort_pairs(tc_context* ctx, uint32_t* d_keys, uint32_t* d_vals, uint32_t* d_keys_alt,
                         uint32_t* d_vals_alt, uint32_t n, int key_bits, uint32_t** keys_out,
                         uint32_t** vals_out) {
  // d_vals may be nullptr on entry => identity values; the result then lands in the alt buffers
  // or (d_keys, d_vals_alt2) — to keep it simple the caller always provides both value buffers.
  const uint32_t n_tiles = (n + kRsTile - 1) / kRsTile;
  uint32_t* d_hist = nullptr;
  TC_TRY(tc_alloc(ctx, &d_hist, (uint64_t)256 * n_tiles + 1));
  uint32_t *kin = d_keys, *kout = d_keys_alt, *vin = nullptr, *vout = d_vals_alt;
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

int tci_bbox(tc_context* ctx, const float* d_xyz, uint64_t n, float mn[3], float mx[3]) {
  k_scratch_init<<<1, 64, 0, ctx->stream>>>(ctx->d_scratch);
  TC_LAUNCHED(ctx);
  k_bbox<<<grid_for(ctx, n, kThreads * 4), kThreads, 0, ctx->stream>>>(d_xyz, n, ctx->d_scratch);
  TC_LAUNCHED(ctx);
  TC_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 8 * sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, ctx->stream));
  TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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

// histogram of `cloud` on grid g (+ keys); returns occupied / max_pop / low-population points
int level_histogram(tc_context* ctx, const tc_cloud* cloud, const GridParams& g, uint32_t low_thr,
                    uint32_t* d_keys, uint32_t** d_counts_out, uint32_t stats[6]) {
  const uint64_t n = cloud->n;
  const uint64_t n_cells = (uint64_t)g.nx * g.ny * g.nz;
  uint32_t* d_counts = nullptr;
  TC_TRY(tc_alloc(ctx, &d_counts, n_cells + 1));
  TC_CUDA(ctx, cudaMemsetAsync(d_counts, 0, (n_cells + 1) * sizeof(uint32_t), ctx->stream));
  k_cell_keys<0><<<grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(
      cloud->d_xyz, (uint32_t)n, g, d_keys, d_counts);
  TC_LAUNCHED(ctx);
  if (stats) {
    TC_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch + 8, 0, 6 * sizeof(uint32_t), ctx->stream));
    k_cell_stats<<<grid_for(ctx, n_cells, kThreads * 4), kThreads, 0, ctx->stream>>>(
        d_counts, n_cells, low_thr, ctx->d_scratch);
    TC_LAUNCHED(ctx);
    TC_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch + 8, ctx->d_scratch + 8, 6 * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int j = 0; j < 6; ++j) stats[j] = ctx->h_scratch[8 + j];
  }
  *d_counts_out = d_counts;
  return TC_OK;
}

// counts -> cell_start (scan), then counting-sort scatter into the level's float4 array
int level_finish(tc_context* ctx, const tc_cloud* cloud, const GridParams& g, uint32_t* d_keys,
                 uint32_t* d_counts, GridLevel* lv) {
  const uint64_t n = cloud->n;
  lv->g = g;
  lv->n_cells = (uint64_t)g.nx * g.ny * g.nz;
  int st = tc_alloc(ctx, &lv->d_cell_start, lv->n_cells + 1);
  if (st == TC_OK) st = tci_exclusive_scan_u32(ctx, d_counts, lv->d_cell_start, lv->n_cells);
  if (st == TC_OK) st = tc_alloc(ctx, &lv->d_pts, n);
  if (st == TC_OK) {
    k_scatter_cells<<<grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(
        cloud->d_xyz, d_keys, (uint32_t)n, lv->d_cell_start, d_counts, lv->d_pts);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "scatter launch failed");
  }
  return st;
}

// One extra resolution, issued on side stream `aux` (allocations included: stream-ordered memory
// may be used on any stream once the join event orders it).
int build_extra_level(tc_context* ctx, int aux, const tc_cloud* cloud, const float mn[3],
                      const float mx[3], float cell, uint64_t table_cap, GridLevel* lv) {
  cudaStream_t main_stream = ctx->stream;
  TC_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[aux], ctx->ev_fork, 0));
  ctx->stream = ctx->aux[aux];
  const GridParams g = make_grid(mn, mx, cell, cloud->n, table_cap);
  uint32_t* d_keys = nullptr;
  uint32_t* d_counts = nullptr;
  int st = tc_alloc(ctx, &d_keys, cloud->n);
  if (st == TC_OK) st = level_histogram(ctx, cloud, g, 0, d_keys, &d_counts, nullptr);
  if (st == TC_OK) st = level_finish(ctx, cloud, g, d_keys, d_counts, lv);
  tc_free(ctx, d_counts);
  tc_free(ctx, d_keys);
  if (st == TC_OK && cudaEventRecord(ctx->ev_join[aux], ctx->stream) != cudaSuccess)
    st = tc_fail(ctx, TC_GPU, "event record failed");
  ctx->stream = main_stream;
  return st;
}

}  // namespace

int g_tc_max_levels = kMaxLevels;
uint64_t g_tc_fine_cap = 64;  // fine-level table cap, cells per point
extern "C" void tc_debug_set_fine_cap(int cells_per_point) { g_tc_fine_cap = (uint64_t)cells_per_point; }
extern "C" void tc_debug_set_max_levels(int n) {
  g_tc_max_levels = n < 1 ? 1 : (n > kMaxLevels ? kMaxLevels : n);
}

extern "C" int tc_index_build(tc_context* ctx, const tc_cloud* cloud, uint32_t k_hint,
                              float cell_size, tc_index** out) {
  if (!ctx || !cloud || !out) return TC_INVALID_DATA;
  *out = nullptr;
  const uint64_t n = cloud->n;
  if (n >= 0xFFFFFFFFull) return tc_fail(ctx, TC_INVALID_DATA, "cloud too large (N must be < 2^32-1)");
  tc_index* ix = new tc_index();
  ix->ctx = ctx;
  ix->cloud = cloud;
  ix->n = n;
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

  uint32_t* d_keys = nullptr;
  uint32_t* d_counts = nullptr;
  st = tc_alloc(ctx, &d_keys, n);
  if (st != TC_OK) {
    delete ix;
    return st;
  }
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
  GridParams g{};
  uint32_t stats[6] = {0, 0, 0, 0, 0, 0};
  // One measured trial (histogram + stats + one host sync).  If the mean population of occupied
  // cells is off target the cell is rescaled ONCE (surface-like scaling: occupied ~ cell^-2) and
  // re-histogrammed without waiting for new statistics; the level decisions below use the
  // measured trial's skew, which is scale-free.
  float stat_scale = 1.0f;  // final cell / measured cell
  g = make_grid(mn, mx, cell, n, table_cap);
  cell = g.cell;
  st = level_histogram(ctx, cloud, g, low_thr, d_keys, &d_counts, stats);
  trace.mark("trial histogram+stats");
  float pop1 = (float)n / (float)std::max(1u, stats[0]);
  if (st == TC_OK && auto_cell && !(pop1 > target * 0.7f && pop1 < target * 1.4f)) {
    float scale = std::sqrt(target / pop1);
    scale = std::min(4.0f, std::max(0.25f, scale));
    float next = cell * scale;
    if (emax > 0 && next > emax) next = (float)emax;
    const GridParams g2 = make_grid(mn, mx, next, n, table_cap);
    if (std::fabs(g2.cell - cell) > 1e-3f * cell) {
      tc_free(ctx, d_counts);
      d_counts = nullptr;
      st = level_histogram(ctx, cloud, g2, 0, d_keys, &d_counts, nullptr);
      stat_scale = g2.cell / cell;
      g = g2;
      cell = g2.cell;
      trace.mark("rescaled histogram");
    }
  }
  if (st != TC_OK) {
    tc_free(ctx, d_keys);
    tc_free(ctx, d_counts);
    delete ix;
    return st;
  }
  // statistics of the final grid, extrapolated from the measured one when it was rescaled
  const float s2 = stat_scale * stat_scale;
  const uint32_t occ_est = (uint32_t)std::max(1.0f, (float)stats[0] / s2);
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
  const bool want_fine = auto_cell && g_tc_max_levels > 1 && skew > target_ratio && n > 4096;
  const bool want_coarse = auto_cell && g_tc_max_levels > (want_fine ? 2 : 1) &&
                           low_thr > 0 && (double)low_pts > 0.01 * (double)n && n > 4096;
  // The extra resolutions only depend on the bbox and the decisions above: they are issued on the
  // two side streams first and run concurrently with the primary scan + scatter.
  GridLevel fine{}, coarse{};
  bool have_fine = false, have_coarse = false;
  if (want_fine || want_coarse) TC_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  if (want_fine) {
    // (a sparse cloud in a big bbox would make a 4x finer dense table mostly empty cells: cap it)
    st = build_extra_level(ctx, 0, cloud, mn, mx, g.cell * 0.25f,
                           std::min<uint64_t>(table_cap, std::max<uint64_t>(g_tc_fine_cap * n, 1u << 18)),
                           &fine);
    have_fine = true;
  }
  if (st == TC_OK && want_coarse) {
    st = build_extra_level(ctx, 1, cloud, mn, mx, g.cell * 4.0f, table_cap, &coarse);
    have_coarse = true;
  }
  GridLevel primary{};
  primary.occupied = occ_est;
  primary.max_pop = maxpop_est;
  if (st == TC_OK) st = level_finish(ctx, cloud, g, d_keys, d_counts, &primary);
  tc_free(ctx, d_counts);
  if (have_fine) cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0);
  if (have_coarse) cudaStreamWaitEvent(ctx->stream, ctx->ev_join[1], 0);
  trace.mark("levels (concurrent)");
  if (st != TC_OK) {  // nothing was handed to the index yet
    cudaStreamSynchronize(ctx->stream);
    for (GridLevel* l : {&fine, &primary, &coarse}) {
      tc_free(ctx, l->d_pts);
      tc_free(ctx, l->d_cell_start);
    }
    tc_free(ctx, d_keys);
    delete ix;
    return st;
  }
  int nl = 0;
  if (have_fine) {
    if (fine.g.cell < g.cell * 0.9f) {
      ix->lv[nl++] = fine;
    } else {  // the table cap refused to refine
      tc_free(ctx, fine.d_pts);
      tc_free(ctx, fine.d_cell_start);
    }
  }
  ix->primary = nl;
  ix->lv[nl++] = primary;
  if (have_coarse) ix->lv[nl++] = coarse;
  ix->n_levels = nl;
  // finer cells nest inside primary cells (same origin, edge / 4): their population is bounded by
  // the primary maximum; the slack covers points that f32 rounding puts across a cell face
  for (int i = 0; i < nl; ++i) ix->lv[i].max_pop_bound = 2 * primary.max_pop + 32;
  tc_free(ctx, d_keys);
  *out = ix;
  return TC_OK;
}

extern "C" void tc_index_free(tc_index* ix) {
  if (!ix) return;
  for (int i = 0; i < kMaxLevels; ++i) {
    tc_free(ix->ctx, ix->lv[i].d_pts);
    tc_free(ix->ctx, ix->lv[i].d_cell_start);
  }
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
  *d_sorted = nullptr;
  if (n == 0) return TC_OK;
  uint32_t *d_keys = nullptr, *d_keys_alt = nullptr, *d_vals = nullptr, *d_vals_alt = nullptr;
  uint32_t *ks = nullptr, *vs = nullptr;
  int st = tc_alloc(ctx, &d_keys, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_keys_alt, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_vals, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_vals_alt, n);
  if (st == TC_OK) {
    k_cell_keys<1><<<grid_for(ctx, n, kThreads), kThreads, 0, ctx->stream>>>(d_xyz, (uint32_t)n, g,
                                                                            d_keys, nullptr);
    ctx->launches++;
    const uint64_t n_cells = (uint64_t)g.nx * g.ny * g.nz;
    st = tci_radix_sort_pairs(ctx, d_keys, d_vals, d_keys_alt, d_vals_alt, (uint32_t)n,
                              key_bits_for(n_cells), &ks, &vs);
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
