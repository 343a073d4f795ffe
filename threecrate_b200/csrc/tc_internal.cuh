// tc_internal.cuh — shared declarations of the threecrate_cuda library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <string>

#include "../../include/threecrate_cuda.h"

// ------------------------------------------------------------------------------------------
// host-side objects behind the opaque handles
// ------------------------------------------------------------------------------------------
struct tc_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;
  // small device scratch (bbox / stats / flags) and its pinned host mirror
  static constexpr int kPlaneWords = 4096;  // per-plane point counts of a slab build (words 64..)
  static constexpr int kScratchWords = 64 + kPlaneWords;  // (host mirror: counts at 64..)
  static constexpr int kPlaneStride = 32;   // device counters 128 B apart (one L2 line each: a
                                            // thousand blocks add to every one of them)
  uint32_t* d_planes = nullptr;    // kPlaneWords x kPlaneStride words, all zero between builds
  uint32_t* d_scratch = nullptr;   // 64 words
  uint32_t* h_scratch = nullptr;   // pinned + device-visible (UVA), 64 words: kernels publish small
                                   // results here and the host polls a sequence word
  uint32_t seq = 0;                // last sequence number handed to a publishing kernel
  // index arenas handed back by tc_index_free, kept for the next build of a similar size: a
  // cudaMallocAsync of a few hundred MB right after a stream synchronisation was measured at
  // 0.5 - 50 ms of HOST time (the pool re-maps the block), which is the whole build time
  static constexpr int kArenaSlots = 2;
  void* arena_cache[kArenaSlots] = {nullptr, nullptr};
  uint64_t arena_words[kArenaSlots] = {0, 0};
  bool stats_on = false;           // device counters of the search kernels (d_scratch[48..55])
  uint64_t stats_queries = 0;      // queries of the last search launch
  tc_stats icp_stats{};            // per-iteration history of the last ICP call (host copy)
  // small grow-only device workspaces reused across calls (index-build histograms, fallback
  // lists): every allocation call costs the host 1-2 us and the LiDAR-frame path is host-bound
  static constexpr int kWsSlots = 3;
  static constexpr uint64_t kWsMaxBytes = 1ull << 30;  // larger requests are not cached (a
                                   // cudaMallocAsync of hundreds of MB right after a stream
                                   // synchronisation costs 0.5 - 50 ms of host time)
  void* ws[kWsSlots] = {nullptr, nullptr, nullptr};
  uint64_t ws_bytes[kWsSlots] = {0, 0, 0};
  bool ws_zero[kWsSlots] = {false, false, false};  // cached buffer known to be all zero
  // distributed normals (tc_estimate_normals_distributed): while chunk != 0 the normals kernels
  // write row i into the buffer of the rank that owns original index i (i / chunk), over NVLink
  // peer memory; base[r] is pre-offset so that base[r] + 6 i is row i of rank r's chunk
  struct OutRoute {
    float* base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t chunk = 0, magic = 0;  // magic = floor(2^32 / chunk)
  } route;
};

struct tc_cloud {
  tc_context* ctx = nullptr;
  uint64_t n = 0;
  float* d_xyz = nullptr;  // n x 3 AoS, original order
};

// Grid geometry, passed to kernels by value.
struct GridParams {
  float ox, oy, oz;     // origin = bbox min
  float inv;            // 1 / cell (the f32 value actually used for cell assignment)
  float cell;           // cell edge
  float ex, ey, ez;     // bbox extents (for the conservative ring bound)
  int nx, ny, nz;
  uint32_t n;           // points in the index
  int flags;            // search variant bits (see tc_search.cuh)
  int ymajor;           // cell id = (cy nz + cz) nx + cx instead of (cz ny + cy) nx + cx: the axis
                        // with more cells is the slowest one, so a contiguous range of the sorted
                        // order is a slab along the long axis (multi-GPU shards: thin halos)
};

// One resolution of the index: a complete uniform grid over all points.
constexpr int kMaxLevels = 3;
struct GridLevel {
  GridParams g{};
  float4* d_pts = nullptr;           // sorted by cell: x, y, z, bits(original index)
  uint32_t* d_cell_start = nullptr;  // n_cells + 1
  uint64_t n_cells = 0;
  uint32_t occupied = 0, max_pop = 0;
  uint32_t max_pop_bound = 0;        // estimate of the largest cell population (informational)
};
// What the search kernels receive (by value): the levels from fine to coarse.
struct LevelSet {
  int n;
  GridParams g[kMaxLevels];
  const float4* pts[kMaxLevels];
  const uint32_t* cs[kMaxLevels];
  // slab-sharded index (tc_index_build_sharded): only the cells of the rank's slab +- `halo`
  // planes hold points; a search whose final block radius exceeds the halo may have missed
  // points of the unbuilt part and is counted in *unsafe (the caller then re-runs unsharded)
  int halo;            // 0: complete index
  uint32_t* unsafe;
  // output routing of the distributed normals (tc_context::OutRoute); chunk == 0: rows go to the
  // launch's own output buffer
  float* route[8];
  uint32_t route_chunk, route_magic;
};

struct tc_index {
  tc_context* ctx = nullptr;
  const tc_cloud* cloud = nullptr;  // borrowed; must outlive the index
  uint64_t n = 0;
  uint64_t arena_alloc_words = 0;   // size of d_arena as allocated (may exceed what is used)
  float bbox_min[3]{}, bbox_max[3]{};
  int n_levels = 0;
  int primary = 0;                  // the level built for the requested / automatic cell size
  GridLevel lv[kMaxLevels];         // fine -> coarse; views into d_arena
  uint32_t* d_arena = nullptr;      // cell_start tables + sorted float4 points of every level
  mutable uint32_t level0_max_pop = 0;  // exact, lazily computed (tci_level0_max_population)
  // slab-sharded build (multi-GPU normals): the index holds only the points of cells
  // [cell_lo, cell_hi) = the rank's planes +- shard_halo; the rank OWNS the queries at sorted
  // positions [own_lo, own_hi) (whole planes).  n_local = points held.
  bool sharded = false;
  int shard_rank = 0, shard_world = 1, shard_halo = 0;
  uint64_t cell_lo = 0, cell_hi = 0, own_cell_lo = 0, own_cell_hi = 0;
  uint64_t own_lo = 0, own_hi = 0, n_local = 0;
  uint32_t k_hint = 0;
  float cell_size_arg = 0.0f;
  LevelSet level_set(int flags) const {
    LevelSet s{};
    s.n = n_levels;
    s.halo = sharded ? shard_halo : 0;
    s.unsafe = sharded ? ctx->d_scratch + 42 : nullptr;
    for (int r = 0; r < 8; ++r) s.route[r] = ctx->route.base[r];
    s.route_chunk = ctx->route.chunk;
    s.route_magic = ctx->route.magic;
    for (int i = 0; i < n_levels; ++i) {
      s.g[i] = lv[i].g;
      // bits 8..15: ring cap of a slab-sharded index (grid_search never reads unbuilt cells)
      s.g[i].flags = (flags & 255) | (sharded ? (shard_halo & 255) << 8 : 0);
      s.pts[i] = lv[i].d_pts;
      s.cs[i] = lv[i].d_cell_start;
    }
    return s;
  }
};

struct tc_comm;  // tc_comm.cu

// ------------------------------------------------------------------------------------------
// NVTX ranges per phase (SURVEY §5 tracing): header-only nvtx3, a no-op unless a profiler
// (nsys / ncu --nvtx) is attached
// ------------------------------------------------------------------------------------------
#include <nvtx3/nvToolsExt.h>
struct TcRange {
  explicit TcRange(const char* name) { nvtxRangePushA(name); }
  ~TcRange() { nvtxRangePop(); }
  TcRange(const TcRange&) = delete;
  TcRange& operator=(const TcRange&) = delete;
};

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
inline int tc_fail(tc_context* ctx, int status, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return status;
}
#define TC_CUDA(ctx, call)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return tc_fail((ctx), TC_GPU,                                                         \
                     std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                         ":" + std::to_string(__LINE__) + ")");                             \
  } while (0)
#define TC_TRY(expr)               \
  do {                             \
    int s__ = (expr);              \
    if (s__ != TC_OK) return s__;  \
  } while (0)
// count + check a kernel launch
#define TC_LAUNCHED(ctx)                         \
  do {                                           \
    (ctx)->launches++;                           \
    TC_CUDA((ctx), cudaGetLastError());          \
  } while (0)

// stream-ordered allocation from the device's default pool (cached across calls)
template <typename T>
inline int tc_alloc(tc_context* ctx, T** p, uint64_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  TC_CUDA(ctx, cudaMallocAsync((void**)p, count * sizeof(T), ctx->stream));
  return TC_OK;
}
inline void tc_free(tc_context* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}
// Workspace slot `slot`: returns a buffer of at least `count` T, cached in the context when small
// (release with tc_ws_release, which frees only uncached buffers).  All use is stream-ordered on
// ctx->stream and a context serves one call at a time, so reuse needs no further ordering.
template <typename T>
inline int tc_ws_get(tc_context* ctx, int slot, T** p, uint64_t count) {
  const uint64_t bytes = std::max<uint64_t>(count, 1) * sizeof(T);
  if (bytes > tc_context::kWsMaxBytes) return tc_alloc(ctx, p, count);
  if (ctx->ws_bytes[slot] < bytes) {
    if (ctx->ws[slot]) cudaFreeAsync(ctx->ws[slot], ctx->stream);
    ctx->ws[slot] = nullptr;
    ctx->ws_bytes[slot] = 0;
    ctx->ws_zero[slot] = false;  // the pool may hand back the same base address: contents unknown
    const uint64_t grown = bytes + bytes / 4;
    TC_CUDA(ctx, cudaMallocAsync(&ctx->ws[slot], grown, ctx->stream));
    ctx->ws_bytes[slot] = grown;
  }
  *p = (T*)ctx->ws[slot];
  return TC_OK;
}
inline void tc_ws_release(tc_context* ctx, int slot, void* p) {
  if (p && p != ctx->ws[slot]) cudaFreeAsync(p, ctx->stream);
}
// Zero-initialised variant: the user promises (restored = true on release) that the work it
// enqueued leaves the buffer all zero again - histograms count back down to zero while the
// scatter consumes them - so the next call needs no memset.  Any early exit releases with
// restored = false and the next user clears the buffer.
template <typename T>
inline int tc_ws_get_zeroed(tc_context* ctx, int slot, T** p, uint64_t count) {
  TC_TRY(tc_ws_get(ctx, slot, p, count));  // (a reallocation clears ws_zero)
  const bool cached = ((void*)*p == ctx->ws[slot]);
  if (!cached || !ctx->ws_zero[slot]) {
    const uint64_t bytes = cached ? ctx->ws_bytes[slot] : std::max<uint64_t>(count, 1) * sizeof(T);
    TC_CUDA(ctx, cudaMemsetAsync(*p, 0, bytes, ctx->stream));
  }
  if (cached) ctx->ws_zero[slot] = false;  // in use
  return TC_OK;
}
inline void tc_ws_release_zeroed(tc_context* ctx, int slot, void* p, bool restored) {
  if (p && p == ctx->ws[slot]) ctx->ws_zero[slot] = restored;
  tc_ws_release(ctx, slot, p);
}

// ------------------------------------------------------------------------------------------
// internal entry points shared between translation units
// ------------------------------------------------------------------------------------------
// tc_index.cu
int tci_scratch_arm(tc_context* ctx);  // once per context (see k_bbox)
int tci_bbox(tc_context* ctx, const float* d_xyz, uint64_t n, float mn[3], float mx[3]);
// Spatially sort arbitrary points by the cells of grid `g` (clamped): returns float4
// (x,y,z,bits(orig idx)) in *d_sorted (caller frees with tc_free).
int tci_sort_by_grid(tc_context* ctx, const float* d_xyz, uint64_t n, const GridParams& g,
                     float4** d_sorted);
int tci_radix_sort_pairs(tc_context* ctx, uint32_t* d_keys, uint32_t* d_vals, uint32_t* d_keys_alt,
                         uint32_t* d_vals_alt, uint32_t n, int key_bits, bool vals_given,
                         uint32_t** keys_out, uint32_t** vals_out);
int tci_exclusive_scan_u32(tc_context* ctx, const uint32_t* d_in, uint32_t* d_out, uint64_t n);
struct tc_index;
int tci_level0_max_population(tc_context* ctx, const tc_index* ix, uint32_t* out);

// tc_search.cu
extern int g_tc_search_flags;  // default search variant (tc_debug_set_search_flags)
int tci_knn_launch(tc_context* ctx, const tc_index* index, const float4* d_queries_sorted,
                   uint64_t q_begin, uint64_t q_end, uint32_t k, int exclude_self, bool self_query,
                   uint32_t* d_idx_out, float* d_dist_out, uint32_t* d_count_out);
// exact_range: [q_begin, q_end) are exactly the owned queries (no per-cell ownership test)
int tci_normals_launch(tc_context* ctx, const tc_index* index, uint32_t k, int orient,
                       const float vp[3], uint64_t q_begin, uint64_t q_end, float* d_out_aos,
                       bool exact_range = false);
// like != nullptr: a complete index with the cell edge and cell order of an existing grid
int tci_index_build(tc_context* ctx, const tc_cloud* cloud, uint32_t k_hint, float cell_size,
                    int rank, int world, tc_index** out, const GridParams* like = nullptr);
int tci_normals_radius_launch(tc_context* ctx, const tc_index* index, float radius, uint32_t k,
                              int orient, const float vp[3], uint64_t q_begin, uint64_t q_end,
                              float* d_out_aos);
int tci_radius_search_launch(tc_context* ctx, const tc_index* index, const float q[3], float radius,
                             uint32_t* d_idx, float* d_d2, uint32_t capacity, uint32_t* d_count);

// tc_tile.cu — staged-tile kernels (shape: 16, 17, 32, 33 list slots; flags: search variant bits)
int tci_tile_normals(tc_context* ctx, const LevelSet& ls, int shape, uint32_t q_begin,
                     uint32_t q_end, uint32_t own_begin, uint32_t own_end, uint32_t k, int orient,
                     const float vp[3], float* d_out, uint32_t* d_fb_list, uint32_t* d_fb_count,
                     uint32_t* d_stats, int flags);
int tci_tile_knn(tc_context* ctx, const LevelSet& ls, int shape, const float4* d_queries,
                 uint32_t q_begin, uint32_t q_end, uint32_t k, uint32_t need, int drop_self,
                 uint32_t* d_idx, float* d_dist, uint32_t* d_count, uint32_t* d_fb_list,
                 uint32_t* d_fb_count, uint32_t* d_stats, int flags);

// tc_icp.cu — GICP building blocks (gicp.rs): per-point covariances as two float4
// {xx, xy, xz, yy | yz, zz, 0, 0} by original index, and the Gauss-Newton loop
struct tc_icp_result;
int tci_gicp_covariances(tc_context* ctx, const tc_cloud* cloud, uint32_t k, float4** d_cov);
int tci_gicp_device(tc_context* ctx, const tc_cloud* src, const tc_index* tgt,
                    const float4* d_src_cov, const float4* d_tgt_cov, const float init[7],
                    uint32_t max_iters, float max_corr_dist, float conv_threshold,
                    tc_icp_result* out, uint32_t* d_match_out);

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
// Exact (unfused, round-to-nearest) f32 ops: Rust never contracts a*b+c into an FMA, so the
// distance / covariance / transform arithmetic must not either.
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }

// KdTree::distance_squared (nearest_neighbor.rs:162-167): (dx*dx + dy*dy) + dz*dz
__device__ __forceinline__ float dist2_exact(float px, float py, float pz, float qx, float qy,
                                             float qz) {
  const float dx = xsub(px, qx), dy = xsub(py, qy), dz = xsub(pz, qz);
  return xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
}

// order-preserving float <-> uint (for atomicMin/Max on floats)
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return u ^ ((uint32_t)(-(int32_t)(u >> 31)) | 0x80000000u);
}
__host__ __device__ inline float ord2f(uint32_t o) {
  const uint32_t u = o ^ (((o >> 31) - 1u) | 0x80000000u);
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

// cell coordinate of a coordinate value along one axis; u is also returned (cell units).
__device__ __forceinline__ int cell_coord(float x, float o, float inv, int n, float& u) {
  u = xmul(xsub(x, o), inv);
  int c = __float2int_rd(u);
  c = max(0, min(n - 1, c));
  return c;
}
__device__ __forceinline__ uint32_t cell_id(const GridParams& g, int cx, int cy, int cz) {
  return g.ymajor ? (uint32_t)(((int64_t)cy * g.nz + cz) * g.nx + cx)
                  : (uint32_t)(((int64_t)cz * g.ny + cy) * g.nx + cx);
}
#endif
