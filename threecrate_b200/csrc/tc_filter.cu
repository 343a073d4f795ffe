// tc_filter.cu — the filters either side of the kNN -> normals -> ICP path:
//   tc_voxel_grid_filter             voxel_grid_filter            (filtering.rs:38-133)
//   tc_radius_outlier_removal        radius_outlier_removal       (filtering.rs:167-218)
//   tc_statistical_outlier_removal   statistical_outlier_removal  (filtering.rs:253-321) and
//                                    .._with_threshold            (filtering.rs:335-394)
// All three return a new device-resident cloud.  The outlier filters keep the surviving points in
// their original order (as the reference does); the voxel filter emits one centroid per voxel in
// ascending (z, y, x) voxel order (the reference's order is a HashMap iteration order, i.e.
// arbitrary), each centroid summed in f64 in original point order exactly like the reference.
#include <cmath>

#include "tc_internal.cuh"
#include "tc_search.cuh"

using namespace tcs;

namespace {

constexpr int kThreads = 256;

inline int blocks_for(tc_context* ctx, uint64_t n, int waves = 8) {
  uint64_t b = (n + kThreads - 1) / kThreads;
  const uint64_t cap = (uint64_t)ctx->sm_count * waves;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

int empty_cloud(tc_context* ctx, tc_cloud** out) {
  tc_cloud* c = new tc_cloud();
  c->ctx = ctx;
  c->n = 0;
  const int st = tc_alloc(ctx, &c->d_xyz, 1);
  if (st != TC_OK) {
    delete c;
    return st;
  }
  *out = c;
  return TC_OK;
}

// ------------------------------------------------------------------------- stable compaction
__global__ void __launch_bounds__(kThreads) k_compact(const float* __restrict__ xyz,
                                                      const uint32_t* __restrict__ keep,
                                                      const uint32_t* __restrict__ pos, uint32_t n,
                                                      float* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (keep[i]) {
      const uint64_t o = 3 * (uint64_t)pos[i];
      out[o + 0] = xyz[3 * (uint64_t)i + 0];
      out[o + 1] = xyz[3 * (uint64_t)i + 1];
      out[o + 2] = xyz[3 * (uint64_t)i + 2];
    }
  }
}

// keep[i] in {0,1}, original order -> new cloud holding the kept points in the same order
int compact_cloud(tc_context* ctx, const tc_cloud* cloud, const uint32_t* d_keep, tc_cloud** out) {
  const uint64_t n = cloud->n;
  uint32_t* d_pos = nullptr;
  TC_TRY(tc_alloc(ctx, &d_pos, n + 1));
  int st = tci_exclusive_scan_u32(ctx, d_keep, d_pos, n);
  uint32_t kept = 0;
  if (st == TC_OK &&
      (cudaMemcpyAsync(ctx->h_scratch + 32, d_pos + n, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                       ctx->stream) != cudaSuccess ||
       cudaStreamSynchronize(ctx->stream) != cudaSuccess))
    st = tc_fail(ctx, TC_GPU, "compaction count read-back failed");
  if (st == TC_OK) kept = ctx->h_scratch[32];
  tc_cloud* c = nullptr;
  if (st == TC_OK) {
    c = new tc_cloud();
    c->ctx = ctx;
    c->n = kept;
    st = tc_alloc(ctx, &c->d_xyz, 3 * (uint64_t)kept);
  }
  if (st == TC_OK && kept > 0) {
    k_compact<<<blocks_for(ctx, n), kThreads, 0, ctx->stream>>>(cloud->d_xyz, d_keep, d_pos,
                                                                (uint32_t)n, c->d_xyz);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "compaction launch failed");
  }
  tc_free(ctx, d_pos);
  if (st != TC_OK) {
    if (c) tc_cloud_free(c);
    return st;
  }
  *out = c;
  return TC_OK;
}

// -------------------------------------------------------------------------------- voxel grid
// voxel coordinate = floor((p - min) / voxel_size) as i32 (filtering.rs:95-100): f32 subtract and
// IEEE divide, then a saturating cast like Rust's `as i32`
__device__ __forceinline__ uint32_t voxel_coord(float v, float mn, float vs) {
  const float q = floorf(xdiv(xsub(v, mn), vs));
  return (uint32_t)max(__float2int_rz(q), 0);  // cvt saturates; (p - min) >= 0 by construction
}
float host_voxel_coord(float v, float mn, float vs) {
  volatile float d = v - mn;
  volatile float q = d / vs;
  const float f = std::floor(q);
  if (!(f < 2147483648.0f)) return 2147483647.0f;
  return f < 0.0f ? 0.0f : f;
}

// mode 0: the three coordinates; mode 1: also the packed key (z << (bx+by) | y << bx | x)
__global__ void __launch_bounds__(kThreads) k_voxel_coords(const float* __restrict__ xyz, uint32_t n,
                                                           float mnx, float mny, float mnz, float vs,
                                                           int bx, int by, int packed,
                                                           uint32_t* __restrict__ cx,
                                                           uint32_t* __restrict__ cy,
                                                           uint32_t* __restrict__ cz,
                                                           uint32_t* __restrict__ key) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t x = voxel_coord(xyz[3 * (uint64_t)i + 0], mnx, vs);
    const uint32_t y = voxel_coord(xyz[3 * (uint64_t)i + 1], mny, vs);
    const uint32_t z = voxel_coord(xyz[3 * (uint64_t)i + 2], mnz, vs);
    cx[i] = x;
    cy[i] = y;
    cz[i] = z;
    if (packed) key[i] = (z << (bx + by)) | (y << bx) | x;
    else key[i] = x;
  }
}

__global__ void __launch_bounds__(kThreads) k_gather_u32(const uint32_t* __restrict__ src,
                                                         const uint32_t* __restrict__ perm,
                                                         uint32_t n, uint32_t* __restrict__ dst) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dst[i] = src[perm[i]];
}

// head[i] = 1 when sorted position i starts a new voxel
__global__ void __launch_bounds__(kThreads) k_voxel_heads(const uint32_t* __restrict__ perm,
                                                          const uint32_t* __restrict__ cx,
                                                          const uint32_t* __restrict__ cy,
                                                          const uint32_t* __restrict__ cz, uint32_t n,
                                                          uint32_t* __restrict__ head) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t h = 1;
    if (i > 0) {
      const uint32_t a = perm[i], b = perm[i - 1];
      h = (cx[a] != cx[b] || cy[a] != cy[b] || cz[a] != cz[b]) ? 1u : 0u;
    }
    head[i] = h;
  }
}

__global__ void __launch_bounds__(kThreads) k_voxel_starts(const uint32_t* __restrict__ head,
                                                           const uint32_t* __restrict__ seg,
                                                           uint32_t n, uint32_t* __restrict__ starts) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (head[i]) starts[seg[i]] = i;
}

// One thread per voxel: f64 sums in original point order (the sort is stable), then
// (sum * (1 / count)) as f32 — the reference's arithmetic (filtering.rs:108-130).
__global__ void __launch_bounds__(kThreads) k_voxel_centroids(const float* __restrict__ xyz,
                                                              const uint32_t* __restrict__ perm,
                                                              const uint32_t* __restrict__ starts,
                                                              uint32_t n_vox, uint32_t n,
                                                              float* __restrict__ out) {
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < n_vox; v += gridDim.x * blockDim.x) {
    const uint32_t a = starts[v], b = (v + 1 < n_vox) ? starts[v + 1] : n;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (uint32_t j = a; j < b; ++j) {
      const float* p = xyz + 3 * (uint64_t)perm[j];
      sx += (double)p[0];
      sy += (double)p[1];
      sz += (double)p[2];
    }
    const double inv = 1.0 / (double)(b - a);
    out[3 * (uint64_t)v + 0] = (float)(sx * inv);
    out[3 * (uint64_t)v + 1] = (float)(sy * inv);
    out[3 * (uint64_t)v + 2] = (float)(sz * inv);
  }
}

int bits_for(double count) {  // bits needed for values 0 .. count-1
  int b = 1;
  while (b < 32 && std::ldexp(1.0, b) < count) ++b;
  return b;
}

// ---------------------------------------------------------------------- radius outlier count
// keep[i] = (|{j : d2(i,j) <= r^2}| - 1 >= min_neighbors); the scan stops as soon as that is
// known (nearest_neighbor.rs:254-298 counts every hit, including the point itself and duplicates)
__global__ void __launch_bounds__(128) k_radius_keep(LevelSet ls, int level, uint32_t n, float radius,
                                                     uint32_t min_neighbors,
                                                     uint32_t* __restrict__ keep) {
  const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= n) return;
  const float4 q = __ldg(&ls.pts[0][qi]);
  const float4* __restrict__ pts = ls.pts[level];
  const float r2 = xmul(radius, radius);  // nearest_neighbor.rs:259
  const uint32_t want = min_neighbors + 1u;
  uint32_t cnt = 0;
  box_visit(ls.g[level], ls.cs[level], q.x, q.y, q.z, r2, [&](uint32_t lo, uint32_t hi) {
    if (cnt >= want) return;
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      cnt += (dist2_exact(c.x, c.y, c.z, q.x, q.y, q.z) <= r2) ? 1u : 0u;
    }
  });
  keep[__float_as_uint(q.w)] = (cnt >= want) ? 1u : 0u;
}

// ------------------------------------------------------------------------ statistical outlier
// mean distance to the k+1 nearest neighbours, skipping every neighbour whose coordinates equal
// the point's (filtering.rs:288-300): f32 sum in ascending-distance order, then one divide
__global__ void __launch_bounds__(kThreads) k_sor_mean(const float* __restrict__ xyz, uint32_t n,
                                                       uint32_t k1, const uint32_t* __restrict__ idx,
                                                       const float* __restrict__ dist,
                                                       float* __restrict__ mean) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float px = xyz[3 * (uint64_t)i], py = xyz[3 * (uint64_t)i + 1], pz = xyz[3 * (uint64_t)i + 2];
    float sum = 0.0f;
    uint32_t cnt = 0;
    for (uint32_t j = 0; j < k1; ++j) {
      const uint32_t id = idx[(uint64_t)i * k1 + j];
      if (id == TC_NO_INDEX) break;
      const float* p = xyz + 3 * (uint64_t)id;
      if (p[0] == px && p[1] == py && p[2] == pz) continue;
      sum = xadd(sum, dist[(uint64_t)i * k1 + j]);
      ++cnt;
    }
    mean[i] = cnt ? xdiv(sum, (float)cnt) : 0.0f;
  }
}

// Global mean and variance of mean[] in the reference's arithmetic: SEQUENTIAL f32 sums over the
// cloud (filtering.rs:304-312).  A float sum is not associative, so the only bit-exact way is
// to add in the same order: the block stages 4096 values at a time in shared memory (coalesced)
// and one thread runs the dependent chain of adds over them (~4 cycles per point per pass);
// `fast` mode below replaces it with an f64 tree reduction.
constexpr int kSeqChunk = 4096;
__global__ void __launch_bounds__(256) k_sor_stats_sequential(const float* __restrict__ mean,
                                                              uint32_t n, float std_mult,
                                                              float* __restrict__ out /*mean, std, thr*/) {
  __shared__ float buf[kSeqChunk];
  __shared__ float s_gm;
  float s = 0.0f;
  for (int pass = 0; pass < 2; ++pass) {
    const float gm = pass ? s_gm : 0.0f;
    s = 0.0f;
    for (uint32_t base = 0; base < n; base += kSeqChunk) {
      const uint32_t m = min((uint32_t)kSeqChunk, n - base);
      for (uint32_t t = threadIdx.x; t < m; t += blockDim.x) {
        const float v = mean[base + t];
        const float d = xsub(v, gm);
        buf[t] = pass ? xmul(d, d) : v;  // powi(2) = d * d
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (; t + 16 <= m; t += 16) {
#pragma unroll
          for (int u = 0; u < 16; ++u) s = xadd(s, buf[t + u]);
        }
        for (; t < m; ++t) s = xadd(s, buf[t]);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (pass == 0) {
        s_gm = xdiv(s, (float)n);
      } else {
        const float sd = xsqrt(xdiv(s, (float)n));
        out[0] = s_gm;
        out[1] = sd;
        out[2] = xadd(s_gm, xmul(std_mult, sd));
      }
    }
    __syncthreads();
  }
}

// fast mode: f64 sums (two passes), any order
__global__ void __launch_bounds__(kThreads) k_sor_sum(const float* __restrict__ mean, uint32_t n,
                                                      const double* __restrict__ centre,
                                                      double* __restrict__ acc) {
  const double c = centre ? centre[0] / (double)n : 0.0;
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = (double)mean[i] - c;
    s += centre ? d * d : d;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}
__global__ void k_sor_stats_fast(const double* __restrict__ acc, uint32_t n, float std_mult,
                                 float* __restrict__ out) {
  const double gm = acc[0] / (double)n;
  const double sd = sqrt(acc[1] / (double)n);
  out[0] = (float)gm;
  out[1] = (float)sd;
  out[2] = (float)(gm + (double)std_mult * sd);
}

__global__ void __launch_bounds__(kThreads) k_sor_keep(const float* __restrict__ mean, uint32_t n,
                                                       const float* __restrict__ thr_dev,
                                                       float thr_host, uint32_t* __restrict__ keep) {
  const float thr = thr_dev ? thr_dev[2] : thr_host;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    keep[i] = (mean[i] <= thr) ? 1u : 0u;  // filtering.rs:317
}

}  // namespace

// ==========================================================================================
extern "C" int tc_voxel_grid_filter(tc_context* ctx, const tc_cloud* cloud, float voxel_size,
                                    tc_cloud** out) {
  if (!ctx || !cloud || !out) return TC_INVALID_DATA;
  *out = nullptr;
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (cloud->n == 0) return empty_cloud(ctx, out);  // filtering.rs:42-44
  if (voxel_size <= 0.0f) return tc_fail(ctx, TC_INVALID_DATA, "voxel_size must be positive");
  const uint32_t n = (uint32_t)cloud->n;
  float mn[3], mx[3];
  TC_TRY(tci_bbox(ctx, cloud->d_xyz, n, mn, mx));
  // largest coordinate per axis (monotone in p, so it is the one of the bbox maximum)
  double dim[3];
  int bits[3];
  for (int a = 0; a < 3; ++a) {
    dim[a] = (double)host_voxel_coord(mx[a], mn[a], voxel_size) + 1.0;
    bits[a] = bits_for(dim[a]);
  }
  const bool packed = bits[0] + bits[1] + bits[2] <= 32;

  uint32_t *d_c[3] = {nullptr, nullptr, nullptr}, *d_key = nullptr, *d_key2 = nullptr,
           *d_val = nullptr, *d_val2 = nullptr, *d_head = nullptr, *d_seg = nullptr,
           *d_starts = nullptr;
  uint32_t *ks = nullptr, *perm = nullptr;
  tc_cloud* c = nullptr;
  int st = TC_OK;
  for (int a = 0; a < 3 && st == TC_OK; ++a) st = tc_alloc(ctx, &d_c[a], n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_key, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_key2, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_val, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_val2, n);
  const int grid = blocks_for(ctx, n);
  if (st == TC_OK) {
    k_voxel_coords<<<grid, kThreads, 0, ctx->stream>>>(cloud->d_xyz, n, mn[0], mn[1], mn[2],
                                                       voxel_size, bits[0], bits[1], packed ? 1 : 0,
                                                       d_c[0], d_c[1], d_c[2], d_key);
    ctx->launches++;
    // stable LSD sort: one pass set over the packed key, or x then y then z
    st = tci_radix_sort_pairs(ctx, d_key, d_val, d_key2, d_val2, n,
                              packed ? bits[0] + bits[1] + bits[2] : bits[0], false, &ks, &perm);
    for (int a = 1; a < 3 && st == TC_OK && !packed; ++a) {
      // (ks, perm) live in one buffer of each ping-pong pair; gather the next key into the other
      uint32_t* kfree = (ks == d_key) ? d_key2 : d_key;
      uint32_t* vfree = (perm == d_val) ? d_val2 : d_val;
      k_gather_u32<<<grid, kThreads, 0, ctx->stream>>>(d_c[a], perm, n, kfree);
      ctx->launches++;
      // sort (kfree, perm) with perm as the given values: keys ping-pong kfree <-> ks
      st = tci_radix_sort_pairs(ctx, kfree, perm, ks, vfree, n, bits[a], true, &ks, &perm);
    }
  }
  if (st == TC_OK) st = tc_alloc(ctx, &d_head, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_seg, (uint64_t)n + 1);
  uint32_t n_vox = 0;
  if (st == TC_OK) {
    k_voxel_heads<<<grid, kThreads, 0, ctx->stream>>>(perm, d_c[0], d_c[1], d_c[2], n, d_head);
    ctx->launches++;
    st = tci_exclusive_scan_u32(ctx, d_head, d_seg, n);
  }
  if (st == TC_OK &&
      (cudaMemcpyAsync(ctx->h_scratch + 32, d_seg + n, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                       ctx->stream) != cudaSuccess ||
       cudaStreamSynchronize(ctx->stream) != cudaSuccess))
    st = tc_fail(ctx, TC_GPU, "voxel count read-back failed");
  if (st == TC_OK) {
    n_vox = ctx->h_scratch[32];
    st = tc_alloc(ctx, &d_starts, n_vox);
  }
  if (st == TC_OK) {
    c = new tc_cloud();
    c->ctx = ctx;
    c->n = n_vox;
    st = tc_alloc(ctx, &c->d_xyz, 3 * (uint64_t)n_vox);
  }
  if (st == TC_OK) {
    k_voxel_starts<<<grid, kThreads, 0, ctx->stream>>>(d_head, d_seg, n, d_starts);
    k_voxel_centroids<<<blocks_for(ctx, n_vox), kThreads, 0, ctx->stream>>>(
        cloud->d_xyz, perm, d_starts, n_vox, n, c->d_xyz);
    ctx->launches += 2;
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "voxel filter launch failed");
  }
  for (int a = 0; a < 3; ++a) tc_free(ctx, d_c[a]);
  tc_free(ctx, d_key);
  tc_free(ctx, d_key2);
  tc_free(ctx, d_val);
  tc_free(ctx, d_val2);
  tc_free(ctx, d_head);
  tc_free(ctx, d_seg);
  tc_free(ctx, d_starts);
  if (st != TC_OK) {
    if (c) tc_cloud_free(c);
    return st;
  }
  *out = c;
  return TC_OK;
}

extern "C" int tc_radius_outlier_removal(tc_context* ctx, const tc_cloud* cloud, float radius,
                                         uint32_t min_neighbors, tc_cloud** out) {
  if (!ctx || !cloud || !out) return TC_INVALID_DATA;
  *out = nullptr;
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (cloud->n == 0) return empty_cloud(ctx, out);  // filtering.rs:172-174
  if (radius <= 0.0f) return tc_fail(ctx, TC_INVALID_DATA, "radius must be positive");
  if (min_neighbors == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "min_neighbors must be greater than 0");
  const uint32_t n = (uint32_t)cloud->n;
  tc_index* ix = nullptr;
  // cells of one radius: the box of a query spans at most 3 cells per axis
  TC_TRY(tc_index_build(ctx, cloud, min_neighbors, radius, &ix));
  uint32_t* d_keep = nullptr;
  int st = tc_alloc(ctx, &d_keep, n);
  if (st == TC_OK) {
    const LevelSet ls = ix->level_set(0);
    k_radius_keep<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ls, 0, n, radius, min_neighbors, d_keep);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "radius filter launch failed");
  }
  if (st == TC_OK) st = compact_cloud(ctx, cloud, d_keep, out);
  tc_free(ctx, d_keep);
  tc_index_free(ix);
  return st;
}

// mode: 0 = threshold from the cloud's statistics in the reference's sequential f32 arithmetic
//           (bit-exact, ~2 ns/point serial tail), 1 = same statistics from f64 tree sums (fast;
//           points whose mean distance is within f32 rounding of the threshold may differ),
//       2 = `value` is the threshold itself (statistical_outlier_removal_with_threshold)
extern "C" int tc_statistical_outlier_removal(tc_context* ctx, const tc_cloud* cloud,
                                              uint32_t k_neighbors, float value, int mode,
                                              float* stats_out, tc_cloud** out) {
  if (!ctx || !cloud || !out) return TC_INVALID_DATA;
  *out = nullptr;
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (mode < 0 || mode > 2) return tc_fail(ctx, TC_INVALID_DATA, "mode must be 0, 1 or 2");
  if (cloud->n == 0) return empty_cloud(ctx, out);  // filtering.rs:258-260
  if (k_neighbors == 0) return tc_fail(ctx, TC_INVALID_DATA, "k_neighbors must be greater than 0");
  if (value <= 0.0f)
    return tc_fail(ctx, TC_INVALID_DATA,
                   mode == 2 ? "threshold must be positive" : "std_dev_multiplier must be positive");
  const uint32_t n = (uint32_t)cloud->n;
  const uint32_t k1 = k_neighbors + 1;
  tc_index* ix = nullptr;
  TC_TRY(tc_index_build(ctx, cloud, k1, 0.0f, &ix));
  uint32_t *d_idx = nullptr, *d_keep = nullptr;
  float *d_dist = nullptr, *d_mean = nullptr, *d_stats = nullptr;
  double* d_acc = nullptr;
  int st = tc_alloc(ctx, &d_idx, (uint64_t)n * k1);
  if (st == TC_OK) st = tc_alloc(ctx, &d_dist, (uint64_t)n * k1);
  if (st == TC_OK) st = tc_alloc(ctx, &d_mean, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_keep, n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_stats, 4);
  if (st == TC_OK) st = tc_alloc(ctx, &d_acc, 2);
  if (st == TC_OK)
    st = tci_knn_launch(ctx, ix, nullptr, 0, n, k1, 0, true, d_idx, d_dist, nullptr);
  if (st == TC_OK) {
    const int grid = blocks_for(ctx, n);
    k_sor_mean<<<grid, kThreads, 0, ctx->stream>>>(cloud->d_xyz, n, k1, d_idx, d_dist, d_mean);
    ctx->launches++;
    if (mode == 0) {
      k_sor_stats_sequential<<<1, 256, 0, ctx->stream>>>(d_mean, n, value, d_stats);
      ctx->launches++;
    } else if (mode == 1) {
      cudaMemsetAsync(d_acc, 0, 2 * sizeof(double), ctx->stream);
      k_sor_sum<<<grid, kThreads, 0, ctx->stream>>>(d_mean, n, nullptr, d_acc);
      k_sor_sum<<<grid, kThreads, 0, ctx->stream>>>(d_mean, n, d_acc, d_acc + 1);
      k_sor_stats_fast<<<1, 1, 0, ctx->stream>>>(d_acc, n, value, d_stats);
      ctx->launches += 3;
    }
    k_sor_keep<<<grid, kThreads, 0, ctx->stream>>>(d_mean, n, mode == 2 ? nullptr : d_stats, value,
                                                   d_keep);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "outlier filter launch failed");
  }
  if (st == TC_OK) st = compact_cloud(ctx, cloud, d_keep, out);  // (synchronises the stream)
  if (st == TC_OK && stats_out) {
    if (mode == 2) {
      stats_out[0] = stats_out[1] = 0.0f;
      stats_out[2] = value;
    } else if (cudaMemcpyAsync(stats_out, d_stats, 3 * sizeof(float), cudaMemcpyDeviceToHost,
                               ctx->stream) != cudaSuccess ||
               cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      st = tc_fail(ctx, TC_GPU, "statistics read-back failed");
    }
  }
  tc_free(ctx, d_idx);
  tc_free(ctx, d_dist);
  tc_free(ctx, d_mean);
  tc_free(ctx, d_keep);
  tc_free(ctx, d_stats);
  tc_free(ctx, d_acc);
  tc_index_free(ix);
  if (st != TC_OK && *out) {
    tc_cloud_free(*out);
    *out = nullptr;
  }
  return st;
}
