// tc_search.cuh — the exact grid search shared by the kNN / normals kernels (tc_search.cu) and
// the ICP correspondence kernel (tc_icp.cu).  See tc_search.cu for the design notes.
#pragma once
#include "tc_internal.cuh"

namespace tcs {

constexpr uint64_t kEmpty = ~0ull;

template <int K>
struct TopK {
  uint64_t key[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; ++i) key[i] = kEmpty;
  }
  // precondition: c < key[K-1]
  __device__ __forceinline__ void insert(uint64_t c) {
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      const uint64_t prev = key[i - 1];
      key[i] = (c < prev) ? prev : ((c < key[i]) ? c : key[i]);
    }
    key[0] = (c < key[0]) ? c : key[0];
  }
  __device__ __forceinline__ bool full() const { return key[K - 1] != kEmpty; }
  __device__ __forceinline__ float kth() const {
    return __uint_as_float((uint32_t)(key[K - 1] >> 32));
  }
  __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t lo, uint32_t hi,
                                       float qx, float qy, float qz) {
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
      const uint64_t k2 = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c.w);
      if (k2 < key[K - 1]) insert(k2);
    }
  }
};

// Conservative lower bound on the distance from the query to any indexed point whose cell lies
// outside the block [c-R, c+R]^3.  In cell units a point beyond the +side of an axis differs by
// more than (R + 1 - f), beyond the -side by more than (f + R) (f = u - c, the query's fractional
// cell coordinate, computed with the same f32 expression as the cell assignment, which is
// monotone).  The f32 rounding of u is bounded by 2^-23 (|p - o| + |q - o|) / cell; we subtract
// twice that and shave 1e-5 relative.  Sides with no cells left contribute +inf.
__device__ __forceinline__ float axis_bound(float du, float cell, float mag) {
  return fmaxf(0.0f, du * cell * 0.99999f - mag * 4.8e-7f);
}
__device__ __forceinline__ float ring_bound(const GridParams& g, int R, int cx, int cy, int cz,
                                            float fx, float fy, float fz, float mx, float my,
                                            float mz) {
  float b = INFINITY;
  if (cx - R > 0) b = fminf(b, axis_bound(fx + (float)R, g.cell, mx));
  if (cx + R < g.nx - 1) b = fminf(b, axis_bound((float)(R + 1) - fx, g.cell, mx));
  if (cy - R > 0) b = fminf(b, axis_bound(fy + (float)R, g.cell, my));
  if (cy + R < g.ny - 1) b = fminf(b, axis_bound((float)(R + 1) - fy, g.cell, my));
  if (cz - R > 0) b = fminf(b, axis_bound(fz + (float)R, g.cell, mz));
  if (cz + R < g.nz - 1) b = fminf(b, axis_bound((float)(R + 1) - fz, g.cell, mz));
  return b;
}

// 1-NN accumulator (ICP correspondences): best (d2, original index) key plus its sorted position.
struct Best1 {
  uint64_t key;
  uint32_t pos;
  __device__ __forceinline__ void init() {
    key = kEmpty;
    pos = 0;
  }
  __device__ __forceinline__ bool full() const { return key != kEmpty; }
  __device__ __forceinline__ float kth() const { return __uint_as_float((uint32_t)(key >> 32)); }
  __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t lo, uint32_t hi,
                                       float qx, float qy, float qz) {
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
      const uint64_t k2 = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c.w);
      if (k2 < key) {
        key = k2;
        pos = j;
      }
    }
  }
};

// Exact nearest-neighbour search for one query; Acc is TopK<K> or Best1.
template <class Acc>
__device__ __forceinline__ void grid_search(const GridParams& g, const float4* __restrict__ pts,
                                            const uint32_t* __restrict__ cell_start, float qx,
                                            float qy, float qz, Acc& tk) {
  tk.init();
  float ux, uy, uz;
  const int cx = cell_coord(qx, g.ox, g.inv, g.nx, ux);
  const int cy = cell_coord(qy, g.oy, g.inv, g.ny, uy);
  const int cz = cell_coord(qz, g.oz, g.inv, g.nz, uz);
  const float fx = ux - (float)cx, fy = uy - (float)cy, fz = uz - (float)cz;
  const float mx = g.ex + fabsf(qx - g.ox), my = g.ey + fabsf(qy - g.oy),
              mz = g.ez + fabsf(qz - g.oz);

  // rings 0 and 1: the 3x3 rows around the query's row, x-range [cx-1, cx+1] (contiguous)
  {
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    for (int dz = -1; dz <= 1; ++dz) {
      const int z = cz + dz;
      if (z < 0 || z >= g.nz) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = cy + dy;
        if (y < 0 || y >= g.ny) continue;
        const uint32_t row = cell_id(g, 0, y, z);
        const uint32_t lo = __ldg(&cell_start[row + x0]);
        const uint32_t hi = __ldg(&cell_start[row + x1 + 1]);
        tk.scan(pts, lo, hi, qx, qy, qz);
      }
    }
  }
  int R = 1;
  while (true) {
    const float b = ring_bound(g, R, cx, cy, cz, fx, fy, fz, mx, my, mz);
    if (b == INFINITY) break;  // the searched block already covers the whole grid
    if (tk.full()) {
      if (tk.kth() < b * b * 0.99999f) break;
    }
    ++R;
    // shell R: full x-range on rows with max(|dy|,|dz|) == R, the two end cells elsewhere
    const int zlo = max(cz - R, 0), zhi = min(cz + R, g.nz - 1);
    const int ylo = max(cy - R, 0), yhi = min(cy + R, g.ny - 1);
    const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
    for (int z = zlo; z <= zhi; ++z) {
      for (int y = ylo; y <= yhi; ++y) {
        const uint32_t row = cell_id(g, 0, y, z);
        const bool outer = (abs(z - cz) == R) || (abs(y - cy) == R);
        if (outer) {
          const uint32_t lo = __ldg(&cell_start[row + x0]);
          const uint32_t hi = __ldg(&cell_start[row + x1 + 1]);
          tk.scan(pts, lo, hi, qx, qy, qz);
        } else {
          if (cx - R >= 0) {
            const uint32_t lo = __ldg(&cell_start[row + cx - R]);
            const uint32_t hi = __ldg(&cell_start[row + cx - R + 1]);
            tk.scan(pts, lo, hi, qx, qy, qz);
          }
          if (cx + R <= g.nx - 1) {
            const uint32_t lo = __ldg(&cell_start[row + cx + R]);
            const uint32_t hi = __ldg(&cell_start[row + cx + R + 1]);
            tk.scan(pts, lo, hi, qx, qy, qz);
          }
        }
      }
    }
  }
}


}  // namespace tcs
