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
  // Candidate scan.  flags bit 0 selects the deferred form: pass 1 only evaluates distances and
  // records, per lane, which of up to 32 candidates beat the current K-th distance (one bit
  // each); pass 2 walks the set bits and runs the insertion chain.  In the direct form the warp
  // pays for the chain whenever ANY lane inserts (nearly every candidate); deferred, the chain
  // runs max-over-lanes(popcount) times per 32 candidates with most lanes active.
  __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t lo, uint32_t hi,
                                       float qx, float qy, float qz, int flags) {
    if (!(flags & 1)) {
      for (uint32_t j = lo; j < hi; ++j) {
        const float4 c = __ldg(&pts[j]);
        const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
        const uint64_t k2 = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c.w);
        if (k2 < key[K - 1]) insert(k2);
      }
      return;
    }
    for (uint32_t base = lo; base < hi; base += 32) {
      const uint32_t len = min(32u, hi - base);
      const float tau = full() ? kth() : INFINITY;
      uint32_t mask = 0;
      for (uint32_t t = 0; t < len; ++t) {
        const float4 c = __ldg(&pts[base + t]);
        const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
        if (d2 <= tau) mask |= (1u << t);  // <=: equal d2 may still win on the index
      }
      while (mask) {
        const uint32_t t = __ffs(mask) - 1;
        mask &= mask - 1;
        const float4 c = __ldg(&pts[base + t]);
        const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
        const uint64_t k2 = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c.w);
        if (k2 < key[K - 1]) insert(k2);
      }
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
                                       float qx, float qy, float qz, int) {
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

// Conservative lower bound (squared) on the distance from the query to any point in row
// (y, z): per axis the gap, in cell units, between the query and the nearest face of that row.
__device__ __forceinline__ float row_gap(int d, float f) {
  // d = row coordinate - query cell coordinate along the axis; f = fractional cell coordinate
  return d == 0 ? 0.0f : (d < 0 ? (f + (float)(-d - 1)) : ((float)d - f));
}

// Exact nearest-neighbour search for one query; Acc is TopK<K> or Best1.
// flags: bit 0 deferred insertion (TopK), bit 1 prune rows that cannot beat the K-th distance.
template <class Acc>
__device__ __forceinline__ void grid_search(const GridParams& g, const float4* __restrict__ pts,
                                            const uint32_t* __restrict__ cell_start, float qx,
                                            float qy, float qz, Acc& tk) {
  tk.init();
  const int flags = g.flags;
  float ux, uy, uz;
  const int cx = cell_coord(qx, g.ox, g.inv, g.nx, ux);
  const int cy = cell_coord(qy, g.oy, g.inv, g.ny, uy);
  const int cz = cell_coord(qz, g.oz, g.inv, g.nz, uz);
  const float fx = ux - (float)cx, fy = uy - (float)cy, fz = uz - (float)cz;
  const float mx = g.ex + fabsf(qx - g.ox), my = g.ey + fabsf(qy - g.oy),
              mz = g.ez + fabsf(qz - g.oz);

  auto scan_row = [&](int y, int z, int xa, int xb) {
    if ((flags & 2) && tk.full()) {
      const float by = axis_bound(row_gap(y - cy, fy), g.cell, my);
      const float bz = axis_bound(row_gap(z - cz, fz), g.cell, mz);
      if ((by * by + bz * bz) * 0.99999f > tk.kth()) return;  // every point here is farther
    }
    const uint32_t row = cell_id(g, 0, y, z);
    const uint32_t lo = __ldg(&cell_start[row + xa]);
    const uint32_t hi = __ldg(&cell_start[row + xb + 1]);
    tk.scan(pts, lo, hi, qx, qy, qz, flags);
  };

  // rings 0 and 1: the 3x3 rows around the query's row, x-range [cx-1, cx+1] (contiguous);
  // the query's own row goes first so the K-th distance tightens before the others are tested
  {
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    scan_row(cy, cz, x0, x1);
    for (int dz = -1; dz <= 1; ++dz) {
      const int z = cz + dz;
      if (z < 0 || z >= g.nz) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = cy + dy;
        if (y < 0 || y >= g.ny || (dy == 0 && dz == 0)) continue;
        scan_row(y, z, x0, x1);
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
        const bool outer = (abs(z - cz) == R) || (abs(y - cy) == R);
        if (outer) {
          scan_row(y, z, x0, x1);
        } else {
          if (cx - R >= 0) scan_row(y, z, cx - R, cx - R);
          if (cx + R <= g.nx - 1) scan_row(y, z, cx + R, cx + R);
        }
      }
    }
  }
}

}  // namespace tcs
