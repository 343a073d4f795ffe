// tc_search.cuh — the exact grid search shared by the kNN / normals kernels (tc_search.cu) and
// the ICP correspondence kernel (tc_icp.cu).  See tc_search.cu for the design notes.
#pragma once
#include <type_traits>
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

// ------------------------------------------------------------------------------------------
// Register sorting networks (all indices compile-time, so arrays stay in registers).
// A float compare-exchange is two ALU ops (FMNMX min/max) and branch-free.
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void bitonic_sort_f(float (&a)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const float lo = fminf(a[i], a[l]), hi = fmaxf(a[i], a[l]);
          if ((i & k) == 0) {
            a[i] = lo;
            a[l] = hi;
          } else {
            a[i] = hi;
            a[l] = lo;
          }
        }
      }
    }
  }
}
// a is bitonic on entry, ascending on exit
template <int N>
__device__ __forceinline__ void bitonic_merge_f(float (&a)[N]) {
#pragma unroll
  for (int j = N >> 1; j > 0; j >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int l = i ^ j;
      if (l > i) {
        const float lo = fminf(a[i], a[l]), hi = fmaxf(a[i], a[l]);
        a[i] = lo;
        a[l] = hi;
      }
    }
  }
}
template <int N>
__device__ __forceinline__ void bitonic_sort_u64(uint64_t (&a)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const bool sw = ((i & k) == 0) ? (a[l] < a[i]) : (a[i] < a[l]);
          const uint64_t x = a[i], y = a[l];
          a[i] = sw ? y : x;
          a[l] = sw ? x : y;
        }
      }
    }
  }
}

// Pass-1 accumulator of the two-pass selection: the smallest squared distances seen so far,
// ascending, indices NOT carried.  The list always has L entries; to select the `need`-th
// smallest with need <= L, the front is pre-filled with (L - need) sentinels of -1 (below any
// d2), which never move: the real values live in v[L-need .. L-1] and the K-th smallest is the
// STATIC register v[L-1] (a runtime index into v[] would be lowered to select chains).
// Candidates are taken L at a time; a batch holding something below the current K-th distance
// is sorted and merged:  m[i] = min(v[i], b[L-1-i]) keeps the L smallest of the 2L and is
// bitonic -> log2(L) merge stages.  A float compare-exchange is 2 FMNMX, branch-free.
template <int L, bool X = false, int B = (L < 16 ? L : 16)>
struct SelF {
  // X adds one scalar slot `x` ranked after v[L-1] (the (L+1)-th smallest): k = 16 needs 17
  // entries, and a 16-wide network plus one min-reduction is far cheaper than the 32-wide one.
  // B = candidates per batch (<= L).  B = 16 with L = 32 costs the same compare-exchanges per
  // candidate as B = 32 (sort16 + merge32 per 16 vs sort32 + merge32 per 32) in half the code
  // and with 16 fewer live registers, and the skip test below is finer grained.
  static_assert(B <= L, "batch larger than the list");
  static constexpr int kSlots = L + (X ? 1 : 0);
  float v[L];
  float x;
  int pad;  // kSlots - need
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < L; ++i) v[i] = (i < pad) ? -1.0f : INFINITY;
    x = INFINITY;
  }
  __device__ __forceinline__ float kth() const { return X ? x : v[L - 1]; }
  __device__ __forceinline__ bool full() const { return kth() < INFINITY; }
  __device__ __forceinline__ void merge(float (&b)[B]) {
    bitonic_sort_f<B>(b);
    // v (ascending) vs b padded with +inf (ascending): m[i] = min(v[i], b'[L-1-i]) keeps the L
    // smallest and is bitonic; only the top B entries of v meet a finite partner
    if (X) {
      // the discarded values are max(v[L-B+i], b[B-1-i]); their minimum is the (L+1)-th smallest
      // of v u b, and every kept value is <= v[L-1] <= x, so x only competes with that minimum
      float dm = INFINITY;
#pragma unroll
      for (int i = 0; i < B; ++i) dm = fminf(dm, fmaxf(v[L - B + i], b[B - 1 - i]));
      x = fminf(x, dm);
    }
#pragma unroll
    for (int i = 0; i < B; ++i) v[L - B + i] = fminf(v[L - B + i], b[B - 1 - i]);
    bitonic_merge_f<L>(v);
  }
  // number of list entries strictly below d2 (sentinels included)
  __device__ __forceinline__ int count_below(float d2) const {
    int r = 0;
#pragma unroll
    for (int i = 0; i < L; ++i) r += (v[i] < d2) ? 1 : 0;
    if (X) r += (x < d2) ? 1 : 0;
    return r;
  }
  // true when two real entries are bit-equal
  __device__ __forceinline__ bool has_equal() const {
    bool t = false;
#pragma unroll
    for (int i = 0; i + 1 < L; ++i)
      if (v[i] == v[i + 1] && v[i] >= 0.0f && v[i] < INFINITY) t = true;
    if (X && v[L - 1] == x && x < INFINITY) t = true;
    return t;
  }
  __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t lo, uint32_t hi,
                                       float qx, float qy, float qz, int) {
    if (lo >= hi) return;
    const uint32_t last = hi - 1;
#pragma unroll 1
    for (uint32_t base = lo; base < hi; base += B) {
      float b[B];
      float bm = INFINITY;
#pragma unroll
      for (int t = 0; t < B; ++t) {
        // unconditional load from a clamped index: the loads of a batch are independent and
        // can all be in flight together; out-of-range slots are masked to +inf afterwards
        const uint32_t j = min(base + t, last);
        const float4 c = __ldg(&pts[j]);
        const float d = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
        b[t] = (base + t <= last) ? d : INFINITY;
        bm = fminf(bm, b[t]);
      }
      if (bm < kth()) merge(b);  // strict: a candidate equal to the K-th distance cannot lower it
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
// `seeded`: key/pos were preset from a known candidate (the previous ICP iteration's match).
struct Best1 {
  static constexpr bool kOutside = true;  // ICP sources need not lie inside the target's grid
  uint64_t key;
  uint32_t pos;
  bool seeded = false;
  // smallest d2 among the scanned candidates OTHER than the current best (a candidate bit-equal
  // to the best in d2 and index - the seed met again - does not count; an exact duplicate with
  // another index does).  With the radius of the fully searched ball it bounds the distance to
  // every other point from below, which lets the next ICP iteration keep the match without a
  // search while the query has moved less than the gap (tc_icp.cu).
  float second = INFINITY;
  __device__ __forceinline__ void init() {
    if (seeded) return;
    key = kEmpty;
    pos = 0;
  }
  __device__ __forceinline__ bool full() const { return key != kEmpty; }
  __device__ __forceinline__ float kth() const { return __uint_as_float((uint32_t)(key >> 32)); }
  // far mode (set by the caller for queries many cells from the surface, where a search scans a
  // thousand candidates and improves its best a handful of times): a candidate strictly farther
  // than the best costs one compare and one min; everything else takes the exact key path
  bool far = false;
  __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t lo, uint32_t hi,
                                       float qx, float qy, float qz, int) {
    if (lo >= hi) return;
    const uint32_t last = hi - 1;
    if (far) {
      float bd = kth();  // NaN while empty: `d2 > bd` is then false and the key path decides
#pragma unroll 1
      for (uint32_t base = lo; base < hi; base += 4) {
        float4 c[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) c[t] = __ldg(&pts[min(base + t, last)]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float d2 = dist2_exact(c[t].x, c[t].y, c[t].z, qx, qy, qz);
          if (base + t > last) continue;
          if (d2 > bd) {
            second = fminf(second, d2);
            continue;
          }
          const uint64_t k2 =
              ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c[t].w);
          if (k2 < key) {
            second = fminf(second, kth());
            key = k2;
            pos = base + t;
            bd = d2;
          } else if (k2 != key) {
            second = fminf(second, d2);
          }
        }
      }
      return;
    }
#pragma unroll 1
    for (uint32_t base = lo; base < hi; base += 4) {
      // four independent loads in flight (clamped: a repeated last candidate cannot win twice)
      float4 c[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) c[t] = __ldg(&pts[min(base + t, last)]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float d2 = dist2_exact(c[t].x, c[t].y, c[t].z, qx, qy, qz);
        const uint64_t k2 =
            ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c[t].w);
        if (base + t <= last) {  // (clamped repeats of the last candidate are not candidates)
          if (k2 < key) {
            second = fminf(second, kth());  // the displaced best (NaN when there was none)
            key = k2;
            pos = base + t;
          } else if (k2 != key) {
            second = fminf(second, d2);
          }
        }
      }
    }
  }
};

// Conservative lower bound (squared) on the distance from the query to any point in row
// (y, z): per axis the gap, in cell units, between the query and the nearest face of that row.
// Accumulators whose queries may lie outside the grid (ICP sources against a target's index)
// declare `static constexpr bool kOutside = true`: the query's OWN row then gets the distance
// from the query to the grid face as its gap, not zero.  A source point 3 m above the target's
// bounding box otherwise treats every row of the top plane within 3 m horizontally as a
// candidate row, although none of it can be nearer than 3 m.
template <class A, class = void>
struct acc_outside : std::false_type {};
template <class A>
struct acc_outside<A, std::void_t<decltype(A::kOutside)>> : std::bool_constant<A::kOutside> {};
template <bool OUTSIDE>
__device__ __forceinline__ float own_row_gap(float f) {
  // f = fractional cell coordinate relative to the CLAMPED cell: < 0 below the grid, > 1 above
  return OUTSIDE ? fmaxf(0.0f, fmaxf(f - 1.0f, -f)) : 0.0f;
}
__device__ __forceinline__ float row_gap(int d, float f) {
  // d = row coordinate - query cell coordinate along the axis; f = fractional cell coordinate
  return d == 0 ? 0.0f : (d < 0 ? (f + (float)(-d - 1)) : ((float)d - f));
}

// Exact nearest-neighbour search for one query; Acc is TopK<K> or Best1.
// flags: bit 0 deferred insertion (TopK), bit 1 prune rows that cannot beat the K-th distance.
// Searches blocks of radius 1, 2, ... ; `done` tells whether the result is final (otherwise the
// caller escalates to a coarser level: only while the block does not even hold K points).
template <class Acc>
__device__ __forceinline__ int grid_search(const GridParams& g, const float4* __restrict__ pts,
                                           const uint32_t* __restrict__ cell_start, float qx,
                                           float qy, float qz, Acc& tk, int max_R, bool& done) {
  tk.init();
  const int flags = g.flags;
  float ux, uy, uz;
  const int cx = cell_coord(qx, g.ox, g.inv, g.nx, ux);
  const int cy = cell_coord(qy, g.oy, g.inv, g.ny, uy);
  const int cz = cell_coord(qz, g.oz, g.inv, g.nz, uz);
  const float fx = ux - (float)cx, fy = uy - (float)cy, fz = uz - (float)cz;
  const float mx = g.ex + fabsf(qx - g.ox), my = g.ey + fabsf(qy - g.oy),
              mz = g.ez + fabsf(qz - g.oz);

  // One loop nest visits the block of radius R = 1 (rings 0 and 1 together), then thicker and
  // thicker shells (Rp, R] until the K-th best d2 is provably final.  While the list is not even
  // full the radius doubles (isolated points would otherwise pay O(R^3) dependent table look-ups);
  // once full it grows by one.  Rows are taken from the query's own row outwards (offsets
  // 0,-1,+1,-2,+2,...) so the K-th distance tightens early, and there is a SINGLE tk.scan call
  // site: the unrolled selection networks are instantiated once (I-cache).
  int R = 1, Rp = -1;  // current block radius, radius already searched
  while (true) {
    const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
    for (int iz = 0; iz <= 2 * R; ++iz) {
      const int dz = ((iz + 1) >> 1) * ((iz & 1) ? -1 : 1);
      const int z = cz + dz;
      if (z < 0 || z >= g.nz) continue;
      for (int iy = 0; iy <= 2 * R; ++iy) {
        const int dy = ((iy + 1) >> 1) * ((iy & 1) ? -1 : 1);
        const int y = cy + dy;
        if (y < 0 || y >= g.ny) continue;
        // rows outside the previous block get their full x-range, rows inside it only the two
        // new end ranges [cx-R, cx-Rp-1] and [cx+Rp+1, cx+R]
        const bool outer = (abs(dz) > Rp) || (abs(dy) > Rp);
        float byz = 0.0f;
        const bool prune = (flags & 2) && tk.full();
        if (prune) {
          constexpr bool kOut = acc_outside<Acc>::value;
          const float by =
              axis_bound(dy == 0 ? own_row_gap<kOut>(fy) : row_gap(dy, fy), g.cell, my);
          const float bz =
              axis_bound(dz == 0 ? own_row_gap<kOut>(fz) : row_gap(dz, fz), g.cell, mz);
          byz = by * by + bz * bz;
          if (byz * 0.99999f > tk.kth()) continue;  // every point of the row is farther
        }
        const uint32_t row = cell_id(g, 0, y, z);
        for (int seg = 0; seg < (outer ? 1 : 2); ++seg) {
          int xa, xb;
          if (outer) {
            xa = x0;
            xb = x1;
          } else if (seg == 0) {
            xa = x0;
            xb = cx - Rp - 1;
          } else {
            xa = cx + Rp + 1;
            xb = x1;
          }
          if (prune) {  // trim cells that cannot beat the K-th distance from both ends
            const float kth = tk.kth();
            while (xa < cx && xa <= xb) {
              const float bx = axis_bound(row_gap(xa - cx, fx), g.cell, mx);
              if ((byz + bx * bx) * 0.99999f > kth) ++xa; else break;
            }
            while (xb > cx && xb >= xa) {
              const float bx = axis_bound(row_gap(xb - cx, fx), g.cell, mx);
              if ((byz + bx * bx) * 0.99999f > kth) --xb; else break;
            }
          }
          if (xa > xb) continue;
          const uint32_t lo = __ldg(&cell_start[row + xa]);
          const uint32_t hi = __ldg(&cell_start[row + xb + 1]);
          tk.scan(pts, lo, hi, qx, qy, qz, flags);
        }
      }
    }
    const float b = ring_bound(g, R, cx, cy, cz, fx, fy, fz, mx, my, mz);
    done = true;
    if (b == INFINITY) break;  // the searched block already covers the whole grid
    if (tk.full() && tk.kth() < b * b * 0.99999f) break;
    done = false;
    // escalate to a coarser level only while the block does not even hold K points (sparse
    // neighbourhood); a full list that is not yet provably final needs just another ring here
    if (R >= max_R && !tk.full()) break;
    const int Rn = tk.full() ? R + 1 : 2 * R;
    // slab-sharded index (flags bits 8..15 = halo planes): cells beyond the halo are not built, so
    // the search stops here and reports a radius past the halo - the caller counts the query as
    // unsafe and its row is redone on a complete index
    const int r_cap = (flags >> 8) & 255;
    if (r_cap && Rn > r_cap) {
      done = true;
      R = r_cap + 1;
      break;
    }
    Rp = R;
    R = Rn;
  }
  return R;
}

// Number of indexed points in the 3x3x3 block of cells around the query (9 row look-ups).
__device__ __forceinline__ uint32_t block_population(const GridParams& g,
                                                     const uint32_t* __restrict__ cell_start,
                                                     float qx, float qy, float qz) {
  float u;
  const int cx = cell_coord(qx, g.ox, g.inv, g.nx, u);
  const int cy = cell_coord(qy, g.oy, g.inv, g.ny, u);
  const int cz = cell_coord(qz, g.oz, g.inv, g.nz, u);
  const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
  uint32_t pop = 0;
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int z = min(max(cz + dz, 0), g.nz - 1), y = min(max(cy + dy, 0), g.ny - 1);
      const bool in = (z == cz + dz) && (y == cy + dy);
      const uint32_t row = cell_id(g, 0, y, z);
      const uint32_t a = __ldg(&cell_start[row + x0]), b = __ldg(&cell_start[row + x1 + 1]);
      pop += in ? (b - a) : 0u;
    }
  }
  return pop;
}

// Multi-resolution driver.  Start at the finest level whose 3x3x3 block around the query holds
// at least `need` points (flags bit 4; otherwise: whose own cell holds 0.4 `need`), search it,
// and restart on the next coarser level while the block does not even hold `need` points; only
// the coarsest level keeps doubling its radius when starved.  The block test matters on skewed
// clouds: a stray point above dense ground has sparse own cells at every level and would
// otherwise search 27 of the coarsest cells - thousands of candidates, and its whole warp waits.
template <class Acc>
__device__ __forceinline__ int level_search(const LevelSet& ls, float qx, float qy, float qz,
                                            uint32_t need, Acc& tk, int& level,
                                            int start_level = -1) {
  int l = start_level >= 0 ? start_level : 0;
  const bool by_block = (ls.g[0].flags & 16) != 0;
  for (; start_level < 0 && l < ls.n - 1; ++l) {
    const GridParams& g = ls.g[l];
    if (by_block) {
      if (acc_outside<Acc>::value) {
        // a query more than a cell outside the grid: the block around its CLAMPED cell says
        // nothing about its neighbourhood - leave the fine levels to queries that are inside
        const float ux = (qx - g.ox) * g.inv, uy = (qy - g.oy) * g.inv, uz = (qz - g.oz) * g.inv;
        if (ux < -1.0f || uy < -1.0f || uz < -1.0f || ux > (float)g.nx + 1.0f ||
            uy > (float)g.ny + 1.0f || uz > (float)g.nz + 1.0f)
          continue;
      }
      if (block_population(g, ls.cs[l], qx, qy, qz) >= need) break;
      continue;
    }
    float u;
    const int cx = cell_coord(qx, g.ox, g.inv, g.nx, u);
    const int cy = cell_coord(qy, g.oy, g.inv, g.ny, u);
    const int cz = cell_coord(qz, g.oz, g.inv, g.nz, u);
    const uint32_t c = cell_id(g, cx, cy, cz);
    const uint32_t pop = __ldg(&ls.cs[l][c + 1]) - __ldg(&ls.cs[l][c]);
    if (pop * 5u >= need * 2u) break;
  }
  int R;
  while (true) {
    bool done;
    R = grid_search(ls.g[l], ls.pts[l], ls.cs[l], qx, qy, qz, tk, (l == ls.n - 1) ? (1 << 30) : 1,
                    done);
    if (done) break;
    ++l;
  }
  level = l;
  return R;
}

// Second traversal for the two-pass selection: visit every row of the block [c-R, c+R]^3 whose
// conservative distance bound does not exceed tau (the final K-th squared distance) and hand its
// candidate range to `f`.  Exact: a skipped row cannot hold a point with d2 <= tau.
template <class F>
__device__ __forceinline__ void grid_visit(const GridParams& g, const uint32_t* __restrict__ cell_start,
                                           float qx, float qy, float qz, int R, float tau, F&& f) {
  float ux, uy, uz;
  const int cx = cell_coord(qx, g.ox, g.inv, g.nx, ux);
  const int cy = cell_coord(qy, g.oy, g.inv, g.ny, uy);
  const int cz = cell_coord(qz, g.oz, g.inv, g.nz, uz);
  const float fx = ux - (float)cx, fy = uy - (float)cy, fz = uz - (float)cz;
  const float mx = g.ex + fabsf(qx - g.ox), my = g.ey + fabsf(qy - g.oy),
              mz = g.ez + fabsf(qz - g.oz);
  const int zlo = max(cz - R, 0), zhi = min(cz + R, g.nz - 1);
  const int ylo = max(cy - R, 0), yhi = min(cy + R, g.ny - 1);
  const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
  for (int z = zlo; z <= zhi; ++z) {
    const float bz = axis_bound(row_gap(z - cz, fz), g.cell, mz);
    for (int y = ylo; y <= yhi; ++y) {
      const float by = axis_bound(row_gap(y - cy, fy), g.cell, my);
      const float byz = by * by + bz * bz;
      if (byz * 0.99999f > tau) continue;
      int xa = x0, xb = x1;  // trim cells that cannot hold a point with d2 <= tau
      while (xa < cx) {
        const float bx = axis_bound(row_gap(xa - cx, fx), g.cell, mx);
        if ((byz + bx * bx) * 0.99999f > tau) ++xa; else break;
      }
      while (xb > cx) {
        const float bx = axis_bound(row_gap(xb - cx, fx), g.cell, mx);
        if ((byz + bx * bx) * 0.99999f > tau) --xb; else break;
      }
      const uint32_t row = cell_id(g, 0, y, z);
      const uint32_t lo = __ldg(&cell_start[row + xa]);
      const uint32_t hi = __ldg(&cell_start[row + xb + 1]);
      f(lo, hi);
    }
  }
}

// Box query: hand `f` the candidate range of every row of cells that can hold a point within
// squared distance r2 of the query.  Exact by monotonicity: the f32 cell-coordinate expression
// is non-decreasing in the coordinate, so a point with |p - q| <= r on an axis has its cell
// coordinate within [cell(q - r), cell(q + r)] (r is inflated to absorb the rounding of the
// square root and of q -+ r).  Used where a distance bound is already known: ICP correspondences
// seeded with the previous iteration's match (or a probe), radius queries, the radius outlier
// count.  The box is fixed on entry; `f` may tighten its own acceptance test while iterating.
struct BoxCells {
  int xa, xb, ya, yb, za, zb;
};
// cells that can hold a point within squared distance r2 (finite) of the query
__device__ __forceinline__ BoxCells box_cells(const GridParams& g, float qx, float qy, float qz,
                                              float r2) {
  const float r = xsqrt(r2) * 1.00001f + 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + g.cell) +
                  1e-30f;
  float u;
  BoxCells b;
  b.xa = cell_coord(qx - r, g.ox, g.inv, g.nx, u);
  b.xb = cell_coord(qx + r, g.ox, g.inv, g.nx, u);
  b.ya = cell_coord(qy - r, g.oy, g.inv, g.ny, u);
  b.yb = cell_coord(qy + r, g.oy, g.inv, g.ny, u);
  b.za = cell_coord(qz - r, g.oz, g.inv, g.nz, u);
  b.zb = cell_coord(qz + r, g.oz, g.inv, g.nz, u);
  return b;
}
template <class F>
__device__ __forceinline__ void box_visit(const GridParams& g, const uint32_t* __restrict__ cell_start,
                                          float qx, float qy, float qz, float r2, F&& f) {
  if (!(r2 < INFINITY)) {  // unbounded: every cell
    for (int z = 0; z < g.nz; ++z)
      for (int y = 0; y < g.ny; ++y) {
        const uint32_t row = cell_id(g, 0, y, z);
        f(__ldg(&cell_start[row]), __ldg(&cell_start[row + g.nx]));
      }
    return;
  }
  const BoxCells bc = box_cells(g, qx, qy, qz, r2);
  const int xa = bc.xa, xb = bc.xb, ya = bc.ya, yb = bc.yb, za = bc.za, zb = bc.zb;
  for (int z = za; z <= zb; ++z)
    for (int y = ya; y <= yb; ++y) {
      const uint32_t row = cell_id(g, 0, y, z);
      const uint32_t lo = __ldg(&cell_start[row + xa]);
      const uint32_t hi = __ldg(&cell_start[row + xb + 1]);
      f(lo, hi);
    }
}

// Where the row of original index `qid` goes: the launch's own buffer, or (distributed normals)
// the chunk buffer of rank qid / chunk.  __umulhi(qid, floor(2^32 / chunk)) is the quotient or
// one below it.
__device__ __forceinline__ float* route_out(const LevelSet& ls, uint32_t qid, float* out) {
  if (ls.route_chunk == 0) return out;
  uint32_t r = __umulhi(qid, ls.route_magic);
  if ((r + 1) * ls.route_chunk <= qid) ++r;
  return ls.route[r];
}

}  // namespace tcs
