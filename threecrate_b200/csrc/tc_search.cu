// tc_search.cu — exact grid kNN with a per-query top-K in registers, and the fused normals
// epilogue (covariance in the reference's summation order -> symmetric 3x3 eigen -> orientation).
//
// Replaces KdTree::find_k_nearest (threecrate-algorithms/src/nearest_neighbor.rs:177-251),
// PointCloudNeighbors::k_nearest_neighbors (point_cloud_ops.rs:80-105) and the body of the rayon
// loop of estimate_normals_with_config (normals.rs:306-354).
//
// One thread per query, queries taken in cell-sorted order so the lanes of a warp walk the same
// (or adjacent) cells: candidate loads are float4 (x,y,z,index) and mostly warp-uniform.
// Distances use the reference's exact f32 expression (no FMA).  The top-K is a sorted list of
// u64 keys (d2 bits << 32 | original index): one integer compare gives the (d2, index) order.
// The search is exact: ring r of cells is added until the K-th best d2 is below the (conservative)
// squared distance from the query to the boundary of the searched block of cells.
#include "tc_search.cuh"
#include "tc_normal.cuh"

// Per-query cycle counters for tools/qclock.py exist only in a -DTC_QCLOCK build
// (python -m threecrate_b200.build --qclock): the release kernels carry neither the extra
// parameter nor the clock reads.
#ifdef TC_QCLOCK
#define TC_DBG_PARAM , uint32_t* __restrict__ dbg
#define TC_DBG_PARAM_DEF , uint32_t* dbg = nullptr
#define TC_DBG_ARG(x) , x
#define TC_DBG(...) __VA_ARGS__
#else
#define TC_DBG_PARAM
#define TC_DBG_PARAM_DEF
#define TC_DBG_ARG(x)
#define TC_DBG(...)
#endif

namespace {

constexpr int kBlock = 128;
using namespace tcs;

// More candidates at or below the K-th squared distance than the member table holds: bit-equal
// d2 straddling rank `need` (a handful of queries per million on real-valued data, every query
// of a cloud with duplicated points).  The members are then collected exactly, in place: first
// everything strictly below tau (fewer than `need` by the definition of tau), then the points AT
// tau in ascending original index - one traversal per slot still open, each taking the smallest
// index above the last one taken.  Not inlined: this is the cold path of the selection, and the
// extra traversals would otherwise sit in the hot kernels' instruction stream.
__device__ __noinline__ uint32_t resolve_ties(const GridParams g, const uint32_t* __restrict__ cell_start,
                                              const float4* __restrict__ pts, float qx, float qy,
                                              float qz, int R, float tau, uint32_t need,
                                              uint32_t* col /* s_a[.][thread], stride kBlock */) {
  uint32_t n = 0;
  grid_visit(g, cell_start, qx, qy, qz, R, tau, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      if (dist2_exact(c.x, c.y, c.z, qx, qy, qz) < tau) col[(n++) * kBlock] = j;
    }
  });
  long long last = -1;
  while (n < need) {
    long long pick = 1ll << 40;
    uint32_t pick_j = 0;
    grid_visit(g, cell_start, qx, qy, qz, R, tau, [&](uint32_t lo, uint32_t hi) {
      for (uint32_t j = lo; j < hi; ++j) {
        const float4 c = __ldg(&pts[j]);
        const long long id = (long long)__float_as_uint(c.w);
        if (dist2_exact(c.x, c.y, c.z, qx, qy, qz) == tau && id > last && id < pick) {
          pick = id;
          pick_j = j;
        }
      }
    });
    if (pick == (1ll << 40)) break;  // (no point left at tau)
    col[(n++) * kBlock] = pick_j;
    last = pick;
  }
  return n;
}

// Two-pass exact selection for one query (see k_knn2).  On success the sorted positions (into
// `pts`) of the neighbours, ascending by (d2, index), are in s_b[0 .. n)[threadIdx.x] and n is
// returned (bit-equal d2 are ordered by original index; more of them at the K-th d2 than the
// member table holds are resolved by resolve_ties).
//   pass 1  SelF: exact K-th squared distance tau (floats only, sorting networks)
//   pass 2  re-scan the rows that can hold d2 <= tau, append each member's position to s_a
//           (a 4-instruction append, so lanes that accept different candidates cost little)
//   rank    uniform loop over the n members: slot = #{v[i] < d2}; s_b[slot] = position
template <int L, bool X>
__device__ __forceinline__ int select_two_pass(const LevelSet& ls, float qx, float qy, float qz,
                                               uint32_t need, uint32_t (*s_a)[kBlock],
                                               uint32_t (*s_b)[kBlock], int& level,
                                               int* R_out = nullptr TC_DBG_PARAM_DEF) {
  constexpr int T = SelF<L, X>::kSlots;  // rows of s_a / s_b
  SelF<L, X> sel;
  sel.pad = T - (int)need;
  int R = level_search(ls, qx, qy, qz, need, sel, level);
  if (R_out) *R_out = R;
  // (slab-sharded index: a search stopped at the ring cap reports cap + 1 - the caller counts the
  //  query as unsafe - but the second traversal must stay inside the built planes as well)
  if (ls.halo && R > ls.halo) R = ls.halo;
  const GridParams& g = ls.g[level];
  const float4* __restrict__ pts = ls.pts[level];
  const uint32_t* __restrict__ cell_start = ls.cs[level];
  const float tau = sel.kth();  // +inf when fewer than `need` points exist: everything is kept
  TC_DBG(if (dbg) {
    dbg[4] = (uint32_t)clock64();
    dbg[6] = (uint32_t)R;
  })
  uint32_t n = 0;
  grid_visit(g, cell_start, qx, qy, qz, R, tau, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
      if (d2 <= tau) {
        if (n < (uint32_t)T) s_a[n][threadIdx.x] = j;
        ++n;
      }
    }
  });
  if (n > (uint32_t)T)  // more ties at tau than the table holds: resolved in place (cold path)
    n = resolve_ties(g, cell_start, pts, qx, qy, qz, R, tau, need, &s_a[0][threadIdx.x]);
  // rank placement.  A member whose d2 is unique lands at #{v[i] < d2}.  Bit-equal d2 (inside
  // the list, or at the K-th distance when more than `need` points are at or below it) are
  // ordered by original index: such a member also counts the equal members with a smaller index.
  const bool any_tie = n > need /* tie straddling rank `need` */ || sel.has_equal();
#pragma unroll 1
  for (uint32_t m = 0; m < n; ++m) {
    const uint32_t j = s_a[m][threadIdx.x];
    const float4 c = __ldg(&pts[j]);
    const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
    int rank = sel.count_below(d2) - sel.pad;  // the sentinels are always below d2
    if (any_tie) {  // rare: order bit-equal d2 by original index
      const uint32_t id = __float_as_uint(c.w);
      for (uint32_t o = 0; o < n; ++o) {
        const float4 e = __ldg(&pts[s_a[o][threadIdx.x]]);
        if (dist2_exact(e.x, e.y, e.z, qx, qy, qz) == d2 && __float_as_uint(e.w) < id) ++rank;
      }
    }
    if (rank < (int)need) s_b[rank][threadIdx.x] = j;
  }
  TC_DBG(if (dbg) dbg[5] = (uint32_t)clock64();)
  if (n > need) n = need;
  return (int)n;
}

// Shard ownership (multi-GPU): the query at sorted position p belongs to the shard that contains
// the FIRST position of its level-0 cell, so shards own whole cells and the assignment does not
// depend on the (arbitrary) order of points inside a cell.
__device__ __forceinline__ bool owns_query(const LevelSet& ls, const float4& q, uint32_t begin,
                                           uint32_t end) {
  const GridParams& g = ls.g[0];
  float u;
  const int cx = cell_coord(q.x, g.ox, g.inv, g.nx, u);
  const int cy = cell_coord(q.y, g.oy, g.inv, g.ny, u);
  const int cz = cell_coord(q.z, g.oz, g.inv, g.nz, u);
  const uint32_t first = __ldg(&ls.cs[0][cell_id(g, cx, cy, cz)]);
  return first >= begin && first < end;
}

// ---------------------------------------------------------------------------------- kNN kernel
// kNN(k+1), retain idx != i, truncate k  (point_cloud_ops.rs:91-99); plain kNN otherwise.
// `key` is the ascending (d2, index) list; entries == kEmpty are padding.
// Neighbour source for the emit functions.
//  RegKeys  : the u64 (d2, index) register list of the chain kernels (static indices, unrolled);
//             coordinates come from the original-order xyz array.
//  SortedPos: positions into the cell-sorted float4 array (shared-memory column, rolled loop);
//             one 128-bit load yields coordinates and index, d2 is recomputed.
struct Nb {
  float x, y, z, d2;
  uint32_t id;
  bool valid;
};
template <int K>
struct RegKeys {
  const uint64_t (&key)[K];
  const float* __restrict__ xyz;
  static constexpr int kCount = K;
  static constexpr bool kUnroll = true;
  __device__ __forceinline__ int count() const { return K; }
  __device__ __forceinline__ Nb head(int i) const {  // index + d2 only
    Nb r;
    r.valid = key[i] != kEmpty;
    r.id = (uint32_t)key[i];
    r.d2 = __uint_as_float((uint32_t)(key[i] >> 32));
    r.x = r.y = r.z = 0.0f;
    return r;
  }
  __device__ __forceinline__ void coords(Nb& r) const {
    const float* p = xyz + 3 * (uint64_t)r.id;
    r.x = __ldg(p + 0);
    r.y = __ldg(p + 1);
    r.z = __ldg(p + 2);
  }
};
struct SortedPos {
  uint32_t (*col)[kBlock];
  int n;
  const float4* __restrict__ pts;
  float qx, qy, qz;
  static constexpr int kCount = 1;
  static constexpr bool kUnroll = false;
  __device__ __forceinline__ int count() const { return n; }
  __device__ __forceinline__ Nb head(int i) const {
    const float4 c = __ldg(&pts[col[i][threadIdx.x]]);
    Nb r;
    r.valid = true;
    r.x = c.x;
    r.y = c.y;
    r.z = c.z;
    r.id = __float_as_uint(c.w);
    r.d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
    return r;
  }
  __device__ __forceinline__ void coords(Nb&) const {}
};

template <class KS>
__device__ __forceinline__ void knn_emit(const KS& keys, uint32_t qid, uint32_t k, int drop_self,
                                         uint32_t* __restrict__ idx_out,
                                         float* __restrict__ dist_out,
                                         uint32_t* __restrict__ count_out) {
  uint32_t c = 0;
  const uint64_t row = (uint64_t)qid * k;
  auto body = [&](int i) {
    const Nb nb = keys.head(i);
    if (nb.valid && c < k && !(drop_self && nb.id == qid)) {
      idx_out[row + c] = nb.id;
      if (dist_out) dist_out[row + c] = xsqrt(nb.d2);
      ++c;
    }
  };
  if (KS::kUnroll) {
#pragma unroll
    for (int i = 0; i < KS::kCount; ++i) body(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < keys.count(); ++i) body(i);
  }
  if (count_out) count_out[qid] = c;
  for (uint32_t j = c; j < k; ++j) {
    idx_out[row + j] = TC_NO_INDEX;
    if (dist_out) dist_out[row + j] = INFINITY;
  }
}

// Exact u64 insertion-chain kernel.  With `list` == nullptr it processes queries
// [q_begin, q_end); otherwise the query positions listed in list[0 .. *list_count) (the
// tie-overflow fallback of the two-pass kernels), grid-stride.
template <int K>
__global__ void __launch_bounds__(kBlock)
k_knn(LevelSet ls, const float4* __restrict__ queries, uint32_t q_begin, uint32_t q_end,
      uint32_t k, uint32_t need, int drop_self, uint32_t* __restrict__ idx_out,
      float* __restrict__ dist_out, uint32_t* __restrict__ count_out,
      const uint32_t* __restrict__ list, const uint32_t* __restrict__ list_count) {
  const uint32_t total = list ? *list_count : (q_end - q_begin);
  for (uint32_t t = blockIdx.x * kBlock + threadIdx.x; t < total; t += gridDim.x * kBlock) {
    const uint32_t qi = list ? list[t] : q_begin + t;
    const float4 q = __ldg(&queries[qi]);
    const uint32_t qid = __float_as_uint(q.w);  // original query index = output row
    TopK<K> tk;
    int level;
    level_search(ls, q.x, q.y, q.z, need, tk, level);
    knn_emit(RegKeys<K>{tk.key, nullptr}, qid, k, drop_self, idx_out, dist_out, count_out);
  }
}

// Two-pass kernel: pass 1 finds the exact K-th squared distance with the float selection list,
// pass 2 re-scans and collects the positions of the points with d2 <= tau into shared memory and
// ranks them.  More members than the table holds (ties at rank k) are resolved in place
// (resolve_ties); fb_list / fb_count are no longer written by this kernel (the staged-tile
// variant shares the launch signature and still hands its unproven queries to the list).
template <int L, bool X>
__global__ void __launch_bounds__(kBlock)
k_knn2(LevelSet ls, const float4* __restrict__ queries, uint32_t q_begin, uint32_t q_end,
       uint32_t k, uint32_t need, int drop_self, uint32_t* __restrict__ idx_out,
       float* __restrict__ dist_out, uint32_t* __restrict__ count_out,
       uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_count) {
  __shared__ uint32_t s_a[L + (X ? 1 : 0)][kBlock], s_b[L + (X ? 1 : 0)][kBlock];
  const uint32_t qi = q_begin + blockIdx.x * kBlock + threadIdx.x;
  if (qi >= q_end) return;
  const float4 q = __ldg(&queries[qi]);
  const uint32_t qid = __float_as_uint(q.w);
  int level;
  const int n = select_two_pass<L, X>(ls, q.x, q.y, q.z, need, s_a, s_b, level);
  knn_emit(SortedPos{s_b, n, ls.pts[level], q.x, q.y, q.z}, qid, k, drop_self, idx_out, dist_out,
           count_out);
}

// ------------------------------------------------------------------------------ normals kernel
// Normal of one point from its ascending (d2, index) neighbour keys (normals.rs:306-354 body).
template <class KS>
__device__ __forceinline__ void normals_emit(const KS& keys, const float4 q, uint32_t qid,
                                             uint32_t k, int orient, float vpx, float vpy,
                                             float vpz, float* __restrict__ out,
                                             bool fast = false) {
  // neighbourhood = first k of kNN(k+1) with self dropped by index, then self appended last
  // (normals.rs:148-153, 338-340).  Sums are sequential f32 in that order (normals.rs:165-177).
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  uint32_t cnt = 0;
  auto sum_body = [&](int i) {
    Nb nb = keys.head(i);
    if (nb.valid && cnt < k && nb.id != qid) {
      keys.coords(nb);
      sx = xadd(sx, nb.x);
      sy = xadd(sy, nb.y);
      sz = xadd(sz, nb.z);
      ++cnt;
    }
  };
  if (KS::kUnroll) {
#pragma unroll
    for (int i = 0; i < KS::kCount; ++i) sum_body(i);
  } else {
#pragma unroll 1
    for (int i = 0; i < keys.count(); ++i) sum_body(i);
  }
  sx = xadd(sx, q.x);
  sy = xadd(sy, q.y);
  sz = xadd(sz, q.z);
  const uint32_t nn = cnt + 1;
  float nrm[3] = {0.0f, 0.0f, 1.0f};  // < 3 points (normals.rs:159-162)
  if (nn >= 3) {
    const float fn = (float)nn;
    const float cx = xdiv(sx, fn), cy = xdiv(sy, fn), cz = xdiv(sz, fn);
    float c[6] = {0, 0, 0, 0, 0, 0};
    auto acc = [&](float px, float py, float pz) {
      const float dx = xsub(px, cx), dy = xsub(py, cy), dz = xsub(pz, cz);
      c[0] = xadd(c[0], xmul(dx, dx));
      c[1] = xadd(c[1], xmul(dx, dy));
      c[2] = xadd(c[2], xmul(dx, dz));
      c[3] = xadd(c[3], xmul(dy, dy));
      c[4] = xadd(c[4], xmul(dy, dz));
      c[5] = xadd(c[5], xmul(dz, dz));
    };
    uint32_t cnt2 = 0;
    auto cov_body = [&](int i) {
      Nb nb = keys.head(i);
      if (nb.valid && cnt2 < k && nb.id != qid) {
        keys.coords(nb);
        acc(nb.x, nb.y, nb.z);
        ++cnt2;
      }
    };
    if (KS::kUnroll) {
#pragma unroll
      for (int i = 0; i < KS::kCount; ++i) cov_body(i);
    } else {
#pragma unroll 1
      for (int i = 0; i < keys.count(); ++i) cov_body(i);
    }
    acc(q.x, q.y, q.z);
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i] = xdiv(c[i], fn);
    normal_from_cov(c, nrm, fast);
  }
  write_normal(nrm, q, qid, orient, vpx, vpy, vpz, out);
}

template <int K>
__global__ void __launch_bounds__(kBlock)
k_normals(LevelSet ls, const float* __restrict__ xyz, uint32_t q_begin, uint32_t q_end,
          uint32_t own_begin, uint32_t own_end, uint32_t k, int orient, float vpx, float vpy,
          float vpz, float* __restrict__ out, const uint32_t* __restrict__ list,
          const uint32_t* __restrict__ list_count TC_DBG_PARAM) {
  const uint32_t total = list ? *list_count : (q_end - q_begin);
  for (uint32_t t = blockIdx.x * kBlock + threadIdx.x; t < total; t += gridDim.x * kBlock) {
    const uint32_t qi = list ? list[t] : q_begin + t;
    const float4 q = __ldg(&ls.pts[0][qi]);
    if (!list && own_end != 0xFFFFFFFFu && !owns_query(ls, q, own_begin, own_end)) continue;
    const uint32_t qid = __float_as_uint(q.w);
    TC_DBG(const long long t0 = dbg ? clock64() : 0;)
    TopK<K> tk;
    int level;
    const int R = level_search(ls, q.x, q.y, q.z, k + 1, tk, level);
    if (ls.halo && R > ls.halo) atomicAdd(ls.unsafe, 1u);  // slab index: may have missed points
    normals_emit(RegKeys<K>{tk.key, xyz}, q, qid, k, orient, vpx, vpy, vpz,
                 route_out(ls, qid, out));
    TC_DBG(if (dbg) {  // per-query cycles and final block radius (tools/qclock.py)
      dbg[8 * (size_t)qid] = (uint32_t)(clock64() - t0);
      dbg[8 * (size_t)qid + 1] = (uint32_t)R | ((uint32_t)level << 16);
    })
  }
}

// (72 registers for the 16-wide list: 7 blocks per SM, +6 % on the 10M cloud; the 32-wide list
//  keeps ptxas' own choice of 96 - a forced 5 blocks/SM spills and loses)
template <int L, bool X>
__global__ void __launch_bounds__(kBlock, L == 16 ? 7 : 0)
k_normals2(LevelSet ls, uint32_t q_begin, uint32_t q_end, uint32_t own_begin, uint32_t own_end,
           uint32_t k, int orient, float vpx, float vpy, float vpz, float* __restrict__ out,
           uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_count TC_DBG_PARAM) {
  __shared__ uint32_t s_a[L + (X ? 1 : 0)][kBlock], s_b[L + (X ? 1 : 0)][kBlock];
  const uint32_t qi = q_begin + blockIdx.x * kBlock + threadIdx.x;
  if (qi >= q_end) return;
  const float4 q = __ldg(&ls.pts[0][qi]);
  if (own_end != 0xFFFFFFFFu && !owns_query(ls, q, own_begin, own_end)) return;
  const uint32_t qid = __float_as_uint(q.w);
#ifdef TC_QCLOCK
  const long long t0 = dbg ? clock64() : 0;
  if (dbg) {  // 8 words per query (tools/qclock.py): cycles, n|level, sorted position, start ns,
              // clock after pass 1, clock after pass 2 + rank, final block radius, start clock
    dbg += 8 * (size_t)qid;
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    dbg[2] = qi;
    dbg[3] = (uint32_t)ns;
    dbg[7] = (uint32_t)t0;
  }
#endif
  int level, R = 0;
  const int n = select_two_pass<L, X>(ls, q.x, q.y, q.z, k + 1, s_a, s_b, level, &R TC_DBG_ARG(dbg));
  if (ls.halo && R > ls.halo) atomicAdd(ls.unsafe, 1u);  // slab index: may have missed points
  normals_emit(SortedPos{s_b, n, ls.pts[level], q.x, q.y, q.z}, q, qid, k, orient, vpx, vpy, vpz,
               route_out(ls, qid, out), (ls.g[0].flags & 128) != 0);
  TC_DBG(if (dbg) {
    dbg[0] = (uint32_t)(clock64() - t0);
    dbg[1] = (uint32_t)n | ((uint32_t)level << 16);
  })
}

// ------------------------------------------------------------------------------ radius mode
// estimate_normals_with_config with radius = Some(r) (normals.rs:141-146, 315-340): neighbourhood
// = every point with d2 <= r^2 except the query's own index, in ascending-distance order
// (find_radius_neighbors sorts, nearest_neighbor.rs:283-296), then the query itself; the sums are
// sequential f32 in that order (normals.rs:165-177).  Queries with fewer than k such neighbours
// fall back to the kNN rule (normals.rs:315-323): appended to `fb_list`.
// Three steps so that every query can hold ALL its neighbours (a radius has no upper bound on
// their number): count -> exclusive scan -> fill + in-place heapsort of the query's own segment
// of u64 (d2 bits << 32 | index) keys + the reference-order sums.
__global__ void __launch_bounds__(kBlock)
k_radius_count(LevelSet ls, int level, uint32_t q_begin, uint32_t q_end, uint32_t own_begin,
               uint32_t own_end, float radius, uint32_t k, uint32_t* __restrict__ counts,
               uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_count) {
  const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
  const uint32_t qi = q_begin + t;
  if (qi >= q_end) return;
  counts[t] = 0u;
  const float4 q = __ldg(&ls.pts[0][qi]);
  if (own_end != 0xFFFFFFFFu && !owns_query(ls, q, own_begin, own_end)) return;
  const uint32_t qid = __float_as_uint(q.w);
  const float4* __restrict__ pts = ls.pts[level];
  const float r2 = xmul(radius, radius);  // nearest_neighbor.rs:259
  uint32_t cnt = 0;
  box_visit(ls.g[level], ls.cs[level], q.x, q.y, q.z, r2, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      const float d2 = dist2_exact(c.x, c.y, c.z, q.x, q.y, q.z);
      if (d2 <= r2 && __float_as_uint(c.w) != qid) ++cnt;
    }
  });
  if (cnt < k) {  // normals.rs:315-323: the kNN rule instead
    fb_list[atomicAdd(fb_count, 1u)] = qi;
    return;
  }
  counts[t] = cnt;
}

__global__ void __launch_bounds__(kBlock)
k_radius_normals(LevelSet ls, int level, const float* __restrict__ xyz, uint32_t q_begin,
                 uint32_t q_end, float radius, int orient, float vpx, float vpy, float vpz,
                 const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys,
                 float* __restrict__ out) {
  const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
  const uint32_t qi = q_begin + t;
  if (qi >= q_end) return;
  const uint32_t n = offsets[t + 1] - offsets[t];
  if (n == 0) return;  // not owned, or left to the kNN rule
  const float4 q = __ldg(&ls.pts[0][qi]);
  const uint32_t qid = __float_as_uint(q.w);
  const float4* __restrict__ pts = ls.pts[level];
  const float r2 = xmul(radius, radius);
  uint64_t* h = keys + offsets[t];
  uint32_t m = 0;
  box_visit(ls.g[level], ls.cs[level], q.x, q.y, q.z, r2, [&](uint32_t lo, uint32_t hi) {
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      const float d2 = dist2_exact(c.x, c.y, c.z, q.x, q.y, q.z);
      const uint32_t id = __float_as_uint(c.w);
      if (d2 <= r2 && id != qid && m < n) h[m++] = ((uint64_t)__float_as_uint(d2) << 32) | id;
    }
  });
  // in-place heapsort -> ascending (d2, index)
  auto sift = [&](uint32_t i, uint32_t end, uint64_t v) {
    while (true) {
      uint32_t c = 2 * i + 1;
      if (c >= end) break;
      if (c + 1 < end && h[c + 1] > h[c]) ++c;
      if (h[c] <= v) break;
      h[i] = h[c];
      i = c;
    }
    h[i] = v;
  };
  for (uint32_t i = m / 2; i-- > 0;) sift(i, m, h[i]);
  for (uint32_t end = m; end > 1; --end) {
    const uint64_t last = h[end - 1];
    h[end - 1] = h[0];
    sift(0, end - 1, last);
  }
  // reference-order sums (normals.rs:165-177): neighbours ascending, the query itself last
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  for (uint32_t i = 0; i < m; ++i) {
    const float* p = xyz + 3 * (uint64_t)(uint32_t)h[i];
    sx = xadd(sx, __ldg(p));
    sy = xadd(sy, __ldg(p + 1));
    sz = xadd(sz, __ldg(p + 2));
  }
  sx = xadd(sx, q.x);
  sy = xadd(sy, q.y);
  sz = xadd(sz, q.z);
  const float fn = (float)(m + 1);
  const float cx = xdiv(sx, fn), cy = xdiv(sy, fn), cz = xdiv(sz, fn);
  float c[6] = {0, 0, 0, 0, 0, 0};
  auto acc = [&](float px, float py, float pz) {
    const float dx = xsub(px, cx), dy = xsub(py, cy), dz = xsub(pz, cz);
    c[0] = xadd(c[0], xmul(dx, dx));
    c[1] = xadd(c[1], xmul(dx, dy));
    c[2] = xadd(c[2], xmul(dx, dz));
    c[3] = xadd(c[3], xmul(dy, dy));
    c[4] = xadd(c[4], xmul(dy, dz));
    c[5] = xadd(c[5], xmul(dz, dz));
  };
  for (uint32_t i = 0; i < m; ++i) {
    const float* p = xyz + 3 * (uint64_t)(uint32_t)h[i];
    acc(__ldg(p), __ldg(p + 1), __ldg(p + 2));
  }
  acc(q.x, q.y, q.z);
  float nrm[3] = {0.0f, 0.0f, 1.0f};
  if (m + 1 >= 3) {  // normals.rs:159-162
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i] = xdiv(c[i], fn);
    normal_from_cov(c, nrm, (ls.g[0].flags & 128) != 0);
  }
  write_normal(nrm, q, qid, orient, vpx, vpy, vpz, out);
}

// KdTree::find_radius_neighbors for ONE query (nearest_neighbor.rs:254-298): one block; warps take
// rows of the box, lanes take candidates; hits are appended (unsorted) and ordered on the host.
__global__ void __launch_bounds__(256)
k_radius_search(LevelSet ls, int level, float qx, float qy, float qz, float radius,
                uint32_t* __restrict__ idx_out, float* __restrict__ d2_out, uint32_t capacity,
                uint32_t* __restrict__ n_found) {
  const GridParams& g = ls.g[level];
  const float4* __restrict__ pts = ls.pts[level];
  const float r2 = xmul(radius, radius);
  const float r = xsqrt(r2) * 1.00001f + 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + g.cell) + 1e-30f;
  float u;
  const int xa = cell_coord(qx - r, g.ox, g.inv, g.nx, u), xb = cell_coord(qx + r, g.ox, g.inv, g.nx, u);
  const int ya = cell_coord(qy - r, g.oy, g.inv, g.ny, u), yb = cell_coord(qy + r, g.oy, g.inv, g.ny, u);
  const int za = cell_coord(qz - r, g.oz, g.inv, g.nz, u), zb = cell_coord(qz + r, g.oz, g.inv, g.nz, u);
  const int ny = yb - ya + 1, nrows = ny * (zb - za + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < nrows; t += 8) {
    const int z = za + t / ny, y = ya + t % ny;
    const uint32_t row = cell_id(g, 0, y, z);
    const uint32_t lo = __ldg(&ls.cs[level][row + xa]), hi = __ldg(&ls.cs[level][row + xb + 1]);
    for (uint32_t j = lo + lane; j < hi; j += 32) {
      const float4 c = __ldg(&pts[j]);
      const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
      if (d2 <= r2) {  // inclusive (nearest_neighbor.rs:270)
        const uint32_t slot = atomicAdd(n_found, 1u);
        if (slot < capacity) {
          idx_out[slot] = __float_as_uint(c.w);
          d2_out[slot] = d2;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------- any k
// k + 1 beyond the register lists (> 64): a binary max-heap of u64 (d2, index) keys per query in
// global memory (the query's `need` slots are contiguous, so the hot top of the heap stays in
// L1), the same exact ring search, then an in-place heapsort -> ascending (d2, index).
struct HeapK {
  uint64_t* h;
  uint32_t cap, n;
  __device__ __forceinline__ void init() { n = 0; }
  __device__ __forceinline__ bool full() const { return n == cap; }
  __device__ __forceinline__ float kth() const {
    return n == cap ? __uint_as_float((uint32_t)(h[0] >> 32)) : INFINITY;
  }
  __device__ __forceinline__ void sift_down(uint32_t i, uint32_t end, uint64_t v) {
    while (true) {
      uint32_t c = 2 * i + 1;
      if (c >= end) break;
      if (c + 1 < end && h[c + 1] > h[c]) ++c;
      if (h[c] <= v) break;
      h[i] = h[c];
      i = c;
    }
    h[i] = v;
  }
  __device__ __forceinline__ void scan(const float4* __restrict__ pts, uint32_t lo, uint32_t hi,
                                       float qx, float qy, float qz, int) {
    for (uint32_t j = lo; j < hi; ++j) {
      const float4 c = __ldg(&pts[j]);
      const float d2 = dist2_exact(c.x, c.y, c.z, qx, qy, qz);
      const uint64_t key = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c.w);
      if (n < cap) {  // sift up
        uint32_t i = n++;
        while (i > 0) {
          const uint32_t p = (i - 1) >> 1;
          if (h[p] >= key) break;
          h[i] = h[p];
          i = p;
        }
        h[i] = key;
      } else if (key < h[0]) {
        sift_down(0, cap, key);
      }
    }
  }
  __device__ __forceinline__ void sort_ascending() {
    for (uint32_t end = n; end > 1; --end) {
      const uint64_t last = h[end - 1];
      h[end - 1] = h[0];
      sift_down(0, end - 1, last);
    }
  }
};
struct GlobalKeys {
  const uint64_t* key;
  int n;
  const float* __restrict__ xyz;
  static constexpr int kCount = 1;
  static constexpr bool kUnroll = false;
  __device__ __forceinline__ int count() const { return n; }
  __device__ __forceinline__ Nb head(int i) const {
    Nb r;
    r.valid = true;
    r.id = (uint32_t)key[i];
    r.d2 = __uint_as_float((uint32_t)(key[i] >> 32));
    r.x = r.y = r.z = 0.0f;
    return r;
  }
  __device__ __forceinline__ void coords(Nb& r) const {
    const float* p = xyz + 3 * (uint64_t)r.id;
    r.x = __ldg(p + 0);
    r.y = __ldg(p + 1);
    r.z = __ldg(p + 2);
  }
};

__global__ void __launch_bounds__(kBlock)
k_knn_big(LevelSet ls, const float4* __restrict__ queries, uint32_t q_begin, uint32_t q_end,
          uint32_t k, uint32_t need, int drop_self, uint32_t* __restrict__ idx_out,
          float* __restrict__ dist_out, uint32_t* __restrict__ count_out,
          uint64_t* __restrict__ heaps) {
  const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= q_end - q_begin) return;
  const float4 q = __ldg(&queries[q_begin + t]);
  HeapK hk{heaps + (uint64_t)t * need, need, 0};
  int level;
  level_search(ls, q.x, q.y, q.z, need, hk, level);
  hk.sort_ascending();
  knn_emit(GlobalKeys{hk.h, (int)hk.n, nullptr}, __float_as_uint(q.w), k, drop_self, idx_out,
           dist_out, count_out);
}

// (`list` != nullptr: the query positions list[q_begin .. q_end) instead of the range itself)
__global__ void __launch_bounds__(kBlock)
k_normals_big(LevelSet ls, const float* __restrict__ xyz, uint32_t q_begin, uint32_t q_end,
              uint32_t own_begin, uint32_t own_end, uint32_t k, int orient, float vpx, float vpy,
              float vpz, float* __restrict__ out, uint64_t* __restrict__ heaps,
              const uint32_t* __restrict__ list = nullptr) {
  const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= q_end - q_begin) return;
  const float4 q = __ldg(&ls.pts[0][list ? list[q_begin + t] : q_begin + t]);
  if (!list && own_end != 0xFFFFFFFFu && !owns_query(ls, q, own_begin, own_end)) return;
  HeapK hk{heaps + (uint64_t)t * (k + 1), k + 1, 0};
  int level;
  const int R = level_search(ls, q.x, q.y, q.z, k + 1, hk, level);
  if (ls.halo && R > ls.halo) atomicAdd(ls.unsafe, 1u);
  hk.sort_ascending();
  normals_emit(GlobalKeys{hk.h, (int)hk.n, xyz}, q, __float_as_uint(q.w), k, orient, vpx, vpy, vpz,
               route_out(ls, __float_as_uint(q.w), out));
}

constexpr uint32_t kBigChunk = 1u << 18;  // queries per launch of the any-k kernels (heap memory)

// smallest instantiated list size >= need (0 if unsupported)
constexpr int kSizes[] = {1, 2, 4, 6, 8, 11, 13, 17, 21, 25, 31, 33, 40, 48, 64};
inline int pick_size(uint32_t need) {
  for (int s : kSizes)
    if ((uint32_t)s >= need) return s;
  return 0;
}

}  // namespace

// bit 5 (32): staged-tile kernels (tc_tile.cu); bit 6 (64): TMA bulk staging (else LDG/STS);
// bit 7 (128): Newton eigen solver with Jacobi fallback
int g_tc_search_flags = 159;  // per-lane kernels + Newton; the staged-tile variant measured 13-17 % slower
#ifdef TC_QCLOCK
static uint32_t* g_tc_dbg = nullptr;  // optional per-query {cycles, R or n} buffer (8 u32 / point)
#endif
extern "C" void tc_debug_set_search_flags(int flags) { g_tc_search_flags = flags; }
extern "C" void tc_debug_set_query_clock_buffer(void* d_buf) {
#ifdef TC_QCLOCK
  g_tc_dbg = (uint32_t*)d_buf;
#else
  (void)d_buf;  // release build: no clock plumbing in the kernels (build with --qclock)
#endif
}

#define TC_DISPATCH_K(SZ, CALL)                     \
  switch (SZ) {                                     \
    case 1: { constexpr int KK = 1; CALL; } break;   \
    case 2: { constexpr int KK = 2; CALL; } break;   \
    case 4: { constexpr int KK = 4; CALL; } break;   \
    case 6: { constexpr int KK = 6; CALL; } break;   \
    case 8: { constexpr int KK = 8; CALL; } break;   \
    case 11: { constexpr int KK = 11; CALL; } break; \
    case 13: { constexpr int KK = 13; CALL; } break; \
    case 17: { constexpr int KK = 17; CALL; } break; \
    case 21: { constexpr int KK = 21; CALL; } break; \
    case 25: { constexpr int KK = 25; CALL; } break; \
    case 31: { constexpr int KK = 31; CALL; } break; \
    case 33: { constexpr int KK = 33; CALL; } break; \
    case 40: { constexpr int KK = 40; CALL; } break; \
    case 48: { constexpr int KK = 48; CALL; } break; \
    case 64: { constexpr int KK = 64; CALL; } break; \
    default: break;                                 \
  }

// flags bit 2 (value 4): two-pass float selection (k_knn2 / k_normals2) when need <= 33.
// flags bit 3 (value 8): need == 17 / 33 use the 16- / 32-wide list plus one scalar slot.
// Returns the selection shape (list slots) for `need`, 0 when the two-pass kernels do not apply.
static int two_pass_shape(uint32_t need, int flags) {
  if (!(flags & 4) || need < 2) return 0;
  if (need <= 16) return 16;
  if (need == 17 && (flags & 8)) return 17;
  if (need <= 32) return 32;
  if (need == 33 && (flags & 8)) return 33;
  return 0;
}
int tci_knn_launch(tc_context* ctx, const tc_index* ix, const float4* d_queries_sorted,
                   uint64_t q_begin, uint64_t q_end, uint32_t k, int exclude_self, bool self_query,
                   uint32_t* d_idx_out, float* d_dist_out, uint32_t* d_count_out) {
  if (q_end <= q_begin) return TC_OK;
  TcRange nvtx_range("tc:knn");
  const int drop_self = (exclude_self && self_query) ? 1 : 0;
  const uint32_t need = k + (drop_self ? 1u : 0u);
  const int sz = pick_size(need);
  const uint32_t nq = (uint32_t)(q_end - q_begin);
  const dim3 grid((nq + kBlock - 1) / kBlock);
  int flags = g_tc_search_flags;
  const bool two_pass = two_pass_shape(need, flags) != 0;
  if (two_pass) flags |= 2;  // pruning is always worth it with the batch scan
  const LevelSet ls = ix->level_set(flags);
  if (self_query) d_queries_sorted = ix->lv[0].d_pts;
  if (sz == 0) {  // beyond the register lists: global-memory heaps, a chunk of queries at a time
    const uint32_t chunk = std::min(nq, kBigChunk);
    uint64_t* d_heaps = nullptr;
    TC_TRY(tc_alloc(ctx, &d_heaps, (uint64_t)chunk * need));
    for (uint64_t b = q_begin; b < q_end; b += chunk) {
      const uint64_t e = std::min<uint64_t>(q_end, b + chunk);
      k_knn_big<<<(uint32_t)((e - b + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(
          ls, d_queries_sorted, (uint32_t)b, (uint32_t)e, k, need, drop_self, d_idx_out,
          d_dist_out, d_count_out, d_heaps);
      ctx->launches++;
    }
    tc_free(ctx, d_heaps);
    TC_CUDA(ctx, cudaGetLastError());
    return TC_OK;
  }
  if (!two_pass) {
    TC_DISPATCH_K(sz, (k_knn<KK><<<grid, kBlock, 0, ctx->stream>>>(
                          ls, d_queries_sorted, (uint32_t)q_begin, (uint32_t)q_end, k, need,
                          drop_self, d_idx_out, d_dist_out, d_count_out, nullptr, nullptr)));
    TC_LAUNCHED(ctx);
    return TC_OK;
  }
  // (a slab-sharded index keeps the per-lane kernel: its ring cap lives in grid_search)
  const bool tile = (flags & 32) && !ix->sharded;
  // the per-lane kernels resolve every query themselves (ties included, resolve_ties); only the
  // staged-tile variant hands what it cannot prove to a list for the chain kernel
  uint32_t* d_fb = nullptr;  // [0] = count, [1..] = query positions
  if (tile) {
    TC_TRY(tc_ws_get(ctx, 2, &d_fb, (uint64_t)nq + 1));
    TC_CUDA(ctx, cudaMemsetAsync(d_fb, 0, sizeof(uint32_t), ctx->stream));
  }
#define TC_KNN2(LL, XX)                                                                       \
  k_knn2<LL, XX><<<grid, kBlock, 0, ctx->stream>>>(ls, d_queries_sorted, (uint32_t)q_begin,   \
                                                   (uint32_t)q_end, k, need, drop_self,       \
                                                   d_idx_out, d_dist_out, d_count_out, d_fb + 1, d_fb)
  if (tile) {  // staged-tile kernel (tc_tile.cu)
    uint32_t* d_stats = nullptr;
    ctx->stats_queries = nq;
    if (ctx->stats_on) {
      d_stats = ctx->d_scratch + 48;
      TC_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 8 * sizeof(uint32_t), ctx->stream));
    }
    TC_TRY(tci_tile_knn(ctx, ls, two_pass_shape(need, flags), d_queries_sorted, (uint32_t)q_begin,
                        (uint32_t)q_end, k, need, drop_self, d_idx_out, d_dist_out, d_count_out,
                        d_fb + 1, d_fb, d_stats, flags));
  } else {
    switch (two_pass_shape(need, flags)) {
      case 16: TC_KNN2(16, false); break;
      case 17: TC_KNN2(16, true); break;
      case 32: TC_KNN2(32, false); break;
      default: TC_KNN2(32, true); break;
    }
    TC_LAUNCHED(ctx);
    return TC_OK;
  }
#undef TC_KNN2
  const dim3 fgrid(std::min<uint32_t>((nq + kBlock - 1) / kBlock, (uint32_t)ctx->sm_count));
  TC_DISPATCH_K(sz, (k_knn<KK><<<fgrid, kBlock, 0, ctx->stream>>>(
                        ls, d_queries_sorted, 0u, 0u, k, need, drop_self, d_idx_out, d_dist_out,
                        d_count_out, d_fb + 1, d_fb)));
  TC_LAUNCHED(ctx);
  tc_ws_release(ctx, 2, d_fb);
  return TC_OK;
}

// Normals for the shard [q_begin, q_end) of the level-0 sorted order.  Shards own whole cells
// (owns_query), so the launch covers up to max_pop extra positions past q_end.
int tci_normals_launch(tc_context* ctx, const tc_index* ix, uint32_t k, int orient,
                       const float vp[3], uint64_t q_begin, uint64_t q_end, float* d_out_aos,
                       bool exact_range) {
  if (q_end <= q_begin) return TC_OK;
  TcRange nvtx_range("tc:normals");
  const int sz = pick_size(k + 1);
  const bool whole = exact_range || (q_begin == 0 && q_end >= ix->n);
  const uint32_t own_begin = (uint32_t)q_begin;
  const uint32_t own_end = whole ? 0xFFFFFFFFu : (uint32_t)q_end;
  if (!whole) {  // cover every position of a cell that starts inside the shard (exact bound)
    uint32_t max_pop = 0;
    TC_TRY(tci_level0_max_population(ctx, ix, &max_pop));
    q_end = std::min<uint64_t>(ix->n, q_end + max_pop);
  }
  const uint32_t nq = (uint32_t)(q_end - q_begin);
  const dim3 grid((nq + kBlock - 1) / kBlock);
  int flags = g_tc_search_flags;
  const bool two_pass = two_pass_shape(k + 1, flags) != 0;
  if (two_pass) flags |= 2;
  const LevelSet ls = ix->level_set(flags);
  if (sz == 0) {  // k + 1 > 64: global-memory heaps, a chunk of queries at a time
    const uint32_t chunk = std::min(nq, kBigChunk);
    uint64_t* d_heaps = nullptr;
    TC_TRY(tc_alloc(ctx, &d_heaps, (uint64_t)chunk * (k + 1)));
    for (uint64_t b = q_begin; b < q_end; b += chunk) {
      const uint64_t e = std::min<uint64_t>(q_end, b + chunk);
      k_normals_big<<<(uint32_t)((e - b + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(
          ls, ix->cloud->d_xyz, (uint32_t)b, (uint32_t)e, own_begin, own_end, k, orient, vp[0],
          vp[1], vp[2], d_out_aos, d_heaps);
      ctx->launches++;
    }
    tc_free(ctx, d_heaps);
    TC_CUDA(ctx, cudaGetLastError());
    return TC_OK;
  }
  if (!two_pass) {
    TC_DISPATCH_K(sz, (k_normals<KK><<<grid, kBlock, 0, ctx->stream>>>(
                          ls, ix->cloud->d_xyz, (uint32_t)q_begin, (uint32_t)q_end, own_begin,
                          own_end, k, orient, vp[0], vp[1], vp[2], d_out_aos, nullptr,
                          nullptr TC_DBG_ARG(g_tc_dbg))));
    TC_LAUNCHED(ctx);
    return TC_OK;
  }
  // (a slab-sharded index keeps the per-lane kernel: its ring cap lives in grid_search)
  const bool tile = (flags & 32) && !ix->sharded;
  uint32_t* d_fb = nullptr;  // the staged-tile variant's list for the chain kernel (see tci_knn_launch)
  if (tile) {
    TC_TRY(tc_ws_get(ctx, 2, &d_fb, (uint64_t)nq + 1));
    TC_CUDA(ctx, cudaMemsetAsync(d_fb, 0, sizeof(uint32_t), ctx->stream));
  }
#define TC_NORMALS2(LL, XX)                                                                  \
  k_normals2<LL, XX><<<grid, kBlock, 0, ctx->stream>>>(                                      \
      ls, (uint32_t)q_begin, (uint32_t)q_end, own_begin, own_end, k, orient, vp[0], vp[1],   \
      vp[2], d_out_aos, d_fb + 1, d_fb TC_DBG_ARG(g_tc_dbg))
  if (tile) {  // staged-tile kernel (tc_tile.cu)
    uint32_t* d_stats = nullptr;
    ctx->stats_queries = nq;
    if (ctx->stats_on) {
      d_stats = ctx->d_scratch + 48;
      TC_CUDA(ctx, cudaMemsetAsync(d_stats, 0, 8 * sizeof(uint32_t), ctx->stream));
    }
    TC_TRY(tci_tile_normals(ctx, ls, two_pass_shape(k + 1, flags), (uint32_t)q_begin,
                            (uint32_t)q_end, own_begin, own_end, k, orient, vp, d_out_aos, d_fb + 1,
                            d_fb, d_stats, flags));
  } else {
    switch (two_pass_shape(k + 1, flags)) {
      case 16: TC_NORMALS2(16, false); break;
      case 17: TC_NORMALS2(16, true); break;
      case 32: TC_NORMALS2(32, false); break;
      default: TC_NORMALS2(32, true); break;
    }
    TC_LAUNCHED(ctx);
    return TC_OK;
  }
#undef TC_NORMALS2
  const dim3 fgrid(std::min<uint32_t>((nq + kBlock - 1) / kBlock, (uint32_t)ctx->sm_count));
  TC_DISPATCH_K(sz, (k_normals<KK><<<fgrid, kBlock, 0, ctx->stream>>>(
                        ls, ix->cloud->d_xyz, 0u, 0u, 0u, 0xFFFFFFFFu, k, orient, vp[0], vp[1], vp[2],
                        d_out_aos, d_fb + 1, d_fb TC_DBG_ARG(nullptr))));
  TC_LAUNCHED(ctx);
  tc_ws_release(ctx, 2, d_fb);
  return TC_OK;
}

// Radius-mode normals for the shard [q_begin, q_end) of the level-0 sorted order (whole cells are
// owned, as in tci_normals_launch).  Queries are taken a chunk at a time so the key segments of a
// chunk fit a bounded buffer.
int tci_normals_radius_launch(tc_context* ctx, const tc_index* ix, float radius, uint32_t k,
                              int orient, const float vp[3], uint64_t q_begin, uint64_t q_end,
                              float* d_out_aos) {
  if (q_end <= q_begin || ix->n == 0) return TC_OK;
  TcRange nvtx_range("tc:normals (radius)");
  const bool whole = (q_begin == 0 && q_end >= ix->n);
  const uint32_t own_begin = (uint32_t)q_begin;
  const uint32_t own_end = whole ? 0xFFFFFFFFu : (uint32_t)q_end;
  if (!whole) {
    uint32_t max_pop = 0;
    TC_TRY(tci_level0_max_population(ctx, ix, &max_pop));
    q_end = std::min<uint64_t>(ix->n, q_end + max_pop);
  }
  const LevelSet ls = ix->level_set(g_tc_search_flags | 2);
  // the level whose cell edge is closest to (and not far below) the radius keeps the box small
  int level = 0;
  for (int l = 0; l < ix->n_levels; ++l)
    if (ix->lv[l].g.cell <= 2.0f * radius) level = l;
  const uint32_t nq = (uint32_t)(q_end - q_begin);
  constexpr uint32_t kChunk = 1u << 20;
  const uint32_t chunk = std::min(nq, kChunk);
  uint32_t *d_fb = nullptr, *d_cnt = nullptr;
  int st = tc_alloc(ctx, &d_fb, (uint64_t)nq + 1);
  if (st == TC_OK) st = tc_alloc(ctx, &d_cnt, (uint64_t)chunk + 1);
  if (st == TC_OK && cudaMemsetAsync(d_fb, 0, sizeof(uint32_t), ctx->stream) != cudaSuccess)
    st = tc_fail(ctx, TC_GPU, "radius normals: memset failed");
  for (uint64_t b = q_begin; b < q_end && st == TC_OK; b += chunk) {
    const uint32_t e = (uint32_t)std::min<uint64_t>(q_end, b + chunk), m = e - (uint32_t)b;
    const dim3 grid((m + kBlock - 1) / kBlock);
    k_radius_count<<<grid, kBlock, 0, ctx->stream>>>(ls, level, (uint32_t)b, e, own_begin, own_end,
                                                     radius, k, d_cnt, d_fb + 1, d_fb);
    ctx->launches++;
    st = tci_exclusive_scan_u32(ctx, d_cnt, d_cnt, m);  // d_cnt[m] = total keys of the chunk
    uint32_t total = 0;
    if (st == TC_OK) {
      cudaError_t err = cudaMemcpyAsync(&total, d_cnt + m, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                        ctx->stream);
      if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
      if (err != cudaSuccess) st = tc_fail(ctx, TC_GPU, "radius normals: count readback failed");
    }
    if (st != TC_OK || total == 0) continue;
    uint64_t* d_keys = nullptr;
    st = tc_alloc(ctx, &d_keys, total);
    if (st == TC_OK) {
      k_radius_normals<<<grid, kBlock, 0, ctx->stream>>>(ls, level, ix->cloud->d_xyz, (uint32_t)b, e,
                                                         radius, orient, vp[0], vp[1], vp[2], d_cnt,
                                                         d_keys, d_out_aos);
      ctx->launches++;
    }
    tc_free(ctx, d_keys);
  }
  // queries with fewer than k neighbours in the radius: the kNN rule, any k
  uint32_t n_fb = 0;
  if (st == TC_OK) {
    cudaError_t err = cudaMemcpyAsync(&n_fb, d_fb, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    if (err != cudaSuccess) st = tc_fail(ctx, TC_GPU, "radius normals: fallback readback failed");
  }
  if (st == TC_OK && n_fb > 0) {
    const int sz = pick_size(k + 1);
    if (sz != 0) {
      const dim3 fgrid(std::min<uint32_t>((n_fb + kBlock - 1) / kBlock, (uint32_t)ctx->sm_count * 4));
      TC_DISPATCH_K(sz, (k_normals<KK><<<fgrid, kBlock, 0, ctx->stream>>>(
                            ls, ix->cloud->d_xyz, 0u, 0u, 0u, 0xFFFFFFFFu, k, orient, vp[0], vp[1],
                            vp[2], d_out_aos, d_fb + 1, d_fb TC_DBG_ARG(nullptr))));
      ctx->launches++;
    } else {  // k + 1 beyond the register lists: global-memory heaps over the list, in chunks
      const uint32_t hchunk = std::min(n_fb, kBigChunk);
      uint64_t* d_heaps = nullptr;
      st = tc_alloc(ctx, &d_heaps, (uint64_t)hchunk * (k + 1));
      for (uint32_t b = 0; b < n_fb && st == TC_OK; b += hchunk) {
        const uint32_t e = std::min(n_fb, b + hchunk);
        k_normals_big<<<(e - b + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(
            ls, ix->cloud->d_xyz, b, e, 0u, 0xFFFFFFFFu, k, orient, vp[0], vp[1], vp[2], d_out_aos,
            d_heaps, d_fb + 1);
        ctx->launches++;
      }
      tc_free(ctx, d_heaps);
    }
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "radius normals launch failed");
  }
  tc_free(ctx, d_cnt);
  tc_free(ctx, d_fb);
  return st;
}

int tci_radius_search_launch(tc_context* ctx, const tc_index* ix, const float q[3], float radius,
                             uint32_t* d_idx, float* d_d2, uint32_t capacity, uint32_t* d_count) {
  const LevelSet ls = ix->level_set(0);
  int level = 0;
  for (int l = 0; l < ix->n_levels; ++l)
    if (ix->lv[l].g.cell <= 2.0f * radius) level = l;
  k_radius_search<<<1, 256, 0, ctx->stream>>>(ls, level, q[0], q[1], q[2], radius, d_idx, d_d2,
                                              capacity, d_count);
  TC_LAUNCHED(ctx);
  return TC_OK;
}
