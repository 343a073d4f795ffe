// tc_tile.cu — staged-tile kNN / normals kernels (the hot kernels of the path).
//
// Replaces the per-query loops of KdTree::find_k_nearest (nearest_neighbor.rs:177-251),
// PointCloudNeighbors::k_nearest_neighbors (point_cloud_ops.rs:80-105) and the rayon body of
// estimate_normals_with_config (normals.rs:306-354).
//
// One WARP owns 32 consecutive queries of the cell-sorted order.  Because cell ids are row-major,
// the cells those queries live in, dilated by one cell, are a handful of CONTIGUOUS runs of the
// sorted point array (one per (y, z) row of the box).  The warp
//   1. finds the box (warp min/max of the lanes' cell coordinates), lets lane r fetch the
//      cell_start pair of row r, prefix-sums the run lengths;
//   2. stages every run into shared memory with one 1-D TMA bulk copy per row
//      (cp.async.bulk.shared.global + mbarrier complete_tx; `UBLKCP` in SASS), rows that contain
//      queries first so the K-th distance tightens early;
//   3. scans the tile in a warp-UNIFORM loop: every lane reads the same candidate (one broadcast
//      LDS.128), evaluates the reference's exact f32 squared distance to ITS query, and, when the
//      candidate is at or below the lane's current K-th distance, appends the distance to a
//      16-entry per-lane column in shared memory and sets the candidate's bit in a per-lane mask;
//      when any lane's column is more than half full the warp merges the columns into the sorted
//      register lists (sort16 + bitonic merge) - ~7 merges per 300 candidates instead of one per
//      batch;
//   4. proves exactness per lane (K-th distance below the distance to the box faces, computed
//      from the same f32 cell coordinates that binned the points); lanes that fail retry with a
//      wider box / a coarser level through a small warp-level work stack, boxes that do not fit
//      the tile are split in halves, and only what still fails goes to the exact chain kernel;
//   5. re-visits ONLY the masked candidates to collect the members (d2 <= tau), ranks them by
//      binary search in the sorted distance list (shared memory), and runs the epilogue:
//      kNN rows, or centroid / covariance in the reference's summation order -> eigenvector ->
//      orientation -> NormalPoint3f.
// No lane ever walks its own candidate list: loops are uniform, loads are broadcasts, and the
// 18 table look-ups + bound arithmetic per query of the per-lane kernels (tc_search.cu) are paid
// once per warp.
#include "tc_normal.cuh"
#include "tc_search.cuh"

namespace {

using namespace tcs;

constexpr int kTileWarps = 1;  // a block's shared memory and registers return as soon as ITS warp is
                                // done (warps differ 2-3x in rounds); nothing is shared between warps
constexpr int kTileBlock = kTileWarps * 32;
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kRunGap = 4;  // lanes whose cells are farther apart in x start a new run

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// 1-D TMA bulk copy global -> shared; completion is signalled on the mbarrier as `bytes` of tx
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar),
      "r"(phase)
      : "memory");
}

// ------------------------------------------------------------------------- per-warp shared memory
template <int L, bool X, int CAP>
struct TileLayout {
  static constexpr int T = L + (X ? 1 : 0);   // list slots = most members a query can have
  static constexpr int kTile = 0;             // float4[CAP + 8]  (8 NaN pad entries)
  static constexpr int kBar = 16 * (CAP + 8); // mbarrier (8 B, padded to 16)
  static constexpr int kU = kBar + 16;        // union: float buf[16][32] (scan) | float vs[T][32]
  static constexpr int kUBytesA = 2048;
  static constexpr int kUBytesB = 128 * ((T + 1) & ~1);
  static constexpr int kUBytes = kUBytesA > kUBytesB ? kUBytesA : kUBytesB;
  static constexpr int kMemb = kU + kUBytes;  // u16 memb[T][32]
  static constexpr int kMembBytes = 64 * ((T + 1) & ~1);
  static constexpr int kOrd = kMemb + kMembBytes;  // u16 ord[T][32]
  static constexpr int kStack = kOrd + kMembBytes; // u32 stack[16][2]
  static constexpr int kWarpBytes = kStack + 128;
  static_assert(kWarpBytes % 16 == 0, "per-warp region must keep 16-byte alignment");
  static_assert(CAP % 32 == 0 && CAP <= 65536 - 64, "positions are stored as u16");
};

struct TileStats {  // device counters (tc_last_stats): [0] queries sent to the chain kernel,
  uint32_t* p;      // [1] rounds, [2] box splits, [3] retries (wider box / coarser level),
};                  // [4] candidates staged (sum), [5] merges (sum over warps)

enum { kModeNormals = 0, kModeKnn = 1 };

struct TileArgs {
  const float4* queries;   // sorted queries (self query: the level-0 array)
  uint32_t q_begin, q_end;
  uint32_t own_begin, own_end;  // shard ownership (own_end == 0xFFFFFFFF: everything)
  uint32_t k, need;
  int drop_self;           // kNN: remove the query's own index (self query + exclude_self)
  int orient;
  float vpx, vpy, vpz;
  float* out;              // normals: n x 6 by original index
  uint32_t* idx_out;       // kNN
  float* dist_out;
  uint32_t* count_out;
  uint32_t* fb_list;       // queries (positions in `queries`) for the exact chain kernel
  uint32_t* fb_count;
  uint32_t* stats;         // may be null
  int flags;               // bit 6: TMA staging (else LDG/STS), bit 7: Newton eigen solver
};

// Shard ownership (multi-GPU): see tc_search.cu owns_query.
__device__ __forceinline__ bool tile_owns(const LevelSet& ls, const float4& q, uint32_t begin,
                                          uint32_t end) {
  const GridParams& g = ls.g[0];
  float u;
  const int cx = cell_coord(q.x, g.ox, g.inv, g.nx, u);
  const int cy = cell_coord(q.y, g.oy, g.inv, g.ny, u);
  const int cz = cell_coord(q.z, g.oz, g.inv, g.nz, u);
  const uint32_t first = __ldg(&ls.cs[0][cell_id(g, cx, cy, cz)]);
  return first >= begin && first < end;
}

__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(kFull, v); }
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(kFull, v); }

template <int L, bool X, int MODE, int CAP>
__global__ void __launch_bounds__(kTileBlock, 13)
k_tile(const __grid_constant__ LevelSet ls, const __grid_constant__ TileArgs a) {
  using LY = TileLayout<L, X, CAP>;
  constexpr int T = LY::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ws = smem_raw + (size_t)warp * LY::kWarpBytes;
  float4* tile = reinterpret_cast<float4*>(ws + LY::kTile);
  float* buf = reinterpret_cast<float*>(ws + LY::kU);                 // [16][32]
  float* vs = reinterpret_cast<float*>(ws + LY::kU);                  // [T][32] (aliases buf)
  uint16_t* memb = reinterpret_cast<uint16_t*>(ws + LY::kMemb);       // [T][32]
  uint16_t* ord = reinterpret_cast<uint16_t*>(ws + LY::kOrd);         // [T][32]
  uint32_t* stack = reinterpret_cast<uint32_t*>(ws + LY::kStack);     // [16][2]
  const uint32_t bar = smem_u32(ws + LY::kBar);
  const uint32_t tile_s = smem_u32(tile);
  const bool use_tma = (a.flags & 64) != 0;

  const uint32_t q0 = a.q_begin + (blockIdx.x * kTileWarps + warp) * 32;
  if (q0 >= a.q_end) return;  // whole warp (no block-level barrier is used below)
  const uint32_t qi = q0 + lane;
  bool active = qi < a.q_end;
  const float4 q = __ldg(&a.queries[active ? qi : a.q_end - 1]);
  if (active && a.own_end != 0xFFFFFFFFu) active = tile_owns(ls, q, a.own_begin, a.own_end);
  if (!__any_sync(kFull, active)) return;
  const uint32_t qid = __float_as_uint(q.w);
  const uint32_t need = a.need;

  if (lane == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
  }
  __syncwarp();
  uint32_t phase = 0;

  // ---- level choice per lane: the finest level whose 3x3x3 block holds `need` points ----------
  int l_own = 0;
  if (ls.n > 1) {
    for (; l_own < ls.n - 1; ++l_own)
      if (block_population(ls.g[l_own], ls.cs[l_own], q.x, q.y, q.z) >= need) break;
  }
  // work stack: (lane mask, level | dilation << 8); one entry per level present in the warp
  int sp = 0;
  for (int l = ls.n - 1; l >= 0; --l) {  // pushed coarse -> fine so the fine rounds run first
    const uint32_t m = __ballot_sync(kFull, active && l_own == l);
    if (m) {
      if (lane == 0) {
        stack[2 * sp] = m;
        stack[2 * sp + 1] = (uint32_t)l | (1u << 8);
      }
      ++sp;
    }
  }
  __syncwarp();
  uint32_t fb_mask = 0;   // lanes that go to the exact chain kernel
  uint32_t st_rounds = 0, st_splits = 0, st_retries = 0, st_cands = 0, st_merges = 0;

  while (sp > 0) {
    --sp;
    const uint32_t rmask = stack[2 * sp];
    const uint32_t rinfo = stack[2 * sp + 1];
    __syncwarp();
    const int level = (int)(rinfo & 0xffu);
    const int dil = (int)(rinfo >> 8);
    const bool in_round = (rmask >> lane) & 1u;
    const GridParams& g = ls.g[level];
    const float4* __restrict__ pts = ls.pts[level];
    const uint32_t* __restrict__ cs = ls.cs[level];
    ++st_rounds;

    // ---- 1. runs and boxes -------------------------------------------------------------------
    // The round's lanes are clustered into RUNS: consecutive lanes (sorted order) whose cells are
    // within kRunGap cells in x and one cell in y and z.  A tilted surface leaves a (y, z) row
    // every few cells, so a warp of the sorted order often holds the end of one x-run and the
    // start of the next, tens of cells apart: one bounding box would span the gap.  Instead the
    // round takes the first TWO runs, each with its own dilated box (B clipped against A where
    // they share a row, so no candidate is staged twice); later runs go back on the stack.
    float ux, uy, uz;
    const int cx = cell_coord(q.x, g.ox, g.inv, g.nx, ux);
    const int cy = cell_coord(q.y, g.oy, g.inv, g.ny, uy);
    const int cz = cell_coord(q.z, g.oz, g.inv, g.nz, uz);
    uint32_t rm = rmask;  // lanes actually processed by this round
    uint32_t maskA, maskB;
    {
      const uint32_t below = rmask & ((1u << lane) - 1u);
      const int plane = below ? (31 - __clz(below)) : lane;  // previous lane of the round
      const int pcx = __shfl_sync(kFull, cx, plane), pcy = __shfl_sync(kFull, cy, plane),
                pcz = __shfl_sync(kFull, cz, plane);
      const bool bnd = in_round && (below == 0u || abs(cx - pcx) > kRunGap || abs(cy - pcy) > 1 ||
                                    abs(cz - pcz) > 1);
      const uint32_t bmask = __ballot_sync(kFull, bnd);  // first lane of every run
      const int nruns = __popc(bmask);
      const uint32_t s1 = nruns >= 2 ? (uint32_t)__fns(bmask, 0, 2) : 32u;  // start of run 1
      const uint32_t s2 = nruns >= 3 ? (uint32_t)__fns(bmask, 0, 3) : 32u;  // start of run 2
      const uint32_t lt1 = s1 >= 32u ? kFull : ((1u << s1) - 1u);
      const uint32_t lt2 = s2 >= 32u ? kFull : ((1u << s2) - 1u);
      maskA = rmask & lt1;
      maskB = dil == 1 ? (rmask & lt2 & ~lt1) : 0u;  // a wider box needs every row slot itself
      const uint32_t rest = rmask & ~(maskA | maskB);
      if (rest) {
        if (sp < 16) {
          if (lane == 0) {
            stack[2 * sp] = rest;
            stack[2 * sp + 1] = rinfo;
          }
          ++sp;
        } else {
          fb_mask |= rest;
        }
        rm = maskA | maskB;
      }
    }
    const bool inA = (maskA >> lane) & 1u, inB = (maskB >> lane) & 1u;
    const bool in_rm = inA || inB;
    constexpr int kBig = 0x3fffffff;
    // box A / box B: bounding boxes of the runs' cells, dilated and clamped
    const int ax0 = max(warp_min(inA ? cx : kBig) - dil, 0), ax1 = min(warp_max(inA ? cx : -kBig) + dil, g.nx - 1);
    const int ay0q = warp_min(inA ? cy : kBig), ay1q = warp_max(inA ? cy : -kBig);
    const int az0q = warp_min(inA ? cz : kBig), az1q = warp_max(inA ? cz : -kBig);
    const int ay0 = max(ay0q - dil, 0), ay1 = min(ay1q + dil, g.ny - 1);
    const int az0 = max(az0q - dil, 0), az1 = min(az1q + dil, g.nz - 1);
    const int anyr = ay1 - ay0 + 1, anr = anyr * (az1 - az0 + 1);
    int bx0 = 0, bx1 = -1, by0q = 0, by1q = -1, bz0q = 0, bz1q = -1, by0 = 0, by1 = -1, bz0 = 0,
        bz1 = -1, bnyr = 1, bnr = 0;
    if (maskB) {  // uniform
      bx0 = max(warp_min(inB ? cx : kBig) - dil, 0);
      bx1 = min(warp_max(inB ? cx : -kBig) + dil, g.nx - 1);
      by0q = warp_min(inB ? cy : kBig);
      by1q = warp_max(inB ? cy : -kBig);
      bz0q = warp_min(inB ? cz : kBig);
      bz1q = warp_max(inB ? cz : -kBig);
      by0 = max(by0q - dil, 0);
      by1 = min(by1q + dil, g.ny - 1);
      bz0 = max(bz0q - dil, 0);
      bz1 = min(bz1q + dil, g.nz - 1);
      bnyr = by1 - by0 + 1;
      bnr = bnyr * (bz1 - bz0 + 1);
    }
    const int nrows = anr + bnr;

    // ---- 2. row runs: lane r owns row r of box A, then of box B ------------------------------
    uint32_t lo = 0, cnt = 0;
    int cls = 1;
    bool weird = false;  // A's interval strictly inside B's on a shared row: not representable
    if (nrows <= 32 && lane < nrows) {
      int y, z, xl, xh;
      if (lane < anr) {
        y = ay0 + lane % anyr;
        z = az0 + lane / anyr;
        xl = ax0;
        xh = ax1;
      } else {
        const int r = lane - anr;
        y = by0 + r % bnyr;
        z = bz0 + r / bnyr;
        xl = bx0;
        xh = bx1;
        if (y >= ay0 && y <= ay1 && z >= az0 && z <= az1 && xl <= ax1 && xh >= ax0) {
          // overlaps A's interval on this row: keep only what A does not already stage
          if (xl >= ax0 && xh <= ax1) xh = xl - 1;      // inside A: nothing left
          else if (xl >= ax0) xl = ax1 + 1;             // sticks out on the right
          else if (xh <= ax1) xh = ax0 - 1;             // sticks out on the left
          else weird = true;                            // sticks out on both sides
        }
      }
      if (xl <= xh) {
        const uint32_t row = cell_id(g, 0, y, z);
        lo = __ldg(&cs[row + xl]);
        cnt = __ldg(&cs[row + xh + 1]) - lo;
      }
      // rows that hold queries go first: the K-th distance tightens early
      cls = ((y >= ay0q && y <= ay1q && z >= az0q && z <= az1q) ||
             (y >= by0q && y <= by1q && z >= bz0q && z <= bz1q)) ? 0 : 1;
    }
    uint32_t c0 = cls == 0 ? cnt : 0u, c1 = cls == 1 ? cnt : 0u;
    uint32_t i0 = c0, i1 = c1;  // inclusive scans
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t0 = __shfl_up_sync(kFull, i0, o), t1 = __shfl_up_sync(kFull, i1, o);
      if (lane >= o) {
        i0 += t0;
        i1 += t1;
      }
    }
    const uint32_t total0 = __shfl_sync(kFull, i0, 31), total1 = __shfl_sync(kFull, i1, 31);
    const uint32_t total = total0 + total1;
    const uint32_t off = cls == 0 ? (i0 - c0) : (total0 + i1 - c1);

    if (nrows > 32 || total > (uint32_t)CAP || __any_sync(kFull, weird)) {
      // does not fit.  Two runs: take them one at a time.  One run: halve it (lanes are in sorted
      // order, so the halves are spatially coherent); a single query that does not fit goes to
      // the chain kernel.
      uint32_t first = maskA, second = maskB;
      if (!maskB) {
        const int np = __popc(maskA);
        if (np > 1) {
          const int cut = __fns(maskA, 0, np / 2 + 1);  // position of the (np/2 + 1)-th set bit
          first = maskA & ((1u << cut) - 1u);
          second = maskA & ~first;
        } else {
          first = 0u;
        }
      }
      if (first && sp + 2 <= 16) {
        if (lane == 0) {
          stack[2 * sp] = second;
          stack[2 * sp + 1] = rinfo;
          stack[2 * sp + 2] = first;
          stack[2 * sp + 3] = rinfo;
        }
        sp += 2;
        ++st_splits;
      } else {
        fb_mask |= rm;
      }
      __syncwarp();
      continue;
    }
    st_cands += total;

    // ---- 3. stage the runs into shared memory ------------------------------------------------
    if (use_tma) {
      fence_proxy_async();  // earlier generic-proxy reads of the tile vs. the async-proxy writes
      __syncwarp();
      if (total > 0) {
        if (lane == 0) mbar_expect_tx(bar, total * 16u);
        __syncwarp();
        if (cnt > 0) bulk_g2s(tile_s + off * 16u, pts + lo, cnt * 16u, bar);
      }
    } else {
      for (int r = 0; r < nrows; ++r) {
        const uint32_t rlo = __shfl_sync(kFull, lo, r), rcnt = __shfl_sync(kFull, cnt, r),
                       roff = __shfl_sync(kFull, off, r);
        for (uint32_t j = lane; j < rcnt; j += 32) tile[roff + j] = __ldg(&pts[rlo + j]);
      }
    }
    // NaN pad behind the last candidate: the scan reads whole groups of 8, and a NaN distance
    // fails every comparison
    if (lane < 8) tile[total + lane] = make_float4(NAN, NAN, NAN, 0.0f);
    // per-lane columns: empty merge buffer
#pragma unroll
    for (int t = 0; t < 16; ++t) buf[t * 32 + lane] = INFINITY;
    if (use_tma && total > 0) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    }
    __syncwarp();

    // ---- 4. uniform scan ---------------------------------------------------------------------
    // lanes outside the round carry a NaN query: nothing ever passes their filter
    const float sx = in_rm ? q.x : NAN, sy = q.y, sz = q.z;
    SelF<L, X> sel;
    sel.pad = T - (int)need;
    sel.init();
    const uint32_t bcol = smem_u32(buf + lane);  // this lane's column of the merge buffer
    uint32_t bptr = bcol;                        // next free slot (shared-memory byte address)
    auto flush = [&]() {
      float b[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) b[t] = buf[t * 32 + lane];
#pragma unroll
      for (int t = 0; t < 16; ++t) buf[t * 32 + lane] = INFINITY;
      bptr = bcol;
      sel.merge(b);
      ++st_merges;
    };
    // groups of 8 candidates; one extra trip performs the final merge (a single flush site keeps
    // the unrolled networks out of the instruction cache's way)
    const uint32_t ngroups = (total + 7) >> 3;
#pragma unroll 1
    for (uint32_t gi = 0; gi <= ngroups; ++gi) {
      const bool last = gi == ngroups;
      if (__any_sync(kFull, last ? (bptr != bcol) : (bptr > bcol + 8 * 128))) flush();
      if (last) break;
      const float tau = sel.kth();
      float4 c[8];  // all eight broadcast loads in flight before the first distance
#pragma unroll
      for (int u = 0; u < 8; ++u) c[u] = tile[gi * 8 + u];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float d2 = dist2_exact(c[u].x, c[u].y, c[u].z, sx, sy, sz);
        if (d2 <= tau) {  // <=: an equal distance may still win on the index (tie rule)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bptr), "f"(d2) : "memory");
          bptr += 128;
        }
      }
    }
    const float tau = sel.kth();  // +inf when fewer than `need` candidates exist
    const bool full = tau < INFINITY;

    // ---- 5. exactness: K-th distance below the distance to every open face of the box --------
    float bound = INFINITY;
    {
      // the lane's own run's box is fully staged (B's clipped parts are inside A's rows)
      const int X0 = inB ? bx0 : ax0, X1 = inB ? bx1 : ax1, Y0 = inB ? by0 : ay0,
                Y1 = inB ? by1 : ay1, Z0 = inB ? bz0 : az0, Z1 = inB ? bz1 : az1;
      const float mx = g.ex + fabsf(q.x - g.ox), my = g.ey + fabsf(q.y - g.oy),
                  mz = g.ez + fabsf(q.z - g.oz);
      if (X0 > 0) bound = fminf(bound, axis_bound(ux - (float)X0, g.cell, mx));
      if (X1 < g.nx - 1) bound = fminf(bound, axis_bound((float)(X1 + 1) - ux, g.cell, mx));
      if (Y0 > 0) bound = fminf(bound, axis_bound(uy - (float)Y0, g.cell, my));
      if (Y1 < g.ny - 1) bound = fminf(bound, axis_bound((float)(Y1 + 1) - uy, g.cell, my));
      if (Z0 > 0) bound = fminf(bound, axis_bound(uz - (float)Z0, g.cell, mz));
      if (Z1 < g.nz - 1) bound = fminf(bound, axis_bound((float)(Z1 + 1) - uz, g.cell, mz));
    }
    const bool exact = (bound == INFINITY) || (full && tau < bound * bound * 0.99999f);
    const bool ok = in_rm && exact;
    {
      // retries: not even `need` candidates -> next coarser level; full but unproven -> wider box
      const bool fail = in_rm && !exact;
      const bool coarser = fail && !full && level < ls.n - 1;
      const bool wider = fail && !coarser && dil < 2;
      const uint32_t mc = __ballot_sync(kFull, coarser), mw = __ballot_sync(kFull, wider);
      fb_mask |= __ballot_sync(kFull, fail && !coarser && !wider);
      if (mc) {
        if (sp < 16) {
          if (lane == 0) {
            stack[2 * sp] = mc;
            stack[2 * sp + 1] = (uint32_t)(level + 1) | (1u << 8);
          }
          ++sp;
          ++st_retries;
        } else {
          fb_mask |= mc;
        }
      }
      if (mw) {
        if (sp < 16) {
          if (lane == 0) {
            stack[2 * sp] = mw;
            stack[2 * sp + 1] = (uint32_t)level | ((uint32_t)(dil + 1) << 8);
          }
          ++sp;
          ++st_retries;
        } else {
          fb_mask |= mw;
        }
      }
    }
    if (!__any_sync(kFull, ok)) {
      __syncwarp();
      continue;
    }

    // ---- 6. members: a second uniform pass keeps the candidates with d2 <= tau ----------------
    uint32_t n = 0;
    {
      const float mqx = ok ? q.x : NAN;  // lanes without a proven result collect nothing
      const uint32_t mcol = smem_u32(memb + lane);
      uint32_t mptr = mcol;
#pragma unroll 1
      for (uint32_t gi = 0; gi < ngroups; ++gi) {
        float4 c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u] = tile[gi * 8 + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float d2 = dist2_exact(c[u].x, c[u].y, c[u].z, mqx, q.y, q.z);
          if (d2 <= tau) {
            if (mptr < mcol + T * 64)
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(mptr), "h"((uint16_t)(gi * 8 + u))
                           : "memory");
            mptr += 64;
          }
        }
      }
      n = (mptr - mcol) >> 6;
    }
    const bool over = ok && n > (uint32_t)T;  // more ties than the table holds -> chain kernel
    fb_mask |= __ballot_sync(kFull, over);
    const bool emit = ok && !over;

    // ---- 7. rank: binary search in the sorted distance list (shared memory column) -----------
    __syncwarp();
#pragma unroll
    for (int i = 0; i < L; ++i) vs[i * 32 + lane] = sel.v[i];
    uint32_t nmax = __reduce_max_sync(kFull, emit ? n : 0u);
    uint64_t seen = 0;
    bool tie = emit && n > need;  // a tie straddling rank `need`
    for (uint32_t mi = 0; mi < nmax; ++mi) {
      if (emit && mi < n) {
        const uint32_t pos = memb[mi * 32 + lane];
        const float4 c = tile[pos];
        const float d2 = dist2_exact(c.x, c.y, c.z, q.x, q.y, q.z);
        int r = 0;  // #{v[i] < d2}; every member has d2 <= kth, so r <= L - 1 (+1 with X)
#pragma unroll
        for (int s = L / 2; s >= 1; s >>= 1)
          if (vs[(r + s - 1) * 32 + lane] < d2) r += s;
        if (vs[r * 32 + lane] < d2) ++r;  // r <= L - 1 here
        r -= sel.pad;
        if (r >= 0 && r < T) {
          if ((seen >> r) & 1u) tie = true;  // two members with bit-equal d2
          seen |= 1ull << r;
          if (r < (int)need) ord[r * 32 + lane] = (uint16_t)pos;
        }
      }
    }
    if (__any_sync(kFull, tie)) {
      // rare: order bit-equal d2 by original index (tie-aware rank), only for the lanes with ties
      for (uint32_t mi = 0; mi < nmax; ++mi) {
        if (tie && mi < n) {
          const uint32_t pos = memb[mi * 32 + lane];
          const float4 c = tile[pos];
          const float d2 = dist2_exact(c.x, c.y, c.z, q.x, q.y, q.z);
          const uint32_t id = __float_as_uint(c.w);
          int r = 0;
          for (uint32_t o = 0; o < n; ++o) {
            const float4 e = tile[memb[o * 32 + lane]];
            const float e2 = dist2_exact(e.x, e.y, e.z, q.x, q.y, q.z);
            if (e2 < d2 || (e2 == d2 && __float_as_uint(e.w) < id)) ++r;
          }
          if (r < (int)need) ord[r * 32 + lane] = (uint16_t)pos;
        }
      }
    }
    if (n > need) n = need;
    nmax = __reduce_max_sync(kFull, emit ? n : 0u);

    // ---- 8. epilogue -------------------------------------------------------------------------
    if (MODE == kModeKnn) {
      // kNN(k+1), retain idx != i, truncate k (point_cloud_ops.rs:91-99); plain kNN otherwise
      const uint32_t k = a.k;
      const uint64_t row = (uint64_t)qid * k;
      uint32_t c = 0;
      for (uint32_t i = 0; i < nmax; ++i) {
        if (emit && i < n) {
          const float4 p = tile[ord[i * 32 + lane]];
          const uint32_t id = __float_as_uint(p.w);
          if (c < k && !(a.drop_self && id == qid)) {
            a.idx_out[row + c] = id;
            if (a.dist_out)
              a.dist_out[row + c] = xsqrt(dist2_exact(p.x, p.y, p.z, q.x, q.y, q.z));
            ++c;
          }
        }
      }
      if (emit) {
        if (a.count_out) a.count_out[qid] = c;
        for (uint32_t j = c; j < k; ++j) {
          a.idx_out[row + j] = TC_NO_INDEX;
          if (a.dist_out) a.dist_out[row + j] = INFINITY;
        }
      }
    } else {
      // neighbourhood = first k of kNN(k+1) with self dropped by index, then self appended last
      // (normals.rs:148-153, 338-340); sums are sequential f32 in that order (normals.rs:165-177)
      const uint32_t k = a.k;
      float mx = 0.0f, my = 0.0f, mz = 0.0f;
      uint32_t cntn = 0;
      for (uint32_t i = 0; i < nmax; ++i) {
        if (emit && i < n) {
          const float4 p = tile[ord[i * 32 + lane]];
          if (cntn < k && __float_as_uint(p.w) != qid) {
            mx = xadd(mx, p.x);
            my = xadd(my, p.y);
            mz = xadd(mz, p.z);
            ++cntn;
          }
        }
      }
      mx = xadd(mx, q.x);
      my = xadd(my, q.y);
      mz = xadd(mz, q.z);
      const uint32_t nn = cntn + 1;
      const float fn = (float)nn;
      const float ccx = xdiv(mx, fn), ccy = xdiv(my, fn), ccz = xdiv(mz, fn);
      float cv[6] = {0, 0, 0, 0, 0, 0};
      auto acc = [&](float px, float py, float pz) {
        const float dx = xsub(px, ccx), dy = xsub(py, ccy), dz = xsub(pz, ccz);
        cv[0] = xadd(cv[0], xmul(dx, dx));
        cv[1] = xadd(cv[1], xmul(dx, dy));
        cv[2] = xadd(cv[2], xmul(dx, dz));
        cv[3] = xadd(cv[3], xmul(dy, dy));
        cv[4] = xadd(cv[4], xmul(dy, dz));
        cv[5] = xadd(cv[5], xmul(dz, dz));
      };
      uint32_t cnt2 = 0;
      for (uint32_t i = 0; i < nmax; ++i) {
        if (emit && i < n) {
          const float4 p = tile[ord[i * 32 + lane]];
          if (cnt2 < k && __float_as_uint(p.w) != qid) {
            acc(p.x, p.y, p.z);
            ++cnt2;
          }
        }
      }
      if (emit) {
        float nrm[3] = {0.0f, 0.0f, 1.0f};  // < 3 points (normals.rs:159-162)
        if (nn >= 3) {
          acc(q.x, q.y, q.z);
#pragma unroll
          for (int i = 0; i < 6; ++i) cv[i] = xdiv(cv[i], fn);
          normal_from_cov(cv, nrm, (a.flags & 128) != 0);
        }
        write_normal(nrm, q, qid, a.orient, a.vpx, a.vpy, a.vpz, a.out);
      }
    }
    __syncwarp();
  }

  // ---- queries left for the exact chain kernel ---------------------------------------------
  if (fb_mask) {
    uint32_t basei = 0;
    if (lane == 0) basei = atomicAdd(a.fb_count, (uint32_t)__popc(fb_mask));
    basei = __shfl_sync(kFull, basei, 0);
    if ((fb_mask >> lane) & 1u)
      a.fb_list[basei + __popc(fb_mask & ((1u << lane) - 1u))] = qi;
  }
  if (a.stats && lane == 0) {
    if (fb_mask) atomicAdd(&a.stats[0], (uint32_t)__popc(fb_mask));
    atomicAdd(&a.stats[1], st_rounds);
    if (st_splits) atomicAdd(&a.stats[2], st_splits);
    if (st_retries) atomicAdd(&a.stats[3], st_retries);
    atomicAdd(&a.stats[4], st_cands);
    atomicAdd(&a.stats[5], st_merges);
  }
}

constexpr int kCap = 512;

template <int L, bool X, int MODE>
int launch_tile(tc_context* ctx, const LevelSet& ls, const TileArgs& a) {
  using LY = TileLayout<L, X, kCap>;
  constexpr int smem = LY::kWarpBytes * kTileWarps;
  static bool configured[64] = {};  // per device; the attribute is per function
  auto kern = k_tile<L, X, MODE, kCap>;
  const int dev = ctx->device & 63;
  if (!configured[dev]) {
    TC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev] = true;
  }
  const uint32_t nq = a.q_end - a.q_begin;
  const uint32_t blocks = (nq + kTileBlock - 1) / kTileBlock;
  kern<<<blocks, kTileBlock, smem, ctx->stream>>>(ls, a);
  TC_LAUNCHED(ctx);
  return TC_OK;
}

}  // namespace

// shape: 16, 17 (16 + scalar slot), 32, 33 (see two_pass_shape in tc_search.cu)
int tci_tile_normals(tc_context* ctx, const LevelSet& ls, int shape, uint32_t q_begin,
                     uint32_t q_end, uint32_t own_begin, uint32_t own_end, uint32_t k, int orient,
                     const float vp[3], float* d_out, uint32_t* d_fb_list, uint32_t* d_fb_count,
                     uint32_t* d_stats, int flags) {
  TileArgs a{};
  a.queries = ls.pts[0];
  a.q_begin = q_begin;
  a.q_end = q_end;
  a.own_begin = own_begin;
  a.own_end = own_end;
  a.k = k;
  a.need = k + 1;
  a.orient = orient;
  a.vpx = vp[0];
  a.vpy = vp[1];
  a.vpz = vp[2];
  a.out = d_out;
  a.fb_list = d_fb_list;
  a.fb_count = d_fb_count;
  a.stats = d_stats;
  a.flags = flags;
  switch (shape) {
    case 16: return launch_tile<16, false, kModeNormals>(ctx, ls, a);
    case 17: return launch_tile<16, true, kModeNormals>(ctx, ls, a);
    case 32: return launch_tile<32, false, kModeNormals>(ctx, ls, a);
    default: return launch_tile<32, true, kModeNormals>(ctx, ls, a);
  }
}

int tci_tile_knn(tc_context* ctx, const LevelSet& ls, int shape, const float4* d_queries,
                 uint32_t q_begin, uint32_t q_end, uint32_t k, uint32_t need, int drop_self,
                 uint32_t* d_idx, float* d_dist, uint32_t* d_count, uint32_t* d_fb_list,
                 uint32_t* d_fb_count, uint32_t* d_stats, int flags) {
  TileArgs a{};
  a.queries = d_queries;
  a.q_begin = q_begin;
  a.q_end = q_end;
  a.own_begin = 0;
  a.own_end = 0xFFFFFFFFu;
  a.k = k;
  a.need = need;
  a.drop_self = drop_self;
  a.idx_out = d_idx;
  a.dist_out = d_dist;
  a.count_out = d_count;
  a.fb_list = d_fb_list;
  a.fb_count = d_fb_count;
  a.stats = d_stats;
  a.flags = flags;
  switch (shape) {
    case 16: return launch_tile<16, false, kModeKnn>(ctx, ls, a);
    case 17: return launch_tile<16, true, kModeKnn>(ctx, ls, a);
    case 32: return launch_tile<32, false, kModeKnn>(ctx, ls, a);
    default: return launch_tile<32, true, kModeKnn>(ctx, ls, a);
  }
}
