// tc_icp.cu — point-to-plane ICP with the whole iteration loop resident on the device.
//
// Replaces icp_point_to_plane_detailed (threecrate-algorithms/src/registration.rs:508-602):
//   transform (serial map, :540-544) + kd-tree 1-NN (rayon, :87-107) + serial gather (:553-565) +
//   serial 6x6 accumulation (:412-428) + Cholesky/LU (:432-438) + compose (:441-449,576) + mse
//   (:453-471) + convergence test (:579-589).
//
// Per iteration two launches and no host round-trip:
//   k_icp_correspond : one thread per source point (grid-stride): T*p with nalgebra's f32 op order,
//                      exact grid 1-NN (tie rule (d2, index)), linearise, accumulate the 29 scalars
//                      (21 AtA + 6 Atb + sum b^2 + n_valid) in f64 registers, warp-shuffle + block
//                      reduce, last block sums the block partials in block order (deterministic).
//   [multi-GPU: NCCL all-reduce of the 29 f64 on the same stream]
//   k_icp_solve      : one thread: 6x6 Cholesky (LU fallback), delta = Rz Ry Rx, T <- delta o T,
//                      mse, convergence flag.  Later iterations early-exit once `done` is set.
#include "tc_search.cuh"

struct tc_comm;
int tci_comm_allreduce(tc_comm* comm, double* d_buf, uint64_t count);  // tc_comm.cu
// peer exchange (tc_comm.cu): returns false when the communicator has no mapped peer buffers
bool tci_comm_peers(tc_comm* comm, double** peers8, int* world, int* rank,
                    unsigned long long* epoch_base, unsigned long long epochs_needed);
void tci_comm_commit_epochs(tc_comm* comm, unsigned long long used);
int g_tc_icp_solve_fused = 1;  // debug: 0 = separate solve launch
int g_tc_icp_keep = 1;         // debug: 0 = search every point every iteration (A/B of the skip)
extern "C" void tc_debug_set_icp_keep(int on) { g_tc_icp_keep = on; }
extern "C" void tc_debug_set_icp_solve_fused(int on) { g_tc_icp_solve_fused = on; }
extern int g_tc_icp_fuse;  // 1 (default): fused tail; 0: separate all-reduce + solve launches

namespace {

using namespace tcs;

constexpr int kIcpBlock = 256;
constexpr int kNumSums = 29;   // point-to-plane: 21 AtA + 6 Atb + sum b^2 + n_valid
constexpr int kNumSumsP2P = 17; // point-to-point: sum s (3) + sum t (3) + sum s t^T (9) + sum |s-t|^2 + n
constexpr int kPlane = 0, kPoint = 1, kGicp = 2;  // kGicp: same 29 sums as kPlane (H, g, sum d^2, n)

struct IcpState {
  float T[7];          // tx,ty,tz, qi,qj,qk,qw
  float prev_mse;
  float mse;
  uint32_t iterations; // iterations executed so far
  int32_t converged;
  int32_t status;      // 0 ok, 2 insufficient correspondences, 3 ill-conditioned
  int32_t done;
  uint32_t ticket;     // block completion counter
  double n_valid;
  // per-iteration history for tc_last_stats (first TC_STATS_MAX_ITERS iterations)
  float mse_hist[TC_STATS_MAX_ITERS];
  double valid_hist[TC_STATS_MAX_ITERS];
};

struct V3 {
  float x, y, z;
};

// Level of a multi-resolution target index for a query whose nearest candidate so far is sqrt(d2)
// away: the finest one on which that is at most four cells (a box or ring search then touches
// tens of rows, not thousands); the coarsest otherwise.
__device__ __forceinline__ int pick_level(const LevelSet& ls, float d2) {
  int l = 0;
  while (l < ls.n - 1 && !(d2 <= 16.0f * ls.g[l].cell * ls.g[l].cell)) ++l;
  return l;
}

// Fused all-reduce over NVLink peer memory (tc_comm peer exchange, tc_comm.cu): every rank's
// exchange buffer is IPC-mapped into all ranks.  slot(parity, src) = base + (parity*world+src)*32
// doubles: [0..28] payload, [31] epoch flag.
struct PeerXchg {
  double* peer[8];     // peer[r] = rank r's buffer as mapped into THIS process (peer[rank] local)
  int world, rank;
  unsigned long long epoch_base;
  unsigned long long spin_limit;  // polls (64 ns apart) before a missing peer fails the call
};
constexpr int kXSlot = 32;

// (not inlined: called once per launch by one thread of the last block; inlined into the
//  correspondence kernel they would set its register count)
__device__ __noinline__ void icp_solve_plane(IcpState* st, const double* sums, float conv);
__device__ __noinline__ void icp_solve_point(IcpState* st, const double* sums, float conv);
__device__ __forceinline__ V3 xcross(const V3& a, const V3& b) {
  return V3{xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)),
            xsub(xmul(a.x, b.y), xmul(a.y, b.x))};
}
// nalgebra UnitQuaternion * Vector3:  t = 2 (qv x v);  (t*w + qv x t) + v   [upstream nalgebra]
__device__ __forceinline__ V3 quat_rotate(const float q[4] /*i,j,k,w*/, const V3& v) {
  const V3 qv{q[0], q[1], q[2]};
  V3 t = xcross(qv, v);
  t = V3{xmul(t.x, 2.0f), xmul(t.y, 2.0f), xmul(t.z, 2.0f)};
  const V3 c = xcross(qv, t);
  return V3{xadd(xadd(xmul(t.x, q[3]), c.x), v.x), xadd(xadd(xmul(t.y, q[3]), c.y), v.y),
            xadd(xadd(xmul(t.z, q[3]), c.z), v.z)};
}
__device__ __forceinline__ void quat_mul(const float a[4], const float b[4], float r[4]) {
  // Hamilton product, left-to-right evaluation, no renormalisation  [upstream nalgebra]
  r[3] = xsub(xsub(xsub(xmul(a[3], b[3]), xmul(a[0], b[0])), xmul(a[1], b[1])), xmul(a[2], b[2]));
  r[0] = xsub(xadd(xadd(xmul(a[3], b[0]), xmul(a[0], b[3])), xmul(a[1], b[2])), xmul(a[2], b[1]));
  r[1] = xadd(xadd(xsub(xmul(a[3], b[1]), xmul(a[0], b[2])), xmul(a[1], b[3])), xmul(a[2], b[0]));
  r[2] = xadd(xsub(xadd(xmul(a[3], b[2]), xmul(a[0], b[1])), xmul(a[1], b[0])), xmul(a[2], b[3]));
}

__global__ void k_icp_init(IcpState* st, const float* init7_src, float t0, float t1, float t2,
                           float q0, float q1, float q2, float q3) {
  (void)init7_src;
  st->T[0] = t0;
  st->T[1] = t1;
  st->T[2] = t2;
  st->T[3] = q0;
  st->T[4] = q1;
  st->T[5] = q2;
  st->T[6] = q3;
  st->prev_mse = INFINITY;
  st->mse = INFINITY;
  st->iterations = 0;
  st->converged = 0;
  st->status = 0;
  st->done = 0;
  st->ticket = 0;
  st->n_valid = 0.0;
}

// target normals AoS (12 B) -> float4 (one aligned 128-bit load per match), original order
__global__ void __launch_bounds__(kIcpBlock)
k_pad_normals(const float* __restrict__ nrm, uint32_t n, float4* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = nrm + 3 * (uint64_t)i;
    out[i] = make_float4(p[0], p[1], p[2], 0.0f);
  }
}

// tgt_nrm: target normals (kPlane) or target covariances, two float4 per point (kGicp);
// src_cov: source covariances, two float4 per point by ORIGINAL source index (kGicp)
template <int MODE>
__global__ void __launch_bounds__(kIcpBlock, MODE == kGicp ? 2 : 4)  // 64 registers (the called
                                                                      // solve may spill freely)
k_icp_correspond(LevelSet ls, const float4* __restrict__ tgt_nrm,
                 const float4* __restrict__ src_cov, const float4* __restrict__ src, uint32_t ns,
                 float max_dist, IcpState* __restrict__ st, double* __restrict__ partials,
                 double* __restrict__ sums, uint32_t* __restrict__ match_out,
                 uint32_t* __restrict__ prev, float4* __restrict__ cache, int fuse, PeerXchg px,
                 int solve_here, float conv) {
  if (st->done) return;
  float T[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T[i] = st->T[i];
  // Sums: the products a_i a_j are f32 in the reference too (registration.rs:426-427).  Each
  // trip of the loop a warp transposes-and-reduces its 32 x NS products with 31 shuffles (lane j
  // ends up with the warp's total of sum j) and adds that to ONE f64 register per lane - instead
  // of carrying NS accumulators per thread through the search, which cost 29 registers and a
  // third of the occupancy.  Everything across warps, blocks and ranks is reduced in f64.
  constexpr int NS = MODE == kPoint ? kNumSumsP2P : kNumSums;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0.0;
  float R[3][3];  // rotation of T (UnitQuaternion::to_rotation_matrix), kGicp only
  if (MODE == kGicp) {
    const float qi = T[3], qj = T[4], qk = T[5], qw = T[6];
    const float ww = xmul(qw, qw), ii = xmul(qi, qi), jj = xmul(qj, qj), kk = xmul(qk, qk);
    const float ij = xmul(xmul(qi, qj), 2.0f), wk = xmul(xmul(qw, qk), 2.0f),
                wj = xmul(xmul(qw, qj), 2.0f), ik = xmul(xmul(qi, qk), 2.0f),
                jk = xmul(xmul(qj, qk), 2.0f), wi = xmul(xmul(qw, qi), 2.0f);
    R[0][0] = xsub(xsub(xadd(ww, ii), jj), kk);
    R[0][1] = xsub(ij, wk);
    R[0][2] = xadd(wj, ik);
    R[1][0] = xadd(wk, ij);
    R[1][1] = xsub(xadd(xsub(ww, ii), jj), kk);
    R[1][2] = xsub(jk, wi);
    R[2][0] = xsub(ik, wj);
    R[2][1] = xadd(wi, jk);
    R[2][2] = xadd(xsub(xsub(ww, ii), jj), kk);
  }

  for (uint32_t i0 = blockIdx.x * kIcpBlock + warp * 32; i0 < ns; i0 += gridDim.x * kIcpBlock) {
    const uint32_t i = i0 + lane;
    bool valid = false;
    V3 s{0.0f, 0.0f, 0.0f};
    float4 d4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    uint32_t sid = 0, tid = TC_NO_INDEX;  // original indices of the source point and its match
    float nn_d2 = 0.0f;
    if (i < ns) {
      const float4 s4 = __ldg(&src[i]);
      // current_transform * p  (registration.rs:540-544)
      s = quat_rotate(T + 3, V3{s4.x, s4.y, s4.z});
      s = V3{xadd(s.x, T[0]), xadd(s.y, T[1]), xadd(s.z, T[2])};
      // Seed the search with last iteration's match (level << 30 | sorted position): the pose
      // moves little between iterations, so this candidate is almost always the answer again and
      // bounds the search to the one or two cells around the query.  Exactness is unaffected.
      Best1 best;
      int level, start = -1;
      const uint32_t seed = prev ? prev[i] : 0xFFFFFFFFu;
      bool keep = false;  // last iteration's match is provably still the nearest: no search
      if (seed != 0xFFFFFFFFu) {
        start = (int)(seed >> 30);
        best.pos = seed & 0x3FFFFFFFu;
        const float4 c = __ldg(&ls.pts[start][best.pos]);
        const float d2 = dist2_exact(c.x, c.y, c.z, s.x, s.y, s.z);
        best.key = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(c.w);
        best.seeded = true;
        // cache = (query position at the last full search, lower bound `a` on the distance from
        // THAT position to every target point other than the match).  By the triangle inequality
        // every other point is now at least a - |s - q_full| away; if the match is strictly closer
        // than that (with slack for f32 rounding of the coordinates, 3e-7 |s|, and of the
        // distances, 2e-5 relative) the reference's 1-NN cannot be anything else, ties included.
        if (cache) {
          const float4 cf = cache[i];  // (all NaN before the point's first search)
          const float mx = s.x - cf.x, my = s.y - cf.y, mz = s.z - cf.z;
          const float moved = sqrtf(mx * mx + my * my + mz * mz);
          const float dist = sqrtf(d2);
          const float lhs = (dist + moved) * 1.00002f +
                            3e-7f * (fabsf(s.x) + fabsf(s.y) + fabsf(s.z));
          keep = cf.w > 0.0f && lhs < cf.w * 0.99998f;
          level = start;
        }
      }
      if (!keep) {
        // Every search starts on the finest level (the probe and the box query are cheapest
        // there); only a candidate that is still several cells away after the probe moves the
        // search to a coarser level (pick_level, below).  A candidate found on another level
        // keeps its distance as the bound, but its position belongs to that level's array: its
        // key is raised by one, so the same point - met again on the new level, inside the ball
        // that is searched in full - replaces it with the position there.
        if (best.seeded && start != 0) {
          best.key += 1;
          start = 0;
        }
        // Seeded (level << 30 | position of last iteration's match): a box query around the seed
        // distance - one or two rows of one or two cells once the pose has settled.  Without a
        // seed (first iteration), or when the pose jumped and the old match is cells away (second
        // iteration of a badly aligned pair), the 2x2 rows nearest to the query are probed first:
        // they usually hold the neighbour and shrink the box from dozens of cells to a few.
        level = max(start, 0);
        const GridParams& g = ls.g[level];
        const float4* __restrict__ pts = ls.pts[level];
        const uint32_t* __restrict__ cs = ls.cs[level];
        auto scan = [&](uint32_t lo, uint32_t hi) { best.scan(pts, lo, hi, s.x, s.y, s.z, 0); };
        float covered = -1.0f;  // radius of the ball around s whose cells were all scanned
        bool probed = false;
        int px0 = 0, px1 = 0, py0 = 0, py1 = 0, pz0 = 0, pz1 = 0;  // cells the probe scanned
        if (!best.seeded || best.kth() > g.cell * g.cell * 0.5f) {
          best.init();
          float ux, uy, uz;
          const int cx = cell_coord(s.x, g.ox, g.inv, g.nx, ux);
          const int cy = cell_coord(s.y, g.oy, g.inv, g.ny, uy);
          const int cz = cell_coord(s.z, g.oz, g.inv, g.nz, uz);
          const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
          const int y2 = min(max(cy + ((uy - (float)cy < 0.5f) ? -1 : 1), 0), g.ny - 1);
          const int z2 = min(max(cz + ((uz - (float)cz < 0.5f) ? -1 : 1), 0), g.nz - 1);
#pragma unroll 1
          for (int t = 0; t < 4; ++t) {
            const int y = (t & 1) ? y2 : cy, z = (t & 2) ? z2 : cz;
            if (((t & 1) && y2 == cy) || ((t & 2) && z2 == cz)) continue;  // clamped: same row
            const uint32_t row = cell_id(g, 0, y, z);
            scan(__ldg(&cs[row + x0]), __ldg(&cs[row + x1 + 1]));
          }
          px0 = x0;
          px1 = x1;
          py0 = min(cy, y2);
          py1 = max(cy, y2);
          pz0 = min(cz, z2);
          pz1 = max(cz, z2);
          probed = !best.seeded;  // (a far seed may lie outside the probed rows: not "all scanned")
        }
        // Nothing in hand (the probe found no point), or the match is more than four cells away
        // (first iterations of a badly aligned pair): the ring search grows cell by cell from the
        // query, nearest rows first, and prunes rows / cells that cannot beat the best distance -
        // a fixed cube of that half-width would distance-test every point of the surface patch
        // inside it.  One call site for both cases (the search code is large).
        const bool far = !best.full() || best.kth() > 16.0f * g.cell * g.cell;
        if (far) {
          const bool have = best.full();
          if (have) {
            best.seeded = true;  // keep the candidate in hand as the pruning bound
            if (ls.n > 1) {      // (a far candidate: continue on the level that suits its distance)
              const int tl = pick_level(ls, best.kth());
              if (tl != level) {
                best.key += 1;
                level = tl;
              }
            }
          }
          best.far = true;
          level_search(ls, s.x, s.y, s.z, 1u, best, level, have ? level : -1);
          if (have && best.full()) covered = sqrtf(best.kth());
        } else {
          float r2 = best.kth();
          // Seeded and within one cell of the match: the ball is taken 1.25x wider than the
          // distance in hand.  What that costs in extra cells buys the bound (distance to every
          // OTHER point) that lets the following iterations keep the match without any search.
          // Unseeded (first iteration): no margin, and the 2x2 probed rows usually contain the
          // whole ball already - nothing left to visit.
          bool inside = false;
          if (best.seeded && cache && r2 < g.cell * g.cell) {
            r2 *= 1.5625f;
          } else if (probed) {
            const BoxCells bc = box_cells(g, s.x, s.y, s.z, r2);
            inside = bc.xa >= px0 && bc.xb <= px1 && bc.ya >= py0 && bc.yb <= py1 && bc.za >= pz0 &&
                     bc.zb <= pz1;
          }
          if (!inside) box_visit(g, cs, s.x, s.y, s.z, r2, scan);
          covered = sqrtf(r2);
        }
        if (cache)
          cache[i] = make_float4(s.x, s.y, s.z,
                                 (best.full() && covered > 0.0f) ? fminf(sqrtf(best.second), covered)
                                                                 : -1.0f);
      }
      valid = best.full();
      if (prev) prev[i] = valid ? (((uint32_t)level << 30) | best.pos) : 0xFFFFFFFFu;
      if (valid && max_dist >= 0.0f) {  // reject iff distance > max (registration.rs:100)
        if (xsqrt(best.kth()) > max_dist) valid = false;
      }
      sid = __float_as_uint(s4.w);
      if (valid) {
        d4 = __ldg(&ls.pts[level][best.pos]);
        tid = (uint32_t)best.key;
        nn_d2 = best.kth();
      }
    }
    float v[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] = 0.0f;
    if (MODE == kPlane) {
      float4 n4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (valid) n4 = __ldg(&tgt_nrm[__float_as_uint(d4.w)]);
      const V3 n{n4.x, n4.y, n4.z};
      const V3 c = xcross(s, n);  // registration.rs:418
      const float dx = xsub(d4.x, s.x), dy = xsub(d4.y, s.y), dz = xsub(d4.z, s.z);
      const float b = xadd(xadd(xmul(n.x, dx), xmul(n.y, dy)), xmul(n.z, dz));  // :424
      const float a[6] = {c.x, c.y, c.z, n.x, n.y, n.z};  // all zero when there is no match
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = r; cc < 6; ++cc) v[r * 6 - (r * (r - 1)) / 2 + (cc - r)] = xmul(a[r], a[cc]);
#pragma unroll
      for (int r = 0; r < 6; ++r) v[21 + r] = xmul(a[r], b);
      v[27] = xmul(b, b);
      v[28] = valid ? 1.0f : 0.0f;
    } else if (MODE == kGicp) {
      // gicp.rs:204-252: M = C_t + R C_s R^T, M^-1 by the adjugate formula (a singular M skips
      // the pair), J = [-skew(T s) | I]; sums = lower triangle of H, g, sum dist^2, n.
      // 3x3 products accumulate (a0 b0 + a1 b1) + a2 b2 like nalgebra's.
      float Cs[3][3], M[3][3];
      {
        float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
        if (valid) {
          a0 = __ldg(&src_cov[2 * (uint64_t)sid]);
          a1 = __ldg(&src_cov[2 * (uint64_t)sid + 1]);
          b0 = __ldg(&tgt_nrm[2 * (uint64_t)tid]);
          b1 = __ldg(&tgt_nrm[2 * (uint64_t)tid + 1]);
        }
        Cs[0][0] = a0.x; Cs[0][1] = a0.y; Cs[0][2] = a0.z;
        Cs[1][0] = a0.y; Cs[1][1] = a0.w; Cs[1][2] = a1.x;
        Cs[2][0] = a0.z; Cs[2][1] = a1.x; Cs[2][2] = a1.y;
        float t[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            t[r][c] = xadd(xadd(xmul(R[r][0], Cs[0][c]), xmul(R[r][1], Cs[1][c])), xmul(R[r][2], Cs[2][c]));
        const float Ct[3][3] = {{b0.x, b0.y, b0.z}, {b0.y, b0.w, b1.x}, {b0.z, b1.x, b1.y}};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            M[r][c] = xadd(Ct[r][c], xadd(xadd(xmul(t[r][0], R[c][0]), xmul(t[r][1], R[c][1])),
                                          xmul(t[r][2], R[c][2])));
      }
      float mi[3][3];
      {
        const float m11 = M[0][0], m12 = M[0][1], m13 = M[0][2], m21 = M[1][0], m22 = M[1][1],
                    m23 = M[1][2], m31 = M[2][0], m32 = M[2][1], m33 = M[2][2];
        const float n1223 = xsub(xmul(m22, m33), xmul(m32, m23));
        const float n1123 = xsub(xmul(m21, m33), xmul(m31, m23));
        const float n1122 = xsub(xmul(m21, m32), xmul(m31, m22));
        const float det = xadd(xsub(xmul(m11, n1223), xmul(m12, n1123)), xmul(m13, n1122));
        if (det == 0.0f) valid = false;  // try_inverse() == None (all-zero input when unmatched)
        mi[0][0] = xdiv(n1223, det);
        mi[0][1] = xdiv(xsub(xmul(m13, m32), xmul(m33, m12)), det);
        mi[0][2] = xdiv(xsub(xmul(m12, m23), xmul(m22, m13)), det);
        mi[1][0] = xdiv(-n1123, det);
        mi[1][1] = xdiv(xsub(xmul(m11, m33), xmul(m31, m13)), det);
        mi[1][2] = xdiv(xsub(xmul(m13, m21), xmul(m23, m11)), det);
        mi[2][0] = xdiv(n1122, det);
        mi[2][1] = xdiv(xsub(xmul(m12, m31), xmul(m32, m11)), det);
        mi[2][2] = xdiv(xsub(xmul(m11, m22), xmul(m21, m12)), det);
      }
      const float rs[3] = {xsub(d4.x, s.x), xsub(d4.y, s.y), xsub(d4.z, s.z)};  // t_i - T s_i
      const float A[3][3] = {{-0.0f, s.z, -s.y}, {-s.z, -0.0f, s.x}, {s.y, -s.x, -0.0f}};
      float ma[3][3], hrr[3][3], hrt[3][3], wr[3], gr[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          ma[r][c] = xadd(xadd(xmul(mi[r][0], A[0][c]), xmul(mi[r][1], A[1][c])), xmul(mi[r][2], A[2][c]));
        wr[r] = xadd(xadd(xmul(mi[r][0], rs[0]), xmul(mi[r][1], rs[1])), xmul(mi[r][2], rs[2]));
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {  // A^T (.)  : A^T[r][k] = A[k][r]
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          hrr[r][c] = xadd(xadd(xmul(A[0][r], ma[0][c]), xmul(A[1][r], ma[1][c])), xmul(A[2][r], ma[2][c]));
          hrt[r][c] = xadd(xadd(xmul(A[0][r], mi[0][c]), xmul(A[1][r], mi[1][c])), xmul(A[2][r], mi[2][c]));
        }
        gr[r] = xadd(xadd(xmul(A[0][r], wr[0]), xmul(A[1][r], wr[1])), xmul(A[2][r], wr[2]));
      }
      // upper-triangle slot (r <= c) of the 21 <- the LOWER element H[c][r] the Cholesky reads
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) {
          const float h = c < 3 ? hrr[c][r] : (r < 3 ? hrt[r][c - 3] : mi[c - 3][r - 3]);
          v[r * 6 - (r * (r - 1)) / 2 + (c - r)] = valid ? h : 0.0f;
        }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        v[21 + r] = valid ? gr[r] : 0.0f;
        v[24 + r] = valid ? wr[r] : 0.0f;
      }
      const float dist = xsqrt(nn_d2);
      v[27] = valid ? xmul(dist, dist) : 0.0f;  // mse_sum += dist * dist (gicp.rs:256)
      v[28] = valid ? 1.0f : 0.0f;
      if (!valid) tid = TC_NO_INDEX;
    } else {
      // raw moments for Kabsch (registration.rs:157-172) and the mse (:206-218)
      const float m = valid ? 1.0f : 0.0f;
      const float sv[3] = {s.x * m, s.y * m, s.z * m}, tv[3] = {d4.x, d4.y, d4.z};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        v[r] = sv[r];
        v[3 + r] = tv[r];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) v[6 + 3 * r + cc] = xmul(sv[r], tv[cc]);
      }
      const float dx = xsub(sv[0], d4.x), dy = xsub(sv[1], d4.y), dz = xsub(sv[2], d4.z);
      v[15] = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
      v[16] = m;
    }
    if (match_out && i < ns) match_out[sid] = valid ? tid : TC_NO_INDEX;
    // transpose-reduce: after the step with offset o a lane holds the values whose index has the
    // same o-bit as its lane id, summed over the lane pair; five steps leave lane j with sum j
#pragma unroll
    for (int o = 16, cnt = 16; o >= 1; o >>= 1, cnt >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int t = 0; t < cnt; ++t) {
        const float give = up ? v[t] : v[t + cnt];
        const float keep = up ? v[t + cnt] : v[t];
        v[t] = keep + __shfl_xor_sync(0xffffffffu, give, o);
      }
    }
    acc += (double)v[0];
  }
  // one row per warp in shared memory
  __shared__ double sm[kIcpBlock / 32][32];
  sm[warp][lane] = acc;
  __syncthreads();
  if (threadIdx.x < NS) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kIcpBlock / 32; ++w) v += sm[w][threadIdx.x];
    partials[(uint64_t)blockIdx.x * NS + threadIdx.x] = v;
  }
  __threadfence();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    // Block partials -> totals with the whole block: warp w takes blocks w, w+8, ... (lane j =
    // sum j, four independent accumulators so the L2 loads overlap), the eight slices are then
    // added in warp order.  The association is fixed by (gridDim, block size) alone, so the
    // result does not depend on scheduling.  (Walking all blocks with 29 threads in one dependent
    // chain was most of the per-iteration tail: ~600 L2 round trips.)
    double v = 0.0;
    {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if (lane < NS) {
        uint32_t b = warp;
        for (; b + 24 < gridDim.x; b += 32) {
          a0 += __ldcg(&partials[(uint64_t)b * NS + lane]);
          a1 += __ldcg(&partials[(uint64_t)(b + 8) * NS + lane]);
          a2 += __ldcg(&partials[(uint64_t)(b + 16) * NS + lane]);
          a3 += __ldcg(&partials[(uint64_t)(b + 24) * NS + lane]);
        }
        for (; b < gridDim.x; b += 8) a0 += __ldcg(&partials[(uint64_t)b * NS + lane]);
      }
      __syncthreads();  // sm[][] was read by the first NS threads above
      sm[warp][lane] = (a0 + a1) + (a2 + a3);
      __syncthreads();
      if (threadIdx.x < NS) {
#pragma unroll
        for (int w = 0; w < kIcpBlock / 32; ++w) v += sm[w][threadIdx.x];
        sums[threadIdx.x] = v;
      }
    }
    if (threadIdx.x == 0) st->ticket = 0;
    if (fuse && px.world > 1) {
      // ---- fused all-reduce: push this rank's sums into every rank's buffer, then flag --------
      const unsigned long long epoch = px.epoch_base + st->iterations;
      const int parity = (int)(epoch & 1ull);
      const size_t slot = ((size_t)parity * px.world + px.rank) * kXSlot;
      if (threadIdx.x < NS)
        for (int r = 0; r < px.world; ++r) px.peer[r][slot + threadIdx.x] = v;
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int r = 0; r < px.world; ++r)
          *(volatile double*)&px.peer[r][slot + kXSlot - 1] = (double)epoch;
        // wait (bounded) until every rank's contribution for this epoch has landed here
        const double* mine = px.peer[px.rank];
        bool ok = true;
        for (int r = 0; r < px.world && ok; ++r) {
          const volatile double* f = &mine[((size_t)parity * px.world + r) * kXSlot + kXSlot - 1];
          unsigned long long spins = 0;
          while (*f != (double)epoch) {
            __nanosleep(64);
            if (++spins > px.spin_limit) {  // a peer died: fail instead of hanging the GPU
              ok = false;
              break;
            }
          }
        }
        if (!ok) {
          st->status = 4;
          st->done = 1;
        }
      }
      __syncthreads();
      if (st->status == 4) return;
      __threadfence_system();
      if (threadIdx.x < NS) {  // sum in rank order: identical bits on every rank
        const double* mine = px.peer[px.rank];
        double t = 0.0;
        for (int r = 0; r < px.world; ++r)
          t += __ldcv(&mine[((size_t)parity * px.world + r) * kXSlot + threadIdx.x]);
        sums[threadIdx.x] = t;
      }
    }
    // The 6x6 / Horn solve runs here too (one thread of the last block, through a real call so
    // that its registers do not count against the search) unless an NCCL all-reduce has to
    // happen between the sums and the solve.
    if (solve_here) {
      __syncthreads();
      if (threadIdx.x == 0) {
        if (MODE == kPoint) icp_solve_point(st, sums, conv);
        else icp_solve_plane(st, sums, conv);  // kGicp: same 6x6 system layout and update
      }
    }
  }
}

// 6x6 SPD solve in f64: Cholesky, falling back to LU with partial pivoting
// (registration.rs:432-438).  Returns 0 ok, 1 singular.
__device__ int solve6(const double* s /*21 upper-tri row-major*/, const double* rhs, double x[6]) {
  double A[6][6];
  int t = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) {
      A[r][c] = s[t];
      A[c][r] = s[t];
      ++t;
    }
  double L[6][6];
  bool ok = true;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) L[r][c] = A[r][c];
  for (int j = 0; j < 6 && ok; ++j) {
    for (int k = 0; k < j; ++k) {
      const double f = L[j][k];
      for (int r = j; r < 6; ++r) L[r][j] -= f * L[r][k];
    }
    const double d = L[j][j];
    if (!(d > 0.0)) {
      ok = false;
      break;
    }
    const double sd = sqrt(d);
    L[j][j] = sd;
    for (int r = j + 1; r < 6; ++r) L[r][j] /= sd;
  }
  if (ok) {
    for (int i = 0; i < 6; ++i) x[i] = rhs[i];
    for (int i = 0; i < 6; ++i) {
      x[i] /= L[i][i];
      for (int r = i + 1; r < 6; ++r) x[r] -= x[i] * L[r][i];
    }
    for (int i = 5; i >= 0; --i) {
      double d = 0.0;
      for (int r = i + 1; r < 6; ++r) d += L[r][i] * x[r];
      x[i] = (x[i] - d) / L[i][i];
    }
    return 0;
  }
  // LU with partial pivoting
  double b[6];
  for (int i = 0; i < 6; ++i) b[i] = rhs[i];
  for (int i = 0; i < 6; ++i) {
    int piv = i;
    double best = fabs(A[i][i]);
    for (int r = i + 1; r < 6; ++r)
      if (fabs(A[r][i]) > best) {
        best = fabs(A[r][i]);
        piv = r;
      }
    if (A[piv][i] == 0.0 || !isfinite(A[piv][i])) return 1;
    if (piv != i) {
      for (int c = 0; c < 6; ++c) {
        const double tmp = A[i][c];
        A[i][c] = A[piv][c];
        A[piv][c] = tmp;
      }
      const double tb = b[i];
      b[i] = b[piv];
      b[piv] = tb;
    }
    for (int r = i + 1; r < 6; ++r) {
      const double f = A[r][i] / A[i][i];
      for (int c = i + 1; c < 6; ++c) A[r][c] -= f * A[i][c];
      b[r] -= f * b[i];
    }
  }
  for (int i = 5; i >= 0; --i) {
    for (int c = i + 1; c < 6; ++c) b[i] -= A[i][c] * b[c];
    b[i] /= A[i][i];
    x[i] = b[i];
  }
  return 0;
}

__device__ void icp_apply_delta(IcpState* st, const float dq[4], const float dt[3], float mse,
                                float conv);

__global__ void k_icp_solve(IcpState* __restrict__ st, const double* __restrict__ sums, float conv) {
  if (threadIdx.x != 0 || st->done) return;
  icp_solve_plane(st, sums, conv);
}

__device__ __noinline__ void icp_solve_plane(IcpState* st, const double* sums, float conv) {
  const double n_valid = sums[28];
  st->n_valid = n_valid;
  if (n_valid < 6.0) {  // registration.rs:568-572
    st->status = 2;
    st->done = 1;
    return;
  }
  double x[6];
  if (solve6(sums, sums + 21, x) != 0) {  // registration.rs:435-437
    st->status = 3;
    st->done = 1;
    return;
  }
  // delta = (Rz * Ry * Rx, t)  (registration.rs:441-449), quaternions as from_axis_angle
  const float ax = (float)x[0], ay = (float)x[1], az = (float)x[2];
  float sx, cx, sy, cy, sz, cz;
  sincosf(xdiv(ax, 2.0f), &sx, &cx);
  sincosf(xdiv(ay, 2.0f), &sy, &cy);
  sincosf(xdiv(az, 2.0f), &sz, &cz);
  const float qx[4] = {sx, 0.0f, 0.0f, cx};
  const float qy[4] = {0.0f, sy, 0.0f, cy};
  const float qz[4] = {0.0f, 0.0f, sz, cz};
  float qzy[4], dq[4];
  quat_mul(qz, qy, qzy);
  quat_mul(qzy, qx, dq);
  const float dt[3] = {(float)x[3], (float)x[4], (float)x[5]};
  // mean b^2 over valid pairs, residuals taken BEFORE delta (registration.rs:453-471, 578);
  // T <- delta * T = (dq * q, dt + dq . t) (registration.rs:576) [nalgebra Isometry3 product]
  const float mse = (float)(sums[27] / n_valid);
  icp_apply_delta(st, dq, dt, mse, conv);
}


// Shared tail of both solve kernels: T <- delta o T, mse, convergence (registration.rs:321-338,
// 576-589).  dq = [i,j,k,w].
__device__ void icp_apply_delta(IcpState* st, const float dq[4], const float dt[3], float mse,
                                float conv) {
  float q_old[4] = {st->T[3], st->T[4], st->T[5], st->T[6]};
  const V3 sh = quat_rotate(dq, V3{st->T[0], st->T[1], st->T[2]});
  float qn[4];
  quat_mul(dq, q_old, qn);
  st->T[0] = xadd(dt[0], sh.x);
  st->T[1] = xadd(dt[1], sh.y);
  st->T[2] = xadd(dt[2], sh.z);
  st->T[3] = qn[0];
  st->T[4] = qn[1];
  st->T[5] = qn[2];
  st->T[6] = qn[3];
  if (st->iterations < TC_STATS_MAX_ITERS) {
    st->mse_hist[st->iterations] = mse;
    st->valid_hist[st->iterations] = st->n_valid;
  }
  st->iterations += 1;
  st->mse = mse;
  if (fabsf(st->prev_mse - mse) < conv) {
    st->converged = 1;
    st->done = 1;
    return;
  }
  st->prev_mse = mse;
}

// Point-to-point step (compute_transformation, registration.rs:144-203): the rotation maximising
// trace(R H) with det R = +1.  The reference takes V U^T from an SVD of H and flips the last
// singular vector on a reflection; the same proper rotation is the dominant eigenvector of Horn's
// symmetric 4x4 matrix built from H, found here by cyclic Jacobi in f64.
__global__ void k_icp_solve_p2p(IcpState* __restrict__ st, const double* __restrict__ sums,
                                float conv) {
  if (threadIdx.x != 0 || st->done) return;
  icp_solve_point(st, sums, conv);
}

__device__ __noinline__ void icp_solve_point(IcpState* st, const double* sums, float conv) {
  const double n = sums[16];
  st->n_valid = n;
  if (n < 3.0) {  // registration.rs:311-315
    st->status = 2;
    st->done = 1;
    return;
  }
  double cs[3], ct[3], H[3][3];
  for (int a = 0; a < 3; ++a) {
    cs[a] = sums[a] / n;
    ct[a] = sums[3 + a] / n;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) H[r][c] = sums[6 + 3 * r + c] - n * cs[r] * ct[c];
  double N[4][4];
  N[0][0] = H[0][0] + H[1][1] + H[2][2];
  N[0][1] = H[1][2] - H[2][1];
  N[0][2] = H[2][0] - H[0][2];
  N[0][3] = H[0][1] - H[1][0];
  N[1][1] = H[0][0] - H[1][1] - H[2][2];
  N[1][2] = H[0][1] + H[1][0];
  N[1][3] = H[2][0] + H[0][2];
  N[2][2] = -H[0][0] + H[1][1] - H[2][2];
  N[2][3] = H[1][2] + H[2][1];
  N[3][3] = -H[0][0] - H[1][1] + H[2][2];
  for (int r = 1; r < 4; ++r)
    for (int c = 0; c < r; ++c) N[r][c] = N[c][r];
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) (r == c ? dg : off) += N[r][c] * N[r][c];
    if (off <= 1e-30 * dg || off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (N[p][q] == 0.0) continue;
        const double theta = (N[q][q] - N[p][p]) / (2.0 * N[p][q]);
        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double a = N[k][p], b = N[k][q];
          N[k][p] = c * a - s * b;
          N[k][q] = s * a + c * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = N[p][k], b = N[q][k];
          N[p][k] = c * a - s * b;
          N[q][k] = s * a + c * b;
        }
        for (int k = 0; k < 4; ++k) {
          const double a = V[k][p], b = V[k][q];
          V[k][p] = c * a - s * b;
          V[k][q] = s * a + c * b;
        }
      }
  }
  int m = 0;
  for (int i = 1; i < 4; ++i)
    if (N[i][i] > N[m][m]) m = i;
  double qw = V[0][m], qx = V[1][m], qy = V[2][m], qz = V[3][m];
  const double qn = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  if (!(qn > 0.0) || !isfinite(qn)) {
    st->status = 3;
    st->done = 1;
    return;
  }
  const double sg = qw < 0.0 ? -1.0 / qn : 1.0 / qn;  // w >= 0, like a matrix -> quaternion conversion
  qw *= sg;
  qx *= sg;
  qy *= sg;
  qz *= sg;
  // translation = target_centroid - rotation * source_centroid (registration.rs:197)
  const double R[3][3] = {
      {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)},
      {2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)},
      {2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)}};
  float dt[3];
  for (int r = 0; r < 3; ++r)
    dt[r] = (float)(ct[r] - (R[r][0] * cs[0] + R[r][1] * cs[1] + R[r][2] * cs[2]));
  const float dq[4] = {(float)qx, (float)qy, (float)qz, (float)qw};
  const float mse = (float)(sums[15] / n);  // pre-delta residuals (registration.rs:324)
  icp_apply_delta(st, dq, dt, mse, conv);
}

// Not converged: the reference reports the mse of the LAST correspondences under the FINAL
// transform (registration.rs:342-361).  prev[] holds the last matches (level << 30 | position).
__global__ void __launch_bounds__(kIcpBlock)
k_icp_p2p_final_mse(LevelSet ls, const float4* __restrict__ src, uint32_t ns,
                    const uint32_t* __restrict__ prev, const uint32_t* __restrict__ match,
                    const IcpState* __restrict__ st,
                    double* __restrict__ out2 /* [sum, count], zeroed */) {
  float T[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) T[i] = st->T[i];
  double sum = 0.0, cnt = 0.0;
  for (uint32_t i = blockIdx.x * kIcpBlock + threadIdx.x; i < ns; i += gridDim.x * kIcpBlock) {
    const uint32_t m = prev[i];
    if (m == 0xFFFFFFFFu) continue;
    const float4 s4 = __ldg(&src[i]);
    // prev[] keeps the nearest point as next iteration's seed even when the pair was rejected by
    // max_correspondence_distance; the reference averages over the ACCEPTED pairs only
    // (final_correspondences, registration.rs:342-361): `match` (by original source index) has them
    if (match && match[__float_as_uint(s4.w)] == TC_NO_INDEX) continue;
    V3 s = quat_rotate(T + 3, V3{s4.x, s4.y, s4.z});
    s = V3{xadd(s.x, T[0]), xadd(s.y, T[1]), xadd(s.z, T[2])};
    const float4 t4 = __ldg(&ls.pts[m >> 30][m & 0x3FFFFFFFu]);
    const float dx = xsub(s.x, t4.x), dy = xsub(s.y, t4.y), dz = xsub(s.z, t4.z);
    sum += (double)xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
    cnt += 1.0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt > 0.0) {
    atomicAdd(&out2[0], sum);
    atomicAdd(&out2[1], cnt);
  }
}

}  // namespace

namespace {

// Shared driver of the device-resident ICP loops (mode: kPlane / kPoint).
// d_tgt_normals_aos: target normals n x 3 (kPlane).  d_src_cov / d_tgt_cov: per-point covariances
// as two float4 {xx, xy, xz, yy | yz, zz, 0, 0} by original index (kGicp).
int icp_device_impl(int mode, tc_context* ctx, tc_comm* comm, const tc_cloud* src,
                    const tc_index* tgt, const float* d_tgt_normals_aos, const float4* d_src_cov,
                    const float4* d_tgt_cov, const float init[7], uint32_t max_iters,
                    float max_corr_dist, float conv_threshold, tc_icp_result* out,
                    uint32_t* d_match_out) {
  if (!ctx || !src || !tgt || !out || !init) return TC_INVALID_DATA;
  TcRange nvtx_range(mode == kPlane ? "tc:icp point-to-plane" : mode == kPoint ? "tc:icp point-to-point" : "tc:gicp");
  // validation order of registration.rs:266-276 / 517-531 (normals length: host wrapper)
  if (tgt->sharded)
    return tc_fail(ctx, TC_INVALID_DATA, "ICP needs a complete target index (not a slab-sharded one)");
  if ((src->n == 0 && !comm) || tgt->n == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "Source or target point cloud is empty");
  if (mode == kPlane && !d_tgt_normals_aos)
    return tc_fail(ctx, TC_INVALID_DATA,
                   "target_normals length must equal the number of target points");
  if (mode == kGicp && (!d_src_cov || !d_tgt_cov)) return TC_INVALID_DATA;
  if (max_iters == 0) return tc_fail(ctx, TC_INVALID_DATA, "Max iterations must be positive");
  const uint32_t ns = (uint32_t)src->n;
  const uint32_t nt = (uint32_t)tgt->n;
  if (mode == kPoint && nt >= (1u << 30))
    return tc_fail(ctx, TC_INVALID_DATA, "point-to-point ICP supports targets below 2^30 points");
  const int n_sums = mode == kPoint ? kNumSumsP2P : kNumSums;

  // sort the source spatially (own bbox, target-sized cells) so warps walk coherent target cells
  float4* d_src = nullptr;
  if (ns > 0) {
    float mn[3], mx[3];
    TC_TRY(tci_bbox(ctx, src->d_xyz, ns, mn, mx));
    GridParams sg{};
    sg.ox = mn[0];
    sg.oy = mn[1];
    sg.oz = mn[2];
    float cell = tgt->lv[tgt->primary].g.cell * 2.0f;
    const float emax = std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
    if (!(cell > 0)) cell = 1.0f;
    if (emax / cell > 1000.0f) cell = emax / 1000.0f;  // <= ~2^30 cells
    sg.cell = cell;
    sg.inv = 1.0f / cell;
    sg.nx = (int)std::floor((mx[0] - mn[0]) / cell) + 1;
    sg.ny = (int)std::floor((mx[1] - mn[1]) / cell) + 1;
    sg.nz = (int)std::floor((mx[2] - mn[2]) / cell) + 1;
    sg.n = ns;
    sg.ymajor = tgt->lv[tgt->primary].g.ymajor;  // walk the target's planes in its own order
    TC_TRY(tci_sort_by_grid(ctx, src->d_xyz, ns, sg, &d_src));
  }
  const int grid = std::max(1, std::min((int)((ns + kIcpBlock - 1) / kIcpBlock), ctx->sm_count * 4));

  float4* d_nrm = nullptr;
  uint32_t* d_prev = nullptr;
  uint32_t* d_match_own = nullptr;  // kPoint + max distance: acceptance record for the final mse
  float4* d_cache = nullptr;        // per source point: position + bound of its last full search
  IcpState* d_state = nullptr;
  double *d_partials = nullptr, *d_sums = nullptr, *d_final = nullptr;
  int st = TC_OK;
  if (mode == kPlane) st = tc_alloc(ctx, &d_nrm, nt);
  if (st == TC_OK) st = tc_alloc(ctx, &d_state, 1);
  if (st == TC_OK && nt < (1u << 30) && ns > 0) {
    st = tc_alloc(ctx, &d_prev, ns);
    if (st == TC_OK) cudaMemsetAsync(d_prev, 0xFF, (size_t)ns * sizeof(uint32_t), ctx->stream);
    if (st == TC_OK && g_tc_icp_keep) st = tc_alloc(ctx, &d_cache, ns);  // bound = NaN: unknown
    if (st == TC_OK && d_cache) cudaMemsetAsync(d_cache, 0xFF, (size_t)ns * sizeof(float4), ctx->stream);
  }
  if (st == TC_OK && mode == kPoint && max_corr_dist >= 0.0f && !d_match_out && ns > 0) {
    st = tc_alloc(ctx, &d_match_own, ns);
    d_match_out = d_match_own;
  }
  if (st == TC_OK) st = tc_alloc(ctx, &d_partials, (uint64_t)grid * n_sums);
  if (st == TC_OK) st = tc_alloc(ctx, &d_sums, n_sums);
  if (st == TC_OK) st = tc_alloc(ctx, &d_final, 2);
  IcpState h_state{};
  double h_final[2] = {0.0, 0.0};
  bool fuse_used = false;
  if (st == TC_OK) {
    if (mode == kPlane) {
      k_pad_normals<<<std::max(1, std::min((int)((nt + kIcpBlock - 1) / kIcpBlock),
                                           ctx->sm_count * 8)),
                      kIcpBlock, 0, ctx->stream>>>(d_tgt_normals_aos, nt, d_nrm);
      ctx->launches++;
    }
    k_icp_init<<<1, 1, 0, ctx->stream>>>(d_state, nullptr, init[0], init[1], init[2], init[3],
                                         init[4], init[5], init[6]);
    ctx->launches++;
    const LevelSet ls = tgt->level_set(g_tc_search_flags);
    // Multi-GPU: with peers mapped over NVLink the all-reduce happens inside the correspondence
    // kernel (correspond+exchange -> solve); otherwise correspond -> ncclAllReduce -> solve.
    PeerXchg px{};
    px.world = 1;
    int fuse = 0;
    if (comm && g_tc_icp_fuse &&
        tci_comm_peers(comm, px.peer, &px.world, &px.rank, &px.epoch_base, 0))
      fuse = 1;
    {
      // how long a rank waits for its peers' sums before giving up (TC_GPU instead of a hang):
      // generous by default - ranks legitimately arrive seconds apart (first-call set-up, an
      // uneven index build); TC_PEER_TIMEOUT_S overrides
      double secs = 30.0;
      if (const char* e = std::getenv("TC_PEER_TIMEOUT_S")) secs = std::max(0.1, atof(e));
      px.spin_limit = (unsigned long long)(secs / 100e-9);  // ~100 ns per poll incl. the sleep
    }
    fuse_used = fuse != 0;
    const int solve_here = (!comm || fuse) && g_tc_icp_solve_fused ? 1 : 0;
    for (uint32_t it = 0; it < max_iters && st == TC_OK; ++it) {
      if (mode == kGicp)
        k_icp_correspond<kGicp><<<grid, kIcpBlock, 0, ctx->stream>>>(
            ls, d_tgt_cov, d_src_cov, d_src, ns, max_corr_dist, d_state, d_partials, d_sums,
            d_match_out, d_prev, d_cache, fuse, px, solve_here, conv_threshold);
      else if (mode == kPlane)
        k_icp_correspond<kPlane><<<grid, kIcpBlock, 0, ctx->stream>>>(
            ls, d_nrm, nullptr, d_src, ns, max_corr_dist, d_state, d_partials, d_sums, d_match_out, d_prev,
            d_cache, fuse, px, solve_here, conv_threshold);
      else
        k_icp_correspond<kPoint><<<grid, kIcpBlock, 0, ctx->stream>>>(
            ls, nullptr, nullptr, d_src, ns, max_corr_dist, d_state, d_partials, d_sums, d_match_out,
            d_prev, d_cache, fuse, px, solve_here, conv_threshold);
      ctx->launches++;
      if (solve_here) continue;
      st = tci_comm_allreduce(comm, d_sums, n_sums);
      if (mode == kPoint)
        k_icp_solve_p2p<<<1, 32, 0, ctx->stream>>>(d_state, d_sums, conv_threshold);
      else
        k_icp_solve<<<1, 32, 0, ctx->stream>>>(d_state, d_sums, conv_threshold);
      ctx->launches++;
    }
    if (st == TC_OK && mode == kPoint) {
      // mse of the last correspondences under the final transform (only used when not converged)
      cudaMemsetAsync(d_final, 0, 2 * sizeof(double), ctx->stream);
      if (ns > 0) {
        k_icp_p2p_final_mse<<<grid, kIcpBlock, 0, ctx->stream>>>(
            ls, d_src, ns, d_prev, max_corr_dist >= 0.0f ? d_match_out : nullptr, d_state, d_final);
        ctx->launches++;
      }
      if (comm) st = tci_comm_allreduce(comm, d_final, 2);
    }
    if (st == TC_OK && cudaGetLastError() != cudaSuccess)
      st = tc_fail(ctx, TC_GPU, "ICP kernel launch failed");
    if (st == TC_OK) {
      cudaError_t e = cudaMemcpyAsync(&h_state, d_state, sizeof(IcpState), cudaMemcpyDeviceToHost,
                                      ctx->stream);
      if (e == cudaSuccess && mode == kPoint)
        e = cudaMemcpyAsync(h_final, d_final, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                            ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess)
        st = tc_fail(ctx, TC_GPU, std::string("ICP failed: ") + cudaGetErrorString(e));
    }
  }
  if (fuse_used) {
    // one epoch per executed exchange; an iteration that stopped in the solve (too few pairs,
    // singular system) had exchanged already.  Every rank sees the same sums, hence the same count.
    const bool stopped_in_solve = h_state.status == 2 || h_state.status == 3;
    tci_comm_commit_epochs(comm, (unsigned long long)h_state.iterations + (stopped_in_solve ? 1 : 0));
  }
  tc_free(ctx, d_src);
  tc_free(ctx, d_prev);
  tc_free(ctx, d_cache);
  tc_free(ctx, d_match_own);
  tc_free(ctx, d_nrm);
  tc_free(ctx, d_state);
  tc_free(ctx, d_partials);
  tc_free(ctx, d_sums);
  tc_free(ctx, d_final);
  if (st != TC_OK) return st;
  if (h_state.status == 2)
    return tc_fail(ctx, TC_ALGORITHM,
                   mode == kPlane  ? "Insufficient correspondences for point-to-plane ICP (need >= 6)"
                   : mode == kGicp ? "GICP: insufficient correspondences (need >= 6)"
                                   : "Insufficient correspondences found");
  if (h_state.status == 4)
    return tc_fail(ctx, TC_GPU, "ICP peer exchange timed out (a rank did not arrive)");
  if (h_state.status == 3)
    return tc_fail(ctx, TC_ALGORITHM, mode == kPlane  ? "Point-to-plane system is ill-conditioned"
                                      : mode == kGicp ? "GICP: Gauss-Newton system is ill-conditioned"
                                                      : "SVD of the cross-covariance failed");
  {  // per-iteration history (tc_last_stats)
    tc_stats& hs = ctx->icp_stats;
    hs.icp_iterations = std::min<uint32_t>(h_state.iterations, TC_STATS_MAX_ITERS);
    for (uint32_t i = 0; i < TC_STATS_MAX_ITERS; ++i) {
      hs.icp_mse[i] = i < hs.icp_iterations ? h_state.mse_hist[i] : 0.0f;
      hs.icp_valid[i] = i < hs.icp_iterations ? (uint64_t)h_state.valid_hist[i] : 0;
    }
  }
  for (int i = 0; i < 7; ++i) out->transform[i] = h_state.T[i];
  if (h_state.converged) {
    out->mse = h_state.mse;
  } else if (mode != kPoint) {
    out->mse = h_state.prev_mse;  // previous_mse == the last iteration's mse (:595-601)
  } else {
    out->mse = h_final[1] > 0.0 ? (float)(h_final[0] / h_final[1]) : h_state.prev_mse;
  }
  out->iterations = h_state.iterations;
  out->converged = h_state.converged;
  out->n_correspondences = (uint64_t)h_state.n_valid;
  return TC_OK;
}

}  // namespace

int g_tc_icp_fuse = 1;
extern "C" void tc_debug_set_icp_fuse(int on) { g_tc_icp_fuse = on; }

extern "C" int tc_icp_point_to_plane_device(tc_context* ctx, tc_comm* comm, const tc_cloud* src,
                                            const tc_index* tgt, const float* d_tgt_normals_aos,
                                            const float init[7], uint32_t max_iters,
                                            float max_corr_dist, float conv_threshold,
                                            tc_icp_result* out, uint32_t* d_match_out) {
  return icp_device_impl(kPlane, ctx, comm, src, tgt, d_tgt_normals_aos, nullptr, nullptr, init,
                         max_iters, max_corr_dist, conv_threshold, out, d_match_out);
}

extern "C" int tc_icp_point_to_point_device(tc_context* ctx, tc_comm* comm, const tc_cloud* src,
                                            const tc_index* tgt, const float init[7],
                                            uint32_t max_iters, float max_corr_dist,
                                            float conv_threshold, tc_icp_result* out,
                                            uint32_t* d_match_out) {
  return icp_device_impl(kPoint, ctx, comm, src, tgt, nullptr, nullptr, nullptr, init, max_iters,
                         max_corr_dist, conv_threshold, out, d_match_out);
}

// ---------------------------------------------------------------------------------- GICP
// compute_covariances (gicp.rs:58-95) from kNN rows (k neighbours INCLUDING the point itself,
// ascending (d2, index)): mean by sequential f32 adds, outer products accumulated in neighbour
// order, / max(n - 1, 1), + 1e-4 I; fewer than 3 neighbours -> 1e-3 I.
__global__ void __launch_bounds__(256) k_gicp_cov(const float* __restrict__ xyz, uint32_t n,
                                                  uint32_t k, const uint32_t* __restrict__ idx,
                                                  float4* __restrict__ cov) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t* row = idx + (uint64_t)i * k;
    uint32_t m = 0;
    float sx = 0.0f, sy = 0.0f, sz = 0.0f;
    for (uint32_t j = 0; j < k; ++j) {
      const uint32_t id = row[j];
      if (id == TC_NO_INDEX) break;
      sx = xadd(sx, xyz[3 * (uint64_t)id]);
      sy = xadd(sy, xyz[3 * (uint64_t)id + 1]);
      sz = xadd(sz, xyz[3 * (uint64_t)id + 2]);
      ++m;
    }
    float c[6] = {1e-3f, 0.0f, 0.0f, 1e-3f, 0.0f, 1e-3f};  // xx xy xz yy yz zz
    if (m >= 3) {
      const float nf = (float)m;
      const float mx = xdiv(sx, nf), my = xdiv(sy, nf), mz = xdiv(sz, nf);
#pragma unroll
      for (int t = 0; t < 6; ++t) c[t] = 0.0f;
      for (uint32_t j = 0; j < m; ++j) {
        const uint32_t id = row[j];
        const float dx = xsub(xyz[3 * (uint64_t)id], mx), dy = xsub(xyz[3 * (uint64_t)id + 1], my),
                    dz = xsub(xyz[3 * (uint64_t)id + 2], mz);
        c[0] = xadd(c[0], xmul(dx, dx));
        c[1] = xadd(c[1], xmul(dx, dy));
        c[2] = xadd(c[2], xmul(dx, dz));
        c[3] = xadd(c[3], xmul(dy, dy));
        c[4] = xadd(c[4], xmul(dy, dz));
        c[5] = xadd(c[5], xmul(dz, dz));
      }
      const float den = fmaxf(xsub(nf, 1.0f), 1.0f);
#pragma unroll
      for (int t = 0; t < 6; ++t) c[t] = xdiv(c[t], den);
      c[0] = xadd(c[0], 1e-4f);
      c[3] = xadd(c[3], 1e-4f);
      c[5] = xadd(c[5], 1e-4f);
    }
    cov[2 * (uint64_t)i] = make_float4(c[0], c[1], c[2], c[3]);
    cov[2 * (uint64_t)i + 1] = make_float4(c[4], c[5], 0.0f, 0.0f);
  }
}

int tci_gicp_covariances(tc_context* ctx, const tc_cloud* cloud, uint32_t k, float4** d_cov) {
  *d_cov = nullptr;
  const uint32_t n = (uint32_t)cloud->n;
  tc_index* ix = nullptr;
  TC_TRY(tc_index_build(ctx, cloud, k, 0.0f, &ix));
  uint32_t* d_idx = nullptr;
  int st = tc_alloc(ctx, &d_idx, (uint64_t)n * k);
  if (st == TC_OK) st = tc_alloc(ctx, d_cov, 2 * (uint64_t)n);
  if (st == TC_OK) st = tci_knn_launch(ctx, ix, nullptr, 0, n, k, 0, true, d_idx, nullptr, nullptr);
  if (st == TC_OK) {
    k_gicp_cov<<<std::max(1, std::min((int)((n + 255) / 256), ctx->sm_count * 8)), 256, 0,
                 ctx->stream>>>(cloud->d_xyz, n, k, d_idx, *d_cov);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) st = tc_fail(ctx, TC_GPU, "covariance launch failed");
  }
  tc_free(ctx, d_idx);
  tc_index_free(ix);
  if (st != TC_OK) {
    tc_free(ctx, *d_cov);
    *d_cov = nullptr;
  }
  return st;
}

int tci_gicp_device(tc_context* ctx, const tc_cloud* src, const tc_index* tgt,
                    const float4* d_src_cov, const float4* d_tgt_cov, const float init[7],
                    uint32_t max_iters, float max_corr_dist, float conv_threshold,
                    tc_icp_result* out, uint32_t* d_match_out) {
  return icp_device_impl(kGicp, ctx, nullptr, src, tgt, nullptr, d_src_cov, d_tgt_cov, init,
                         max_iters, max_corr_dist, conv_threshold, out, d_match_out);
}
