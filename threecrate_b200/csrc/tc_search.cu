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

namespace {

constexpr int kBlock = 128;
using namespace tcs;

// ---------------------------------------------------------------------------------- kNN kernel
template <int K>
__global__ void __launch_bounds__(kBlock)
k_knn(GridParams g, const float4* __restrict__ pts, const uint32_t* __restrict__ cell_start,
      const float4* __restrict__ queries, uint32_t q_begin, uint32_t q_end, uint32_t k,
      int drop_self, uint32_t* __restrict__ idx_out, float* __restrict__ dist_out,
      uint32_t* __restrict__ count_out) {
  const uint32_t qi = q_begin + blockIdx.x * kBlock + threadIdx.x;
  if (qi >= q_end) return;
  const float4 q = __ldg(&queries[qi]);
  const uint32_t qid = __float_as_uint(q.w);  // original query index = output row
  TopK<K> tk;
  grid_search(g, pts, cell_start, q.x, q.y, q.z, tk);
  // kNN(k+1), retain idx != i, truncate k  (point_cloud_ops.rs:91-99); plain kNN otherwise
  uint32_t c = 0;
  const uint64_t row = (uint64_t)qid * k;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const uint64_t key = tk.key[i];
    const uint32_t id = (uint32_t)key;
    if (key != kEmpty && c < k && !(drop_self && id == qid)) {
      idx_out[row + c] = id;
      if (dist_out) dist_out[row + c] = xsqrt(__uint_as_float((uint32_t)(key >> 32)));
      ++c;
    }
  }
  if (count_out) count_out[qid] = c;
  for (uint32_t j = c; j < k; ++j) {
    idx_out[row + j] = TC_NO_INDEX;
    if (dist_out) dist_out[row + j] = INFINITY;
  }
}

// ------------------------------------------------------------------------- symmetric 3x3 eigen
// Cyclic Jacobi in f64 on the f32 covariance the reference would hand to nalgebra's
// symmetric_eigen (normals.rs:181).  Returns the unit eigenvector of the smallest eigenvalue
// (first strict minimum, normals.rs:186-191).
__device__ __forceinline__ void smallest_eigvec(const float cov[6] /*xx,xy,xz,yy,yz,zz*/,
                                                float n[3]) {
  double a00 = cov[0], a01 = cov[1], a02 = cov[2], a11 = cov[3], a12 = cov[4], a22 = cov[5];
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};  // v[row][col]
#define TC_JACOBI(app, aqq, apq, arp, arq, P, Q)                                   \
  if (apq != 0.0) {                                                                \
    const double theta = (aqq - app) / (2.0 * apq);                                \
    const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0)); \
    const double c = rsqrt(t * t + 1.0), s = t * c;                                \
    app = app - t * apq;                                                           \
    aqq = aqq + t * apq;                                                           \
    apq = 0.0;                                                                     \
    const double rp = arp, rq = arq;                                               \
    arp = c * rp - s * rq;                                                         \
    arq = s * rp + c * rq;                                                         \
    _Pragma("unroll") for (int r = 0; r < 3; ++r) {                                \
      const double vp = v[r][P], vq = v[r][Q];                                     \
      v[r][P] = c * vp - s * vq;                                                   \
      v[r][Q] = s * vp + c * vq;                                                   \
    }                                                                              \
  }
  for (int sweep = 0; sweep < 10; ++sweep) {
    const double off = a01 * a01 + a02 * a02 + a12 * a12;
    const double dg = a00 * a00 + a11 * a11 + a22 * a22;
    if (off <= 1e-30 * dg || off == 0.0) break;
    TC_JACOBI(a00, a11, a01, a02, a12, 0, 1)  // (p,q) = (0,1); third index r = 2
    TC_JACOBI(a00, a22, a02, a01, a12, 0, 2)  // (0,2); r = 1  (a01 = a_{r p}, a12 = a_{r q})
    TC_JACOBI(a11, a22, a12, a01, a02, 1, 2)  // (1,2); r = 0
  }
#undef TC_JACOBI
  int m = 0;
  double lm = a00;
  if (a11 < lm) {
    lm = a11;
    m = 1;
  }
  if (a22 < lm) {
    lm = a22;
    m = 2;
  }
  const double ex = m == 0 ? v[0][0] : (m == 1 ? v[0][1] : v[0][2]);
  const double ey = m == 0 ? v[1][0] : (m == 1 ? v[1][1] : v[1][2]);
  const double ez = m == 0 ? v[2][0] : (m == 1 ? v[2][1] : v[2][2]);
  n[0] = (float)ex;
  n[1] = (float)ey;
  n[2] = (float)ez;
}

// ------------------------------------------------------------------------------ normals kernel
template <int K>
__global__ void __launch_bounds__(kBlock)
k_normals(GridParams g, const float4* __restrict__ pts, const uint32_t* __restrict__ cell_start,
          const float* __restrict__ xyz, uint32_t q_begin, uint32_t q_end, uint32_t k, int orient,
          float vpx, float vpy, float vpz, float* __restrict__ out) {
  const uint32_t qi = q_begin + blockIdx.x * kBlock + threadIdx.x;
  if (qi >= q_end) return;
  const float4 q = __ldg(&pts[qi]);
  const uint32_t qid = __float_as_uint(q.w);
  TopK<K> tk;
  grid_search(g, pts, cell_start, q.x, q.y, q.z, tk);

  // neighbourhood = first k of kNN(k+1) with self dropped by index, then self appended last
  // (normals.rs:148-153, 338-340).  Sums are sequential f32 in that order (normals.rs:165-177).
  float sx = 0.0f, sy = 0.0f, sz = 0.0f;
  uint32_t cnt = 0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const uint64_t key = tk.key[i];
    const uint32_t id = (uint32_t)key;
    if (key != kEmpty && cnt < k && id != qid) {
      const float* p = xyz + 3 * (uint64_t)id;
      sx = xadd(sx, __ldg(p + 0));
      sy = xadd(sy, __ldg(p + 1));
      sz = xadd(sz, __ldg(p + 2));
      ++cnt;
    }
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
    uint32_t cnt2 = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const uint64_t key = tk.key[i];
      const uint32_t id = (uint32_t)key;
      if (key != kEmpty && cnt2 < k && id != qid) {
        const float* p = xyz + 3 * (uint64_t)id;
        const float dx = xsub(__ldg(p + 0), cx), dy = xsub(__ldg(p + 1), cy),
                    dz = xsub(__ldg(p + 2), cz);
        c[0] = xadd(c[0], xmul(dx, dx));
        c[1] = xadd(c[1], xmul(dx, dy));
        c[2] = xadd(c[2], xmul(dx, dz));
        c[3] = xadd(c[3], xmul(dy, dy));
        c[4] = xadd(c[4], xmul(dy, dz));
        c[5] = xadd(c[5], xmul(dz, dz));
        ++cnt2;
      }
    }
    {
      const float dx = xsub(q.x, cx), dy = xsub(q.y, cy), dz = xsub(q.z, cz);
      c[0] = xadd(c[0], xmul(dx, dx));
      c[1] = xadd(c[1], xmul(dx, dy));
      c[2] = xadd(c[2], xmul(dx, dz));
      c[3] = xadd(c[3], xmul(dy, dy));
      c[4] = xadd(c[4], xmul(dy, dz));
      c[5] = xadd(c[5], xmul(dz, dz));
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) c[i] = xdiv(c[i], fn);
    smallest_eigvec(c, nrm);
    // renormalise, fall back to +z for a vanishing vector (normals.rs:197-202)
    const float mag =
        xsqrt(xadd(xadd(xmul(nrm[0], nrm[0]), xmul(nrm[1], nrm[1])), xmul(nrm[2], nrm[2])));
    if (mag > 1e-6f) {
      nrm[0] = xdiv(nrm[0], mag);
      nrm[1] = xdiv(nrm[1], mag);
      nrm[2] = xdiv(nrm[2], mag);
    } else {
      nrm[0] = 0.0f;
      nrm[1] = 0.0f;
      nrm[2] = 1.0f;
    }
  }
  if (orient) {  // normals.rs:208-222: flip iff n . normalize(vp - p) < 0
    float tx = xsub(vpx, q.x), ty = xsub(vpy, q.y), tz = xsub(vpz, q.z);
    const float mag = xsqrt(xadd(xadd(xmul(tx, tx), xmul(ty, ty)), xmul(tz, tz)));
    tx = xdiv(tx, mag);
    ty = xdiv(ty, mag);
    tz = xdiv(tz, mag);
    const float d = xadd(xadd(xmul(nrm[0], tx), xmul(nrm[1], ty)), xmul(nrm[2], tz));
    if (d < 0.0f) {
      nrm[0] = -nrm[0];
      nrm[1] = -nrm[1];
      nrm[2] = -nrm[2];
    }
  }
  float* o = out + 6 * (uint64_t)qid;
  o[0] = q.x;
  o[1] = q.y;
  o[2] = q.z;
  o[3] = nrm[0];
  o[4] = nrm[1];
  o[5] = nrm[2];
}

// smallest instantiated list size >= need (0 if unsupported)
constexpr int kSizes[] = {1, 2, 4, 6, 8, 11, 13, 17, 21, 25, 31, 33, 40, 48, 64};
inline int pick_size(uint32_t need) {
  for (int s : kSizes)
    if ((uint32_t)s >= need) return s;
  return 0;
}

}  // namespace

int g_tc_search_flags = 3;
extern "C" void tc_debug_set_search_flags(int flags) { g_tc_search_flags = flags; }

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

int tci_knn_launch(tc_context* ctx, const tc_index* ix, const float4* d_queries_sorted,
                   uint64_t q_begin, uint64_t q_end, uint32_t k, int exclude_self, bool self_query,
                   uint32_t* d_idx_out, float* d_dist_out, uint32_t* d_count_out) {
  if (q_end <= q_begin) return TC_OK;
  const int drop_self = (exclude_self && self_query) ? 1 : 0;
  const uint32_t need = k + (drop_self ? 1u : 0u);
  const int sz = pick_size(need);
  if (sz == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "k too large for the device top-k (max 64 incl. self)");
  const uint32_t nq = (uint32_t)(q_end - q_begin);
  const dim3 grid((nq + kBlock - 1) / kBlock);
  GridParams gp = ix->g;
  gp.flags = g_tc_search_flags;
  TC_DISPATCH_K(sz, (k_knn<KK><<<grid, kBlock, 0, ctx->stream>>>(
                        gp, ix->d_pts, ix->d_cell_start, d_queries_sorted, (uint32_t)q_begin,
                        (uint32_t)q_end, k, drop_self, d_idx_out, d_dist_out, d_count_out)));
  TC_LAUNCHED(ctx);
  return TC_OK;
}

int tci_normals_launch(tc_context* ctx, const tc_index* ix, uint32_t k, int orient,
                       const float vp[3], uint64_t q_begin, uint64_t q_end, float* d_out_aos) {
  if (q_end <= q_begin) return TC_OK;
  const int sz = pick_size(k + 1);
  if (sz == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "k too large for the device top-k (max 63 for normals)");
  const uint32_t nq = (uint32_t)(q_end - q_begin);
  const dim3 grid((nq + kBlock - 1) / kBlock);
  GridParams gp = ix->g;
  gp.flags = g_tc_search_flags;
  TC_DISPATCH_K(sz, (k_normals<KK><<<grid, kBlock, 0, ctx->stream>>>(
                        gp, ix->d_pts, ix->d_cell_start, ix->cloud->d_xyz, (uint32_t)q_begin,
                        (uint32_t)q_end, k, orient, vp[0], vp[1], vp[2], d_out_aos)));
  TC_LAUNCHED(ctx);
  return TC_OK;
}
