// tc_normal.cuh — the PCA-normal epilogue shared by the per-lane kernels (tc_search.cu) and the
// staged-tile kernels (tc_tile.cu): covariance -> eigenvector of the smallest eigenvalue ->
// renormalise -> orient -> NormalPoint3f row (normals.rs:158-222).
#pragma once
#include "tc_internal.cuh"

namespace tcs {

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

// Fast path for the same eigenvector: the smallest root of the characteristic cubic
// f(l) = l^3 - c2 l^2 + c1 l - c0 by Newton's iteration started at 0 in f64 (f is concave and
// increasing left of its smallest root, so the iterates approach it monotonically from below;
// a start right of the root - a covariance that rounding made slightly indefinite - lands left
// of it after one step), then the eigenvector as the largest cross product of two rows of
// A - l I.  ~150 f64 instructions instead of ~900 for the Jacobi sweeps.  Returns false when
// the answer is not trustworthy (the two smallest eigenvalues nearly coincide, so every cross
// product vanishes, or the iteration did not settle): the caller then runs Jacobi.
__device__ __forceinline__ bool smallest_eigvec_newton(const float cov[6], float n[3]) {
  const double a00 = cov[0], a01 = cov[1], a02 = cov[2], a11 = cov[3], a12 = cov[4], a22 = cov[5];
  const double c2 = a00 + a11 + a22;
  const double m0 = a11 * a22 - a12 * a12, m1 = a01 * a22 - a12 * a02, m2 = a01 * a12 - a11 * a02;
  const double c1 = m0 + (a00 * a22 - a02 * a02) + (a00 * a11 - a01 * a01);
  const double c0 = a00 * m0 - a01 * m1 + a02 * m2;
  if (!(c2 > 0.0)) return false;
  const double tol = 1e-13 * c2;
  double l = 0.0;
  bool settled = false;
#pragma unroll 1
  for (int it = 0; it < 12; ++it) {
    const double f = ((l - c2) * l + c1) * l - c0;
    const double fp = (3.0 * l - 2.0 * c2) * l + c1;
    if (!(fp > 0.0)) break;
    const double d = f / fp;
    l -= d;
    if (fabs(d) <= tol) {
      settled = true;
      break;
    }
  }
  if (!settled) return false;
  const double r00 = a00 - l, r11 = a11 - l, r22 = a22 - l;
  // rows r0 = (r00, a01, a02), r1 = (a01, r11, a12), r2 = (a02, a12, r22)
  const double x01 = a01 * a12 - a02 * r11, y01 = a02 * a01 - r00 * a12, z01 = r00 * r11 - a01 * a01;
  const double x02 = a01 * r22 - a02 * a12, y02 = a02 * a02 - r00 * r22, z02 = r00 * a12 - a01 * a02;
  const double x12 = r11 * r22 - a12 * a12, y12 = a12 * a02 - a01 * r22, z12 = a01 * a12 - r11 * a02;
  const double n01 = x01 * x01 + y01 * y01 + z01 * z01;
  const double n02 = x02 * x02 + y02 * y02 + z02 * z02;
  const double n12 = x12 * x12 + y12 * y12 + z12 * z12;
  double ex = x01, ey = y01, ez = z01, nn = n01;
  if (n02 > nn) {
    ex = x02;
    ey = y02;
    ez = z02;
    nn = n02;
  }
  if (n12 > nn) {
    ex = x12;
    ey = y12;
    ez = z12;
    nn = n12;
  }
  const double c22 = c2 * c2;
  if (!(nn > 1e-16 * c22 * c22)) return false;
  const double inv = rsqrt(nn);
  n[0] = (float)(ex * inv);
  n[1] = (float)(ey * inv);
  n[2] = (float)(ez * inv);
  return true;
}

// covariance (f32, already divided by n) -> unit eigenvector of the smallest eigenvalue,
// renormalised, +z for a vanishing vector (normals.rs:181-202)
__device__ __forceinline__ void normal_from_cov(const float c[6], float nrm[3], bool fast = false) {
  if (!fast || !smallest_eigvec_newton(c, nrm)) smallest_eigvec(c, nrm);
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
// orientation (normals.rs:208-222: flip iff n . normalize(vp - p) < 0) and the NormalPoint3f row
__device__ __forceinline__ void write_normal(float nrm[3], const float4 q, uint32_t qid, int orient,
                                             float vpx, float vpy, float vpz,
                                             float* __restrict__ out) {
  if (orient) {
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

}  // namespace tcs
