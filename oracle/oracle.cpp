// oracle.cpp — CPU restatement of threecrate's kNN -> normals -> point-to-plane ICP path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under threecrate_b200/ may include, link, import or
// execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs use it, and only as the checker / reported CPU baseline.
//
// PARITY UNPINNED: the reference is pure Rust (no cargo/rustc in this image), ships no golden
// vectors for this path, and its linear algebra lives in nalgebra 0.34
// (not vendored under /root/reference).  What IS pinned here: every assertion of the
// reference's own inline tests for the path (tests/test_oracle_*.py), the kd-tree vs
// brute-force agreement the reference itself tests, and independent numpy/scipy cross-checks.
//
// Each function cites the reference lines it follows (paths relative to /root/reference).
// Arithmetic follows Rust semantics: f32 throughout, no FMA contraction (build with
// -ffp-contract=off), IEEE sqrt/div.
//
// Parallelism mirrors the reference: OpenMP only over the loops the reference hands to rayon
// (normals.rs:306-307, registration.rs:92-93); everything else is serial as in the reference.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct P3 {
  float x, y, z;
  float operator[](int a) const { return a == 0 ? x : (a == 1 ? y : z); }
};

constexpr uint32_t NIL = 0xFFFFFFFFu;  // nearest_neighbor.rs:8

// ------------------------------------------------------------------------------------------
// KdTree — threecrate-algorithms/src/nearest_neighbor.rs:17-168
// ------------------------------------------------------------------------------------------
struct KdNode {  // nearest_neighbor.rs:17-23
  P3 point;
  uint64_t original_index;
  uint32_t left, right;
  uint8_t axis;
};

struct PI {
  P3 p;
  uint64_t i;
};

struct KdTree {
  std::vector<KdNode> nodes;
  bool has_root = false;
  uint64_t n_points = 0;
};

// nearest_neighbor.rs:134-159 — Lomuto partition, pivot = last element, `<=` comparison.
size_t kd_partition(std::vector<PI>& pts, size_t start, size_t end, int axis) {
  const float pivot_value = pts[end].p[axis];
  size_t i = start;
  for (size_t j = start; j < end; ++j) {
    if (pts[j].p[axis] <= pivot_value) {
      std::swap(pts[i], pts[j]);
      ++i;
    }
  }
  std::swap(pts[i], pts[end]);
  return i;
}

// nearest_neighbor.rs:112-131
void kd_select_median(std::vector<PI>& pts, size_t start, size_t end, size_t target, int axis) {
  size_t left = start, right = end;
  while (left < right) {
    size_t pivot_idx = kd_partition(pts, left, right, axis);
    if (pivot_idx == target) return;
    if (pivot_idx < target)
      left = pivot_idx + 1;
    else
      right = pivot_idx - 1;
  }
}

// nearest_neighbor.rs:66-109 — pre-order slots, axis = depth % 3, median index (start+end)/2.
uint32_t kd_build(std::vector<KdNode>& nodes, std::vector<PI>& pts, size_t depth, size_t start,
                  size_t end) {
  const int axis = (int)(depth % 3);
  const size_t median_idx = (start + end) / 2;
  kd_select_median(pts, start, end, median_idx, axis);
  const uint32_t my_idx = (uint32_t)nodes.size();
  nodes.push_back(KdNode{pts[median_idx].p, pts[median_idx].i, NIL, NIL, (uint8_t)axis});
  uint32_t left = NIL, right = NIL;
  if (median_idx > start) left = kd_build(nodes, pts, depth + 1, start, median_idx - 1);
  if (median_idx < end) right = kd_build(nodes, pts, depth + 1, median_idx + 1, end);
  nodes[my_idx].left = left;
  nodes[my_idx].right = right;
  return my_idx;
}

// nearest_neighbor.rs:37-60
KdTree* kd_new(const float* xyz, uint64_t n) {
  KdTree* t = new KdTree();
  t->n_points = n;
  if (n == 0) return t;
  std::vector<PI> pts(n);
  for (uint64_t i = 0; i < n; ++i) pts[i] = PI{P3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, i};
  t->nodes.reserve(n);
  kd_build(t->nodes, pts, 0, 0, n - 1);
  t->has_root = true;
  return t;
}

// nearest_neighbor.rs:162-167 — (dx*dx + dy*dy) + dz*dz, f32, unfused.
inline float dist_sq(const P3& a, const P3& b) {
  const float dx = a.x - b.x;
  const float dy = a.y - b.y;
  const float dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}

// ------------------------------------------------------------------------------------------
// Rust std::collections::BinaryHeap<Neighbor> emulation [upstream std; SURVEY.md App. A.1].
// Neighbor::cmp compares distance only (nearest_neighbor.rs:316-324), NaN == Equal.
// ------------------------------------------------------------------------------------------
struct Neighbor {
  float distance;  // squared distance during traversal
  uint64_t index;
};
// partial_cmp-derived operators (distance only)
inline bool nb_le(const Neighbor& a, const Neighbor& b) { return !(a.distance > b.distance); }
inline bool nb_ge(const Neighbor& a, const Neighbor& b) { return !(a.distance < b.distance); }
inline bool nb_lt(const Neighbor& a, const Neighbor& b) { return a.distance < b.distance; }

struct RustHeap {
  std::vector<Neighbor> data;
  size_t len() const { return data.size(); }
  // sift_up(start, pos): move hole up while element > parent
  size_t sift_up(size_t start, size_t pos) {
    Neighbor elem = data[pos];
    while (pos > start) {
      size_t parent = (pos - 1) / 2;
      if (nb_le(elem, data[parent])) break;
      data[pos] = data[parent];
      pos = parent;
    }
    data[pos] = elem;
    return pos;
  }
  void push(const Neighbor& n) {
    size_t old_len = data.size();
    data.push_back(n);
    sift_up(0, old_len);
  }
  void sift_down_to_bottom(size_t pos) {
    const size_t end = data.size();
    const size_t start = pos;
    Neighbor elem = data[pos];
    size_t child = 2 * pos + 1;
    const size_t lim = end >= 2 ? end - 2 : 0;  // end.saturating_sub(2)
    while (child <= lim && end >= 2) {
      child += nb_le(data[child], data[child + 1]) ? 1 : 0;
      data[pos] = data[child];
      pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1) {
      data[pos] = data[child];
      pos = child;
    }
    data[pos] = elem;
    sift_up(start, pos);
  }
  void pop() {
    Neighbor item = data.back();
    data.pop_back();
    if (!data.empty()) {
      std::swap(item, data[0]);
      sift_down_to_bottom(0);
    }
  }
  void sift_down_range(size_t pos, size_t end) {
    Neighbor elem = data[pos];
    size_t child = 2 * pos + 1;
    const size_t lim = end >= 2 ? end - 2 : 0;
    while (child <= lim && end >= 2) {
      child += nb_le(data[child], data[child + 1]) ? 1 : 0;
      if (nb_ge(elem, data[child])) {
        data[pos] = elem;
        return;
      }
      data[pos] = data[child];
      pos = child;
      child = 2 * pos + 1;
    }
    if (child == end - 1 && nb_lt(elem, data[child])) {
      data[pos] = data[child];
      pos = child;
    }
    data[pos] = elem;
  }
  // into_sorted_vec: in-place heapsort, ascending
  void into_sorted() {
    size_t end = data.size();
    while (end > 1) {
      --end;
      std::swap(data[0], data[end]);
      sift_down_range(0, end);
    }
  }
};

// nearest_neighbor.rs:177-251 — iterative traversal with explicit LIFO stack.
// Output: (original index, squared distance) ascending; caller takes sqrt (line 249).
void kd_find_k_nearest(const KdTree& t, const P3& q, size_t k, std::vector<Neighbor>& out,
                       std::vector<uint32_t>& stack) {
  out.clear();
  if (k == 0 || t.n_points == 0) return;
  RustHeap heap;
  heap.data.swap(out);
  heap.data.clear();
  heap.data.reserve(k + 1);
  stack.clear();
  if (t.has_root) stack.push_back(0);
  while (!stack.empty()) {
    const uint32_t idx = stack.back();
    stack.pop_back();
    const KdNode& node = t.nodes[idx];
    const float d2 = dist_sq(node.point, q);
    if (heap.len() < k) {
      heap.push(Neighbor{d2, node.original_index});
    } else {
      if (d2 < heap.data[0].distance) {  // strict <, nearest_neighbor.rs:206
        heap.pop();
        heap.push(Neighbor{d2, node.original_index});
      }
    }
    const float query_val = q[node.axis];
    const float node_val = node.point[node.axis];
    const float axis_dist = query_val - node_val;
    const float axis_dist_sq = axis_dist * axis_dist;
    uint32_t near, far;
    if (query_val <= node_val) {
      near = node.left;
      far = node.right;
    } else {
      near = node.right;
      far = node.left;
    }
    // heap is never empty here (we pushed or it was full), nearest_neighbor.rs:232-236
    const bool search_far = heap.len() < k || axis_dist_sq < heap.data[0].distance;
    if (search_far && far != NIL) stack.push_back(far);
    if (near != NIL) stack.push_back(near);
  }
  heap.into_sorted();
  out.swap(heap.data);
}

// nearest_neighbor.rs:254-298 — radius search; stable sort by sqrt distance.
void kd_find_radius(const KdTree& t, const P3& q, float radius,
                    std::vector<std::pair<uint64_t, float>>& out, std::vector<uint32_t>& stack) {
  out.clear();
  if (radius <= 0.0f || t.n_points == 0) return;
  const float radius_sq = radius * radius;
  stack.clear();
  if (t.has_root) stack.push_back(0);
  while (!stack.empty()) {
    const uint32_t idx = stack.back();
    stack.pop_back();
    const KdNode& node = t.nodes[idx];
    const float d2 = dist_sq(node.point, q);
    if (d2 <= radius_sq) out.emplace_back(node.original_index, std::sqrt(d2));
    const float query_val = q[node.axis];
    const float node_val = node.point[node.axis];
    const float axis_dist = query_val - node_val;
    uint32_t near, far;
    if (query_val <= node_val) {
      near = node.left;
      far = node.right;
    } else {
      near = node.right;
      far = node.left;
    }
    if (axis_dist * axis_dist <= radius_sq) {
      if (far != NIL) stack.push_back(far);
    }
    if (near != NIL) stack.push_back(near);
  }
  // Rust sort_by is a stable sort (line 296)
  std::stable_sort(out.begin(), out.end(),
                   [](const std::pair<uint64_t, float>& a, const std::pair<uint64_t, float>& b) {
                     return a.second < b.second;
                   });
}

// ------------------------------------------------------------------------------------------
// nalgebra 0.34 restatements [upstream, not under /root/reference; written from the published
// algorithm, SURVEY.md Appendix A.3-A.5].  All f32.
// ------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
inline V3 cross(const V3& a, const V3& b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Quat {  // storage order [i, j, k, w]
  float i, j, k, w;
};
struct Iso {
  Quat q;
  V3 t;
};

// UnitQuaternion * Vector3:  t = 2 (qv × v);  (t*w + qv × t) + v
inline V3 quat_rotate(const Quat& q, const V3& v) {
  const V3 qv{q.i, q.j, q.k};
  V3 t = cross(qv, v);
  t = V3{t.x * 2.0f, t.y * 2.0f, t.z * 2.0f};
  const V3 c = cross(qv, t);
  return V3{t.x * q.w + c.x + v.x, t.y * q.w + c.y + v.y, t.z * q.w + c.z + v.z};
}
// Isometry3 * Point3 = rotation * p + translation
inline V3 iso_apply(const Iso& T, const V3& p) {
  const V3 r = quat_rotate(T.q, p);
  return V3{r.x + T.t.x, r.y + T.t.y, r.z + T.t.z};
}
// Hamilton product, no renormalisation
inline Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k;
  r.i = a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j;
  r.j = a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i;
  r.k = a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w;
  return r;
}
// Isometry3 * Isometry3 = (q1 q2, t1 + q1·t2)
inline Iso iso_mul(const Iso& a, const Iso& b) {
  const V3 s = quat_rotate(a.q, b.t);
  return Iso{quat_mul(a.q, b.q), V3{a.t.x + s.x, a.t.y + s.y, a.t.z + s.z}};
}
// UnitQuaternion::from_axis_angle(unit axis, angle): (w = cos(a/2), ijk = axis * sin(a/2))
inline Quat quat_axis_angle(int axis, float angle) {
  const float h = angle / 2.0f;
  const float s = std::sin(h), c = std::cos(h);
  Quat q{0, 0, 0, c};
  if (axis == 0) q.i = s;
  if (axis == 1) q.j = s;
  if (axis == 2) q.k = s;
  return q;
}

// --- Matrix3::symmetric_eigen (SymmetricEigen::new): scale by max-abs, Householder
// tridiagonalisation, implicit symmetric QR with Wilkinson shifts to f32 epsilon.
// Eigenvalues unsorted; eigenvectors as columns of q (column-major q[c][r]).
struct Eig3 {
  float val[3];
  float vec[3][3];  // vec[c][r] = component r of eigenvector c
};

inline float rsign(float x) { return std::copysign(1.0f, x); }  // Rust f32::signum
// ComplexField::to_exp for reals: (|x|, x/|x|), or (0, 1) when x == 0
inline void to_exp(float x, float& modulus, float& sign) {
  modulus = std::fabs(x);
  sign = modulus != 0.0f ? x / modulus : 1.0f;
}

// GivensRotation::cancel_y(v): rotation (c, s) and r with R v = (r, 0); None when v.y == 0.
inline bool givens_cancel_y(float vx, float vy, float& c, float& s, float& norm) {
  if (vy == 0.0f) return false;
  float mod0, sign0;
  to_exp(vx, mod0, sign0);
  const float denom = std::sqrt(mod0 * mod0 + vy * vy);
  c = mod0 / denom;
  s = -vy / (sign0 * denom);
  norm = sign0 * denom;
  return true;
}
// GivensRotation::try_new(c, s, eps)
inline bool givens_try_new(float cc, float ss, float eps, float& c, float& s) {
  float mod0, sign0;
  to_exp(cc, mod0, sign0);
  const float denom = std::sqrt(mod0 * mod0 + ss * ss);
  if (denom > eps) {
    const float norm = sign0 * denom;
    c = mod0 / denom;
    s = ss / norm;
    return true;
  }
  return false;
}

inline float wilkinson_shift(float tmm, float tnn, float tmn) {
  const float sq_tmn = tmn * tmn;
  if (sq_tmn != 0.0f) {
    const float d = (tmm - tnn) * 0.5f;
    return tnn - sq_tmn / (d + rsign(d) * std::sqrt(d * d + sq_tmn));
  }
  return tnn;
}

// SymmetricEigen::delimit_subproblem(diag, off_diag, end, eps) -> (start, end)
inline void delimit_subproblem(const float* diag, float* off, int end_in, float eps, int& start,
                               int& end) {
  int n = end_in;
  while (n > 0) {
    const int m = n - 1;
    if (std::fabs(off[m]) > eps * (std::fabs(diag[n]) + std::fabs(diag[m]))) break;
    n -= 1;
  }
  if (n == 0) {
    start = 0;
    end = 0;
    return;
  }
  int new_start = n - 1;
  while (new_start > 0) {
    const int m = new_start - 1;
    if (off[m] == 0.0f ||
        std::fabs(off[m]) <= eps * (std::fabs(diag[new_start]) + std::fabs(diag[m]))) {
      off[m] = 0.0f;
      break;
    }
    new_start -= 1;
  }
  start = new_start;
  end = n;
}

// GivensRotation{c,s}.rotate_rows(q.fixed_columns_mut::<2>(i)):  for every row r,
//   (a, b) = (q[r,i], q[r,i+1]);  q[r,i] = a c + s b;  q[r,i+1] = -s a + b c
inline void rotate_cols(float q[3][3] /*q[col][row]*/, int i, float c, float s) {
  for (int r = 0; r < 3; ++r) {
    const float a = q[i][r];
    const float b = q[i + 1][r];
    q[i][r] = a * c + s * b;
    q[i + 1][r] = -s * a + b * c;
  }
}

Eig3 symmetric_eigen3(const float m_in[3][3]) {  // m_in[r][c]
  Eig3 out;
  float m[3][3];
  float amax = 0.0f;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      m[r][c] = m_in[r][c];
      amax = std::max(amax, std::fabs(m_in[r][c]));
    }
  if (amax != 0.0f)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m[r][c] = m[r][c] / amax;  // unscale_mut

  // --- SymmetricTridiagonal::new + unpack for dim 3 (lower triangle is the one read).
  // Step i=0: Householder axis u from column 0 rows 1..2 (householder::reflection_axis_mut),
  // symmetric rank-2 update of the trailing 2x2.  Step i=1 is a 1-element reflection (u=±1),
  // which only negates the (2,1) entry.  unpack() returns |off_diagonal| and folds the signs
  // into Q (assemble_q + reflect_with_sign); here Q = H * diag(1, s1, s2).
  float diag[3], off[2];
  float q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};  // q[c][r]
  {
    float u0 = m[1][0], u1 = m[2][0];
    const float sq_norm = u0 * u0 + u1 * u1;
    const float norm = std::sqrt(sq_norm);
    float modulus, sign;
    to_exp(u0, modulus, sign);
    const float signed_norm = sign * norm;
    const float factor = (sq_norm + modulus * norm) * 2.0f;
    u0 = u0 + signed_norm;
    float a00 = m[1][1], a10 = m[2][1], a11 = m[2][2];  // trailing block, lower triangle
    float off0;
    bool not_zero = false;
    if (factor != 0.0f) {
      const float fs = std::sqrt(factor);
      u0 = u0 / fs;
      u1 = u1 / fs;
      const float un = std::sqrt(u0 * u0 + u1 * u1);  // second normalisation
      u0 = u0 / un;
      u1 = u1 / un;
      off0 = -signed_norm;
      not_zero = true;
    } else {
      off0 = signed_norm;
    }
    if (not_zero) {
      // p = 2 M u ; dot = u.p ; M -= p u^T ; M -= u p^T ; M += 2 dot u u^T
      const float p0 = (a00 * u0 + a10 * u1) * 2.0f;
      const float p1 = (a10 * u0 + a11 * u1) * 2.0f;
      const float dt = u0 * p0 + u1 * p1;
      a00 = a00 - p0 * u0;
      a10 = a10 - p1 * u0;
      a11 = a11 - p1 * u1;
      a00 = a00 - u0 * p0;
      a10 = a10 - u1 * p0;
      a11 = a11 - u1 * p1;
      const float d2 = dt * 2.0f;
      a00 = a00 + d2 * u0 * u0;
      a10 = a10 + d2 * u1 * u0;
      a11 = a11 + d2 * u1 * u1;
      // H = I - 2 u u^T on rows/cols 1..2
      q[1][1] = 1.0f - 2.0f * u0 * u0;
      q[1][2] = -2.0f * u1 * u0;
      q[2][1] = -2.0f * u0 * u1;
      q[2][2] = 1.0f - 2.0f * u1 * u1;
    }
    const float off1 = -a10;  // 1-element reflection negates the entry
    diag[0] = m[0][0];
    diag[1] = a00;
    diag[2] = a11;
    // T has off-diagonals (off0, a10) w.r.t. basis H; take moduli and fold signs into Q.
    const float s1 = rsign(off0);
    const float s2 = s1 * rsign(a10);
    off[0] = std::fabs(off0);
    off[1] = std::fabs(off1);
    for (int r = 0; r < 3; ++r) {
      q[1][r] = q[1][r] * s1;
      q[2][r] = q[2][r] * s2;
    }
  }

  const float eps = std::numeric_limits<float>::epsilon();
  int start, end;
  delimit_subproblem(diag, off, 2, eps, start, end);
  int niter = 0;
  while (end != start) {
    const int subdim = end - start + 1;
    if (subdim > 2) {
      const int mi = end - 1;
      const int n = end;
      float vx = diag[start] - wilkinson_shift(diag[mi], diag[n], off[mi]);
      float vy = off[start];
      for (int i = start; i < n; ++i) {
        const int j = i + 1;
        float c, s, norm;
        if (givens_cancel_y(vx, vy, c, s, norm)) {
          if (i > start) off[i - 1] = norm;
          const float mii = diag[i], mjj = diag[j], mij = off[i];
          const float cc = c * c, ss = s * s, cs = c * s;
          const float b = cs * 2.0f * mij;
          diag[i] = (cc * mii + ss * mjj) - b;
          diag[j] = (ss * mii + cc * mjj) + b;
          off[i] = cs * (mii - mjj) + mij * (cc - ss);
          if (i != n - 1) {
            vx = off[i];
            vy = -s * off[i + 1];
            off[i + 1] *= c;
          }
          rotate_cols(q, i, c, -s);  // rot.inverse().rotate_rows(...)
        } else {
          break;
        }
      }
      if (std::fabs(off[mi]) <= eps * (std::fabs(diag[mi]) + std::fabs(diag[n]))) end -= 1;
    } else if (subdim == 2) {
      // compute_2x2_eigvals on [[a, b],[b, d]]
      const float a = diag[start], b = off[start], d = diag[start + 1];
      const float val = (a - d) * 0.5f;
      const float discr = b * b + val * val;
      const float sq = std::sqrt(discr);
      const float half_tra = (a + d) * 0.5f;
      float e0 = half_tra + sq, e1 = half_tra - sq;
      // Matrix2::eigenvalues() goes through Schur::do_decompose's 2x2 special case
      // (compute_2x2_basis), which orders the pair so that the FIRST eigenvalue is the one
      // with the larger |eigval - m11| ("choose the one that yields a larger x component").
      if (!(std::fabs(e0 - d) > std::fabs(e1 - d))) std::swap(e0, e1);
      const float bx = e0 - d, by = b;
      diag[start] = e0;
      diag[start + 1] = e1;
      float c, s;
      if (givens_try_new(bx, by, eps, c, s)) rotate_cols(q, start, c, s);
      end -= 1;
    }
    delimit_subproblem(diag, off, end, eps, start, end);
    if (++niter == 100000) break;  // SymmetricEigen::new is unbounded (max_niter = 0)
  }
  for (int i = 0; i < 3; ++i) {
    out.val[i] = diag[i] * amax;
    for (int r = 0; r < 3; ++r) out.vec[i][r] = q[i][r];
  }
  return out;
}

// f64 cyclic Jacobi on a symmetric 3x3 (independent cross-check + eigengap analysis).
void jacobi3_f64(const double a_in[3][3], double val[3], double vec[3][3] /*vec[c][r]*/) {
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) a[r][c] = a_in[r][c];
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double offn = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double dn = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (offn <= 1e-40 * dn || offn == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {  // V <- V J   (v[c][r]: column c)
          const double vkp = v[p][k], vkq = v[q][k];
          v[p][k] = c * vkp - s * vkq;
          v[q][k] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) {
    val[i] = a[i][i];
    for (int r = 0; r < 3; ++r) vec[i][r] = v[i][r];
  }
}

// --- Matrix6::cholesky() / solve (column-oriented, None on non-positive pivot), f32.
bool cholesky6(float a[6][6] /*a[r][c], overwritten with L in lower triangle*/) {
  for (int j = 0; j < 6; ++j) {
    for (int k = 0; k < j; ++k) {
      const float factor = -a[j][k];
      for (int r = j; r < 6; ++r) a[r][j] = factor * a[r][k] + a[r][j];
    }
    const float diag = a[j][j];
    if (!(diag > 0.0f)) return false;  // zero, negative or NaN pivot -> None
    const float denom = std::sqrt(diag);
    a[j][j] = denom;
    for (int r = j + 1; r < 6; ++r) a[r][j] = a[r][j] / denom;
  }
  return true;
}
void cholesky6_solve(const float L[6][6], float b[6]) {
  // L y = b (forward, column-oriented)
  for (int i = 0; i < 6; ++i) {
    const float coeff = b[i] / L[i][i];
    b[i] = coeff;
    for (int r = i + 1; r < 6; ++r) b[r] = -coeff * L[r][i] + b[r];
  }
  // L^T x = y (backward, dot-oriented)
  for (int i = 5; i >= 0; --i) {
    float d = 0.0f;
    for (int r = i + 1; r < 6; ++r) d = d + L[r][i] * b[r];
    b[i] = (b[i] - d) / L[i][i];
  }
}
// --- Matrix6::lu().solve: partial pivoting; None if a pivot is exactly zero.
bool lu6_solve(float a[6][6], float b[6]) {
  for (int i = 0; i < 6; ++i) {
    int piv = i;
    float best = std::fabs(a[i][i]);
    for (int r = i + 1; r < 6; ++r)
      if (std::fabs(a[r][i]) > best) {
        best = std::fabs(a[r][i]);
        piv = r;
      }
    if (a[piv][i] == 0.0f) return false;
    if (piv != i) {
      for (int c = 0; c < 6; ++c) std::swap(a[i][c], a[piv][c]);
      std::swap(b[i], b[piv]);
    }
    const float inv = 1.0f / a[i][i];
    for (int r = i + 1; r < 6; ++r) {
      a[r][i] = a[r][i] * inv;
      for (int c = i + 1; c < 6; ++c) a[r][c] = a[r][c] - a[r][i] * a[i][c];
    }
  }
  for (int i = 0; i < 6; ++i)
    for (int r = i + 1; r < 6; ++r) b[r] = b[r] - a[r][i] * b[i];
  for (int i = 5; i >= 0; --i) {
    for (int c = i + 1; c < 6; ++c) b[i] = b[i] - a[i][c] * b[c];
    b[i] = b[i] / a[i][i];
  }
  return true;
}

// ------------------------------------------------------------------------------------------
// normals.rs
// ------------------------------------------------------------------------------------------
// normals.rs:158-205
V3 compute_normal_pca(const float* xyz, const std::vector<uint64_t>& idx) {
  if (idx.size() < 3) return V3{0.0f, 0.0f, 1.0f};
  float cx = 0.0f, cy = 0.0f, cz = 0.0f;
  for (uint64_t i : idx) {
    cx += xyz[3 * i];
    cy += xyz[3 * i + 1];
    cz += xyz[3 * i + 2];
  }
  const float n = (float)idx.size();
  cx /= n;
  cy /= n;
  cz /= n;
  float cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (uint64_t i : idx) {
    const float d[3] = {xyz[3 * i] - cx, xyz[3 * i + 1] - cy, xyz[3 * i + 2] - cz};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[r][c] += d[r] * d[c];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cov[r][c] /= n;
  const Eig3 e = symmetric_eigen3(cov);
  int min_idx = 0;
  for (int i = 1; i < 3; ++i)
    if (e.val[i] < e.val[min_idx]) min_idx = i;  // strict <, first wins (normals.rs:186-191)
  V3 nrm{e.vec[min_idx][0], e.vec[min_idx][1], e.vec[min_idx][2]};
  const float mag = std::sqrt(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
  if (mag > 1e-6f) {
    nrm = V3{nrm.x / mag, nrm.y / mag, nrm.z / mag};
  } else {
    nrm = V3{0.0f, 0.0f, 1.0f};
  }
  return nrm;
}

// normals.rs:208-222
V3 orient_normal(V3 n, V3 p, V3 vp) {
  V3 tv{vp.x - p.x, vp.y - p.y, vp.z - p.z};
  const float mag = std::sqrt(tv.x * tv.x + tv.y * tv.y + tv.z * tv.z);
  tv = V3{tv.x / mag, tv.y / mag, tv.z / mag};
  const float d = dot(n, tv);
  if (d < 0.0f) return V3{-n.x, -n.y, -n.z};
  return n;
}

// normals.rs:148-153 — kNN(k+1), drop self by index, take k
void knn_minus_self(const KdTree& t, const float* xyz, uint64_t i, size_t k,
                    std::vector<uint64_t>& out, std::vector<Neighbor>& tmp,
                    std::vector<uint32_t>& stack) {
  const P3 q{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  kd_find_k_nearest(t, q, k + 1, tmp, stack);
  out.clear();
  for (const Neighbor& nb : tmp) {
    if (nb.index == i) continue;
    if (out.size() >= k) break;
    out.push_back(nb.index);
  }
}

}  // namespace

// ==========================================================================================
// C API (ctypes)
// ==========================================================================================
extern "C" {

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void* orc_kdtree_new(const float* xyz, uint64_t n) { return kd_new(xyz, n); }
void orc_kdtree_free(void* t) { delete (KdTree*)t; }

// KdTree::find_k_nearest (nearest_neighbor.rs:177-251). Returns count; dist_out = sqrt(d2).
uint64_t orc_kdtree_knn(void* tree, const float* q3, uint64_t k, uint64_t* idx_out,
                        float* dist_out, float* d2_out) {
  std::vector<Neighbor> out;
  std::vector<uint32_t> stack;
  kd_find_k_nearest(*(KdTree*)tree, P3{q3[0], q3[1], q3[2]}, k, out, stack);
  for (size_t j = 0; j < out.size(); ++j) {
    idx_out[j] = out[j].index;
    if (dist_out) dist_out[j] = std::sqrt(out[j].distance);
    if (d2_out) d2_out[j] = out[j].distance;
  }
  return out.size();
}

// Batch of queries through one tree, OpenMP over queries. Row stride k; pad idx = UINT64_MAX.
void orc_kdtree_knn_batch(void* tree, const float* q, uint64_t nq, uint64_t k, uint64_t* idx_out,
                          float* d2_out, uint64_t* count_out, int threads) {
  const KdTree& t = *(KdTree*)tree;
#pragma omp parallel num_threads(threads > 0 ? threads : orc_max_threads())
  {
    std::vector<Neighbor> out;
    std::vector<uint32_t> stack;
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)nq; ++i) {
      kd_find_k_nearest(t, P3{q[3 * i], q[3 * i + 1], q[3 * i + 2]}, k, out, stack);
      for (size_t j = 0; j < k; ++j) {
        idx_out[i * k + j] = j < out.size() ? out[j].index : UINT64_MAX;
        d2_out[i * k + j] = j < out.size() ? out[j].distance : INFINITY;
      }
      if (count_out) count_out[i] = out.size();
    }
  }
}

// KdTree::find_radius_neighbors (nearest_neighbor.rs:254-298). cap = capacity of outputs.
uint64_t orc_kdtree_radius(void* tree, const float* q3, float radius, uint64_t* idx_out,
                           float* dist_out, uint64_t cap) {
  std::vector<std::pair<uint64_t, float>> out;
  std::vector<uint32_t> stack;
  kd_find_radius(*(KdTree*)tree, P3{q3[0], q3[1], q3[2]}, radius, out, stack);
  for (size_t j = 0; j < out.size() && j < cap; ++j) {
    idx_out[j] = out[j].first;
    dist_out[j] = out[j].second;
  }
  return out.size();
}

// Canonical brute force: the k smallest under ascending (d2 bits, index) — the tie rule the
// GPU path adopts (SimdBruteForceSearch, simd_distance.rs:370-385,444-452; BruteForceSearch
// nearest_neighbor.rs:340-362 modulo sqrt). OpenMP over queries.
void orc_brute_knn_batch(const float* xyz, uint64_t n, const float* q, uint64_t nq, uint64_t k,
                         uint64_t* idx_out, float* d2_out, int threads) {
#pragma omp parallel num_threads(threads > 0 ? threads : orc_max_threads())
  {
    std::vector<std::pair<float, uint64_t>> all(n);
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)nq; ++i) {
      const P3 qp{q[3 * i], q[3 * i + 1], q[3 * i + 2]};
      for (uint64_t j = 0; j < n; ++j)
        all[j] = {dist_sq(P3{xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]}, qp), j};
      const size_t kk = (size_t)std::min<uint64_t>(k, n);
      std::partial_sort(all.begin(), all.begin() + kk, all.end());
      for (size_t j = 0; j < k; ++j) {
        idx_out[i * k + j] = j < kk ? all[j].second : UINT64_MAX;
        d2_out[i * k + j] = j < kk ? all[j].first : INFINITY;
      }
    }
  }
}

// PointCloudNeighbors::k_nearest_neighbors (point_cloud_ops.rs:80-105): kd kNN(k+1), retain
// idx != i, truncate k.  The reference loop is serial; `threads` only speeds the checker up.
void orc_k_nearest_neighbors(const float* xyz, uint64_t n, uint64_t k, uint64_t* idx_out,
                             float* dist_out, uint64_t* count_out, int threads) {
  if (n == 0 || k == 0) return;
  KdTree* t = kd_new(xyz, n);
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
  {
    std::vector<Neighbor> out;
    std::vector<uint32_t> stack;
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      kd_find_k_nearest(*t, P3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, k + 1, out, stack);
      size_t c = 0;
      for (const Neighbor& nb : out) {
        if (nb.index == (uint64_t)i) continue;
        if (c >= k) break;
        idx_out[i * k + c] = nb.index;
        dist_out[i * k + c] = std::sqrt(nb.distance);
        ++c;
      }
      count_out[i] = c;
      for (; c < k; ++c) {
        idx_out[i * k + c] = UINT64_MAX;
        dist_out[i * k + c] = INFINITY;
      }
    }
  }
  delete t;
}

// estimate_normals_with_config (normals.rs:257-357).
// radius < 0 => None.  viewpoint == NULL => default (bbox centre + (0,0,extent)).
// out: N x 6 f32 (NormalPoint3f: position, normal).  Returns 0 OK, 1 InvalidData.
// OpenMP mirrors rayon's into_par_iter over points (normals.rs:306-307); kd build is serial.
int orc_estimate_normals(const float* xyz, uint64_t n, uint64_t k, float radius,
                         int consistent_orientation, const float* viewpoint, float* out,
                         int threads) {
  if (n == 0) return 0;   // normals.rs:261-263
  if (k < 3) return 1;    // normals.rs:265-269
  KdTree* t = kd_new(xyz, n);
  V3 vp;
  if (viewpoint) {
    vp = V3{viewpoint[0], viewpoint[1], viewpoint[2]};
  } else {  // normals.rs:275-303
    float mnx = xyz[0], mny = xyz[1], mnz = xyz[2], mxx = xyz[0], mxy = xyz[1], mxz = xyz[2];
    for (uint64_t i = 0; i < n; ++i) {
      mnx = std::fmin(mnx, xyz[3 * i]);
      mny = std::fmin(mny, xyz[3 * i + 1]);
      mnz = std::fmin(mnz, xyz[3 * i + 2]);
      mxx = std::fmax(mxx, xyz[3 * i]);
      mxy = std::fmax(mxy, xyz[3 * i + 1]);
      mxz = std::fmax(mxz, xyz[3 * i + 2]);
    }
    const float ex = mxx - mnx, ey = mxy - mny, ez = mxz - mnz;
    const float extent = std::sqrt(ex * ex + ey * ey + ez * ez);  // powi(2) == x*x
    vp = V3{(mnx + mxx) / 2.0f, (mny + mxy) / 2.0f, (mnz + mxz) / 2.0f + extent};
  }
#pragma omp parallel num_threads(threads > 0 ? threads : orc_max_threads())
  {
    std::vector<uint64_t> nb;
    std::vector<Neighbor> tmp;
    std::vector<uint32_t> stack;
    std::vector<std::pair<uint64_t, float>> rad;
#pragma omp for schedule(dynamic, 256)
    for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
      const uint64_t i = (uint64_t)ii;
      const P3 p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
      if (radius >= 0.0f) {  // normals.rs:141-146
        kd_find_radius(*t, p, radius, rad, stack);
        nb.clear();
        for (auto& pr : rad)
          if (pr.first != i) nb.push_back(pr.first);
        if (nb.size() < k) knn_minus_self(*t, xyz, i, k, nb, tmp, stack);  // :315-323
      } else {
        knn_minus_self(*t, xyz, i, k, nb, tmp, stack);
      }
      if (nb.size() < 3) {  // normals.rs:326-336
        const size_t fk = std::max<size_t>(k, 5);
        knn_minus_self(*t, xyz, i, fk, nb, tmp, stack);
      }
      if (std::find(nb.begin(), nb.end(), i) == nb.end()) nb.push_back(i);  // :338-340
      V3 nrm = compute_normal_pca(xyz, nb);
      if (consistent_orientation) nrm = orient_normal(nrm, V3{p.x, p.y, p.z}, vp);
      out[6 * i + 0] = p.x;
      out[6 * i + 1] = p.y;
      out[6 * i + 2] = p.z;
      out[6 * i + 3] = nrm.x;
      out[6 * i + 4] = nrm.y;
      out[6 * i + 5] = nrm.z;
    }
  }
  delete t;
  return 0;
}

// Conditioning analysis for the parity harness: f64 covariance of the SAME neighbourhood the
// reference uses (kd kNN), f64 Jacobi eigen-decomposition.  Outputs per point: the f64 normal
// (unoriented, unit), and relative eigengap (l1 - l0) / l2 (ascending l0<=l1<=l2; 0 if l2==0).
void orc_normals_f64(const float* xyz, uint64_t n, uint64_t k, double* normal_out,
                     double* relgap_out, int threads) {
  if (n == 0) return;
  KdTree* t = kd_new(xyz, n);
#pragma omp parallel num_threads(threads > 0 ? threads : orc_max_threads())
  {
    std::vector<uint64_t> nb;
    std::vector<Neighbor> tmp;
    std::vector<uint32_t> stack;
#pragma omp for schedule(dynamic, 256)
    for (int64_t ii = 0; ii < (int64_t)n; ++ii) {
      const uint64_t i = (uint64_t)ii;
      knn_minus_self(*t, xyz, i, k, nb, tmp, stack);
      if (nb.size() < 3) knn_minus_self(*t, xyz, i, std::max<size_t>(k, 5), nb, tmp, stack);
      if (std::find(nb.begin(), nb.end(), i) == nb.end()) nb.push_back(i);
      double c[3] = {0, 0, 0};
      for (uint64_t j : nb)
        for (int a = 0; a < 3; ++a) c[a] += xyz[3 * j + a];
      for (int a = 0; a < 3; ++a) c[a] /= (double)nb.size();
      double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (uint64_t j : nb) {
        const double d[3] = {xyz[3 * j] - c[0], xyz[3 * j + 1] - c[1], xyz[3 * j + 2] - c[2]};
        for (int r = 0; r < 3; ++r)
          for (int cc = 0; cc < 3; ++cc) cov[r][cc] += d[r] * d[cc];
      }
      for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) cov[r][cc] /= (double)nb.size();
      double val[3], vec[3][3];
      jacobi3_f64(cov, val, vec);
      int o[3] = {0, 1, 2};
      std::sort(o, o + 3, [&](int a, int b) { return val[a] < val[b]; });
      if (nb.size() < 3) {
        normal_out[3 * i] = 0;
        normal_out[3 * i + 1] = 0;
        normal_out[3 * i + 2] = 1;
        relgap_out[i] = 0;
        continue;
      }
      for (int a = 0; a < 3; ++a) normal_out[3 * i + a] = vec[o[0]][a];
      relgap_out[i] = val[o[2]] > 0 ? (val[o[1]] - val[o[0]]) / val[o[2]] : 0.0;
    }
  }
  delete t;
}

// Expose the f32 symmetric eigen restatement and the f64 Jacobi for unit tests.
void orc_symmetric_eigen3(const float* m9_rowmajor, float* val3, float* vec9_cols) {
  float m[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[r][c] = m9_rowmajor[3 * r + c];
  const Eig3 e = symmetric_eigen3(m);
  for (int i = 0; i < 3; ++i) {
    val3[i] = e.val[i];
    for (int r = 0; r < 3; ++r) vec9_cols[3 * i + r] = e.vec[i][r];
  }
}

struct orc_icp_result {
  float t[3];      // translation
  float q[4];      // rotation quaternion, [i, j, k, w]
  float mse;
  uint64_t iterations;
  int32_t converged;
  uint64_t n_corr;  // number of correspondence pairs written
};

// Diagnostic only (not part of the restated algorithm): the mean of the same f32 squared
// residuals the reference sums, accumulated in f64 - tells how much of an mse difference is the
// reference's own sequential-f32 accumulation error (5e-4 relative at 1M pairs).
static double g_last_p2plane_mse_f64 = 0.0;
double orc_last_icp_mse_f64() { return g_last_p2plane_mse_f64; }

// icp_point_to_plane_detailed (registration.rs:508-602).
// init7 = [tx,ty,tz, qi,qj,qk,qw].  max_dist < 0 => None.  pairs_out: capacity ns x 2 (u64).
// Returns 0 OK, 1 InvalidData, 2 Algorithm.  OpenMP only over the correspondence search
// (registration.rs:92-93); transform, gather, accumulate, mse are serial as in the reference.
int orc_icp_point_to_plane(const float* src, uint64_t ns, const float* tgt, uint64_t nt,
                           const float* nrm, uint64_t nn, const float* init7, uint64_t max_iters,
                           float max_dist, float conv, orc_icp_result* res, uint64_t* pairs_out,
                           int threads) {
  if (ns == 0 || nt == 0) return 1;  // :517-521
  if (nn != nt) return 1;            // :522-526
  if (max_iters == 0) return 1;      // :527-531
  Iso T{Quat{init7[3], init7[4], init7[5], init7[6]}, V3{init7[0], init7[1], init7[2]}};
  float previous_mse = INFINITY;
  double previous_mse64 = 0.0;
  std::vector<std::pair<uint64_t, uint64_t>> final_corr;
  KdTree* tree = kd_new(tgt, nt);  // :536
  std::vector<V3> ts(ns);
  std::vector<int64_t> corr_idx(ns);
  std::vector<V3> vs, vt, vn;
  std::vector<std::pair<uint64_t, uint64_t>> corr_pairs;
  const int nth = threads > 0 ? threads : orc_max_threads();
  int status = 0;
  auto finish = [&](float mse, uint64_t iters, int converged,
                    const std::vector<std::pair<uint64_t, uint64_t>>& pairs) {
    res->t[0] = T.t.x;
    res->t[1] = T.t.y;
    res->t[2] = T.t.z;
    res->q[0] = T.q.i;
    res->q[1] = T.q.j;
    res->q[2] = T.q.k;
    res->q[3] = T.q.w;
    res->mse = mse;
    res->iterations = iters;
    res->converged = converged;
    res->n_corr = pairs.size();
    if (pairs_out)
      for (size_t i = 0; i < pairs.size(); ++i) {
        pairs_out[2 * i] = pairs[i].first;
        pairs_out[2 * i + 1] = pairs[i].second;
      }
  };
  for (uint64_t iteration = 0; iteration < max_iters; ++iteration) {
    for (uint64_t i = 0; i < ns; ++i)  // :540-544 serial map
      ts[i] = iso_apply(T, V3{src[3 * i], src[3 * i + 1], src[3 * i + 2]});
#pragma omp parallel num_threads(nth)
    {
      std::vector<Neighbor> out;
      std::vector<uint32_t> stack;
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < (int64_t)ns; ++i) {  // :87-107
        kd_find_k_nearest(*tree, P3{ts[i].x, ts[i].y, ts[i].z}, 1, out, stack);
        if (out.empty()) {
          corr_idx[i] = -1;
          continue;
        }
        const float distance = std::sqrt(out[0].distance);
        if (max_dist >= 0.0f && distance > max_dist)
          corr_idx[i] = -1;
        else
          corr_idx[i] = (int64_t)out[0].index;
      }
    }
    vs.clear();
    vt.clear();
    vn.clear();
    corr_pairs.clear();
    for (uint64_t i = 0; i < ns; ++i) {  // :553-565
      if (corr_idx[i] < 0) continue;
      const uint64_t j = (uint64_t)corr_idx[i];
      vs.push_back(ts[i]);
      vt.push_back(V3{tgt[3 * j], tgt[3 * j + 1], tgt[3 * j + 2]});
      vn.push_back(V3{nrm[3 * j], nrm[3 * j + 1], nrm[3 * j + 2]});
      corr_pairs.emplace_back(i, j);
    }
    if (vs.size() < 6) {  // :568-572
      status = 2;
      break;
    }
    // compute_transformation_point_to_plane (:395-450)
    float ata[6][6];
    float atb[6];
    std::memset(ata, 0, sizeof(ata));
    std::memset(atb, 0, sizeof(atb));
    for (size_t i = 0; i < vs.size(); ++i) {
      const V3 c = cross(vs[i], vn[i]);
      const float a[6] = {c.x, c.y, c.z, vn[i].x, vn[i].y, vn[i].z};
      const V3 d{vt[i].x - vs[i].x, vt[i].y - vs[i].y, vt[i].z - vs[i].z};
      const float b = dot(vn[i], d);
      for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc < 6; ++cc) ata[r][cc] += a[r] * a[cc];
      for (int r = 0; r < 6; ++r) atb[r] += a[r] * b;
    }
    float x[6];
    {
      float L[6][6];
      std::memcpy(L, ata, sizeof(L));
      std::memcpy(x, atb, sizeof(x));
      if (cholesky6(L)) {
        cholesky6_solve(L, x);
      } else {
        std::memcpy(L, ata, sizeof(L));
        std::memcpy(x, atb, sizeof(x));
        if (!lu6_solve(L, x)) {  // :435-437
          status = 2;
          break;
        }
      }
    }
    const Quat rx = quat_axis_angle(0, x[0]);
    const Quat ry = quat_axis_angle(1, x[1]);
    const Quat rz = quat_axis_angle(2, x[2]);
    const Iso delta{quat_mul(quat_mul(rz, ry), rx), V3{x[3], x[4], x[5]}};  // :441-449
    T = iso_mul(delta, T);                                                  // :576
    // compute_point_to_plane_mse (:453-471) — residuals before delta
    float sum = 0.0f;
    double sum64 = 0.0;
    for (size_t i = 0; i < vs.size(); ++i) {
      const V3 d{vt[i].x - vs[i].x, vt[i].y - vs[i].y, vt[i].z - vs[i].z};
      const float e = dot(vn[i], d);
      sum += e * e;
      sum64 += (double)(e * e);
    }
    const float current_mse = sum / (float)vs.size();
    const float mse_change = std::fabs(previous_mse - current_mse);
    if (mse_change < conv) {  // :581-589
      g_last_p2plane_mse_f64 = sum64 / (double)vs.size();
      finish(current_mse, iteration + 1, 1, corr_pairs);
      delete tree;
      return 0;
    }
    previous_mse = current_mse;
    previous_mse64 = sum64 / (double)vs.size();
    final_corr = corr_pairs;
  }
  delete tree;
  if (status != 0) return status;
  g_last_p2plane_mse_f64 = previous_mse64;
  finish(previous_mse, max_iters, 0, final_corr);  // :595-601
  return 0;
}

// ------------------------------------------------------------------------------------------
// Point-to-point ICP: icp_detailed (registration.rs:258-370), compute_transformation (:144-203),
// compute_mse (:206-218).
// Matrix3::svd and UnitQuaternion::from_matrix live in nalgebra [upstream, not under
// /root/reference]: the SVD is restated with a one-sided Jacobi SVD in f32 (any backward-stable
// SVD yields the same V U^T to ~1e-6), from_matrix with the standard matrix -> quaternion
// conversion (the input is already a rotation, so nalgebra's iterative extraction is a no-op).
// ------------------------------------------------------------------------------------------
}  // extern "C"
namespace {

// one-sided Jacobi SVD of a 3x3: A = U diag(s) V^T  (f32)
void svd3(const float a_in[3][3], float U[3][3], float S[3], float V[3][3]) {
  float a[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      a[r][c] = a_in[r][c];
      V[r][c] = r == c ? 1.0f : 0.0f;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    float off = 0.0f;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        float alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < 3; ++r) {
          alpha += a[r][p] * a[r][p];
          beta += a[r][q] * a[r][q];
          gamma += a[r][p] * a[r][q];
        }
        off = std::max(off, std::fabs(gamma) / std::sqrt(std::max(alpha * beta, 1e-30f)));
        if (gamma == 0.0f) continue;
        const float zeta = (beta - alpha) / (2.0f * gamma);
        const float t = std::copysign(1.0f, zeta) / (std::fabs(zeta) + std::sqrt(1.0f + zeta * zeta));
        const float c = 1.0f / std::sqrt(1.0f + t * t), sn = c * t;
        for (int r = 0; r < 3; ++r) {
          const float x = a[r][p], y = a[r][q];
          a[r][p] = c * x - sn * y;
          a[r][q] = sn * x + c * y;
          const float vx = V[r][p], vy = V[r][q];
          V[r][p] = c * vx - sn * vy;
          V[r][q] = sn * vx + c * vy;
        }
      }
    if (off < 1e-7f) break;
  }
  for (int c = 0; c < 3; ++c) {
    float n = 0;
    for (int r = 0; r < 3; ++r) n += a[r][c] * a[r][c];
    S[c] = std::sqrt(n);
  }
  // order descending so that a rank-deficient H puts its null direction last
  int o[3] = {0, 1, 2};
  std::sort(o, o + 3, [&](int x, int y) { return S[x] > S[y]; });
  float a2[3][3], V2[3][3], S2[3];
  for (int c = 0; c < 3; ++c) {
    S2[c] = S[o[c]];
    for (int r = 0; r < 3; ++r) {
      a2[r][c] = a[r][o[c]];
      V2[r][c] = V[r][o[c]];
    }
  }
  for (int c = 0; c < 3; ++c) {
    S[c] = S2[c];
    for (int r = 0; r < 3; ++r) {
      V[r][c] = V2[r][c];
      U[r][c] = S2[c] > 0 ? a2[r][c] / S2[c] : 0.0f;
    }
  }
  // complete U to an orthonormal basis when a singular value vanished
  if (!(S[2] > 0)) {
    const V3 u0{U[0][0], U[1][0], U[2][0]}, u1{U[0][1], U[1][1], U[2][1]};
    const V3 u2 = cross(u0, u1);
    U[0][2] = u2.x;
    U[1][2] = u2.y;
    U[2][2] = u2.z;
  }
}

float det3(const float m[3][3]) {
  return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
         m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

Quat quat_from_rotation(const float R[3][3]) {  // Shepperd
  Quat q;
  const float tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0.0f) {
    const float s = std::sqrt(tr + 1.0f) * 2.0f;
    q.w = 0.25f * s;
    q.i = (R[2][1] - R[1][2]) / s;
    q.j = (R[0][2] - R[2][0]) / s;
    q.k = (R[1][0] - R[0][1]) / s;
  } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
    const float s = std::sqrt(1.0f + R[0][0] - R[1][1] - R[2][2]) * 2.0f;
    q.w = (R[2][1] - R[1][2]) / s;
    q.i = 0.25f * s;
    q.j = (R[0][1] + R[1][0]) / s;
    q.k = (R[0][2] + R[2][0]) / s;
  } else if (R[1][1] > R[2][2]) {
    const float s = std::sqrt(1.0f + R[1][1] - R[0][0] - R[2][2]) * 2.0f;
    q.w = (R[0][2] - R[2][0]) / s;
    q.i = (R[0][1] + R[1][0]) / s;
    q.j = 0.25f * s;
    q.k = (R[1][2] + R[2][1]) / s;
  } else {
    const float s = std::sqrt(1.0f + R[2][2] - R[0][0] - R[1][1]) * 2.0f;
    q.w = (R[1][0] - R[0][1]) / s;
    q.i = (R[0][2] + R[2][0]) / s;
    q.j = (R[1][2] + R[2][1]) / s;
    q.k = 0.25f * s;
  }
  return q;
}

// compute_transformation (registration.rs:144-203)
Iso kabsch(const std::vector<V3>& src, const std::vector<V3>& tgt) {
  const float n = (float)src.size();
  V3 cs{0, 0, 0}, ct{0, 0, 0};
  for (const V3& p : src) cs = V3{cs.x + p.x, cs.y + p.y, cs.z + p.z};
  for (const V3& p : tgt) ct = V3{ct.x + p.x, ct.y + p.y, ct.z + p.z};
  cs = V3{cs.x / n, cs.y / n, cs.z / n};
  ct = V3{ct.x / n, ct.y / n, ct.z / n};
  float H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (size_t i = 0; i < src.size(); ++i) {
    const float p[3] = {src[i].x - cs.x, src[i].y - cs.y, src[i].z - cs.z};
    const float q[3] = {tgt[i].x - ct.x, tgt[i].y - ct.y, tgt[i].z - ct.z};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) H[r][c] += p[r] * q[c];
  }
  float U[3][3], S[3], V[3][3];
  svd3(H, U, S, V);
  auto vut = [&](float R[3][3]) {  // R = V U^T
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        R[r][c] = 0;
        for (int k = 0; k < 3; ++k) R[r][c] += V[r][k] * U[c][k];
      }
  };
  float R[3][3];
  vut(R);
  if (det3(R) < 0.0f) {  // flip the last row of V^T == last column of V (:187-191)
    for (int r = 0; r < 3; ++r) V[r][2] = -V[r][2];
    vut(R);
  }
  const Quat q = quat_from_rotation(R);
  const V3 rc = quat_rotate(q, cs);
  return Iso{q, V3{ct.x - rc.x, ct.y - rc.y, ct.z - rc.z}};
}

float mse_pairs(const std::vector<V3>& a, const std::vector<V3>& b) {  // :206-218
  if (a.empty()) return 0.0f;
  float sum = 0.0f;
  for (size_t i = 0; i < a.size(); ++i) {
    const float dx = a[i].x - b[i].x, dy = a[i].y - b[i].y, dz = a[i].z - b[i].z;
    sum += dx * dx + dy * dy + dz * dz;
  }
  return sum / (float)a.size();
}

}  // namespace
extern "C" {

// icp_detailed (registration.rs:258-370).  Returns 0 OK, 1 InvalidData, 2 Algorithm.
int orc_icp_point_to_point(const float* src, uint64_t ns, const float* tgt, uint64_t nt,
                           const float* init7, uint64_t max_iters, float max_dist, float conv,
                           orc_icp_result* res, uint64_t* pairs_out, int threads) {
  if (ns == 0 || nt == 0) return 1;  // :266-270
  if (max_iters == 0) return 1;      // :272-276
  Iso T{Quat{init7[3], init7[4], init7[5], init7[6]}, V3{init7[0], init7[1], init7[2]}};
  float previous_mse = INFINITY;
  std::vector<std::pair<uint64_t, uint64_t>> final_corr, corr_pairs;
  KdTree* tree = kd_new(tgt, nt);
  std::vector<V3> ts(ns), vs, vt;
  std::vector<int64_t> corr_idx(ns);
  const int nth = threads > 0 ? threads : orc_max_threads();
  auto finish = [&](float mse, uint64_t iters, int converged,
                    const std::vector<std::pair<uint64_t, uint64_t>>& pairs) {
    res->t[0] = T.t.x; res->t[1] = T.t.y; res->t[2] = T.t.z;
    res->q[0] = T.q.i; res->q[1] = T.q.j; res->q[2] = T.q.k; res->q[3] = T.q.w;
    res->mse = mse;
    res->iterations = iters;
    res->converged = converged;
    res->n_corr = pairs.size();
    if (pairs_out)
      for (size_t i = 0; i < pairs.size(); ++i) {
        pairs_out[2 * i] = pairs[i].first;
        pairs_out[2 * i + 1] = pairs[i].second;
      }
  };
  for (uint64_t iteration = 0; iteration < max_iters; ++iteration) {
    for (uint64_t i = 0; i < ns; ++i)
      ts[i] = iso_apply(T, V3{src[3 * i], src[3 * i + 1], src[3 * i + 2]});
#pragma omp parallel num_threads(nth)
    {
      std::vector<Neighbor> out;
      std::vector<uint32_t> stack;
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < (int64_t)ns; ++i) {
        kd_find_k_nearest(*tree, P3{ts[i].x, ts[i].y, ts[i].z}, 1, out, stack);
        if (out.empty()) { corr_idx[i] = -1; continue; }
        const float distance = std::sqrt(out[0].distance);
        corr_idx[i] = (max_dist >= 0.0f && distance > max_dist) ? -1 : (int64_t)out[0].index;
      }
    }
    vs.clear(); vt.clear(); corr_pairs.clear();
    for (uint64_t i = 0; i < ns; ++i) {
      if (corr_idx[i] < 0) continue;
      const uint64_t j = (uint64_t)corr_idx[i];
      vs.push_back(ts[i]);
      vt.push_back(V3{tgt[3 * j], tgt[3 * j + 1], tgt[3 * j + 2]});
      corr_pairs.emplace_back(i, j);
    }
    if (vs.size() < 3) { delete tree; return 2; }  // :311-315
    const Iso delta = kabsch(vs, vt);
    T = iso_mul(delta, T);
    const float current_mse = mse_pairs(vs, vt);
    if (std::fabs(previous_mse - current_mse) < conv) {  // :327-336
      finish(current_mse, iteration + 1, 1, corr_pairs);
      delete tree;
      return 0;
    }
    previous_mse = current_mse;
    final_corr = corr_pairs;
  }
  delete tree;
  // :342-361 — mse of the LAST correspondences under the FINAL transform
  float final_mse = previous_mse;
  if (!final_corr.empty()) {
    vs.clear(); vt.clear();
    for (auto& pr : final_corr) {
      vs.push_back(iso_apply(T, V3{src[3 * pr.first], src[3 * pr.first + 1], src[3 * pr.first + 2]}));
      vt.push_back(V3{tgt[3 * pr.second], tgt[3 * pr.second + 1], tgt[3 * pr.second + 2]});
    }
    final_mse = mse_pairs(vs, vt);
  }
  finish(final_mse, max_iters, 0, final_corr);
  return 0;
}

// Helpers exposed for unit tests of the nalgebra restatements.
void orc_iso_apply(const float* iso7, const float* p3, float* out3) {
  const Iso T{Quat{iso7[3], iso7[4], iso7[5], iso7[6]}, V3{iso7[0], iso7[1], iso7[2]}};
  const V3 r = iso_apply(T, V3{p3[0], p3[1], p3[2]});
  out3[0] = r.x;
  out3[1] = r.y;
  out3[2] = r.z;
}
void orc_iso_mul(const float* a7, const float* b7, float* out7) {
  const Iso A{Quat{a7[3], a7[4], a7[5], a7[6]}, V3{a7[0], a7[1], a7[2]}};
  const Iso B{Quat{b7[3], b7[4], b7[5], b7[6]}, V3{b7[0], b7[1], b7[2]}};
  const Iso C = iso_mul(A, B);
  out7[0] = C.t.x;
  out7[1] = C.t.y;
  out7[2] = C.t.z;
  out7[3] = C.q.i;
  out7[4] = C.q.j;
  out7[5] = C.q.k;
  out7[6] = C.q.w;
}
int orc_solve6(const float* ata36_rowmajor, const float* atb6, float* x6) {
  float L[6][6];
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) L[r][c] = ata36_rowmajor[6 * r + c];
  for (int i = 0; i < 6; ++i) x6[i] = atb6[i];
  if (cholesky6(L)) {
    cholesky6_solve(L, x6);
    return 0;
  }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) L[r][c] = ata36_rowmajor[6 * r + c];
  for (int i = 0; i < 6; ++i) x6[i] = atb6[i];
  return lu6_solve(L, x6) ? 1 : 2;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// GICP (threecrate-algorithms/src/gicp.rs).  nalgebra pieces restated [upstream, not under
// /root/reference]: Matrix3 products accumulate sum_k a(i,k) b(k,j) in k order; Matrix3::
// try_inverse is the adjugate / determinant formula with a zero-determinant test;
// UnitQuaternion::to_rotation_matrix is the usual ww+ii-jj-kk form.
// ------------------------------------------------------------------------------------------
namespace {
struct M3 {
  float m[3][3];
};
inline M3 m3_mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
  return r;
}
inline M3 m3_t(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
inline V3 m3_vec(const M3& a, const V3& v) {
  return V3{(a.m[0][0] * v.x + a.m[0][1] * v.y) + a.m[0][2] * v.z,
            (a.m[1][0] * v.x + a.m[1][1] * v.y) + a.m[1][2] * v.z,
            (a.m[2][0] * v.x + a.m[2][1] * v.y) + a.m[2][2] * v.z};
}
inline bool m3_try_inverse(const M3& a, M3& out) {
  const float m11 = a.m[0][0], m12 = a.m[0][1], m13 = a.m[0][2];
  const float m21 = a.m[1][0], m22 = a.m[1][1], m23 = a.m[1][2];
  const float m31 = a.m[2][0], m32 = a.m[2][1], m33 = a.m[2][2];
  const float minor_m12_m23 = m22 * m33 - m32 * m23;
  const float minor_m11_m23 = m21 * m33 - m31 * m23;
  const float minor_m11_m22 = m21 * m32 - m31 * m22;
  const float det = (m11 * minor_m12_m23 - m12 * minor_m11_m23) + m13 * minor_m11_m22;
  if (det == 0.0f) return false;
  out.m[0][0] = minor_m12_m23 / det;
  out.m[0][1] = (m13 * m32 - m33 * m12) / det;
  out.m[0][2] = (m12 * m23 - m22 * m13) / det;
  out.m[1][0] = -minor_m11_m23 / det;
  out.m[1][1] = (m11 * m33 - m31 * m13) / det;
  out.m[1][2] = (m13 * m21 - m23 * m11) / det;
  out.m[2][0] = minor_m11_m22 / det;
  out.m[2][1] = (m12 * m31 - m32 * m11) / det;
  out.m[2][2] = (m11 * m22 - m21 * m12) / det;
  return true;
}
inline M3 quat_to_matrix(const Quat& q) {
  const float i = q.i, j = q.j, k = q.k, w = q.w;
  const float ww = w * w, ii = i * i, jj = j * j, kk = k * k;
  const float ij = i * j * 2.0f, wk = w * k * 2.0f, wj = w * j * 2.0f, ik = i * k * 2.0f,
              jk = j * k * 2.0f, wi = w * i * 2.0f;
  M3 r;
  r.m[0][0] = ww + ii - jj - kk;
  r.m[0][1] = ij - wk;
  r.m[0][2] = wj + ik;
  r.m[1][0] = wk + ij;
  r.m[1][1] = ww - ii + jj - kk;
  r.m[1][2] = jk - wi;
  r.m[2][0] = ik - wj;
  r.m[2][1] = wi + jk;
  r.m[2][2] = ww - ii - jj + kk;
  return r;
}
// -skew_sym(v)  (gicp.rs:50-53, 222)
inline M3 neg_skew(const V3& v) {
  M3 r;
  r.m[0][0] = -0.0f;
  r.m[0][1] = v.z;
  r.m[0][2] = -v.y;
  r.m[1][0] = -v.z;
  r.m[1][1] = -0.0f;
  r.m[1][2] = v.x;
  r.m[2][0] = v.y;
  r.m[2][1] = -v.x;
  r.m[2][2] = -0.0f;
  return r;
}

// compute_covariances (gicp.rs:58-95): kNN(k) INCLUDING the point itself, mean by fold, outer
// products accumulated in neighbour order, / max(n-1, 1), + 1e-4 I
void gicp_covariances(const float* xyz, uint64_t n, uint64_t k, std::vector<M3>& covs, int nth) {
  k = std::max<uint64_t>(k, 4);
  KdTree* tree = kd_new(xyz, n);
  covs.resize(n);
#pragma omp parallel num_threads(nth)
  {
    std::vector<Neighbor> out;
    std::vector<uint32_t> stack;
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      kd_find_k_nearest(*tree, P3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, k, out, stack);
      M3 c;
      std::memset(&c, 0, sizeof(c));
      const size_t m = out.size();
      if (m < 3) {
        c.m[0][0] = c.m[1][1] = c.m[2][2] = 1.0f * 1e-3f;
        covs[i] = c;
        continue;
      }
      const float nf = (float)m;
      float mx = 0.0f, my = 0.0f, mz = 0.0f;
      for (const Neighbor& nb : out) {
        mx += xyz[3 * nb.index];
        my += xyz[3 * nb.index + 1];
        mz += xyz[3 * nb.index + 2];
      }
      mx /= nf;
      my /= nf;
      mz /= nf;
      for (const Neighbor& nb : out) {
        const float d[3] = {xyz[3 * nb.index] - mx, xyz[3 * nb.index + 1] - my,
                            xyz[3 * nb.index + 2] - mz};
        for (int r = 0; r < 3; ++r)
          for (int cc = 0; cc < 3; ++cc) c.m[r][cc] += d[r] * d[cc];
      }
      const float den = std::max(nf - 1.0f, 1.0f);
      for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) c.m[r][cc] /= den;
      for (int r = 0; r < 3; ++r) c.m[r][r] += 1.0f * 1e-4f;
      covs[i] = c;
    }
  }
  delete tree;
}
}  // namespace

extern "C" {
// gicp (gicp.rs:117-312).  Returns 0 OK, 1 InvalidData (detail in *why: 1 empty, 2 max_iters,
// 3 too few points, 4 coplanar source, 5 coplanar target), 2 insufficient correspondences,
// 3 ill-conditioned.
int orc_gicp(const float* src, uint64_t ns, const float* tgt, uint64_t nt, const float* init7,
             uint64_t max_iters, float max_dist, float conv, uint64_t k_corr, orc_icp_result* res,
             uint64_t* pairs_out, int* why, int threads) {
  *why = 0;
  if (ns == 0 || nt == 0) {
    *why = 1;
    return 1;
  }
  if (max_iters == 0) {
    *why = 2;
    return 1;
  }
  const uint64_t min_k = std::max<uint64_t>(k_corr, 4);
  if (ns < min_k || nt < min_k) {
    *why = 3;
    return 1;
  }
  for (int which = 0; which < 2; ++which) {  // gicp.rs:148-166
    const float* p = which == 0 ? src : tgt;
    const uint64_t n = which == 0 ? ns : nt;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint64_t i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::fmin(mn[a], p[3 * i + a]);
        mx[a] = std::fmax(mx[a], p[3 * i + a]);
      }
    float min_extent = INFINITY;
    for (int a = 0; a < 3; ++a) min_extent = std::fmin(min_extent, mx[a] - mn[a]);
    if (min_extent < 1e-4f) {
      *why = 4 + which;
      return 1;
    }
  }
  const int nth = threads > 0 ? threads : orc_max_threads();
  std::vector<M3> scov, tcov;
  gicp_covariances(src, ns, k_corr, scov, nth);
  gicp_covariances(tgt, nt, k_corr, tcov, nth);
  KdTree* tree = kd_new(tgt, nt);
  Iso T{Quat{init7[3], init7[4], init7[5], init7[6]}, V3{init7[0], init7[1], init7[2]}};
  float prev_mse = INFINITY;
  std::vector<std::pair<uint64_t, uint64_t>> final_corr, corr_pairs;
  std::vector<V3> ts(ns);
  std::vector<int64_t> corr_idx(ns);
  std::vector<float> corr_dist(ns);
  auto finish = [&](float mse, uint64_t iters, int converged,
                    const std::vector<std::pair<uint64_t, uint64_t>>& pairs) {
    res->t[0] = T.t.x;
    res->t[1] = T.t.y;
    res->t[2] = T.t.z;
    res->q[0] = T.q.i;
    res->q[1] = T.q.j;
    res->q[2] = T.q.k;
    res->q[3] = T.q.w;
    res->mse = mse;
    res->iterations = iters;
    res->converged = converged;
    res->n_corr = pairs.size();
    if (pairs_out)
      for (size_t i = 0; i < pairs.size(); ++i) {
        pairs_out[2 * i] = pairs[i].first;
        pairs_out[2 * i + 1] = pairs[i].second;
      }
  };
  int status = 0;
  for (uint64_t iteration = 0; iteration < max_iters; ++iteration) {
    for (uint64_t i = 0; i < ns; ++i)
      ts[i] = iso_apply(T, V3{src[3 * i], src[3 * i + 1], src[3 * i + 2]});
    const M3 R = quat_to_matrix(T.q);
    const M3 Rt = m3_t(R);
    // the reference's 1-NN loop is serial (gicp.rs:200-255); the queries are independent, so
    // they are run in parallel here and consumed in order
#pragma omp parallel num_threads(nth)
    {
      std::vector<Neighbor> out;
      std::vector<uint32_t> stack;
#pragma omp for schedule(dynamic, 256)
      for (int64_t i = 0; i < (int64_t)ns; ++i) {
        kd_find_k_nearest(*tree, P3{ts[i].x, ts[i].y, ts[i].z}, 1, out, stack);
        corr_idx[i] = out.empty() ? -1 : (int64_t)out[0].index;
        corr_dist[i] = out.empty() ? 0.0f : std::sqrt(out[0].distance);
      }
    }
    float h[6][6];
    float g[6];
    std::memset(h, 0, sizeof(h));
    std::memset(g, 0, sizeof(g));
    uint64_t n_corr = 0;
    float mse_sum = 0.0f;
    corr_pairs.clear();
    for (uint64_t i = 0; i < ns; ++i) {
      if (corr_idx[i] < 0) continue;
      const float dist = corr_dist[i];
      if (dist > max_dist) continue;
      const uint64_t j = (uint64_t)corr_idx[i];
      const M3 rc = m3_mul(m3_mul(R, scov[i]), Rt);
      M3 m;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) m.m[r][c] = tcov[j].m[r][c] + rc.m[r][c];
      M3 mi;
      if (!m3_try_inverse(m, mi)) continue;
      const V3 resid{tgt[3 * j] - ts[i].x, tgt[3 * j + 1] - ts[i].y, tgt[3 * j + 2] - ts[i].z};
      const M3 a = neg_skew(ts[i]);
      const M3 at = m3_t(a);
      const M3 h_rr = m3_mul(at, m3_mul(mi, a));
      const M3 h_rt = m3_mul(at, mi);
      const V3 wr = m3_vec(mi, resid);
      const V3 g_r = m3_vec(at, wr);
      const float grv[3] = {g_r.x, g_r.y, g_r.z}, wrv[3] = {wr.x, wr.y, wr.z};
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) {
          h[r][c] += h_rr.m[r][c];
          h[r][c + 3] += h_rt.m[r][c];
          h[r + 3][c] += h_rt.m[c][r];
          h[r + 3][c + 3] += mi.m[r][c];
        }
        g[r] += grv[r];
        g[r + 3] += wrv[r];
      }
      n_corr += 1;
      mse_sum += dist * dist;
      corr_pairs.emplace_back(i, j);
    }
    if (n_corr < 6) {
      status = 2;
      break;
    }
    const float mse = mse_sum / (float)n_corr;
    float x[6];
    {
      float L[6][6];
      std::memcpy(L, h, sizeof(L));
      std::memcpy(x, g, sizeof(x));
      if (cholesky6(L)) {
        cholesky6_solve(L, x);
      } else {
        std::memcpy(L, h, sizeof(L));
        std::memcpy(x, g, sizeof(x));
        if (!lu6_solve(L, x)) {
          status = 3;
          break;
        }
      }
    }
    const Quat rx = quat_axis_angle(0, x[0]);
    const Quat ry = quat_axis_angle(1, x[1]);
    const Quat rz = quat_axis_angle(2, x[2]);
    const Iso delta{quat_mul(quat_mul(rz, ry), rx), V3{x[3], x[4], x[5]}};
    T = iso_mul(delta, T);
    if (std::fabs(prev_mse - mse) < conv) {
      finish(mse, iteration + 1, 1, corr_pairs);
      delete tree;
      return 0;
    }
    prev_mse = mse;
    final_corr = corr_pairs;
  }
  delete tree;
  if (status != 0) return status;
  finish(prev_mse, max_iters, 0, final_corr);
  return 0;
}

// per-point GICP covariances (row-major 3x3 per point), for kernel-level parity checks
void orc_gicp_covariances(const float* xyz, uint64_t n, uint64_t k, float* cov9_out, int threads) {
  std::vector<M3> covs;
  gicp_covariances(xyz, n, k, covs, threads > 0 ? threads : orc_max_threads());
  for (uint64_t i = 0; i < n; ++i) std::memcpy(cov9_out + 9 * i, covs[i].m, 9 * sizeof(float));
}
}  // extern "C"
