// threecrate_cuda.hpp — C++17 host-side mirror of the reference's operator interface for the
// kNN -> normals -> ICP path, header-only, over the C ABI in threecrate_cuda.h.
//
// The reference is a Rust workspace and there is no Rust toolchain in the build image, so this is
// the compiled-language host side that IS built and run here (tests/cpp/host_mirror_test.cpp);
// rust/threecrate-cuda/ holds the same wrappers as (uncompiled) Rust.  Names, argument meaning,
// validation order and error kinds follow the reference functions cited on each wrapper
// (paths relative to the reference checkout, threecrate-algorithms/src/...).
//
// There is no CPU fallback: every call needs a CUDA device and throws Error{Gpu} without one.
#pragma once

#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "threecrate_cuda.h"

namespace threecrate {

// threecrate-core/src/point.rs:8, :31-36 — same repr(C) layouts
struct Point3f {
  float x = 0, y = 0, z = 0;
};
struct NormalPoint3f {
  Point3f position;
  Point3f normal;
};
static_assert(sizeof(Point3f) == 12 && sizeof(NormalPoint3f) == 24, "AoS layouts cross the ABI");

// Isometry3<f32> as it crosses the ABI: translation + unit quaternion [i, j, k, w]
struct Isometry3f {
  std::array<float, 3> translation{0, 0, 0};
  std::array<float, 4> rotation{0, 0, 0, 1};
  static Isometry3f identity() { return {}; }
  std::array<float, 7> packed() const {
    return {translation[0], translation[1], translation[2], rotation[0], rotation[1], rotation[2],
            rotation[3]};
  }
  Point3f apply(const Point3f& p) const {  // rotation * p + translation
    const float qi = rotation[0], qj = rotation[1], qk = rotation[2], qw = rotation[3];
    const float tx = 2 * (qj * p.z - qk * p.y), ty = 2 * (qk * p.x - qi * p.z),
                tz = 2 * (qi * p.y - qj * p.x);
    return {p.x + qw * tx + (qj * tz - qk * ty) + translation[0],
            p.y + qw * ty + (qk * tx - qi * tz) + translation[1],
            p.z + qw * tz + (qi * ty - qj * tx) + translation[2]};
  }
};

// threecrate-core/src/error.rs:7-28
enum class ErrorKind { InvalidData = TC_INVALID_DATA, Algorithm = TC_ALGORITHM, Gpu = TC_GPU };
class Error : public std::runtime_error {
 public:
  Error(ErrorKind k, const std::string& msg) : std::runtime_error(msg), kind(k) {}
  ErrorKind kind;
};

// One context per thread (the C ABI serialises the calls of a context on its stream).
class Context {
 public:
  explicit Context(int device = 0) {
    if (tc_context_create(device, &ctx_) != TC_OK)
      throw Error(ErrorKind::Gpu, "no CUDA device (this library has no CPU fallback)");
  }
  ~Context() { tc_context_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  tc_context* get() const { return ctx_; }
  void check(int status) const {
    if (status == TC_OK) return;
    const char* m = tc_last_error(ctx_);
    throw Error(static_cast<ErrorKind>(status), m ? m : "");
  }
  static Context& current() {
    thread_local Context c;
    return c;
  }

 private:
  tc_context* ctx_ = nullptr;
};

namespace detail {
inline const float* f(const std::vector<Point3f>& v) {
  return reinterpret_cast<const float*>(v.data());
}
struct Cloud {  // RAII tc_cloud
  tc_cloud* h = nullptr;
  Cloud() = default;
  Cloud(Context& c, const std::vector<Point3f>& pts) {
    c.check(tc_cloud_upload(c.get(), f(pts), pts.size(), &h));
  }
  ~Cloud() { tc_cloud_free(h); }
  Cloud(const Cloud&) = delete;
  Cloud& operator=(const Cloud&) = delete;
  std::vector<Point3f> download(Context& c) const {
    std::vector<Point3f> out(tc_cloud_len(h));
    c.check(tc_cloud_download(c.get(), h, reinterpret_cast<float*>(out.data())));
    return out;
  }
};
}  // namespace detail

// ---------------------------------------------------------------------------------------------
// KdTree (nearest_neighbor.rs:29-299) — here a device-resident uniform-grid index
// ---------------------------------------------------------------------------------------------
class KdTree {
 public:
  explicit KdTree(const std::vector<Point3f>& points, unsigned k_hint = 8,
                  Context& ctx = Context::current())
      : ctx_(ctx), cloud_(ctx, points), n_(points.size()) {
    ctx_.check(tc_index_build(ctx_.get(), cloud_.h, k_hint, 0.0f, &index_));
  }
  ~KdTree() { tc_index_free(index_); }
  KdTree(const KdTree&) = delete;
  KdTree& operator=(const KdTree&) = delete;
  const tc_index* index() const { return index_; }

  // find_k_nearest (nearest_neighbor.rs:177-251): (index, distance) ascending; min(k, n) entries
  std::vector<std::pair<size_t, float>> find_k_nearest(const Point3f& q, size_t k) const {
    std::vector<std::pair<size_t, float>> out;
    if (k == 0 || n_ == 0) return out;
    std::vector<uint32_t> idx(k);
    std::vector<float> dist(k);
    uint32_t cnt = 0;
    ctx_.check(tc_knn(ctx_.get(), index_, &q.x, 1, (uint32_t)k, 0, idx.data(), dist.data(), &cnt));
    for (uint32_t i = 0; i < cnt; ++i) out.emplace_back(idx[i], dist[i]);
    return out;
  }
  // find_radius_neighbors (nearest_neighbor.rs:254-298)
  std::vector<std::pair<size_t, float>> find_radius_neighbors(const Point3f& q, float radius) const {
    std::vector<uint32_t> idx(n_ ? n_ : 1);
    std::vector<float> dist(n_ ? n_ : 1);
    uint64_t found = 0;
    const float qq[3] = {q.x, q.y, q.z};
    ctx_.check(tc_radius_search(ctx_.get(), index_, qq, radius, idx.data(), dist.data(), n_, &found));
    std::vector<std::pair<size_t, float>> out;
    for (uint64_t i = 0; i < found && i < n_; ++i) out.emplace_back(idx[i], dist[i]);
    return out;
  }
  // PointCloudNeighbors::k_nearest_neighbors (point_cloud_ops.rs:80-105): self excluded by index
  std::vector<std::vector<std::pair<size_t, float>>> k_nearest_neighbors(size_t k) const {
    std::vector<std::vector<std::pair<size_t, float>>> out;
    if (k == 0 || n_ == 0) return out;
    std::vector<uint32_t> idx(n_ * k), cnt(n_);
    std::vector<float> dist(n_ * k);
    ctx_.check(tc_knn(ctx_.get(), index_, nullptr, n_, (uint32_t)k, 1, idx.data(), dist.data(),
                      cnt.data()));
    out.resize(n_);
    for (size_t i = 0; i < n_; ++i)
      for (uint32_t j = 0; j < cnt[i]; ++j) out[i].emplace_back(idx[i * k + j], dist[i * k + j]);
    return out;
  }

 private:
  Context& ctx_;
  detail::Cloud cloud_;
  tc_index* index_ = nullptr;
  size_t n_;
};

// ---------------------------------------------------------------------------------------------
// normals (normals.rs:117-380)
// ---------------------------------------------------------------------------------------------
struct NormalEstimationConfig {  // normals.rs:117-133
  size_t k_neighbors = 10;
  std::optional<float> radius;
  bool consistent_orientation = true;
  std::optional<std::array<float, 3>> viewpoint;
};
inline std::vector<NormalPoint3f> estimate_normals_with_config(
    const std::vector<Point3f>& cloud, const NormalEstimationConfig& cfg,
    Context& ctx = Context::current()) {
  std::vector<NormalPoint3f> out(cloud.size());
  ctx.check(tc_estimate_normals(ctx.get(), detail::f(cloud), cloud.size(), (uint32_t)cfg.k_neighbors,
                                cfg.radius ? *cfg.radius : -1.0f, cfg.consistent_orientation ? 1 : 0,
                                cfg.viewpoint ? cfg.viewpoint->data() : nullptr,
                                reinterpret_cast<float*>(out.data())));
  return out;
}
inline std::vector<NormalPoint3f> estimate_normals(const std::vector<Point3f>& cloud, size_t k) {
  NormalEstimationConfig cfg;  // normals.rs:238-246
  cfg.k_neighbors = k;
  return estimate_normals_with_config(cloud, cfg);
}
inline std::vector<NormalPoint3f> estimate_normals_radius(const std::vector<Point3f>& cloud,
                                                          float radius, bool consistent) {
  NormalEstimationConfig cfg;  // normals.rs:368-380
  cfg.radius = radius;
  cfg.consistent_orientation = consistent;
  return estimate_normals_with_config(cloud, cfg);
}

// ---------------------------------------------------------------------------------------------
// registration (registration.rs, gicp.rs)
// ---------------------------------------------------------------------------------------------
struct ICPResult {  // registration.rs:13-24
  Isometry3f transformation;
  float mse = 0;
  size_t iterations = 0;
  bool converged = false;
  std::vector<std::pair<size_t, size_t>> correspondences;
};
namespace detail {
inline ICPResult unpack(const tc_icp_result& r, const std::vector<uint64_t>& pairs) {
  ICPResult o;
  for (int i = 0; i < 3; ++i) o.transformation.translation[i] = r.transform[i];
  for (int i = 0; i < 4; ++i) o.transformation.rotation[i] = r.transform[3 + i];
  o.mse = r.mse;
  o.iterations = r.iterations;
  o.converged = r.converged != 0;
  for (uint64_t i = 0; i < r.n_correspondences; ++i)
    o.correspondences.emplace_back(pairs[2 * i], pairs[2 * i + 1]);
  return o;
}
}  // namespace detail

// icp_point_to_plane_detailed (registration.rs:508-602)
inline ICPResult icp_point_to_plane_detailed(const std::vector<Point3f>& source,
                                             const std::vector<Point3f>& target,
                                             const std::vector<Point3f>& target_normals,
                                             const Isometry3f& init, size_t max_iterations,
                                             std::optional<float> max_correspondence_distance,
                                             float convergence_threshold,
                                             Context& ctx = Context::current()) {
  tc_icp_result r{};
  std::vector<uint64_t> pairs(2 * (source.size() ? source.size() : 1));
  const auto i7 = init.packed();
  ctx.check(tc_icp_point_to_plane(ctx.get(), detail::f(source), source.size(), detail::f(target),
                                  target.size(), detail::f(target_normals), target_normals.size(),
                                  i7.data(), (uint32_t)max_iterations,
                                  max_correspondence_distance ? *max_correspondence_distance : -1.0f,
                                  convergence_threshold, &r, pairs.data()));
  return detail::unpack(r, pairs);
}
inline ICPResult icp_point_to_plane(const std::vector<Point3f>& source,
                                    const std::vector<Point3f>& target,
                                    const std::vector<Point3f>& target_normals,
                                    const Isometry3f& init, size_t max_iterations) {
  return icp_point_to_plane_detailed(source, target, target_normals, init, max_iterations,
                                     std::nullopt, 1e-6f);  // registration.rs:488-496
}
// icp_detailed (registration.rs:258-370)
inline ICPResult icp_detailed(const std::vector<Point3f>& source, const std::vector<Point3f>& target,
                              const Isometry3f& init, size_t max_iterations,
                              std::optional<float> max_correspondence_distance,
                              float convergence_threshold, Context& ctx = Context::current()) {
  tc_icp_result r{};
  std::vector<uint64_t> pairs(2 * (source.size() ? source.size() : 1));
  const auto i7 = init.packed();
  ctx.check(tc_icp_point_to_point(ctx.get(), detail::f(source), source.size(), detail::f(target),
                                  target.size(), i7.data(), (uint32_t)max_iterations,
                                  max_correspondence_distance ? *max_correspondence_distance : -1.0f,
                                  convergence_threshold, &r, pairs.data()));
  return detail::unpack(r, pairs);
}
// icp_point_to_point (registration.rs:644-680)
inline ICPResult icp_point_to_point(const std::vector<Point3f>& source,
                                    const std::vector<Point3f>& target, const Isometry3f& init,
                                    size_t max_iterations, float convergence_threshold,
                                    std::optional<float> max_correspondence_distance) {
  if (source.empty() || target.empty())
    throw Error(ErrorKind::InvalidData, "Source or target point cloud is empty");
  if (max_iterations == 0) throw Error(ErrorKind::InvalidData, "Max iterations must be positive");
  if (convergence_threshold <= 0.0f)
    throw Error(ErrorKind::InvalidData, "Convergence threshold must be positive");
  return icp_detailed(source, target, init, max_iterations, max_correspondence_distance,
                      convergence_threshold);
}
// icp (registration.rs:232-242): the transformation, or `init` on any error
inline Isometry3f icp(const std::vector<Point3f>& source, const std::vector<Point3f>& target,
                      const Isometry3f& init, size_t max_iterations) {
  try {
    return icp_detailed(source, target, init, max_iterations, std::nullopt, 1e-6f).transformation;
  } catch (const Error&) {
    return init;
  }
}

struct IcpScaleLevel {  // registration.rs:28-35
  float voxel_size;
  size_t max_iterations;
  std::optional<float> max_correspondence_distance;
};
struct MultiScaleIcpConfig {  // registration.rs:39-71
  std::vector<IcpScaleLevel> levels{{0.20f, 10, 0.50f}, {0.10f, 10, 0.25f}, {0.05f, 15, 0.15f}};
  size_t final_refinement_iterations = 10;
  std::optional<float> final_max_correspondence_distance = 0.10f;
  float convergence_threshold = 1e-5f;
};
// multiscale_icp_point_to_point (registration.rs:704-789)
inline ICPResult multiscale_icp_point_to_point(const std::vector<Point3f>& source,
                                               const std::vector<Point3f>& target,
                                               const Isometry3f& init,
                                               const MultiScaleIcpConfig& cfg,
                                               Context& ctx = Context::current()) {
  std::vector<tc_icp_scale_level> lv;
  for (const auto& l : cfg.levels)
    lv.push_back({l.voxel_size, (uint32_t)l.max_iterations,
                  l.max_correspondence_distance ? *l.max_correspondence_distance : -1.0f});
  tc_icp_result r{};
  std::vector<uint64_t> pairs(2 * (source.size() ? source.size() : 1));
  const auto i7 = init.packed();
  const tc_icp_scale_level none{};
  ctx.check(tc_multiscale_icp_point_to_point(
      ctx.get(), detail::f(source), source.size(), detail::f(target), target.size(), i7.data(),
      lv.empty() ? &none : lv.data(), (uint32_t)lv.size(),
      (uint32_t)cfg.final_refinement_iterations,
      cfg.final_max_correspondence_distance ? *cfg.final_max_correspondence_distance : -1.0f,
      cfg.convergence_threshold, &r, pairs.data()));
  return detail::unpack(r, pairs);
}

struct GicpConfig {  // gicp.rs:24-46
  size_t max_iterations = 50;
  float max_correspondence_distance = 1.0f;
  float convergence_threshold = 1e-6f;
  size_t k_correspondences = 20;
};
// gicp (gicp.rs:117-312)
inline ICPResult gicp(const std::vector<Point3f>& source, const std::vector<Point3f>& target,
                      const Isometry3f& init, const GicpConfig& cfg = {},
                      Context& ctx = Context::current()) {
  tc_icp_result r{};
  std::vector<uint64_t> pairs(2 * (source.size() ? source.size() : 1));
  const auto i7 = init.packed();
  ctx.check(tc_gicp(ctx.get(), detail::f(source), source.size(), detail::f(target), target.size(),
                    i7.data(), (uint32_t)cfg.max_iterations, cfg.max_correspondence_distance,
                    cfg.convergence_threshold, (uint32_t)cfg.k_correspondences, &r, pairs.data()));
  return detail::unpack(r, pairs);
}

// ---------------------------------------------------------------------------------------------
// filters (filtering.rs)
// ---------------------------------------------------------------------------------------------
// voxel_grid_filter (filtering.rs:38-133); output in ascending (z, y, x) voxel order
inline std::vector<Point3f> voxel_grid_filter(const std::vector<Point3f>& cloud, float voxel_size,
                                              Context& ctx = Context::current()) {
  detail::Cloud in(ctx, cloud), out;
  ctx.check(tc_voxel_grid_filter(ctx.get(), in.h, voxel_size, &out.h));
  return out.download(ctx);
}
// radius_outlier_removal (filtering.rs:167-218)
inline std::vector<Point3f> radius_outlier_removal(const std::vector<Point3f>& cloud, float radius,
                                                   size_t min_neighbors,
                                                   Context& ctx = Context::current()) {
  detail::Cloud in(ctx, cloud), out;
  ctx.check(tc_radius_outlier_removal(ctx.get(), in.h, radius, (uint32_t)min_neighbors, &out.h));
  return out.download(ctx);
}
// statistical_outlier_removal (filtering.rs:253-321): the reference's sequential f32 statistics
inline std::vector<Point3f> statistical_outlier_removal(const std::vector<Point3f>& cloud,
                                                        size_t k_neighbors, float std_dev_multiplier,
                                                        Context& ctx = Context::current()) {
  detail::Cloud in(ctx, cloud), out;
  ctx.check(tc_statistical_outlier_removal(ctx.get(), in.h, (uint32_t)k_neighbors,
                                           std_dev_multiplier, 0, nullptr, &out.h));
  return out.download(ctx);
}
// statistical_outlier_removal_with_threshold (filtering.rs:335-394)
inline std::vector<Point3f> statistical_outlier_removal_with_threshold(
    const std::vector<Point3f>& cloud, size_t k_neighbors, float threshold,
    Context& ctx = Context::current()) {
  detail::Cloud in(ctx, cloud), out;
  ctx.check(tc_statistical_outlier_removal(ctx.get(), in.h, (uint32_t)k_neighbors, threshold, 2,
                                           nullptr, &out.h));
  return out.download(ctx);
}

}  // namespace threecrate
