// host_mirror_test.cpp — exercises include/threecrate_cuda.hpp (the C++ host-side mirror of the
// reference interface) with the scenarios of the reference's own inline tests:
//   nearest_neighbor.rs:408-727, normals.rs:398-624, registration.rs:791-1267, gicp.rs:314-583,
//   filtering.rs:396-660.
// Built by threecrate_b200/build.py (g++ -std=c++17, links libthreecrate_cuda.so); run on a GPU
// box by tests/test_gpu_cpp_mirror.py.  Exit code 0 = every check passed.
#include <cmath>
#include <cstdio>
#include <functional>

#include "threecrate_cuda.hpp"

using namespace threecrate;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond)                                                           \
  do {                                                                        \
    ++g_checks;                                                               \
    if (!(cond)) {                                                            \
      ++g_failed;                                                             \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
    }                                                                         \
  } while (0)

static bool throws(ErrorKind kind, const std::function<void()>& f) {
  try {
    f();
  } catch (const Error& e) {
    return e.kind == kind;
  }
  return false;
}

static std::vector<Point3f> cube8() {  // nearest_neighbor.rs:395-406
  return {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
}
static std::vector<Point3f> sphere(size_t n, float radius) {  // gicp.rs:318-333
  std::vector<Point3f> c;
  const float golden = 3.14159265358979f * (3.0f - std::sqrt(5.0f));
  for (size_t i = 0; i < n; ++i) {
    const float y = 1.0f - ((float)i / std::fmax((float)n - 1.0f, 1.0f)) * 2.0f;
    const float r = std::sqrt(std::fmax(1.0f - y * y, 0.0f));
    const float th = golden * (float)i;
    c.push_back({std::cos(th) * r * radius, y * radius, std::sin(th) * r * radius});
  }
  return c;
}
// smooth terrain with analytic normals: z = 0.3 sin x cos y on a 60 x 60 grid over [0, 6]^2
static void terrain(std::vector<Point3f>& pts, std::vector<Point3f>& nrm) {
  for (int i = 0; i < 60; ++i)
    for (int j = 0; j < 60; ++j) {
      const float x = 0.1f * i + 0.013f * ((i * 7 + j * 3) % 5), y = 0.1f * j + 0.011f * ((i + j * 5) % 7);
      pts.push_back({x, y, 0.3f * std::sin(x) * std::cos(y)});
      const float nx = -0.3f * std::cos(x) * std::cos(y), ny = 0.3f * std::sin(x) * std::sin(y);
      const float m = std::sqrt(nx * nx + ny * ny + 1.0f);
      nrm.push_back({nx / m, ny / m, 1.0f / m});
    }
}

int main() {
  try {
    Context::current();
  } catch (const Error& e) {
    std::fprintf(stderr, "host_mirror_test: %s\n", e.what());
    return 77;  // no device
  }

  {  // ---- KdTree (nearest_neighbor.rs:417-520)
    const auto pts = cube8();
    KdTree tree(pts);
    const auto nn = tree.find_k_nearest({0.5f, 0.5f, 0.5f}, 3);
    CHECK(nn.size() == 3);
    for (size_t i = 0; i < nn.size(); ++i) CHECK(std::fabs(nn[i].second - std::sqrt(0.75f)) < 1e-6f);
    for (size_t i = 1; i < nn.size(); ++i) CHECK(nn[i - 1].second <= nn[i].second);
    const auto rad = tree.find_radius_neighbors({0.5f, 0.5f, 0.5f}, 1.5f);
    CHECK(rad.size() == 8);
    for (const auto& r : rad) CHECK(r.second <= 1.5f);
    CHECK(tree.find_k_nearest({0, 0, 0}, 100).size() == 8);  // k > n -> n results
    CHECK(tree.find_k_nearest({0, 0, 0}, 1)[0].first == 0);
    const auto all = tree.k_nearest_neighbors(3);
    CHECK(all.size() == 8);
    for (size_t i = 0; i < all.size(); ++i) {
      CHECK(all[i].size() == 3);
      for (const auto& nb : all[i]) CHECK(nb.first != i && std::fabs(nb.second - 1.0f) < 1e-6f);
    }
    KdTree empty(std::vector<Point3f>{});
    CHECK(empty.find_k_nearest({0, 0, 0}, 5).empty());
  }

  {  // ---- normals (normals.rs:398-470): plane z = 0 -> +-z; empty -> empty; k < 3 -> error
    std::vector<Point3f> plane;
    for (int i = 0; i < 10; ++i)
      for (int j = 0; j < 10; ++j) plane.push_back({0.1f * i, 0.1f * j, 0.0f});
    const auto n = estimate_normals(plane, 5);
    CHECK(n.size() == plane.size());
    for (const auto& p : n) {
      CHECK(std::fabs(p.normal.z) > 0.9f);
      CHECK(std::fabs(std::sqrt(p.normal.x * p.normal.x + p.normal.y * p.normal.y +
                                p.normal.z * p.normal.z) - 1.0f) < 1e-5f);
    }
    CHECK(n[37].position.x == plane[37].x && n[37].position.y == plane[37].y);
    CHECK(estimate_normals({}, 5).empty());
    CHECK(throws(ErrorKind::InvalidData, [&] { estimate_normals(plane, 2); }));
    const auto nr = estimate_normals_radius(plane, 0.25f, true);
    for (const auto& p : nr) CHECK(std::fabs(p.normal.z) > 0.9f);
    NormalEstimationConfig cfg;
    cfg.k_neighbors = 8;
    cfg.viewpoint = std::array<float, 3>{0.5f, 0.5f, -5.0f};  // below the plane -> normals point down
    for (const auto& p : estimate_normals_with_config(plane, cfg)) CHECK(p.normal.z < -0.9f);
  }

  std::vector<Point3f> tgt, nrm;
  terrain(tgt, nrm);
  Isometry3f truth;
  truth.translation = {0.03f, -0.02f, 0.01f};
  truth.rotation = {0.0f, 0.0f, std::sin(0.004f), std::cos(0.004f)};  // yaw 0.008 rad
  // source = truth^-1 applied to the target, so that truth * source == target
  std::vector<Point3f> src;
  {
    Isometry3f inv;
    inv.rotation = {0.0f, 0.0f, -truth.rotation[2], truth.rotation[3]};
    for (const auto& p : tgt) {
      const Point3f d{p.x - truth.translation[0], p.y - truth.translation[1], p.z - truth.translation[2]};
      src.push_back(inv.apply(d));
    }
  }
  auto t_err = [&](const Isometry3f& T) {
    double e = 0;
    for (int i = 0; i < 3; ++i) e += std::pow((double)T.translation[i] - truth.translation[i], 2);
    return std::sqrt(e);
  };

  {  // ---- point-to-plane ICP (registration.rs:1144-1267)
    const auto r = icp_point_to_plane(src, tgt, nrm, Isometry3f::identity(), 30);
    CHECK(r.iterations > 0 && r.iterations <= 30);
    CHECK(t_err(r.transformation) < 5e-3);
    CHECK(r.mse < 1e-4f);
    CHECK(!r.correspondences.empty() && r.correspondences.size() <= src.size());
    CHECK(throws(ErrorKind::InvalidData, [&] { icp_point_to_plane({}, tgt, nrm, Isometry3f::identity(), 10); }));
    CHECK(throws(ErrorKind::InvalidData, [&] { icp_point_to_plane(src, tgt, {}, Isometry3f::identity(), 10); }));
    CHECK(throws(ErrorKind::InvalidData, [&] { icp_point_to_plane(src, tgt, nrm, Isometry3f::identity(), 0); }));
    std::vector<Point3f> far = src;
    for (auto& p : far) p.x += 1000.0f;
    CHECK(throws(ErrorKind::Algorithm, [&] {
      icp_point_to_plane_detailed(far, tgt, nrm, Isometry3f::identity(), 10, 0.5f, 1e-6f);
    }));
  }

  {  // ---- point-to-point ICP (registration.rs:791-1140)
    const auto same = icp_point_to_point(tgt, tgt, Isometry3f::identity(), 10, 1e-6f, std::nullopt);
    CHECK(same.converged && same.mse < 1e-6f && t_err(same.transformation) < 0.04);
    const auto r = icp_point_to_point(src, tgt, Isometry3f::identity(), 50, 1e-9f, std::nullopt);
    CHECK(t_err(r.transformation) < 2e-2);
    CHECK(throws(ErrorKind::InvalidData, [&] { icp_point_to_point({}, tgt, Isometry3f::identity(), 10, 1e-6f, std::nullopt); }));
    CHECK(throws(ErrorKind::InvalidData, [&] { icp_point_to_point(src, tgt, Isometry3f::identity(), 0, 1e-6f, std::nullopt); }));
    CHECK(throws(ErrorKind::InvalidData, [&] { icp_point_to_point(src, tgt, Isometry3f::identity(), 10, 0.0f, std::nullopt); }));
    Isometry3f guess;
    guess.translation = {0.5f, 0, 0};
    const Isometry3f back = icp({}, tgt, guess, 10);  // any error -> init (registration.rs:232-242)
    CHECK(back.translation[0] == 0.5f);
  }

  {  // ---- multiscale ICP (registration.rs:704-789)
    MultiScaleIcpConfig cfg;
    cfg.levels = {{0.4f, 10, 1.0f}, {0.2f, 10, 0.5f}};
    cfg.final_max_correspondence_distance = 0.3f;
    const auto r = multiscale_icp_point_to_point(src, tgt, Isometry3f::identity(), cfg);
    CHECK(r.iterations > 0 && t_err(r.transformation) < 3e-2);
    MultiScaleIcpConfig none;
    none.levels.clear();
    CHECK(throws(ErrorKind::InvalidData, [&] { multiscale_icp_point_to_point(src, tgt, Isometry3f::identity(), none); }));
    MultiScaleIcpConfig huge;
    huge.levels = {{100.0f, 5, std::nullopt}};
    CHECK(throws(ErrorKind::Algorithm, [&] { multiscale_icp_point_to_point(src, tgt, Isometry3f::identity(), huge); }));
  }

  {  // ---- GICP (gicp.rs:335-380, 556-582)
    const auto c = sphere(100, 3.0f);
    GicpConfig cfg;
    cfg.max_iterations = 30;
    const auto r = gicp(c, c, Isometry3f::identity(), cfg);
    CHECK(r.converged && r.mse < 1e-4f);
    auto shifted = sphere(150, 3.0f);
    const auto base = shifted;
    for (auto& p : shifted) p.x += 0.1f;
    GicpConfig c2;
    c2.max_iterations = 60;
    c2.max_correspondence_distance = 2.0f;
    const auto r2 = gicp(base, shifted, Isometry3f::identity(), c2);
    CHECK(std::fabs(r2.transformation.translation[0] - 0.1f) < 0.05f && r2.mse < 0.1f);
    CHECK(throws(ErrorKind::InvalidData, [&] { gicp({}, c, Isometry3f::identity()); }));
    CHECK(throws(ErrorKind::InvalidData, [&] { gicp(sphere(10, 1.0f), sphere(10, 1.0f), Isometry3f::identity()); }));
    GicpConfig zero;
    zero.max_iterations = 0;
    CHECK(throws(ErrorKind::InvalidData, [&] { gicp(c, c, Isometry3f::identity(), zero); }));
  }

  {  // ---- filters (filtering.rs:396-660)
    const std::vector<Point3f> dup{{0, 0, 0}, {0, 0, 0}, {0.1f, 0, 0}, {0.1f, 0, 0}, {0, 0.1f, 0}};
    CHECK(voxel_grid_filter(dup, 0.05f).size() == 3);
    CHECK(voxel_grid_filter({}, 0.1f).empty());
    CHECK(throws(ErrorKind::InvalidData, [&] { voxel_grid_filter(dup, 0.0f); }));
    CHECK(radius_outlier_removal({{0, 0, 0}}, 0.5f, 1).empty());
    std::vector<Point3f> plane;
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 5; ++j) plane.push_back({0.1f * i, 0.1f * j, 0.0f});
    plane.push_back({10, 10, 10});
    plane.push_back({-10, -10, -10});
    CHECK(radius_outlier_removal(plane, 0.5f, 2).size() == 25);
    CHECK(throws(ErrorKind::InvalidData, [&] { radius_outlier_removal(plane, 0.0f, 3); }));
    CHECK(throws(ErrorKind::InvalidData, [&] { radius_outlier_removal(plane, 0.5f, 0); }));
    const std::vector<Point3f> five{{0, 0, 0}, {0.1f, 0, 0}, {0, 0.1f, 0}, {0, 0, 0.1f}, {10, 10, 10}};
    CHECK(statistical_outlier_removal_with_threshold(five, 3, 0.5f).size() == 4);
    std::vector<Point3f> cluster;
    for (int i = 0; i < 10; ++i)
      for (int j = 0; j < 10; ++j)
        for (int k = 0; k < 10; ++k) cluster.push_back({0.1f * i, 0.1f * j, 0.1f * k});
    cluster.push_back({10, 10, 10});
    cluster.push_back({-10, -10, -10});
    cluster.push_back({5, 5, 5});
    const auto kept = statistical_outlier_removal(cluster, 5, 1.0f);
    CHECK(!kept.empty() && kept.size() < cluster.size());
    for (const auto& p : kept) CHECK(std::fabs(p.x) < 9.0f);
    CHECK(throws(ErrorKind::InvalidData, [&] { statistical_outlier_removal(cluster, 0, 1.0f); }));
    CHECK(throws(ErrorKind::InvalidData, [&] { statistical_outlier_removal(cluster, 5, -1.0f); }));
  }

  std::printf("host_mirror_test: %d checks, %d failed\n", g_checks, g_failed);
  return g_failed == 0 ? 0 : 1;
}
