"""CPU oracle for the kNN -> normals -> point-to-plane ICP path (TEST INFRASTRUCTURE ONLY).

A C++17 restatement of the reference's CPU algorithm (see oracle.cpp for the per-function
reference citations) loaded through ctypes.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing
under ``threecrate_b200/`` does.

PARITY UNPINNED: the reference is Rust-only (no toolchain here) and holds no golden vectors
for this path; the oracle is pinned by the reference's own inline test assertions and by
independent numpy/scipy cross-checks (tests/test_oracle_*.py).
"""
from .filters import (  # noqa: F401
    multiscale_icp_point_to_point,
    radius_outlier_removal,
    sor_mean_distances,
    sor_threshold,
    statistical_outlier_removal,
    statistical_outlier_removal_with_threshold,
    voxel_grid_filter,
)
from .oracle import (  # noqa: F401
    AlgorithmError,
    IcpResult,
    InvalidData,
    OracleKdTree,
    brute_knn,
    build,
    estimate_normals,
    gicp,
    gicp_covariances,
    icp_point_to_plane,
    icp_point_to_point,
    iso_apply,
    iso_mul,
    k_nearest_neighbors,
    last_icp_mse_f64,
    max_threads,
    normals_f64,
    solve6,
    symmetric_eigen3,
)
