"""threecrate_b200 — B200-native (sm_100a) kNN -> normals -> point-to-plane ICP hot path of
rajgandhi1/threecrate, behind the reference's operator interface.

  csrc/      hand-written CUDA kernels + the C ABI (include/threecrate_cuda.h)
  api.py     host-side mirror of the reference functions (ctypes over the C ABI)
  (synthetic clouds / reference fixtures live outside the package: fixtures/synth.py)
  build.py   in-tree nvcc build of lib/libthreecrate_cuda.so

The CUDA library is the only compute path: importing `api` symbols works on a CPU box (for
argument validation and symbol checks), but any compute call without the built library or a
CUDA device raises GpuError.
"""
from .api import (  # noqa: F401
    AlgorithmError,
    Comm,
    Context,
    DeviceCloud,
    GicpConfig,
    GpuError,
    GridIndex,
    ICPResult,
    IDENTITY,
    DEFAULT_SEARCH_FLAGS,
    IcpScaleLevel,
    MultiScaleIcpConfig,
    InvalidData,
    KdTree,
    NormalEstimationConfig,
    ThreecrateError,
    default_context,
    dist_chunk,
    estimate_normals,
    estimate_normals_radius,
    estimate_normals_with_config,
    gicp,
    icp,
    icp_detailed,
    icp_point_to_plane,
    icp_point_to_point,
    icp_point_to_plane_detailed,
    icp_point_to_plane_device,
    icp_point_to_point_device,
    k_nearest_neighbors,
    multiscale_icp_point_to_point,
    pinned_empty,
    radius_outlier_removal,
    statistical_outlier_removal,
    statistical_outlier_removal_with_threshold,
    voxel_grid_filter,
)

__version__ = "0.1.0"
