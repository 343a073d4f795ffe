/* threecrate_cuda.h — C ABI of the B200-native kNN -> normals -> point-to-plane ICP path.
 *
 * This is the boundary the reference's Rust crates would bind (a `threecrate-cuda` crate behind a
 * `cuda` cargo feature on threecrate-algorithms; see INTEGRATION.md for the `extern "C"` block and
 * the #[cfg(feature = "cuda")] routing).  Plain pointers and sizes only; no torch/nalgebra types.
 * Citations are paths relative to the reference checkout.
 *
 * Conventions
 *   - Every function returns a tc_status; the message for the last failure on a context is
 *     available from tc_last_error().  Status codes map 1:1 onto threecrate_core::Error variants
 *     (threecrate-core/src/error.rs:7-28): InvalidData, Algorithm, Gpu.
 *   - Points are AoS f32 triples, i.e. exactly `Vec<Point3f>::as_ptr()` (nalgebra Point3<f32> is a
 *     repr(C) [f32;3]; threecrate-core/src/point.rs:8).  Normals output is AoS f32 sextuples ==
 *     `#[repr(C)] NormalPoint3f {position, normal}` (threecrate-core/src/point.rs:31-36).
 *   - Indices are u32 (the reference's own tree limits N < 2^32: NIL = u32::MAX,
 *     threecrate-algorithms/src/nearest_neighbor.rs:8); TC_NO_INDEX pads short rows.
 *   - Rigid transforms cross the ABI as 7 floats [tx,ty,tz, qi,qj,qk,qw] (nalgebra Isometry3
 *     layout is not relied upon).
 *   - A tc_context owns one CUDA stream on one device; calls on a context are serialised on that
 *     stream; contexts are independent (one per thread, like rayon workers sharing `&KdTree`).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point returns TC_GPU.
 *   - Tie rule for bit-equal squared distances: ascending (d2, original index) — the order of
 *     SimdBruteForceSearch (threecrate-algorithms/src/simd_distance.rs:370-385,444-452).
 */
#ifndef THREECRATE_CUDA_H
#define THREECRATE_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TC_NO_INDEX 0xFFFFFFFFu

typedef enum tc_status {
  TC_OK = 0,
  TC_INVALID_DATA = 1, /* Error::InvalidData */
  TC_ALGORITHM = 2,    /* Error::Algorithm   */
  TC_GPU = 3           /* Error::Gpu         */
} tc_status;

typedef struct tc_context tc_context; /* device + stream + scratch                       */
typedef struct tc_cloud tc_cloud;     /* device-resident PointCloud<Point3f>              */
typedef struct tc_index tc_index;     /* uniform-grid spatial index over a tc_cloud       */
typedef struct tc_comm tc_comm;       /* NCCL communicator for the sharded ICP reduction  */

/* ---- context -------------------------------------------------------------------------- */
int tc_context_create(int device, tc_context** out);
void tc_context_destroy(tc_context* ctx);
const char* tc_last_error(const tc_context* ctx);
/* cudaStream_t of the context (for CUDA-event timing by the caller). */
void* tc_context_stream(tc_context* ctx);
int tc_context_synchronize(tc_context* ctx);
/* Number of kernels this library has launched on the context so far. */
uint64_t tc_launch_count(const tc_context* ctx);
/* CUDA-event timer on the context's stream: start, run calls, stop -> elapsed ms. */
int tc_timer_start(tc_context* ctx);
int tc_timer_stop(tc_context* ctx, float* ms_out);
const char* tc_version(void);

/* Per-call statistics of the last search / ICP call on the context (SURVEY §5 "metrics"): off by
 * default because collecting them costs one memset and a few atomics per call.  The search
 * counters describe the staged-tile kNN / normals kernel; the ICP block is the per-iteration
 * history of the last registration call (first TC_STATS_MAX_ITERS iterations). */
#define TC_STATS_MAX_ITERS 64
typedef struct tc_stats {
  uint64_t queries;            /* queries of the last kNN / normals launch                      */
  uint64_t chain_queries;      /* ... finished by the exact chain kernel (ties, unproven boxes)  */
  uint64_t rounds;             /* warp rounds (one staged box each)                              */
  uint64_t box_splits;         /* boxes halved because they exceeded the tile                    */
  uint64_t retries;            /* rounds repeated with a wider box or a coarser level            */
  uint64_t candidates_staged;  /* candidate points copied into shared memory, summed over rounds */
  uint64_t merges;             /* selection-list merges, summed over warps                       */
  uint32_t icp_iterations;     /* iterations recorded below                                      */
  uint32_t reserved;
  float icp_mse[TC_STATS_MAX_ITERS];        /* mse of each iteration (pre-update residuals)    */
  uint64_t icp_valid[TC_STATS_MAX_ITERS];   /* accepted correspondences of each iteration       */
} tc_stats;
int tc_stats_enable(tc_context* ctx, int on);
int tc_last_stats(tc_context* ctx, tc_stats* out); /* synchronises the context's stream */

/* ---- device-resident cloud (replaces per-call re-pack/upload of the wgpu path) ------------ */
/* Upload `n` points from host AoS f32 (PointCloud<Point3f>.points, point_cloud.rs:11-13). */
int tc_cloud_upload(tc_context* ctx, const float* xyz_aos, uint64_t n, tc_cloud** out);
/* Upload from strided host records, e.g. KITTI .bin x,y,z,intensity stride 16
 * (threecrate-io/src/lidar.rs:310-345). */
int tc_cloud_upload_strided(tc_context* ctx, const void* base, uint64_t n, uint32_t stride_bytes,
                            tc_cloud** out);
/* Wrap points that already live in device memory (copied; AoS f32, n x 3). */
int tc_cloud_from_device(tc_context* ctx, const float* d_xyz_aos, uint64_t n, tc_cloud** out);
void tc_cloud_free(tc_cloud* cloud);
uint64_t tc_cloud_len(const tc_cloud* cloud);
/* Copy the points back to host AoS f32 (n x 3); synchronises the context's stream. */
int tc_cloud_download(tc_context* ctx, const tc_cloud* cloud, float* xyz_aos_out);

/* ---- filters (threecrate-algorithms/src/filtering.rs) -------------------------------------
 * Each returns a NEW device-resident cloud in *out (free with tc_cloud_free); an empty input
 * yields an empty cloud, as in the reference.
 * voxel_grid_filter (filtering.rs:38-133): one centroid per occupied voxel, voxel =
 * floor((p - bbox_min) / voxel_size) per axis, centroid summed in f64 in original point order and
 * rounded to f32 like the reference.  Output order: ascending (z, y, x) voxel coordinate (the
 * reference's order is a HashMap iteration order).  voxel_size <= 0 -> TC_INVALID_DATA. */
int tc_voxel_grid_filter(tc_context* ctx, const tc_cloud* cloud, float voxel_size, tc_cloud** out);
/* radius_outlier_removal (filtering.rs:167-218): keeps, in original order, the points with at
 * least `min_neighbors` OTHER points within `radius` (d2 <= radius^2).  radius <= 0 or
 * min_neighbors == 0 -> TC_INVALID_DATA. */
int tc_radius_outlier_removal(tc_context* ctx, const tc_cloud* cloud, float radius,
                              uint32_t min_neighbors, tc_cloud** out);
/* statistical_outlier_removal (filtering.rs:253-321) and .._with_threshold (:335-394): keeps, in
 * original order, the points whose mean distance to their k nearest neighbours (kNN(k+1), every
 * neighbour with the point's own coordinates skipped, f32 sum in ascending-distance order) is
 * <= threshold.
 *   mode 0: threshold = mean + value * std_dev of those mean distances, accumulated like the
 *           reference (sequential f32 sums over the cloud) - bit-exact, ~6 ns/point serial tail;
 *   mode 1: same statistics from f64 tree sums (fast; a point whose mean distance lies within the
 *           reference's own f32 accumulation error of the threshold may be classified differently);
 *   mode 2: `value` IS the threshold (statistical_outlier_removal_with_threshold).
 * stats_out (may be NULL) receives {global mean, std dev, threshold}.  k_neighbors == 0 or
 * value <= 0 -> TC_INVALID_DATA. */
int tc_statistical_outlier_removal(tc_context* ctx, const tc_cloud* cloud, uint32_t k_neighbors,
                                   float value, int mode, float* stats_out, tc_cloud** out);

/* ---- spatial index (replaces KdTree::new, nearest_neighbor.rs:37-60) ---------------------- */
/* Builds the uniform grid(s): bbox -> cell histogram -> cell-range scan -> counting-sort scatter
 * into float4 (x,y,z,original index) sorted by cell; up to three resolutions when the density is
 * strongly skewed (LiDAR).  cell_size <= 0 selects the cell edge automatically from `k_hint`
 * (the k the index will mostly be queried with; 1 for ICP correspondence search); an explicit
 * cell_size builds exactly one level.  Coordinates must be finite: NaN / Inf anywhere in a cloud
 * (index build, kNN queries, ICP source, filters) is TC_INVALID_DATA - a documented deviation from
 * the reference, whose KdTree::new accepts them and treats NaN as "Equal" in comparisons. */
int tc_index_build(tc_context* ctx, const tc_cloud* cloud, uint32_t k_hint, float cell_size,
                   tc_index** out);
void tc_index_free(tc_index* index);
/* Multi-GPU normals (queries sharded, one process per GPU, no data-path collective): every rank
 * calls this on ITS copy of the same cloud.  The ranks agree on one grid without talking (bbox from
 * the whole cloud; cell size and level decisions from the statistics of every 8th cell plane,
 * which every rank histograms alike), cut it into `world` slabs of whole cell planes along its
 * longest axis - the boundaries placed so that the slabs hold equal numbers of POINTS (per-plane
 * counts of the whole cloud, identical on every rank) - and rank `rank` counts, scans and sorts only
 * the cells of its slab plus a halo of a few planes: the histogram atomics, the cell-range scan and
 * the counting-sort scatter shrink by ~1/world.  tc_estimate_normals_device on such an index computes exactly the rank's own rows
 * (pass shard range [0, UINT64_MAX)); should a query ever need points beyond the halo the call
 * transparently finishes on a complete index, so results never depend on the sharding.  Clouds that
 * need several grid resolutions, or have fewer planes than ranks, get a complete index and the
 * rank's share of its sorted order instead.  kNN / ICP entry points need a complete index. */
int tc_index_build_sharded(tc_context* ctx, const tc_cloud* cloud, uint32_t k_hint, float cell_size,
                           int rank, int world, tc_index** out);

typedef struct tc_index_info {
  uint64_t n_points;
  uint64_t n_cells;
  uint32_t dims[3];
  float cell_size;
  float bbox_min[3];
  float bbox_max[3];
  uint32_t occupied_cells;
  uint32_t max_cell_population;
  uint32_t n_levels; /* grid resolutions built (1 unless the density is strongly skewed) */
  uint32_t reserved;
} tc_index_info;
int tc_index_get_info(const tc_index* index, tc_index_info* out);

/* ---- kNN (KdTree::find_k_nearest nearest_neighbor.rs:177-251;
 *           PointCloudNeighbors::k_nearest_neighbors point_cloud_ops.rs:80-105) ---------------
 * queries_aos == NULL: the indexed cloud queries itself (nq must equal its length) and, with
 * exclude_self != 0, each point's own index is removed (kNN(k+1), retain idx != i, truncate k).
 * Outputs are HOST buffers: idx_out/dist_out are nq x k row-major (TC_NO_INDEX / +inf padded when
 * fewer than k exist), dist_out = sqrt(d2) (may be NULL), count_out[nq] valid entries (may be
 * NULL).  Rows ascend by (d2, index).  k == 0 or an empty cloud yields zero counts. */
int tc_knn(tc_context* ctx, const tc_index* index, const float* queries_aos, uint64_t nq,
           uint32_t k, int exclude_self, uint32_t* idx_out, float* dist_out, uint32_t* count_out);
/* Same with DEVICE output buffers (results stay resident; no readback). */
int tc_knn_device(tc_context* ctx, const tc_index* index, const float* d_queries_aos, uint64_t nq,
                  uint32_t k, int exclude_self, uint32_t* d_idx_out, float* d_dist_out,
                  uint32_t* d_count_out);
/* KdTree::find_radius_neighbors (nearest_neighbor.rs:254-298) for one query: all points with
 * d2 <= radius^2, ascending by distance; at most `capacity` written, total returned in n_found. */
int tc_radius_search(tc_context* ctx, const tc_index* index, const float query[3], float radius,
                     uint32_t* idx_out, float* dist_out, uint64_t capacity, uint64_t* n_found);

/* ---- normals (estimate_normals_with_config, normals.rs:257-357) ---------------------------
 * Drop-in: host AoS in, host AoS NormalPoint3f out (n x 6 f32).  Includes upload + index build,
 * as the reference's call includes KdTree::new (normals.rs:272).
 *   n == 0 -> TC_OK, nothing written (checked BEFORE k, normals.rs:261-269);
 *   k < 3  -> TC_INVALID_DATA "k_neighbors must be at least 3";
 *   radius < 0 means None; viewpoint == NULL means the bbox-derived default (normals.rs:275-303). */
int tc_estimate_normals(tc_context* ctx, const float* xyz_aos, uint64_t n, uint32_t k, float radius,
                        int consistent_orientation, const float* viewpoint3, float* out_aos);
/* Same on a prebuilt index (host output). */
int tc_estimate_normals_indexed(tc_context* ctx, const tc_index* index, uint32_t k, float radius,
                                int consistent_orientation, const float* viewpoint3,
                                float* out_aos);
/* Device-resident output, optionally only for the shard [shard_begin, shard_end) of the index's
 * spatially sorted order (multi-GPU: queries sharded over a replicated grid).  d_out_aos is the
 * full n x 6 buffer indexed by ORIGINAL point index; only the shard's rows are written. */
int tc_estimate_normals_device(tc_context* ctx, const tc_index* index, uint32_t k, float radius,
                               int consistent_orientation, const float* viewpoint3,
                               uint64_t shard_begin, uint64_t shard_end, float* d_out_aos);

/* ---- point-to-plane ICP (icp_point_to_plane_detailed, registration.rs:508-602) ------------ */
typedef struct tc_icp_result {
  float transform[7]; /* tx,ty,tz, qi,qj,qk,qw                        (ICPResult.transformation) */
  float mse;          /*                                              (ICPResult.mse)            */
  uint32_t iterations;/*                                              (ICPResult.iterations)     */
  int32_t converged;  /*                                              (ICPResult.converged)      */
  uint64_t n_correspondences; /* pairs of the last executed iteration (ICPResult.correspondences) */
} tc_icp_result;

/* Drop-in: host buffers in, result out.  Includes target index build (registration.rs:536).
 *   ns == 0 or nt == 0 -> TC_INVALID_DATA; n_normals != nt -> TC_INVALID_DATA;
 *   max_iters == 0 -> TC_INVALID_DATA (registration.rs:517-531);
 *   < 6 valid pairs -> TC_ALGORITHM; singular system -> TC_ALGORITHM (registration.rs:568-572,
 *   432-438).  max_corr_dist < 0 means None.  pairs_out (may be NULL): capacity ns x 2 u64,
 *   (source index, target index) of the last executed iteration in source order. */
int tc_icp_point_to_plane(tc_context* ctx, const float* src_aos, uint64_t ns, const float* tgt_aos,
                          uint64_t nt, const float* tgt_normals_aos, uint64_t n_normals,
                          const float init[7], uint32_t max_iters, float max_corr_dist,
                          float conv_threshold, tc_icp_result* out, uint64_t* pairs_out);

/* Device-resident variant: source cloud, prebuilt target index and target normals (device AoS,
 * nt x 3, original target order) stay in HBM; the whole iteration loop runs without a host
 * round-trip.  With comm != NULL every rank passes ITS source shard and the per-iteration
 * 29-scalar normal-equation sums (21 AtA + 6 Atb + sum b^2 + n_valid) are all-reduced; all ranks
 * then solve the identical 6x6 system.  d_match_out (may be NULL): ns u32, matched target index
 * per source point of this rank (TC_NO_INDEX = rejected), original source order. */
int tc_icp_point_to_plane_device(tc_context* ctx, tc_comm* comm, const tc_cloud* src,
                                 const tc_index* tgt_index, const float* d_tgt_normals_aos,
                                 const float init[7], uint32_t max_iters, float max_corr_dist,
                                 float conv_threshold, tc_icp_result* out, uint32_t* d_match_out);

/* ---- point-to-point ICP (icp_detailed, registration.rs:258-370; the workload of the
 *      reference's published benchmark, examples/threecrate_dataset_bench.rs:155-171) ----------
 * Same conventions as the point-to-plane pair above.  < 3 valid pairs -> TC_ALGORITHM
 * (registration.rs:311-315).  When not converged the reported mse is that of the last
 * correspondences under the final transform (registration.rs:342-361).  The wrappers
 * `icp_point_to_point` (conv_threshold <= 0 -> InvalidData, :665-669) and `icp` (errors -> init,
 * :238-241) are host-side one-liners over this call. */
int tc_icp_point_to_point(tc_context* ctx, const float* src_aos, uint64_t ns, const float* tgt_aos,
                          uint64_t nt, const float init[7], uint32_t max_iters, float max_corr_dist,
                          float conv_threshold, tc_icp_result* out, uint64_t* pairs_out);
int tc_icp_point_to_point_device(tc_context* ctx, tc_comm* comm, const tc_cloud* src,
                                 const tc_index* tgt_index, const float init[7], uint32_t max_iters,
                                 float max_corr_dist, float conv_threshold, tc_icp_result* out,
                                 uint32_t* d_match_out);

/* gicp (threecrate-algorithms/src/gicp.rs:117-312), GicpConfig as plain arguments (defaults:
 * 50 iterations, max distance 1.0, convergence 1e-6, k_correspondences 20).  Per-point
 * covariances from kNN(max(k,4)) including the point itself (+1e-4 I), correspondences by exact
 * 1-NN within max_correspondence_distance, Gauss-Newton on M = C_t + R C_s R^T with the 6x6
 * system reduced in f64 on the device.  Errors as in the reference: empty cloud, zero
 * iterations, fewer points than max(k,4), coplanar/collinear cloud (bbox extent < 1e-4) ->
 * TC_INVALID_DATA; fewer than 6 correspondences or a singular system -> TC_ALGORITHM. */
int tc_gicp(tc_context* ctx, const float* src_aos, uint64_t ns, const float* tgt_aos, uint64_t nt,
            const float init[7], uint32_t max_iterations, float max_correspondence_distance,
            float convergence_threshold, uint32_t k_correspondences, tc_icp_result* out,
            uint64_t* pairs_out);

/* multiscale_icp_point_to_point (registration.rs:704-789): coarse-to-fine point-to-point ICP.
 * Per level both clouds are voxel-downsampled (tc_voxel_grid_filter) and icp_point_to_point runs
 * from the previous level's transform (levels whose downsampled clouds hold < 3 points are
 * skipped); a final refinement runs on the full clouds.  iterations = sum over all stages;
 * pairs_out (may be NULL, 2 x ns u64) = the final stage's correspondences.
 * max_correspondence_distance < 0 means None. */
typedef struct tc_icp_scale_level {
  float voxel_size;
  uint32_t max_iterations;
  float max_correspondence_distance;
} tc_icp_scale_level;
int tc_multiscale_icp_point_to_point(tc_context* ctx, const float* src_aos, uint64_t ns,
                                     const float* tgt_aos, uint64_t nt, const float init[7],
                                     const tc_icp_scale_level* levels, uint32_t n_levels,
                                     uint32_t final_refinement_iterations,
                                     float final_max_correspondence_distance,
                                     float convergence_threshold, tc_icp_result* out,
                                     uint64_t* pairs_out);

/* ---- multi-GPU (one process per GPU; NCCL bootstrap, id exchanged out of band) ------------- */
#define TC_COMM_ID_BYTES 128
int tc_comm_get_unique_id(tc_context* ctx, void* id_out /* TC_COMM_ID_BYTES */);
int tc_comm_init_rank(tc_context* ctx, const void* id, int n_ranks, int rank, tc_comm** out);
void tc_comm_destroy(tc_comm* comm);
/* Fused all-reduce over NVLink peer memory (optional; without it the ICP falls back to
 * ncclAllReduce + a separate solve launch).  Each rank calls tc_comm_peer_handle, the host
 * gathers the TC_IPC_HANDLE_BYTES-byte handles of all ranks (rank order) and passes the
 * concatenation to tc_comm_peer_open on every rank.  Afterwards the last block of the ICP
 * correspondence kernel writes its partial normal equations straight into every peer's buffer,
 * waits for the peers' contributions and solves — one kernel per iteration, no collective call. */
#define TC_IPC_HANDLE_BYTES 64
int tc_comm_peer_handle(tc_comm* comm, void* handle_out /* TC_IPC_HANDLE_BYTES */);
int tc_comm_peer_open(tc_comm* comm, const void* all_handles /* n_ranks * TC_IPC_HANDLE_BYTES */);
/* Sum-all-reduce of `count` f64 on the context's stream (exposed for tests). */
int tc_comm_allreduce_f64(tc_comm* comm, double* d_buf, uint64_t count);

/* ---- distributed estimate_normals (normals.rs:238-268 over the ranks of a tc_comm) ---------
 * Strong scaling of ONE cloud: its rows are split into n_ranks contiguous chunks
 * (tc_dist_chunk: equal lengths, a multiple of 4 rows; the last ranks may hold fewer rows or
 * none).  Rank r passes its chunk of the points (HOST) and receives the NormalPoint3f rows of the
 * same range (HOST), bit-identical to the single-GPU result.  Per call a rank uploads only its
 * chunk, pulls the other chunks out of the peers' windows over NVLink, builds the slab-sharded
 * index (tc_index_build_sharded), and its normals kernel writes every row straight into the
 * window of the rank that owns it; two stream-ordered peer barriers, no collective call.
 * Set-up (once per cloud size, collective): every rank calls tc_comm_window_handle, the host
 * gathers the handles of all ranks (rank order) and passes the concatenation to
 * tc_comm_window_open.  A rank's previous window stays allocated until its next
 * tc_comm_window_open, when no peer can still have it mapped.
 * kNN mode only (radius-mode normals: shard tc_estimate_normals_device by sorted range). */
void tc_dist_chunk(uint64_t n_total, int n_ranks, int rank, uint64_t* lo, uint64_t* hi);
int tc_comm_window_handle(tc_comm* comm, uint64_t n_total, void* handle_out /* TC_IPC_HANDLE_BYTES */);
int tc_comm_window_open(tc_comm* comm, const void* all_handles /* n_ranks * TC_IPC_HANDLE_BYTES */);
int tc_estimate_normals_distributed(tc_context* ctx, tc_comm* comm, const float* chunk_xyz_aos,
                                    uint64_t n_total, uint32_t k, int consistent_orientation,
                                    const float* viewpoint3 /* NULL = reference default */,
                                    float* chunk_out_aos /* (hi - lo) x 6 */);

/* ---- raw device memory helpers for hosts without their own CUDA bindings ------------------- */
int tc_device_alloc(tc_context* ctx, uint64_t bytes, void** d_out);
int tc_device_free(tc_context* ctx, void* d_ptr);
int tc_copy_to_device(tc_context* ctx, void* d_dst, const void* h_src, uint64_t bytes);
int tc_copy_to_host(tc_context* ctx, void* h_dst, const void* d_src, uint64_t bytes);
/* Page-locked host memory (so uploads/readbacks run at PCIe rate and asynchronously). */
int tc_host_alloc_pinned(uint64_t bytes, void** h_out);
int tc_host_free_pinned(void* h_ptr);

#ifdef __cplusplus
}
#endif
#endif /* THREECRATE_CUDA_H */
