//! threecrate-cuda: Rust host code over the `extern "C"` ABI of `include/threecrate_cuda.h`.
//!
//! UNCOMPILED in the build image (no cargo/rustc there) — this is the binding a maintainer adds;
//! the same ABI is exercised from Python/ctypes by the test-suite of the CUDA repo.
//!
//! No wgpu, no multi-backend dispatch, no CPU fallback: with the `cuda` feature on, failures
//! surface as `Error::Gpu`.
use std::ffi::{c_char, c_int, c_void, CStr};
use std::ptr;

use nalgebra::{Isometry3, Quaternion, Translation3, UnitQuaternion};
use threecrate_core::{Error, NormalPoint3f, Point3f, PointCloud, Result, Vector3f};

#[repr(C)]
pub struct TcContext { _p: [u8; 0] }
#[repr(C)]
pub struct TcCloud { _p: [u8; 0] }
#[repr(C)]
pub struct TcIndex { _p: [u8; 0] }

#[repr(C)]
#[derive(Default)]
pub struct TcIcpResult {
    pub transform: [f32; 7], // tx,ty,tz, qi,qj,qk,qw
    pub mse: f32,
    pub iterations: u32,
    pub converged: i32,
    pub n_correspondences: u64,
}

extern "C" {
    fn tc_context_create(device: c_int, out: *mut *mut TcContext) -> c_int;
    fn tc_context_destroy(ctx: *mut TcContext);
    fn tc_last_error(ctx: *const TcContext) -> *const c_char;
    fn tc_cloud_upload(ctx: *mut TcContext, xyz: *const f32, n: u64, out: *mut *mut TcCloud) -> c_int;
    fn tc_cloud_free(c: *mut TcCloud);
    fn tc_index_build(ctx: *mut TcContext, c: *const TcCloud, k_hint: u32, cell: f32,
                      out: *mut *mut TcIndex) -> c_int;
    fn tc_index_free(i: *mut TcIndex);
    fn tc_knn(ctx: *mut TcContext, ix: *const TcIndex, queries: *const f32, nq: u64, k: u32,
              exclude_self: c_int, idx: *mut u32, dist: *mut f32, count: *mut u32) -> c_int;
    fn tc_radius_search(ctx: *mut TcContext, ix: *const TcIndex, query: *const f32, radius: f32,
                        idx: *mut u32, dist: *mut f32, capacity: u64, n_found: *mut u64) -> c_int;
    fn tc_estimate_normals(ctx: *mut TcContext, xyz: *const f32, n: u64, k: u32, radius: f32,
                           consistent: c_int, viewpoint: *const f32, out: *mut f32) -> c_int;
    fn tc_icp_point_to_plane(ctx: *mut TcContext, src: *const f32, ns: u64, tgt: *const f32, nt: u64,
                             nrm: *const f32, nn: u64, init: *const f32, max_iters: u32,
                             max_dist: f32, conv: f32, out: *mut TcIcpResult,
                             pairs: *mut u64) -> c_int;
    fn tc_icp_point_to_point(ctx: *mut TcContext, src: *const f32, ns: u64, tgt: *const f32, nt: u64,
                             init: *const f32, max_iters: u32, max_dist: f32, conv: f32,
                             out: *mut TcIcpResult, pairs: *mut u64) -> c_int;
    fn tc_multiscale_icp_point_to_point(ctx: *mut TcContext, src: *const f32, ns: u64,
                                        tgt: *const f32, nt: u64, init: *const f32,
                                        levels: *const TcIcpScaleLevel, n_levels: u32,
                                        final_iters: u32, final_max_dist: f32, conv: f32,
                                        out: *mut TcIcpResult, pairs: *mut u64) -> c_int;
    fn tc_gicp(ctx: *mut TcContext, src: *const f32, ns: u64, tgt: *const f32, nt: u64,
               init: *const f32, max_iters: u32, max_dist: f32, conv: f32, k: u32,
               out: *mut TcIcpResult, pairs: *mut u64) -> c_int;
    fn tc_cloud_len(c: *const TcCloud) -> u64;
    fn tc_cloud_download(ctx: *mut TcContext, c: *const TcCloud, out: *mut f32) -> c_int;
    fn tc_voxel_grid_filter(ctx: *mut TcContext, c: *const TcCloud, voxel: f32,
                            out: *mut *mut TcCloud) -> c_int;
    fn tc_radius_outlier_removal(ctx: *mut TcContext, c: *const TcCloud, radius: f32,
                                 min_neighbors: u32, out: *mut *mut TcCloud) -> c_int;
    fn tc_statistical_outlier_removal(ctx: *mut TcContext, c: *const TcCloud, k: u32, value: f32,
                                      mode: c_int, stats: *mut f32, out: *mut *mut TcCloud) -> c_int;
    // multi-GPU (one process per GPU)
    fn tc_comm_get_unique_id(ctx: *mut TcContext, id_out: *mut u8) -> c_int;
    fn tc_comm_init_rank(ctx: *mut TcContext, id: *const u8, n_ranks: c_int, rank: c_int,
                         out: *mut *mut TcComm) -> c_int;
    fn tc_comm_destroy(comm: *mut TcComm);
    fn tc_dist_chunk(n_total: u64, n_ranks: c_int, rank: c_int, lo: *mut u64, hi: *mut u64);
    fn tc_comm_window_handle(comm: *mut TcComm, n_total: u64, handle_out: *mut u8) -> c_int;
    fn tc_comm_window_open(comm: *mut TcComm, all_handles: *const u8) -> c_int;
    fn tc_estimate_normals_distributed(ctx: *mut TcContext, comm: *mut TcComm, chunk_xyz: *const f32,
                                       n_total: u64, k: u32, consistent: c_int,
                                       viewpoint: *const f32, chunk_out: *mut f32) -> c_int;
}
#[repr(C)] pub struct TcComm { _p: [u8; 0] }
pub const TC_COMM_ID_BYTES: usize = 128;
pub const TC_IPC_HANDLE_BYTES: usize = 64;

/// `IcpScaleLevel` (registration.rs:28-35) as it crosses the ABI; negative distance = None.
#[repr(C)]
pub struct TcIcpScaleLevel { pub voxel_size: f32, pub max_iterations: u32, pub max_correspondence_distance: f32 }

/// One context per thread (the C ABI serialises calls on a context's stream).
pub struct Context(*mut TcContext);
unsafe impl Send for Context {}

thread_local! {
    static CTX: Context = Context::new(0).expect("no CUDA device (the cuda feature has no CPU fallback)");
}

impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut p = ptr::null_mut();
        match unsafe { tc_context_create(device, &mut p) } {
            0 => Ok(Self(p)),
            _ => Err(Error::Gpu("no usable CUDA device".into())),
        }
    }
    fn check(&self, st: c_int) -> Result<()> {
        if st == 0 { return Ok(()); }
        let msg = unsafe { CStr::from_ptr(tc_last_error(self.0)) }.to_string_lossy().into_owned();
        Err(match st { 1 => Error::InvalidData(msg), 2 => Error::Algorithm(msg), _ => Error::Gpu(msg) })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { tc_context_destroy(self.0) } } }

// `Point3<f32>` is a repr(C) newtype over [f32; 3]: `points.as_ptr() as *const f32` is the AoS the
// ABI expects; `NormalPoint3f` is #[repr(C)] {position, normal} = 6 f32.

/// Drop-in for `threecrate_algorithms::estimate_normals_with_config` (normals.rs:257).
pub fn estimate_normals_with_config(points: &[Point3f], k: usize, radius: Option<f32>,
                                    consistent_orientation: bool, viewpoint: Option<Point3f>)
                                    -> Result<PointCloud<NormalPoint3f>> {
    if points.is_empty() { return Ok(PointCloud::new()); }
    let mut out: Vec<NormalPoint3f> = Vec::with_capacity(points.len());
    CTX.with(|c| {
        let vp = viewpoint.map(|p| [p.x, p.y, p.z]);
        c.check(unsafe {
            tc_estimate_normals(c.0, points.as_ptr() as *const f32, points.len() as u64, k as u32,
                                radius.unwrap_or(-1.0), consistent_orientation as c_int,
                                vp.as_ref().map_or(ptr::null(), |v| v.as_ptr()),
                                out.as_mut_ptr() as *mut f32)
        })
    })?;
    unsafe { out.set_len(points.len()) };
    Ok(PointCloud::from_points(out))
}

/// Drop-in for `estimate_normals(&cloud, k)` (normals.rs:238-247): defaults of
/// `NormalEstimationConfig` (no radius, consistent orientation, bbox-derived viewpoint).
pub fn estimate_normals(points: &[Point3f], k: usize) -> Result<PointCloud<NormalPoint3f>> {
    estimate_normals_with_config(points, k, None, true, None)
}

/// Drop-in for `estimate_normals_radius(&cloud, radius, consistent)` (normals.rs:368-380):
/// `k_neighbors = 10` is the value the reference passes as the fallback count.
pub fn estimate_normals_radius(points: &[Point3f], radius: f32, consistent_orientation: bool)
                               -> Result<PointCloud<NormalPoint3f>> {
    estimate_normals_with_config(points, 10, Some(radius), consistent_orientation, None)
}

/// Device-backed stand-in for `KdTree` (nearest_neighbor.rs:29): build once, query many.
pub struct CudaKdTree { cloud: *mut TcCloud, index: *mut TcIndex, len: usize }

impl CudaKdTree {
    pub fn new(points: &[Point3f]) -> Result<Self> {
        CTX.with(|c| {
            let (mut cl, mut ix) = (ptr::null_mut(), ptr::null_mut());
            c.check(unsafe { tc_cloud_upload(c.0, points.as_ptr() as *const f32, points.len() as u64, &mut cl) })?;
            c.check(unsafe { tc_index_build(c.0, cl, 8, 0.0, &mut ix) })?;
            Ok(Self { cloud: cl, index: ix, len: points.len() })
        })
    }
    /// `NearestNeighborSearch::find_k_nearest` (traits.rs:6-12): (usize, f32) tuples are not
    /// repr(C), so rows are repacked and u32 widened here.
    pub fn find_k_nearest(&self, query: &Point3f, k: usize) -> Vec<(usize, f32)> {
        if k == 0 || self.len == 0 { return Vec::new(); }
        let (mut idx, mut dist, mut cnt) = (vec![0u32; k], vec![0f32; k], [0u32; 1]);
        let q = [query.x, query.y, query.z];
        CTX.with(|c| c.check(unsafe {
            tc_knn(c.0, self.index, q.as_ptr(), 1, k as u32, 0, idx.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr())
        })).expect("tc_knn");
        (0..cnt[0] as usize).map(|j| (idx[j] as usize, dist[j])).collect()
    }
    /// `NearestNeighborSearch::find_radius_neighbors` (traits.rs:6-12; nearest_neighbor.rs:254-298):
    /// every point with d2 <= radius^2, ascending by distance; radius <= 0 finds nothing.
    pub fn find_radius_neighbors(&self, query: &Point3f, radius: f32) -> Vec<(usize, f32)> {
        if !(radius > 0.0) || self.len == 0 { return Vec::new(); }
        let q = [query.x, query.y, query.z];
        let (mut idx, mut dist, mut found) = (vec![0u32; self.len], vec![0f32; self.len], 0u64);
        CTX.with(|c| c.check(unsafe {
            tc_radius_search(c.0, self.index, q.as_ptr(), radius, idx.as_mut_ptr(), dist.as_mut_ptr(),
                             self.len as u64, &mut found)
        })).expect("tc_radius_search");
        (0..(found as usize).min(self.len)).map(|j| (idx[j] as usize, dist[j])).collect()
    }
    /// `PointCloudNeighbors::k_nearest_neighbors` (point_cloud_ops.rs:80-105).
    pub fn k_nearest_neighbors(&self, k: usize) -> Vec<Vec<(usize, f32)>> {
        if k == 0 || self.len == 0 { return Vec::new(); }
        let n = self.len;
        let (mut idx, mut dist, mut cnt) = (vec![0u32; n * k], vec![0f32; n * k], vec![0u32; n]);
        CTX.with(|c| c.check(unsafe {
            tc_knn(c.0, self.index, ptr::null(), n as u64, k as u32, 1, idx.as_mut_ptr(), dist.as_mut_ptr(), cnt.as_mut_ptr())
        })).expect("tc_knn");
        (0..n).map(|i| (0..cnt[i] as usize).map(|j| (idx[i * k + j] as usize, dist[i * k + j])).collect()).collect()
    }
}
impl Drop for CudaKdTree { fn drop(&mut self) { unsafe { tc_index_free(self.index); tc_cloud_free(self.cloud) } } }

/// The trait the reference's generic callers use (`threecrate-core/src/traits.rs:6-12`), so a
/// `CudaKdTree` can stand wherever a `KdTree` is passed as `&dyn NearestNeighborSearch`.
impl threecrate_core::NearestNeighborSearch for CudaKdTree {
    fn find_k_nearest(&self, query: &Point3f, k: usize) -> Vec<(usize, f32)> {
        CudaKdTree::find_k_nearest(self, query, k)
    }
    fn find_radius_neighbors(&self, query: &Point3f, radius: f32) -> Vec<(usize, f32)> {
        CudaKdTree::find_radius_neighbors(self, query, radius)
    }
}

/// Mirror of `threecrate_algorithms::ICPResult` fields (registration.rs:13-24).
pub struct IcpOut {
    pub transformation: Isometry3<f32>, pub mse: f32, pub iterations: usize, pub converged: bool,
    pub correspondences: Vec<(usize, usize)>,
}

/// Drop-in for `icp_point_to_plane_detailed` (registration.rs:508).
pub fn icp_point_to_plane_detailed(source: &[Point3f], target: &[Point3f], normals: &[Vector3f],
                                   init: Isometry3<f32>, max_iters: usize, max_dist: Option<f32>,
                                   conv: f32) -> Result<IcpOut> {
    let q = init.rotation.quaternion().coords; // [i, j, k, w]
    let t = init.translation.vector;
    let init7 = [t.x, t.y, t.z, q[0], q[1], q[2], q[3]]; // never rely on Isometry3's layout
    let mut res = TcIcpResult::default();
    let mut pairs = vec![0u64; 2 * source.len().max(1)];
    CTX.with(|c| c.check(unsafe {
        tc_icp_point_to_plane(c.0, source.as_ptr() as *const f32, source.len() as u64,
                              target.as_ptr() as *const f32, target.len() as u64,
                              normals.as_ptr() as *const f32, normals.len() as u64, init7.as_ptr(),
                              max_iters as u32, max_dist.unwrap_or(-1.0), conv, &mut res, pairs.as_mut_ptr())
    }))?;
    let r = res.transform;
    let iso = Isometry3::from_parts(
        Translation3::new(r[0], r[1], r[2]),
        UnitQuaternion::new_unchecked(Quaternion::new(r[6], r[3], r[4], r[5])),
    );
    let m = res.n_correspondences as usize;
    Ok(IcpOut { transformation: iso, mse: res.mse, iterations: res.iterations as usize,
                converged: res.converged != 0,
                correspondences: (0..m).map(|i| (pairs[2 * i] as usize, pairs[2 * i + 1] as usize)).collect() })
}

fn iso7(init: &Isometry3<f32>) -> [f32; 7] {
    let q = init.rotation.quaternion().coords;
    let t = init.translation.vector;
    [t.x, t.y, t.z, q[0], q[1], q[2], q[3]]
}

fn icp_out(res: &TcIcpResult, pairs: &[u64]) -> IcpOut {
    let r = res.transform;
    let m = res.n_correspondences as usize;
    IcpOut {
        transformation: Isometry3::from_parts(
            Translation3::new(r[0], r[1], r[2]),
            UnitQuaternion::new_unchecked(Quaternion::new(r[6], r[3], r[4], r[5]))),
        mse: res.mse, iterations: res.iterations as usize, converged: res.converged != 0,
        correspondences: (0..m).map(|i| (pairs[2 * i] as usize, pairs[2 * i + 1] as usize)).collect(),
    }
}

/// Drop-in for `icp_detailed` (registration.rs:258): no check of the threshold's sign.
pub fn icp_detailed(source: &[Point3f], target: &[Point3f], init: Isometry3<f32>, max_iters: usize,
                    max_dist: Option<f32>, conv: f32) -> Result<IcpOut> {
    icp_point_to_point_unchecked(source, target, init, max_iters, conv, max_dist)
}

/// Drop-in for `icp_point_to_point` (registration.rs:644-680), including its own validation:
/// `convergence_threshold <= 0` is `InvalidData` there (:665-669) before anything runs.
pub fn icp_point_to_point(source: &[Point3f], target: &[Point3f], init: Isometry3<f32>,
                          max_iters: usize, conv: f32, max_dist: Option<f32>) -> Result<IcpOut> {
    if conv <= 0.0 {
        return Err(Error::InvalidData("Convergence threshold must be positive".into()));
    }
    icp_point_to_point_unchecked(source, target, init, max_iters, conv, max_dist)
}

fn icp_point_to_point_unchecked(source: &[Point3f], target: &[Point3f], init: Isometry3<f32>,
                                max_iters: usize, conv: f32, max_dist: Option<f32>) -> Result<IcpOut> {
    let init7 = iso7(&init);
    let mut res = TcIcpResult::default();
    let mut pairs = vec![0u64; 2 * source.len().max(1)];
    CTX.with(|c| c.check(unsafe {
        tc_icp_point_to_point(c.0, source.as_ptr() as *const f32, source.len() as u64,
                              target.as_ptr() as *const f32, target.len() as u64, init7.as_ptr(),
                              max_iters as u32, max_dist.unwrap_or(-1.0), conv, &mut res, pairs.as_mut_ptr())
    }))?;
    Ok(icp_out(&res, &pairs))
}

/// Drop-in for `multiscale_icp_point_to_point` (registration.rs:704).
pub fn multiscale_icp_point_to_point(source: &[Point3f], target: &[Point3f], init: Isometry3<f32>,
                                     levels: &[(f32, usize, Option<f32>)], final_iters: usize,
                                     final_max_dist: Option<f32>, conv: f32) -> Result<IcpOut> {
    let lv: Vec<TcIcpScaleLevel> = levels.iter().map(|&(v, it, d)| TcIcpScaleLevel {
        voxel_size: v, max_iterations: it as u32, max_correspondence_distance: d.unwrap_or(-1.0) }).collect();
    let init7 = iso7(&init);
    let mut res = TcIcpResult::default();
    let mut pairs = vec![0u64; 2 * source.len().max(1)];
    CTX.with(|c| c.check(unsafe {
        tc_multiscale_icp_point_to_point(c.0, source.as_ptr() as *const f32, source.len() as u64,
                                         target.as_ptr() as *const f32, target.len() as u64,
                                         init7.as_ptr(), lv.as_ptr(), lv.len() as u32, final_iters as u32,
                                         final_max_dist.unwrap_or(-1.0), conv, &mut res, pairs.as_mut_ptr())
    }))?;
    Ok(icp_out(&res, &pairs))
}

/// Drop-in for `gicp` (gicp.rs:117); arguments are the fields of `GicpConfig`.
pub fn gicp(source: &[Point3f], target: &[Point3f], init: Isometry3<f32>, max_iters: usize,
            max_dist: f32, conv: f32, k: usize) -> Result<IcpOut> {
    let init7 = iso7(&init);
    let mut res = TcIcpResult::default();
    let mut pairs = vec![0u64; 2 * source.len().max(1)];
    CTX.with(|c| c.check(unsafe {
        tc_gicp(c.0, source.as_ptr() as *const f32, source.len() as u64, target.as_ptr() as *const f32,
                target.len() as u64, init7.as_ptr(), max_iters as u32, max_dist, conv, k as u32,
                &mut res, pairs.as_mut_ptr())
    }))?;
    Ok(icp_out(&res, &pairs))
}

/// Shared tail of the three filters: upload, run `f`, download the resulting cloud.
fn filter_with(points: &[Point3f],
               f: impl Fn(*mut TcContext, *const TcCloud, *mut *mut TcCloud) -> c_int) -> Result<Vec<Point3f>> {
    CTX.with(|c| {
        let (mut cloud, mut out) = (ptr::null_mut(), ptr::null_mut());
        c.check(unsafe { tc_cloud_upload(c.0, points.as_ptr() as *const f32, points.len() as u64, &mut cloud) })?;
        let st = f(c.0, cloud, &mut out);
        unsafe { tc_cloud_free(cloud) };
        c.check(st)?;
        let n = unsafe { tc_cloud_len(out) } as usize;
        let mut v = vec![Point3f::origin(); n];
        let st = unsafe { tc_cloud_download(c.0, out, v.as_mut_ptr() as *mut f32) };
        unsafe { tc_cloud_free(out) };
        c.check(st)?;
        Ok(v)
    })
}

/// Drop-ins for filtering.rs:38, 167, 253, 335.
pub fn voxel_grid_filter(points: &[Point3f], voxel_size: f32) -> Result<Vec<Point3f>> {
    filter_with(points, |ctx, c, out| unsafe { tc_voxel_grid_filter(ctx, c, voxel_size, out) })
}
pub fn radius_outlier_removal(points: &[Point3f], radius: f32, min_neighbors: usize) -> Result<Vec<Point3f>> {
    filter_with(points, |ctx, c, out| unsafe { tc_radius_outlier_removal(ctx, c, radius, min_neighbors as u32, out) })
}
pub fn statistical_outlier_removal(points: &[Point3f], k: usize, std_dev_multiplier: f32) -> Result<Vec<Point3f>> {
    filter_with(points, |ctx, c, out| unsafe {
        tc_statistical_outlier_removal(ctx, c, k as u32, std_dev_multiplier, 0, ptr::null_mut(), out) })
}
pub fn statistical_outlier_removal_with_threshold(points: &[Point3f], k: usize, threshold: f32) -> Result<Vec<Point3f>> {
    filter_with(points, |ctx, c, out| unsafe {
        tc_statistical_outlier_removal(ctx, c, k as u32, threshold, 2, ptr::null_mut(), out) })
}

/// One rank of a multi-GPU job (one process per GPU).  The 128-byte id (rank 0: `unique_id`) and
/// the 64-byte window handles travel between the processes by whatever the host application has
/// (MPI, a socket, a file): the library never opens a connection of its own.
pub struct Comm { ctx: Context, comm: *mut TcComm, pub n_ranks: usize, pub rank: usize }

impl Comm {
    pub fn unique_id(ctx: &Context) -> Result<[u8; TC_COMM_ID_BYTES]> {
        let mut id = [0u8; TC_COMM_ID_BYTES];
        ctx.check(unsafe { tc_comm_get_unique_id(ctx.0, id.as_mut_ptr()) })?;
        Ok(id)
    }
    pub fn new(device: i32, id: &[u8; TC_COMM_ID_BYTES], n_ranks: usize, rank: usize) -> Result<Self> {
        let ctx = Context::new(device)?;
        let mut p = ptr::null_mut();
        ctx.check(unsafe { tc_comm_init_rank(ctx.0, id.as_ptr(), n_ranks as c_int, rank as c_int, &mut p) })?;
        Ok(Self { ctx, comm: p, n_ranks, rank })
    }
    /// Rows `[lo, hi)` of an `n_total`-point cloud this rank passes in and gets back.
    pub fn chunk(&self, n_total: usize) -> (usize, usize) {
        let (mut lo, mut hi) = (0u64, 0u64);
        unsafe { tc_dist_chunk(n_total as u64, self.n_ranks as c_int, self.rank as c_int, &mut lo, &mut hi) };
        (lo as usize, hi as usize)
    }
    /// Step 1 of the window set-up (once per cloud size): this rank's handle, to be gathered.
    pub fn window_handle(&self, n_total: usize) -> Result<[u8; TC_IPC_HANDLE_BYTES]> {
        let mut h = [0u8; TC_IPC_HANDLE_BYTES];
        self.ctx.check(unsafe { tc_comm_window_handle(self.comm, n_total as u64, h.as_mut_ptr()) })?;
        Ok(h)
    }
    /// Step 2: the handles of ALL ranks, in rank order.
    pub fn open_window(&self, handles: &[[u8; TC_IPC_HANDLE_BYTES]]) -> Result<()> {
        assert_eq!(handles.len(), self.n_ranks);
        let blob: Vec<u8> = handles.iter().flatten().copied().collect();
        self.ctx.check(unsafe { tc_comm_window_open(self.comm, blob.as_ptr()) })
    }
    /// `estimate_normals` (normals.rs:238) of ONE cloud over all ranks: `chunk` = this rank's
    /// rows `self.chunk(n_total)`; returns the `NormalPoint3f` of the same rows, bit-identical
    /// to the single-GPU result.  Collective: every rank calls it.
    pub fn estimate_normals(&self, chunk: &[Point3f], n_total: usize, k: usize) -> Result<Vec<NormalPoint3f>> {
        let (lo, hi) = self.chunk(n_total);
        if chunk.len() != hi - lo {
            return Err(Error::InvalidData(format!("rank {} holds rows [{lo}, {hi})", self.rank)));
        }
        let mut out: Vec<NormalPoint3f> = Vec::with_capacity(chunk.len());
        self.ctx.check(unsafe {
            tc_estimate_normals_distributed(self.ctx.0, self.comm, chunk.as_ptr() as *const f32,
                                            n_total as u64, k as u32, 1, ptr::null(),
                                            out.as_mut_ptr() as *mut f32)
        })?;
        unsafe { out.set_len(chunk.len()) };  // (every row written by the call: repr(C) 6 x f32)
        Ok(out)
    }
}
impl Drop for Comm { fn drop(&mut self) { unsafe { tc_comm_destroy(self.comm) } } }

#[allow(dead_code)]
fn _unused(_: *mut c_void) {}
