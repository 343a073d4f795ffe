//! Reference-side golden dumper (UNCOMPILED here: the build image has no cargo/rustc).
//!
//! Drop this file into the reference checkout as `examples/dump_golden.rs` (next to
//! `examples/threecrate_dataset_bench.rs`, whose workload definitions it follows, :148-171) and run
//!
//!     cargo run --release --example dump_golden -- tests/golden/inputs_v1.bin tests/golden/reference_v1.bin
//!
//! with the `inputs_v1.bin` this repository's `tests/golden/make_reference_inputs.py` writes.  It
//! runs the REFERENCE's own CPU path - `KdTree::find_k_nearest`, `estimate_normals`,
//! `icp_point_to_plane_detailed` - on those inputs and writes the results in the flat little-endian
//! layout `tests/test_golden.py::test_reference_pinned_vectors` reads.  The day a Rust toolchain is
//! available this turns "parity unpinned" into "pinned": the test then checks BOTH the oracle
//! restatement and the CUDA path against the reference's real output.
//!
//! File layouts (all little-endian):
//!   inputs_v1.bin    u32 magic 0x54433031 | u32 n_cases | per case:
//!                      u32 kind (1 = knn+normals, 2 = icp) | u32 k | u32 n | f32[n*3] cloud
//!                      kind 2 additionally: u32 m | f32[m*3] target | f32[m*3] target normals | u32 max_iters
//!   reference_v1.bin u32 magic 0x54433032 | u32 n_cases | per case:
//!                      kind 1: u64[n*(k+1)] knn indices (u64::MAX pad) | f32[n*(k+1)] distances
//!                              | f32[n*6] NormalPoint3f rows
//!                      kind 2: f32[7] transform (tx,ty,tz,qi,qj,qk,qw) | f32 mse | u64 iterations
//!                              | u32 converged | u64 n_pairs | u64[n_pairs*2] correspondences
use std::fs::File;
use std::io::{BufReader, BufWriter, Read, Write};

use threecrate_algorithms::{estimate_normals, icp_point_to_plane_detailed, KdTree};
use threecrate_core::{NearestNeighborSearch, Point3f, PointCloud, Vector3f};

fn rd_u32(r: &mut impl Read) -> u32 { let mut b = [0u8; 4]; r.read_exact(&mut b).unwrap(); u32::from_le_bytes(b) }
fn rd_f32s(r: &mut impl Read, n: usize) -> Vec<f32> {
    let mut b = vec![0u8; 4 * n];
    r.read_exact(&mut b).unwrap();
    b.chunks_exact(4).map(|c| f32::from_le_bytes([c[0], c[1], c[2], c[3]])).collect()
}
fn pts(v: &[f32]) -> Vec<Point3f> { v.chunks_exact(3).map(|c| Point3f::new(c[0], c[1], c[2])).collect() }

fn main() {
    let a: Vec<String> = std::env::args().collect();
    let mut r = BufReader::new(File::open(&a[1]).expect("inputs"));
    let mut w = BufWriter::new(File::create(&a[2]).expect("output"));
    assert_eq!(rd_u32(&mut r), 0x5443_3031);
    let n_cases = rd_u32(&mut r);
    w.write_all(&0x5443_3032u32.to_le_bytes()).unwrap();
    w.write_all(&n_cases.to_le_bytes()).unwrap();
    for _ in 0..n_cases {
        let (kind, k, n) = (rd_u32(&mut r), rd_u32(&mut r) as usize, rd_u32(&mut r) as usize);
        let cloud = pts(&rd_f32s(&mut r, 3 * n));
        if kind == 1 {
            // kNN(k+1) of every point (nearest_neighbor.rs:177-251), then estimate_normals(k)
            let tree = KdTree::new(&cloud).unwrap();
            let (mut idx, mut dist) = (vec![u64::MAX; n * (k + 1)], vec![f32::INFINITY; n * (k + 1)]);
            for (i, p) in cloud.iter().enumerate() {
                for (j, (id, d)) in tree.find_k_nearest(p, k + 1).into_iter().enumerate() {
                    idx[i * (k + 1) + j] = id as u64;
                    dist[i * (k + 1) + j] = d;
                }
            }
            for v in &idx { w.write_all(&v.to_le_bytes()).unwrap(); }
            for v in &dist { w.write_all(&v.to_le_bytes()).unwrap(); }
            let normals = estimate_normals(&PointCloud::from_points(cloud.clone()), k).unwrap();
            for p in &normals.points {
                for v in [p.position.x, p.position.y, p.position.z, p.normal.x, p.normal.y, p.normal.z] {
                    w.write_all(&v.to_le_bytes()).unwrap();
                }
            }
        } else {
            let m = rd_u32(&mut r) as usize;
            let target = pts(&rd_f32s(&mut r, 3 * m));
            let tn: Vec<Vector3f> = rd_f32s(&mut r, 3 * m).chunks_exact(3)
                .map(|c| Vector3f::new(c[0], c[1], c[2])).collect();
            let max_iters = rd_u32(&mut r) as usize;
            let res = icp_point_to_plane_detailed(
                &PointCloud::from_points(cloud), &PointCloud::from_points(target), &tn,
                nalgebra::Isometry3::identity(), max_iters, None, -1.0).unwrap();
            let t = res.transformation.translation.vector;
            let q = res.transformation.rotation.quaternion().coords; // [i, j, k, w]
            for v in [t.x, t.y, t.z, q[0], q[1], q[2], q[3], res.mse] { w.write_all(&v.to_le_bytes()).unwrap(); }
            w.write_all(&(res.iterations as u64).to_le_bytes()).unwrap();
            w.write_all(&(res.converged as u32).to_le_bytes()).unwrap();
            w.write_all(&(res.correspondences.len() as u64).to_le_bytes()).unwrap();
            for (s, d) in &res.correspondences {
                w.write_all(&(*s as u64).to_le_bytes()).unwrap();
                w.write_all(&(*d as u64).to_le_bytes()).unwrap();
            }
        }
    }
}
