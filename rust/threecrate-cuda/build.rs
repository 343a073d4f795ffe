// build.rs — compiles the CUDA sources for sm_100a with nvcc and links the result.
// (Uncompiled in the build image: no Rust toolchain there.)
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("THREECRATE_CUDA_CSRC").unwrap_or_else(|_| "csrc".into()));
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["tc_api.cu", "tc_index.cu", "tc_search.cu", "tc_icp.cu", "tc_comm.cu", "tc_filter.cu"];
    let mut objs = Vec::new();
    for s in sources {
        let o = out.join(s.replace(".cu", ".o"));
        let st = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-Xcompiler", "-fPIC", "-I", "include", "-c"])
            .arg(csrc.join(s)).arg("-o").arg(&o)
            .status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {s}");
        objs.push(o);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let lib = out.join("libthreecrate_cuda.a");
    let st = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=threecrate_cuda");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
}
