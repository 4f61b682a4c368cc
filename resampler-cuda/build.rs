// build.rs -- compiles the hand-written CUDA sources for sm_100a and links them.
// NOTE: this crate is source only in this repository: the build image has no rustc/cargo, so
// it has never been compiled here.  The C ABI it binds (include/resampler_b200.h) is what the
// Python/ctypes tests exercise.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("resampler_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    for name in ["fir_api.cu", "fir_kernels.cu", "fir_fast.cu", "fir_tensor.cu", "fir_tc2.cu", "fir_submit.cu", "filter_design_device.cu", "fft_resampler.cu", "pcm_ingest.cu", "microbench.cu"] {
        let obj = out.join(name).with_extension("o");
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
            .args(["-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-c"])
            .arg(csrc.join(name))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found");
        assert!(status.success(), "nvcc failed on {name}");
        objs.push(obj);
    }
    let obj = out.join("filter_design.o");
    let status = Command::new("g++")
        .args(["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-c"])
        .arg(csrc.join("filter_design.cpp"))
        .arg("-o")
        .arg(&obj)
        .status()
        .expect("g++ not found");
    assert!(status.success());
    objs.push(obj);
    let lib = out.join("libresampler_b200.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=resampler_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
