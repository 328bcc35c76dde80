// build.rs for the `ezpz` crate with the B200 path (copy to ezpz/build.rs).
//
// Builds libezpz_b200.so from an ezpz-b200 checkout (EZPZ_B200_DIR) with nvcc for sm_100a and links it.  The flags are the
// ones __graft_entry__.build() uses: -fmad=false because the constraint formulas must round exactly as the Rust code does
// (rustc never contracts a*b+c); the linear-algebra phases ask for fused multiply-adds explicitly.
use std::{env, path::PathBuf, process::Command};

const SOURCES: &[&str] = &[
    "device.cu", "freedom.cu", "large.cu", "structure.cpp", "sparse_direct.cpp", "textual.cpp", "host_api.cpp", "multi.cpp",
];

fn main() {
    let dir = env::var("EZPZ_B200_DIR").expect("set EZPZ_B200_DIR to the ezpz-b200 checkout");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libezpz_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let status = Command::new(&nvcc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "-fmad=false", "-DEZPZ_NO_FMAD=1",
            "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-mfma",
            "--shared", "-o",
        ])
        .arg(&lib)
        .args(SOURCES.iter().map(|s| format!("{dir}/ezpz_b200/csrc/{s}")))
        .status()
        .expect("nvcc not found (set NVCC)");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ezpz_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
    println!("cargo:rerun-if-env-changed=EZPZ_B200_DIR");
    for s in SOURCES {
        println!("cargo:rerun-if-changed={dir}/ezpz_b200/csrc/{s}");
    }
    println!("cargo:rerun-if-changed={dir}/include/ezpz_b200.h");
}
