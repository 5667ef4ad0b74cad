// UNVERIFIED SOURCE (no Rust toolchain in the authoring environment).
//
// Builds libp25cu.so with nvcc for sm_100a only and links it.  P25CU_SRC points at the root of the
// p25rx_b200 repository (default: three directories up from this crate).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let root = env::var("P25CU_SRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../.."));
    let csrc = root.join("p25rx_b200/csrc");
    // The Makefile compiles every .cu with `-gencode arch=compute_100a,code=sm_100a -lineinfo`; there is no other
    // architecture and no CPU fallback: p25cu_create fails on anything that is not an sm_100 device.
    let status = Command::new("make")
        .arg("-C")
        .arg(&csrc)
        .status()
        .expect("failed to run make (nvcc >= 12.8 with sm_100a support is required)");
    assert!(status.success(), "building libp25cu.so failed");
    println!("cargo:rustc-link-search=native={}", root.join("p25rx_b200").display());
    println!("cargo:rustc-link-lib=dylib=p25cu");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/p25cu.h").display());
}
