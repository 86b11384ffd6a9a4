// Links libzenu_b200.so (built by `make -C zenu_b200/csrc`, sm_100a only).  Replaces, for the hot path, what
// zenu-cuda-kernel-sys/build.rs:5-53 (nvcc build of kernel/*.cu + bindgen) and zenu-cudnn-frontend-wrapper-sys/build.rs:7-41
// (cmake build of the cuDNN-frontend wrapper) do: there is nothing to compile on the Rust side any more.
use std::env;
use std::path::PathBuf;

fn main() {
    // ZENU_B200_LIB_DIR = directory holding libzenu_b200.so (default: the in-tree build output next to this crate)
    let dir = env::var("ZENU_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../zenu_b200/lib")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=zenu_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=ZENU_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/zenu_b200.h");
}
