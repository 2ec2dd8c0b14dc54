// Locates liblzfear_b200.so (built by `python rust-lz-fear_b200/build.py` with nvcc for sm_100a).
//   LZFEAR_B200_LIB_DIR   directory holding the shared library (default: ../../rust-lz-fear_b200 of this checkout)
// The library links only libcudart; there is no CPU fallback behind it — without a CUDA device lzf_create fails.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("LZFEAR_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../rust-lz-fear_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=lzfear_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=LZFEAR_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/lzfear_b200.h");
}
