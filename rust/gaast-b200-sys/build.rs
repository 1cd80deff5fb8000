// UNCOMPILED (no Rust toolchain in the build image).
// Links libgaast_b200.so, built by `python -m gaast_b200.build` in this repository.
// GAAST_B200_LIB_DIR = the directory that holds libgaast_b200.so (default: ../../gaast_b200).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("GAAST_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../gaast_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=gaast_b200");
    // the library finds its cubin cache next to itself; let the test binaries find the library
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=GAAST_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/gaast_b200.h");
}
