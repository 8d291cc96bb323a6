// Links libptau_b200.so (built by `make -C kzg_setup_powersoftau_b200/csrc`).  SOURCE ONLY, never run here.
// PTAU_B200_LIB_DIR overrides the directory that holds the library; an rpath is recorded so that the binaries
// find it without LD_LIBRARY_PATH.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("PTAU_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../kzg_setup_powersoftau_b200")
    });
    let dir = dir.canonicalize().unwrap_or(dir);
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=ptau_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=PTAU_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../include/ptau_b200.h");
}
