//! Drop-in `preprocess-kgz` main() on top of libptau_b200.so.  SOURCE ONLY (no rustc in the
//! build image; never compiled).  Same cwd files and messages as the reference
//! (/root/reference/src/bin/preprocess-kgz.rs:162-200): `powersoftau` -> `powersoftau_uncompressed`
//! -> `kzg_setup` (powers_of_g, powers_of_gamma_g, VerifierKey).  The whole body is one FFI call
//! that streams the ceremony file through pinned slabs; a failed run publishes nothing.
fn main() {
    kzg_setup_powersoftau::preprocess_files(false);
}
