//! Drop-in `preprocess-fastkgz` main() on top of libptau_b200.so.  SOURCE ONLY (no rustc in the
//! build image; never compiled).  Same cwd files and messages as the reference
//! (/root/reference/src/bin/preprocess-fastkgz.rs:180-214): `kzg_setup` in the UniversalParams
//! layout -- powers_of_g, powers_of_gamma_g, h, beta_h, powers_of_h -- with beta_tau_powers_g1
//! read, checked and dropped as the reference does (:156-159).
fn main() {
    kzg_setup_powersoftau::preprocess_files(true);
}
