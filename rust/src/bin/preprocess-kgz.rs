//! Drop-in `preprocess-kgz` main() on top of libptau_b200.so.  SOURCE ONLY (no rustc in the
//! build image; never compiled).  Same cwd files and messages as the reference
//! (/root/reference/src/bin/preprocess-kgz.rs:162-200); the whole body is one FFI call that
//! streams `powersoftau` through pinned slabs.  `preprocess-fastkgz.rs` is identical with
//! PTAU_VARIANT_FASTKGZ.
use std::ffi::CString;
use std::os::raw::{c_char, c_int, c_uint};

#[repr(C)]
pub struct ptau_ctx {
    _private: [u8; 0],
}

#[link(name = "ptau_b200")]
extern "C" {
    fn ptau_create(ctx: *mut *mut ptau_ctx, n_gpus: c_int, device_ids: *const c_int, chunk_points: usize) -> c_int;
    fn ptau_destroy(ctx: *mut ptau_ctx);
    fn ptau_strerror(code: c_int) -> *const c_char;
    fn ptau_preprocess_files(
        ctx: *mut ptau_ctx, variant: c_int, response_path: *const c_char, setup_path: *const c_char,
        uncompressed_path: *const c_char, log2_powers: c_uint, expected_digest_hex: *const c_char, flags: c_uint,
        checks: c_uint, bad_index: *mut u64, bad_kind: *mut c_int, bad_section: *mut c_int,
    ) -> c_int;
}

const PTAU_VARIANT_KGZ: c_int = 1;
const PTAU_CHECKS_STRICT: c_uint = 14;

fn main() {
    let (input, unc, out) = (
        CString::new("powersoftau").unwrap(),
        CString::new("powersoftau_uncompressed").unwrap(),
        CString::new("kzg_setup").unwrap(),
    );
    println!("Checking existing powersoftau file...");
    println!("Started deserializing compressed Powers of Tau...");
    let (mut bad_i, mut bad_k, mut bad_s) = (0u64, 0i32, -1i32);
    let rc = unsafe {
        let mut ctx = std::ptr::null_mut();
        assert_eq!(ptau_create(&mut ctx, 1, std::ptr::null(), 0), 0, "a B200 is required");
        // NULL digest = POWERSOFTAU_DIGEST of the reference (preprocess-kgz.rs:19); 21 = TAU_POWERS_LENGTH
        let rc = ptau_preprocess_files(
            ctx, PTAU_VARIANT_KGZ, input.as_ptr(), out.as_ptr(), unc.as_ptr(), 21, std::ptr::null(), 0,
            PTAU_CHECKS_STRICT, &mut bad_i, &mut bad_k, &mut bad_s,
        );
        ptau_destroy(ctx);
        rc
    };
    if rc != 0 {
        // the reference panics: `read_g1(f).unwrap()` (preprocess-kgz.rs:142), size check (:83), create_new (:118)
        let msg = unsafe { std::ffi::CStr::from_ptr(ptau_strerror(rc)) }.to_string_lossy().into_owned();
        panic!("{} (section {}, point {})", msg, bad_s, bad_i);
    }
    println!("Done serializing. KZG parameters are stored in kzg_setup");
}
