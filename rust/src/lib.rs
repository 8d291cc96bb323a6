//! Drop-in `kzg_setup_powersoftau` lib.rs on top of libptau_b200.so.
//!
//! SOURCE ONLY (never compiled: no rustc in the build image).  The pub surface is the
//! reference's, signature for signature (/root/reference/src/lib.rs):
//!   KZG_SETUP_FILE :20, Phase1Parameters :30-39, read_g1 :41, read_g2 :56, load_phase1 :82,
//!   download_kzg_setup :166, download_fastkzg_setup :170, load_kzg_setup :174, load_fastkzg_setup :197.
//! The per-point loops are replaced by one FFI call per file; one CUDA context lives for the
//! whole process (`context()`), so `commit` / `check` calls do not pay context creation.
mod ffi;

use ark_bls12_381::{Bls12_381, Fq, Fq2};
use ark_ec::PairingEngine;
use ark_ff::{BigInteger384, Fp384};
use ark_poly_commit::kzg10::{Powers, UniversalParams, VerifierKey};
use ark_serialize::SerializationError;
use once_cell::sync::OnceCell;
use std::collections::BTreeMap;
use std::ffi::CString;
use std::{
    fs::File,
    io::{self, BufReader, Read, Write},
    path::Path,
};

type ArkG1Affine = <ark_ec::bls12::Bls12<ark_bls12_381::Parameters> as PairingEngine>::G1Affine;
type ArkG2Affine = <ark_ec::bls12::Bls12<ark_bls12_381::Parameters> as PairingEngine>::G2Affine;

pub const KZG_SETUP_FILE: &str = "kzg_setup";
const KZG_SETUP_FILE_DIGEST: &str = "87932f626204ab9a5d4be67ef2ee479471baf942364ada2f89840a2afec8925911fb88cb77024e66d759b4970b25cf2a7b03d1fc8c15768e021220b8ba21efcf";
const FASTKZG_SETUP_FILE_DIGEST: &str = "d177841ad145c0d526e56a8d2cde473f09e85944f5c5d6b72d8063e4a199f8a6fca0b0f6ee91ef79df48518b5edd8165bbdecf0fe4eb0d29809032878f8b17ce";
const TAU_POWERS_LENGTH: usize = 1 << 21;
const TAU_POWERS_G1_LENGTH: usize = (TAU_POWERS_LENGTH << 1) - 1;
const KZG_SETUP_URL: &str = "https://heliax-ferveo-v1.s3-eu-west-1.amazonaws.com/ferveo-dkg-kzg-setup";
const FASTKZG_SETUP_URL: &str = "https://heliax-ferveo-v1.s3-eu-west-1.amazonaws.com/ferveo-dkg-fastkzg-setup";

// ---------------------------------------------------------------------------------------------
// the process-wide context
// ---------------------------------------------------------------------------------------------
struct Ctx(*mut ffi::ptau_ctx);
// libptau_b200 contexts are driven by one host thread at a time; the mutex below enforces it
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}

static CONTEXT: OnceCell<std::sync::Mutex<Ctx>> = OnceCell::new();

/// One `ptau_ctx` for the life of the process (PTAU_GPUS = number of GPUs it owns, default 1).  Creating a CUDA
/// context and the fixed-base tables of a verifier key costs hundreds of milliseconds; every call below borrows
/// this one instead.
fn context() -> std::sync::MutexGuard<'static, Ctx> {
    CONTEXT
        .get_or_init(|| {
            let gpus: i32 = std::env::var("PTAU_GPUS").ok().and_then(|s| s.parse().ok()).unwrap_or(1);
            let mut ctx = std::ptr::null_mut();
            let rc = unsafe { ffi::ptau_create(&mut ctx, gpus, std::ptr::null(), 0) };
            assert_eq!(rc, 0, "ptau_create({} GPUs): a CUDA sm_100 device is required", gpus);
            std::sync::Mutex::new(Ctx(ctx))
        })
        .lock()
        .unwrap()
}

// ---------------------------------------------------------------------------------------------
// records <-> arkworks points
// ---------------------------------------------------------------------------------------------
fn fq_from_limbs(b: &[u8]) -> Fq {
    let mut l = [0u64; 6];
    for i in 0..6 {
        let mut w = [0u8; 8];
        w.copy_from_slice(&b[8 * i..8 * i + 8]);
        l[i] = u64::from_le_bytes(w);
    }
    // the limbs are already in Montgomery form, which is what `Fp384::new` stores in ark-ff 0.2
    Fp384::new(BigInteger384(l))
}

/// 104-byte PTAU_FMT_ARK_MONT_LIMBS record -> G1Affine.  `GroupAffine` is `repr(Rust)`, so the point is built
/// field by field, not transmuted.
fn g1_from_record(rec: &[u8]) -> ArkG1Affine {
    ArkG1Affine::new(fq_from_limbs(&rec[0..48]), fq_from_limbs(&rec[48..96]), rec[96] != 0)
}

fn g2_from_record(rec: &[u8]) -> ArkG2Affine {
    ArkG2Affine::new(
        Fq2::new(fq_from_limbs(&rec[0..48]), fq_from_limbs(&rec[48..96])),
        Fq2::new(fq_from_limbs(&rec[96..144]), fq_from_limbs(&rec[144..192])),
        rec[192] != 0,
    )
}

/// G1Affine -> 104-byte record: the Montgomery limbs as they sit in memory (`Fp384.0.0`), then the infinity byte.
fn g1_to_record(p: &ArkG1Affine) -> Vec<u8> {
    let mut r = Vec::with_capacity(104);
    for f in [&p.x, &p.y].iter() {
        for limb in (f.0).0.iter() {
            r.extend_from_slice(&limb.to_le_bytes());
        }
    }
    r.extend_from_slice(&[p.infinity as u8, 0, 0, 0, 0, 0, 0, 0]);
    r
}

fn g2_to_record(p: &ArkG2Affine) -> Vec<u8> {
    let mut r = Vec::with_capacity(200);
    for f in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1].iter() {
        for limb in (f.0).0.iter() {
            r.extend_from_slice(&limb.to_le_bytes());
        }
    }
    r.extend_from_slice(&[p.infinity as u8, 0, 0, 0, 0, 0, 0, 0]);
    r
}

/// A `PTAU_BAD_*` kind as the error `deserialize_uncompressed` returns for it in ark-serialize 0.2.
fn serialization_error(kind: i32) -> SerializationError {
    match kind {
        ffi::PTAU_BAD_FLAGS => SerializationError::UnexpectedFlags,
        _ => SerializationError::InvalidData,
    }
}

// ---------------------------------------------------------------------------------------------
// read_g1 / read_g2 (src/lib.rs:41-80)
// ---------------------------------------------------------------------------------------------
fn read_point(reader: &mut BufReader<File>, group: i32) -> Result<Vec<u8>, SerializationError> {
    let (r_in, r_out) = if group == ffi::PTAU_G1 { (96usize, 104usize) } else { (192, 200) };
    let mut repr = vec![0u8; r_in];
    reader.read_exact(&mut repr).unwrap(); // the reference unwraps the read too (:43, :58)
    let mut out = vec![0u8; r_out];
    let (mut bad_i, mut bad_k) = (0u64, 0i32);
    let ctx = context();
    // zcash-uncompressed bytes parsed with arkworks' flag semantics + subgroup check = what the byte reversal
    // followed by `deserialize_uncompressed` does (PTAU_CHECKS_READ is reference-exact on every input)
    let rc = unsafe {
        ffi::ptau_convert(
            ctx.0, group, ffi::PTAU_FMT_ZCASH_UNCOMPRESSED, repr.as_ptr() as *const _, ffi::PTAU_FMT_ARK_MONT_LIMBS,
            out.as_mut_ptr() as *mut _, 1, ffi::PTAU_CHECKS_READ, &mut bad_i, &mut bad_k,
        )
    };
    match rc {
        0 => Ok(out),
        k if k > 0 => Err(serialization_error(k)),
        e => panic!("libptau_b200: error {}", e),
    }
}

/// One zcash-uncompressed G1 point.  Kept for API compatibility: a batch of one is the worst case for a GPU; the
/// loaders and the binaries below stream whole sections instead.
pub fn read_g1(reader: &mut BufReader<File>) -> Result<ArkG1Affine, SerializationError> {
    read_point(reader, ffi::PTAU_G1).map(|r| g1_from_record(&r))
}

pub fn read_g2(reader: &mut BufReader<File>) -> Result<ArkG2Affine, SerializationError> {
    read_point(reader, ffi::PTAU_G2).map(|r| g2_from_record(&r))
}

// ---------------------------------------------------------------------------------------------
// load_phase1 (src/lib.rs:30-39, 82-121)
// ---------------------------------------------------------------------------------------------
#[derive(Debug)]
pub struct Phase1Parameters {
    alpha: ArkG1Affine,
    beta_g1: ArkG1Affine,
    beta_g2: ArkG2Affine,
    coeffs_g1: Vec<ArkG1Affine>,
    coeffs_g2: Vec<ArkG2Affine>,
    alpha_coeffs_g1: Vec<ArkG1Affine>,
    beta_coeffs_g1: Vec<ArkG1Affine>,
}

pub fn load_phase1(exp: u32) -> io::Result<Phase1Parameters> {
    let m = 2_usize.pow(exp);
    let data = match std::fs::read(format!("../phase1radix2m{}", exp)) {
        Ok(d) => d,
        Err(e) => {
            panic!("Couldn't load phase1radix2m{}: {:?}", exp, e);
        }
    };
    // alpha, beta_g1, m coeffs_g1, m alpha_coeffs_g1, m beta_coeffs_g1 -> g1 ; beta_g2, m coeffs_g2 -> g2
    let mut g1 = vec![0u8; (2 + 3 * m) * 104];
    let mut g2 = vec![0u8; (1 + m) * 200];
    let (mut bad_i, mut bad_k) = (0u64, 0i32);
    let rc = unsafe {
        ffi::ptau_load_phase1(
            context().0, data.as_ptr() as *const _, data.len() as u64, m as u64, ffi::PTAU_CHECKS_READ,
            g1.as_mut_ptr() as *mut _, g1.len() as u64, g2.as_mut_ptr() as *mut _, g2.len() as u64, &mut bad_i,
            &mut bad_k,
        )
    };
    // the reference `unwrap()`s every read_g1 / read_g2 (:95-112); a short file panics inside read_exact (:43)
    assert_eq!(rc, 0, "phase1radix2m{}: error {} at point {} (kind {})", exp, rc, bad_i, bad_k);
    let p1: Vec<ArkG1Affine> = g1.chunks(104).map(g1_from_record).collect();
    let p2: Vec<ArkG2Affine> = g2.chunks(200).map(g2_from_record).collect();
    Ok(Phase1Parameters {
        alpha: p1[0],
        beta_g1: p1[1],
        beta_g2: p2[0],
        coeffs_g1: p1[2..2 + m].to_vec(),
        coeffs_g2: p2[1..1 + m].to_vec(),
        alpha_coeffs_g1: p1[2 + m..2 + 2 * m].to_vec(),
        beta_coeffs_g1: p1[2 + 2 * m..2 + 3 * m].to_vec(),
    })
}

// ---------------------------------------------------------------------------------------------
// download_* (src/lib.rs:123-172): network fetch with a BLAKE2b-512 check; behaviour kept, not on the GPU path
// ---------------------------------------------------------------------------------------------
fn blake2b_hex_matches(data: &[u8], want: &str) -> bool {
    blake2b_simd::State::new().update(data).finalize().to_hex().as_str() == want
}

fn download_setup(file_url: &str, file_digest: &str, check_digest: bool) -> Result<(), minreq::Error> {
    if Path::new(KZG_SETUP_FILE).exists() {
        // an existing file is kept whatever its digest says; the check only prints its verdict (:135-146)
        if check_digest {
            println!("Checking existing {} file...", KZG_SETUP_FILE);
            let mut buffer = Vec::new();
            File::open(KZG_SETUP_FILE)?.read_to_end(&mut buffer)?;
            if blake2b_hex_matches(&buffer, file_digest) {
                println!("Checking passed, using existing {} file.", KZG_SETUP_FILE);
            }
        }
        return Ok(());
    }
    println!("Downloading {}", file_url);
    let fetched = minreq::get(file_url).send()?;
    if !blake2b_hex_matches(fetched.as_bytes(), file_digest) {
        return Err(io::Error::new(
            io::ErrorKind::InvalidData,
            format!("failed validation (expected: {}, fetched {} bytes)", file_digest, fetched.as_bytes().len()),
        )
        .into());
    }
    File::create(KZG_SETUP_FILE)?.write_all(fetched.as_bytes())?;
    Ok(())
}

pub fn download_kzg_setup(check_digest: bool) -> Result<(), minreq::Error> {
    download_setup(KZG_SETUP_URL, KZG_SETUP_FILE_DIGEST, check_digest)
}

pub fn download_fastkzg_setup(check_digest: bool) -> Result<(), minreq::Error> {
    download_setup(FASTKZG_SETUP_URL, FASTKZG_SETUP_FILE_DIGEST, check_digest)
}

// ---------------------------------------------------------------------------------------------
// load_kzg_setup / load_fastkzg_setup (src/lib.rs:174-228)
// ---------------------------------------------------------------------------------------------
fn load(variant: i32) -> (Vec<ArkG1Affine>, Vec<ArkG2Affine>) {
    assert!(Path::new(KZG_SETUP_FILE).exists(), "File::open({}).unwrap()", KZG_SETUP_FILE);
    let n = TAU_POWERS_LENGTH as u64;
    let fast = variant == ffi::PTAU_VARIANT_FASTKGZ;
    let n_g1 = 3 * n - 1 + if fast { 0 } else { 2 };
    let n_g2 = if fast { n + 2 } else { 2 };
    let mut g1 = vec![0u8; (n_g1 * 104) as usize];
    let mut g2 = vec![0u8; (n_g2 * 200) as usize];
    let (mut n_out, mut bad_i, mut bad_k) = (0u64, 0u64, 0i32);
    let path = CString::new(KZG_SETUP_FILE).unwrap();
    let rc = unsafe {
        ffi::ptau_load_setup_file(
            context().0, variant, path.as_ptr(), n, ffi::PTAU_CHECKS_LOAD, g1.as_mut_ptr() as *mut _, g1.len() as u64,
            g2.as_mut_ptr() as *mut _, g2.len() as u64, &mut n_out, &mut bad_i, &mut bad_k,
        )
    };
    // the reference `unwrap()`s every deserialize_unchecked (:180, :183, :192)
    assert_eq!(rc, 0, "InvalidData at point {} (kind {})", bad_i, bad_k);
    (g1.chunks(104).map(g1_from_record).collect(), g2.chunks(200).map(g2_from_record).collect())
}

pub fn load_kzg_setup<'a>() -> (Powers<'a, Bls12_381>, VerifierKey<Bls12_381>) {
    let (g1, g2) = load(ffi::PTAU_VARIANT_KGZ);
    let powers = Powers::<Bls12_381> {
        powers_of_g: ark_std::borrow::Cow::Owned(g1[..TAU_POWERS_G1_LENGTH].to_vec()),
        powers_of_gamma_g: ark_std::borrow::Cow::Owned(
            g1[TAU_POWERS_G1_LENGTH..TAU_POWERS_G1_LENGTH + TAU_POWERS_LENGTH].to_vec(),
        ),
    };
    let base = TAU_POWERS_G1_LENGTH + TAU_POWERS_LENGTH;
    let vk = VerifierKey::<Bls12_381> {
        g: g1[base],
        gamma_g: g1[base + 1],
        h: g2[0],
        beta_h: g2[1],
        prepared_h: g2[0].into(),
        prepared_beta_h: g2[1].into(),
    };
    (powers, vk)
}

pub fn load_fastkzg_setup() -> (UniversalParams<Bls12_381>, Vec<ArkG2Affine>) {
    let (g1, g2) = load(ffi::PTAU_VARIANT_FASTKGZ);
    let powers_of_h = g2[2..].to_vec();
    let mut powers_of_gamma_g = BTreeMap::<usize, ArkG1Affine>::new();
    for i in 0..TAU_POWERS_LENGTH {
        powers_of_gamma_g.insert(i, g1[TAU_POWERS_G1_LENGTH + i]);
    }
    // the reference's own quirk, kept: `beta_h` is powers_of_h[1], the prepared one comes from the stored beta_h
    let params = UniversalParams::<Bls12_381> {
        powers_of_g: g1[..TAU_POWERS_G1_LENGTH].to_vec(),
        powers_of_gamma_g,
        h: g2[0],
        beta_h: powers_of_h[1],
        neg_powers_of_h: BTreeMap::new(),
        prepared_h: g2[0].into(),
        prepared_beta_h: g2[1].into(),
    };
    (params, powers_of_h)
}

/// The two binaries' `main()` (src/bin/preprocess-kgz.rs:162-200, preprocess-fastkgz.rs:180-214): one call that
/// streams `powersoftau` through pinned slabs (digest, size check, create_new of the intermediate file included).
pub fn preprocess_files(variant_fast: bool) {
    let (input, unc, out) = (
        CString::new("powersoftau").unwrap(),
        CString::new("powersoftau_uncompressed").unwrap(),
        CString::new(KZG_SETUP_FILE).unwrap(),
    );
    println!("Checking existing powersoftau file...");
    println!("Started deserializing compressed Powers of Tau...");
    let (mut bad_i, mut bad_k, mut bad_s) = (0u64, 0i32, -1i32);
    let variant = if variant_fast { ffi::PTAU_VARIANT_FASTKGZ } else { ffi::PTAU_VARIANT_KGZ };
    // NULL digest = POWERSOFTAU_DIGEST of the reference (preprocess-kgz.rs:19); 21 = log2 TAU_POWERS_LENGTH
    let rc = unsafe {
        ffi::ptau_preprocess_files(
            context().0, variant, input.as_ptr(), out.as_ptr(), unc.as_ptr(), 21, std::ptr::null(), 0,
            ffi::PTAU_CHECKS_STRICT, &mut bad_i, &mut bad_k, &mut bad_s,
        )
    };
    if rc != 0 {
        // the reference panics: `read_g1(f).unwrap()` (preprocess-kgz.rs:142), size check (:83), create_new (:118)
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::ptau_strerror(rc)) }.to_string_lossy().into_owned();
        panic!("{} (section {}, point {})", msg, bad_s, bad_i);
    }
    println!("Done serializing. KZG parameters are stored in {}", KZG_SETUP_FILE);
}

/// ark-poly-commit 0.2 `KZG10::{commit, check}` on the GPU (what the reference's own test exercises,
/// /root/reference/src/lib.rs:266-286).  Points travel as the Montgomery-limb records the loader produces
/// (`Fp384` in memory), scalars as `Fr::into_repr()` little-endian bytes.
pub mod kzg {
    use super::*;
    use ark_bls12_381::Fr;
    use ark_ff::{BigInteger, PrimeField};
    use ark_poly_commit::kzg10::{Commitment, Proof};

    fn le32(v: &[Fr]) -> Vec<u8> {
        v.iter().flat_map(|s| s.into_repr().to_bytes_le()).collect()
    }

    /// Powers kept on the GPU: a commitment then uploads only its scalars.
    pub struct ResidentPowers {
        handle: *mut ffi::ptau_kzg_powers,
        len: usize,
    }
    impl ResidentPowers {
        pub fn new(powers: &[ArkG1Affine]) -> Self {
            let recs: Vec<u8> = powers.iter().flat_map(g1_to_record).collect();
            let mut handle = std::ptr::null_mut();
            let rc = unsafe { ffi::ptau_kzg_powers_upload(context().0, recs.as_ptr() as *const _, powers.len(), &mut handle) };
            assert_eq!(rc, 0);
            ResidentPowers { handle, len: powers.len() }
        }
        /// sum_i coeffs[i] * powers[i]
        pub fn commit(&self, coeffs: &[Fr]) -> ArkG1Affine {
            assert!(coeffs.len() <= self.len);
            let mut out = [0u8; 104];
            let rc = unsafe {
                ffi::ptau_kzg_commit_resident(context().0, self.handle, le32(coeffs).as_ptr() as *const _, coeffs.len(), out.as_mut_ptr() as *mut _)
            };
            assert_eq!(rc, 0);
            g1_from_record(&out)
        }
    }
    impl Drop for ResidentPowers {
        fn drop(&mut self) {
            unsafe { ffi::ptau_kzg_powers_free(self.handle) }
        }
    }

    /// sum_i coeffs[i] * powers[i]  (the MSM inside `KZG10::commit` / `open`)
    pub fn commit(powers: &[ArkG1Affine], coeffs: &[Fr]) -> ArkG1Affine {
        let recs: Vec<u8> = powers[..coeffs.len()].iter().flat_map(g1_to_record).collect();
        let mut out = [0u8; 104];
        let rc = unsafe {
            ffi::ptau_kzg_commit(context().0, recs.as_ptr() as *const _, le32(coeffs).as_ptr() as *const _, coeffs.len(), out.as_mut_ptr() as *mut _)
        };
        assert_eq!(rc, 0);
        g1_from_record(&out)
    }

    /// `KZG10::check(&vk, &comm, point, value, &proof)` for many openings at once
    pub fn check_many(vk: &VerifierKey<Bls12_381>, comms: &[Commitment<Bls12_381>], points: &[Fr], values: &[Fr],
                      proofs: &[Proof<Bls12_381>]) -> Vec<bool> {
        let n = comms.len();
        let g1: Vec<u8> = [vk.g, vk.gamma_g].iter().flat_map(g1_to_record).collect();
        let g2: Vec<u8> = [vk.h, vk.beta_h].iter().flat_map(g2_to_record).collect();
        let c: Vec<u8> = comms.iter().flat_map(|c| g1_to_record(&c.0)).collect();
        let w: Vec<u8> = proofs.iter().flat_map(|p| g1_to_record(&p.w)).collect();
        let rv: Vec<Fr> = proofs.iter().map(|p| p.random_v.unwrap_or_default()).collect(); // None == 0: [0]gamma_g = O
        let mut ok = vec![0u8; n];
        let rc = unsafe {
            ffi::ptau_kzg_check(context().0, g1.as_ptr() as *const _, g2.as_ptr() as *const _, c.as_ptr() as *const _,
                                le32(points).as_ptr() as *const _, le32(values).as_ptr() as *const _,
                                w.as_ptr() as *const _, le32(&rv).as_ptr() as *const _, n, ok.as_mut_ptr())
        };
        assert_eq!(rc, 0);
        ok.into_iter().map(|b| b != 0).collect()
    }

    /// The Miller-loop line coefficients ark-ec 0.2 keeps in `G2Prepared { ell_coeffs, infinity }`
    /// (`prepared_h`, `prepared_beta_h`: src/lib.rs:223-224), computed on the GPU: 68 triples of Fq2 per point.
    pub fn g2_prepare(points: &[ArkG2Affine]) -> Vec<(Vec<(Fq2, Fq2, Fq2)>, bool)> {
        let recs: Vec<u8> = points.iter().flat_map(g2_to_record).collect();
        let mut coeffs = vec![0u8; points.len() * 68 * 288];
        let mut inf = vec![0u8; points.len()];
        let rc = unsafe {
            ffi::ptau_g2_prepare(context().0, recs.as_ptr() as *const _, points.len(), coeffs.as_mut_ptr() as *mut _, inf.as_mut_ptr())
        };
        assert_eq!(rc, 0);
        coeffs
            .chunks(68 * 288)
            .zip(inf.iter())
            .map(|(c, &i)| {
                let f2 = |b: &[u8]| Fq2::new(fq_from_limbs(&b[0..48]), fq_from_limbs(&b[48..96]));
                (c.chunks(288).map(|t| (f2(&t[0..96]), f2(&t[96..192]), f2(&t[192..288]))).collect(), i != 0)
            })
            .collect()
    }
}
