//! Drop-in `kzg_setup_powersoftau` lib.rs on top of libptau_b200.so.
//!
//! SOURCE ONLY (never compiled: no rustc in the build image).  Signatures are the
//! reference's, verbatim (/root/reference/src/lib.rs:20,41,56,82,166,170,174,197);
//! the per-point loops are replaced by one FFI call per section.  `download_*` and
//! `Phase1Parameters` keep the reference's code and are elided here.
mod ffi;

use ark_bls12_381::{Bls12_381, Fq, Fq2, G1Affine as ArkG1Affine, G2Affine as ArkG2Affine};
use ark_ff::{BigInteger384, Fp384};
use ark_poly_commit::kzg10::{Powers, UniversalParams, VerifierKey};
use std::collections::BTreeMap;

pub const KZG_SETUP_FILE: &str = "kzg_setup";
const TAU_POWERS_LENGTH: usize = 1 << 21;
const TAU_POWERS_G1_LENGTH: usize = (TAU_POWERS_LENGTH << 1) - 1;

/// 104-byte PTAU_FMT_ARK_MONT_LIMBS record -> G1Affine.  The limbs are already in
/// Montgomery form, which is what `Fp384::new(BigInteger384(..))` expects in ark-ff 0.2;
/// `GroupAffine` is `repr(Rust)`, so the point is built field by field, not transmuted.
fn g1_from_record(rec: &[u8]) -> ArkG1Affine {
    let fq = |b: &[u8]| -> Fq {
        let mut l = [0u64; 6];
        for i in 0..6 {
            l[i] = u64::from_le_bytes(b[8 * i..8 * i + 8].try_into().unwrap());
        }
        Fp384::new(BigInteger384(l))
    };
    ArkG1Affine::new(fq(&rec[0..48]), fq(&rec[48..96]), rec[96] != 0)
}

fn g2_from_record(rec: &[u8]) -> ArkG2Affine {
    let fq = |b: &[u8]| -> Fq {
        let mut l = [0u64; 6];
        for i in 0..6 {
            l[i] = u64::from_le_bytes(b[8 * i..8 * i + 8].try_into().unwrap());
        }
        Fp384::new(BigInteger384(l))
    };
    ArkG2Affine::new(
        Fq2::new(fq(&rec[0..48]), fq(&rec[48..96])),
        Fq2::new(fq(&rec[96..144]), fq(&rec[144..192])),
        rec[192] != 0,
    )
}

/// G1Affine -> 104-byte record: the Montgomery limbs as they sit in memory (`Fp384.0.0`), then the infinity byte.
fn g1_to_record(p: &ArkG1Affine) -> Vec<u8> {
    let mut r = Vec::with_capacity(104);
    for f in [&p.x, &p.y] {
        for limb in (f.0).0.iter() {
            r.extend_from_slice(&limb.to_le_bytes());
        }
    }
    r.extend_from_slice(&[p.infinity as u8, 0, 0, 0, 0, 0, 0, 0]);
    r
}

fn g2_to_record(p: &ArkG2Affine) -> Vec<u8> {
    let mut r = Vec::with_capacity(200);
    for f in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1] {
        for limb in (f.0).0.iter() {
            r.extend_from_slice(&limb.to_le_bytes());
        }
    }
    r.extend_from_slice(&[p.infinity as u8, 0, 0, 0, 0, 0, 0, 0]);
    r
}

unsafe fn new_ctx() -> *mut ffi::ptau_ctx {
    let mut ctx = std::ptr::null_mut();
    assert_eq!(ffi::ptau_create(&mut ctx, 1, std::ptr::null(), 0), 0, "a B200 is required");
    ctx
}

fn load(variant: i32) -> (Vec<ArkG1Affine>, Vec<ArkG2Affine>) {
    let data = std::fs::read(KZG_SETUP_FILE).unwrap();
    let n = TAU_POWERS_LENGTH as u64;
    let fast = variant == ffi::PTAU_VARIANT_FASTKGZ;
    let n_g1 = 3 * n - 1 + if fast { 0 } else { 2 };
    let n_g2 = if fast { n + 2 } else { 2 };
    let mut g1 = vec![0u8; (n_g1 * 104) as usize];
    let mut g2 = vec![0u8; (n_g2 * 200) as usize];
    let (mut bad_i, mut bad_k) = (0u64, 0i32);
    unsafe {
        let mut ctx = std::ptr::null_mut();
        assert_eq!(ffi::ptau_create(&mut ctx, 1, std::ptr::null(), 0), 0, "a B200 is required");
        let rc = ffi::ptau_load_setup(
            ctx, variant, data.as_ptr() as *const _, data.len() as u64, n, ffi::PTAU_CHECKS_LOAD,
            g1.as_mut_ptr() as *mut _, g1.len() as u64, g2.as_mut_ptr() as *mut _, g2.len() as u64, &mut bad_i,
            &mut bad_k,
        );
        ffi::ptau_destroy(ctx);
        // the reference `unwrap()`s every deserialize_unchecked (src/lib.rs:180)
        assert_eq!(rc, 0, "InvalidData at point {} (kind {})", bad_i, bad_k);
    }
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

    /// sum_i coeffs[i] * powers[i]  (the MSM inside `KZG10::commit` / `open`)
    pub fn commit(powers: &[ArkG1Affine], coeffs: &[Fr]) -> ArkG1Affine {
        let recs: Vec<u8> = powers[..coeffs.len()].iter().flat_map(g1_to_record).collect();
        let mut out = [0u8; 104];
        unsafe {
            let ctx = new_ctx();
            let rc = ffi::ptau_kzg_commit(ctx, recs.as_ptr() as *const _, le32(coeffs).as_ptr() as *const _, coeffs.len(), out.as_mut_ptr() as *mut _);
            ffi::ptau_destroy(ctx);
            assert_eq!(rc, 0);
        }
        g1_from_record(&out)
    }

    /// `KZG10::check(&vk, &comm, point, value, &proof)` for many openings at once (one GPU thread per opening)
    pub fn check_many(vk: &VerifierKey<Bls12_381>, comms: &[Commitment<Bls12_381>], points: &[Fr], values: &[Fr],
                      proofs: &[Proof<Bls12_381>]) -> Vec<bool> {
        let n = comms.len();
        let g1: Vec<u8> = [vk.g, vk.gamma_g].iter().flat_map(g1_to_record).collect();
        let g2: Vec<u8> = [vk.h, vk.beta_h].iter().flat_map(g2_to_record).collect();
        let c: Vec<u8> = comms.iter().flat_map(|c| g1_to_record(&c.0)).collect();
        let w: Vec<u8> = proofs.iter().flat_map(|p| g1_to_record(&p.w)).collect();
        let rv: Vec<Fr> = proofs.iter().map(|p| p.random_v.unwrap_or_default()).collect(); // None == 0: [0]gamma_g = O
        let mut ok = vec![0u8; n];
        unsafe {
            let ctx = new_ctx();
            let rc = ffi::ptau_kzg_check(ctx, g1.as_ptr() as *const _, g2.as_ptr() as *const _, c.as_ptr() as *const _,
                                         le32(points).as_ptr() as *const _, le32(values).as_ptr() as *const _,
                                         w.as_ptr() as *const _, le32(&rv).as_ptr() as *const _, n, ok.as_mut_ptr());
            ffi::ptau_destroy(ctx);
            assert_eq!(rc, 0);
        }
        ok.into_iter().map(|b| b != 0).collect()
    }
}
