"""One test per [dagger] assumption of SURVEY.md 8a: behaviour of crates whose source is NOT in /root/reference
(pairing 0.14.2, ark-serialize / ark-ff / ark-ec 0.2.0, ark-poly-commit 0.2.0) that the oracle restates from memory.

Each test states the assumption, builds the record that isolates it by hand, and asserts the behaviour on the three
CPU implementations (Python big-int oracle, C oracle, the product's limb code compiled for the host).  If an
assumption is wrong, the named test of rust/tests/golden.rs (run against the real crates with `cargo test`) fails,
and the expectation to flip is the one asserted here -- nothing else in the repository encodes it independently.
"""
import ctypes

import pytest

import ptau_oracle as o
from conftest import golden

SZ = {1: {1: 96, 2: 48, 3: 96, 4: 104}, 2: {1: 192, 2: 96, 3: 192, 4: 200}}
OK, NON_CANONICAL, FLAGS, INFINITY, NOT_ON_CURVE, NOT_IN_SUBGROUP = 0, 1, 2, 3, 4, 5
LOAD, READ, STRICT = 0, 4, 14  # PTAU_CHECKS_*


@pytest.fixture(scope="module")
def impls(hostemul, cref):
    """status(group, in_fmt, record, checks) and bytes(group, in_fmt, record, out_fmt, checks) on both native CPU
    implementations; they must agree with each other before they are compared with the expectation."""
    def run(group, in_fmt, rec, out_fmt, checks):
        out = (ctypes.c_uint8 * SZ[group][out_fmt])()
        st = (ctypes.c_uint32 * 1)()
        assert hostemul.hostemul_convert(group, in_fmt, bytes(rec), out_fmt, out, ctypes.c_size_t(1), checks, st) == 0
        got_c, st_c = cref.convert(group, in_fmt, bytes(rec), out_fmt, checks)
        assert st_c[0] == st[0], "C oracle and limb code disagree"
        if st[0] == OK:
            assert got_c == bytes(out)
        return int(st[0]), bytes(out)
    return run


def _g1(k):
    return o.g1_mul(o.G1_GEN, k)


def _g2(k):
    return o.g2_mul(o.G2_GEN, k)


def test_a3_compressed_flag_bits_and_sign_choice(impls):
    """[dagger 8a-3] pairing 0.14.2 G1/G2Compressed: bit 7 of byte 0 = compressed (required), bit 6 = infinity (then
    every other bit must be 0), bit 5 = "y is the lexicographically largest root"; G2 is x.c1 | x.c0 and Fq2 is
    ordered by c1, then c0.  rust/tests/golden.rs::decompression_matches_pairing_crate, ::edge_cases_match_real_primitives."""
    for grp, q, enc, unc in ((1, _g1(7), o.zcash_g1_compressed_encode, o.zcash_g1_uncompressed_encode),
                             (2, _g2(7), o.zcash_g2_compressed_encode, o.zcash_g2_uncompressed_encode)):
        rec = bytearray(enc(q))
        assert rec[0] & 0x80 and not rec[0] & 0x40
        y = q[1]
        largest = o.fq_lex_largest(y) if grp == 1 else o.fq2_lex_largest(y)
        assert bool(rec[0] & 0x20) == largest
        st, out = impls(grp, 2, rec, 1, LOAD)
        assert st == OK and out == unc(q)
        flipped = bytearray(rec)
        flipped[0] ^= 0x20  # the other root: -q, a valid point
        neg = (q[0], (o.P - y) % o.P) if grp == 1 else (q[0], o.fq2_neg(y))
        st, out = impls(grp, 2, flipped, 1, LOAD)
        assert st == OK and out == unc(neg)
        nocomp = bytearray(rec)
        nocomp[0] &= 0x7F
        assert impls(grp, 2, nocomp, 1, LOAD)[0] == FLAGS
        inf = bytearray(SZ[grp][2])
        inf[0] = 0xC0
        assert impls(grp, 2, inf, 3, LOAD)[0] == OK          # a legal encoding of the point at infinity ...
        assert impls(grp, 2, inf, 3, STRICT)[0] == INFINITY  # ... which the read_g1 / read_g2 stage refuses (a1/a2)
        inf[0] = 0xE0
        assert impls(grp, 2, inf, 3, LOAD)[0] == FLAGS
    # Fq2 order: c1 decides, c0 only breaks the tie c1 == 0
    big, small = (o.P - 1) // 2 + 1, 5
    assert o.fq2_lex_largest((small, big)) and not o.fq2_lex_largest((big, small))
    assert o.fq2_lex_largest((big, 0)) and not o.fq2_lex_largest((small, 0))


def test_a3_decompression_does_not_check_the_subgroup(impls):
    """[dagger 8a-3] Accumulator::deserialize(Compressed, CheckForCorrectness::No) = into_affine_unchecked: a curve
    point outside the r-torsion decompresses without error; only the read_g1 stage rejects it.
    rust/tests/golden.rs::edge_cases_match_real_primitives (in_fmt 2, `nocheck`)."""
    x = 1
    while True:
        y = o.fq_sqrt((x * x * x + 4) % o.P)
        if y is not None and not o.g1_in_subgroup_rmul((x, y)):
            break
        x += 1
    rec = o.zcash_g1_compressed_encode((x, y))
    assert impls(1, 2, rec, 1, LOAD)[0] == OK
    assert impls(1, 2, rec, 3, READ)[0] == NOT_IN_SUBGROUP


def test_a1_a2_read_rejects_zcash_infinity_and_flag_bits(impls):
    """[dagger 8a-1/2, 8a-7] read_g1 / read_g2 reverse the bytes and call ark deserialize_uncompressed: x is read with
    EmptyFlags, so the zcash infinity bit (bit 6 of byte 0) and any other flag bit of byte 0 make x >= p
    -> Err (NON_CANONICAL); the reference then panics.  rust/tests/golden.rs::edge_cases_match_real_primitives (`read`)."""
    for grp, q, unc in ((1, _g1(9), o.zcash_g1_uncompressed_encode), (2, _g2(9), o.zcash_g2_uncompressed_encode)):
        good = unc(q)
        assert impls(grp, 1, good, 3, READ)[0] == OK
        inf = bytearray(len(good))
        inf[0] = 0x40
        assert impls(grp, 1, inf, 3, READ)[0] == NON_CANONICAL
        for bit in (0x80, 0x40, 0x20):
            bad = bytearray(good)
            bad[0] |= bit
            assert impls(grp, 1, bad, 3, READ)[0] == NON_CANONICAL


def test_a6_unchecked_flag_semantics(impls):
    """[dagger 8a-6] ark-serialize 0.2 deserialize_unchecked reads the UNCOMPRESSED form: x with EmptyFlags (must be
    < p), y with SWFlags in the top two bits of its last byte: (1,0) PositiveY accepted and stripped, (1,1) an error,
    (0,1) infinity with x, y kept as read.  rust/tests/golden.rs::unchecked_loads_match_golden_limbs,
    ::edge_cases_match_real_primitives (in_fmt 3)."""
    for grp, q, ser, rec_fn in ((1, _g1(11), o.ark_g1_serialize_uncompressed, o.g1_mont_record),
                                (2, _g2(11), o.ark_g2_serialize_uncompressed, o.g2_mont_record)):
        good = ser(q)
        st, limbs = impls(grp, 3, good, 4, LOAD)
        assert st == OK and limbs == rec_fn(q[0], q[1], False)
        pos = bytearray(good)
        pos[-1] |= 0x80
        st, limbs2 = impls(grp, 3, pos, 4, LOAD)
        assert st == OK and limbs2 == limbs                     # flag stripped, same point
        both = bytearray(good)
        both[-1] |= 0xC0
        assert impls(grp, 3, both, 4, LOAD)[0] == FLAGS
        inf = bytearray(good)
        inf[-1] |= 0x40
        st, limbs3 = impls(grp, 3, inf, 4, LOAD)
        assert st == OK and limbs3 == rec_fn(q[0], q[1], True)   # x, y kept as read
        xbad = bytearray(good)
        xbad[47] |= 0x80                                          # a "flag" bit on x: just a value >= p
        assert impls(grp, 3, xbad, 4, LOAD)[0] == NON_CANONICAL
        xp = bytearray(good)
        xp[0:48] = o.P.to_bytes(48, "little")
        assert impls(grp, 3, xp, 4, LOAD)[0] == NON_CANONICAL


def test_a7_no_on_curve_check_in_ark_0_2(impls):
    """[dagger 8a-7] ark-ec 0.2.0 deserialize_uncompressed = unchecked + is_in_correct_subgroup_assuming_on_curve
    (multiplication by r, formulas that never use b): an r-torsion point of an ISOMORPHIC curve is accepted, a point
    with y + 1 is rejected as "not in subgroup", never as "not on curve".  PTAU_CHECKS_READ reproduces that;
    PTAU_CHECKS_STRICT adds the curve equation.  rust/tests/golden.rs::edge_cases_match_real_primitives (`read`)."""
    # y^2 = x^3 + 4 u^6 is isomorphic to E by (x, y) -> (u^2 x, u^3 y); u = 7
    u = 7
    q = _g1(13)
    iso = (q[0] * u * u % o.P, q[1] * u ** 3 % o.P)
    assert not o.g1_on_curve(iso)
    rec = o.zcash_g1_uncompressed_encode(iso)
    assert impls(1, 1, rec, 3, READ)[0] == OK
    assert impls(1, 1, rec, 3, STRICT)[0] == NOT_ON_CURVE
    off = (q[0], (q[1] + 1) % o.P)
    rec = o.zcash_g1_uncompressed_encode(off)
    assert impls(1, 1, rec, 3, READ)[0] == NOT_IN_SUBGROUP
    assert impls(1, 1, rec, 3, STRICT)[0] == NOT_ON_CURVE
    # G2: the twist y^2 = x^3 + b' w^6, w = 5 + 3u
    w = (5, 3)
    w2 = o.fq2_sqr(w)
    w3 = o.fq2_mul(w2, w)
    q2 = _g2(13)
    iso2 = (o.fq2_mul(q2[0], w2), o.fq2_mul(q2[1], w3))
    assert not o.g2_on_curve(iso2)
    rec = o.zcash_g2_uncompressed_encode(iso2)
    assert impls(2, 1, rec, 3, READ)[0] == OK
    assert impls(2, 1, rec, 3, STRICT)[0] == NOT_ON_CURVE


def test_a7_subgroup_predicate_equals_multiplication_by_r(impls):
    """[dagger 8a-7] is_in_correct_subgroup_assuming_on_curve is the plain multiplication by r.  The GPU uses
    phi(P) = -[z^2]P (G1) and psi(P) = [z]P (G2); they must give the same boolean on every curve point, in particular
    on small-order points of the cofactor groups.  (Restated arithmetic only: no crate behaviour involved beyond
    "multiply by r".)"""
    checked = 0
    x = 2
    while checked < 12:
        y = o.fq_sqrt((x * x * x + 4) % o.P)
        x += 1
        if y is None:
            continue
        q = (x - 1, y)
        want = o.g1_in_subgroup_rmul(q)
        assert o.g1_in_subgroup_glv(q) == want
        rec = o.zcash_g1_uncompressed_encode(q)
        assert (impls(1, 1, rec, 3, READ)[0] == OK) == want
        checked += 1
    # cofactor-cleared point: in the subgroup by construction
    h1 = 0x396C8C005555E1568C00AAAB0000AAAB
    q = o.g1_mul((4, o.fq_sqrt((4 ** 3 + 4) % o.P)), h1) if o.fq_sqrt((4 ** 3 + 4) % o.P) else None
    if q is not None:
        assert o.g1_in_subgroup_rmul(q) and impls(1, 1, o.zcash_g1_uncompressed_encode(q), 3, READ)[0] == OK


def test_a5_serialize_uncompressed_layout(impls):
    """[dagger 8a-5] GroupAffine::serialize_uncompressed: x LE | y LE (G2: x.c0 | x.c1 | y.c0 | y.c1), infinity flag
    = bit 6 of the last byte, bit 7 never set in the uncompressed form, zero() = (0, 1, infinity).
    rust/tests/golden.rs::reference_read_and_serialize_reproduce_golden_setups."""
    q = _g1(17)
    st, out = impls(1, 1, o.zcash_g1_uncompressed_encode(q), 3, STRICT)
    assert st == OK and out == q[0].to_bytes(48, "little") + q[1].to_bytes(48, "little") and not out[95] & 0xC0
    q2 = _g2(17)
    st, out = impls(2, 1, o.zcash_g2_uncompressed_encode(q2), 3, STRICT)
    want = b"".join(v.to_bytes(48, "little") for v in (q2[0][0], q2[0][1], q2[1][0], q2[1][1]))
    assert st == OK and out == want and not out[191] & 0xC0
    # infinity through the serialize direction (in-memory record -> file bytes)
    st, out = impls(1, 4, o.g1_mont_record(0, 1, True), 3, LOAD)
    assert st == OK and out == (0).to_bytes(48, "little") + (1).to_bytes(47, "little") + b"\x40"


def test_a9_verifier_key_is_four_uncompressed_points():
    """[dagger 8a-9] ark-poly-commit 0.2.0 VerifierKey (de)serialization = g, gamma_g (G1), h, beta_h (G2) in the
    uncompressed form, prepared_h / prepared_beta_h recomputed on read: a 576-byte tail, so `kzg_setup` of the real
    ceremony is 603,980,256 bytes (src/lib.rs:191-192, preprocess-kgz.rs:177-194).
    rust/tests/golden.rs::unchecked_loads_match_golden_limbs asserts that the real deserializer consumes exactly it."""
    n = 8
    kgz = golden("n8_kzg_setup_kgz.bin")
    assert len(kgz) == (3 * n - 1) * 96 + 576 == o.kgz_size(n)
    assert o.kgz_size(1 << 21) == 603980256
    tail = kgz[-576:]
    unc = golden("n8_powersoftau_uncompressed.bin")
    tau_g1_0 = o.read_g1_bytes(unc[0:96])
    tau_g2 = unc[(2 * n - 1) * 96:]
    alpha_0 = o.read_g1_bytes(unc[(2 * n - 1) * 96 + n * 192:][:96])
    assert tail == tau_g1_0 + alpha_0 + o.read_g2_bytes(tau_g2[0:192]) + o.read_g2_bytes(tau_g2[192:384])


def test_a6_in_memory_limbs_are_montgomery_r_2_384():
    """[dagger 8a-6/8b] ark-ff 0.2 Fp384 keeps v * 2^384 mod p as 6 little-endian u64 limbs; `Fp384::new(BigInteger384)`
    takes those limbs as they are.  The ARK_MONT_LIMBS record is that memory image.
    rust/tests/golden.rs::unchecked_loads_match_golden_limbs compares it with the real in-memory limbs."""
    one = o.fq_to_mont_limbs(1)
    # ark-bls12-381 0.2 publishes R = 2^384 mod p as Fq::R: first limb 0x760900000002fffd
    assert one[:8] == (0x760900000002FFFD).to_bytes(8, "little")
    assert int.from_bytes(one, "little") == (1 << 384) % o.P
    rec = golden("n8_load_kgz_g1.bin")[:104]
    x = int.from_bytes(rec[:48], "little") * pow(1 << 384, -1, o.P) % o.P
    assert x == o.G1_GEN[0] and rec[96] == 0


def test_a4_response_layout_and_sizes():
    """[dagger 8a-4, App. B] powersoftau (heliaxdev fork, rev e3318303) response = 64-byte hash | tau_g1 (2N-1) x 48 |
    tau_g2 N x 96 | alpha_g1 N x 48 | beta_g1 N x 48 | beta_g2 96 | public key 1152; CONTRIBUTION_BYTE_SIZE for 2^21
    powers is 603,981,040 (the value the reference asserts at preprocess-kgz.rs:83 before reading anything)."""
    assert o.response_size(1 << 21) == 64 + (2 * (1 << 21) - 1) * 48 + (1 << 21) * (96 + 48 + 48) + 96 + 1152 == 603981040
    resp = golden("n8_powersoftau.bin")
    assert len(resp) == o.response_size(8)
    first = o.zcash_g1_compressed_decode(resp[64:112])
    assert first == o.G1_GEN                                   # tau^0 G right after the hash
    g2_off = 64 + 15 * 48
    assert o.zcash_g2_compressed_decode(resp[g2_off:g2_off + 96]) == o.G2_GEN
