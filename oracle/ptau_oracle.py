"""CPU oracle (Python big integers) for the Powers-of-Tau -> arkworks KZG path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it, and only as the checker.

PARITY UNPINNED BY THE REFERENCE: the reference (heliaxdev/kzg-setup-powersoftau,
Rust) cannot be built here (no rustc/cargo, crates not vendored, no network) and
its only test needs the network and does not compile (src/lib.rs:252 vs :166).
The three BLAKE2b digests it carries (src/lib.rs:21-22, preprocess-kgz.rs:19)
are over >=604 MB ceremony files that are absent.  Substitute pins used instead
(tests/test_oracle_pins.py): the published BLS12-381 generator encodings (zcash
spec), r*G = O, endomorphism eigenvalue identities, and known-tau ground truth
([tau^i]G computed a second way).

This file restates, function by function:
  * src/lib.rs:41-54   read_g1   (zcash-uncompressed G1 -> byte reversal -> ark
                                  deserialize_uncompressed)
  * src/lib.rs:56-80   read_g2
  * src/lib.rs:82-121  load_phase1
  * src/lib.rs:174-228 load_kzg_setup / load_fastkzg_setup
  * src/bin/preprocess-kgz.rs:69-200, src/bin/preprocess-fastkgz.rs:129-214
and the third-party algorithms those call (not vendored under /root/reference;
pinned in Cargo.lock): pairing 0.14.2 (zcash encodings, Fq sqrt a^((p+1)/4),
Fq2 sqrt Algorithm 9 of eprint 2012/685, lexicographic y sign),
powersoftau@e3318303 (Accumulator layout), ark-serialize/ark-ff/ark-ec 0.2.0
(LE encodings with SWFlags, subgroup check = multiplication by r in Jacobian
coordinates, no on-curve check in deserialize_uncompressed), ark-poly-commit
0.2.0 (Powers / VerifierKey / UniversalParams containers).
"""
from __future__ import annotations

import hashlib
import struct
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------
# constants (BLS12-381)
# --------------------------------------------------------------------------
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Z_ABS = 0xD201000000010000  # z = -Z_ABS
Z = -Z_ABS
B_G1 = 4
B_G2 = (4, 4)  # 4(1+u)

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (
        0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E,
    ),
    (
        0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE,
    ),
)

# GLV: phi(x,y) = (BETA*x, y) acts on G1 as multiplication by -z^2
BETA = 0x5F19672FDF76CE51BA69C6076A0F77EADDB3A93BE6F89688DE17D813620A00022E01FFFFFFFEFFFE

MONT_R = (1 << 384) % P
MONT_R2 = (MONT_R * MONT_R) % P

assert P % 4 == 3
assert R_ORDER == Z**4 - Z**2 + 1
assert P == (Z - 1) ** 2 * R_ORDER // 3 + Z


# --------------------------------------------------------------------------
# Fq, Fq2 = Fq[u]/(u^2+1)
# --------------------------------------------------------------------------
def fq_sqrt(a: int) -> Optional[int]:
    """pairing 0.14.2 fq.rs: sqrt = a^((p+1)/4), valid iff its square is a."""
    s = pow(a, (P + 1) // 4, P)
    return s if (s * s - a) % P == 0 else None


Fq2 = Tuple[int, int]
FQ2_ZERO: Fq2 = (0, 0)
FQ2_ONE: Fq2 = (1, 0)


def fq2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fq2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fq2_neg(a: Fq2) -> Fq2:
    return ((-a[0]) % P, (-a[1]) % P)


def fq2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fq2_sqr(a: Fq2) -> Fq2:
    return fq2_mul(a, a)


def fq2_conj(a: Fq2) -> Fq2:
    return (a[0], (-a[1]) % P)


def fq2_inv(a: Fq2) -> Fq2:
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, -1, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def fq2_pow(a: Fq2, e: int) -> Fq2:
    res = FQ2_ONE
    base = a
    while e:
        if e & 1:
            res = fq2_mul(res, base)
        base = fq2_sqr(base)
        e >>= 1
    return res


def fq2_sqrt_alg9(a: Fq2) -> Optional[Fq2]:
    """pairing 0.14.2 fq2.rs sqrt: Algorithm 9 of eprint 2012/685 (p = 3 mod 4)."""
    if a == FQ2_ZERO:
        return FQ2_ZERO
    a1 = fq2_pow(a, (P - 3) // 4)
    alpha = fq2_mul(fq2_sqr(a1), a)
    a0 = fq2_mul(fq2_conj(alpha), alpha)  # frobenius_map(1) is conjugation
    neg1 = (P - 1, 0)
    if a0 == neg1:
        return None
    a1 = fq2_mul(a1, a)
    if alpha == neg1:
        return fq2_mul(a1, (0, 1))
    alpha = fq2_add(alpha, FQ2_ONE)
    alpha = fq2_pow(alpha, (P - 1) // 2)
    return fq2_mul(a1, alpha)


def fq2_sqrt_complex(a: Fq2) -> Optional[Fq2]:
    """The norm ("complex") method the CUDA kernels use: two Fq exponentiations.

    Returns some square root (sign unspecified) or None.  Both roots are
    equivalent for the path because the zcash y-sign flag picks the root.
    """
    a0, a1 = a
    n = (a0 * a0 + a1 * a1) % P
    s = pow(n, (P + 1) // 4, P)
    if (s * s - n) % P:
        return None
    half = (P + 1) // 2
    d = (a0 + s) * half % P
    if d == 0:
        d = (a0 - s) * half % P
    t = pow(d, (P - 3) // 4, P)
    x0 = d * t % P
    chi = x0 * t % P  # d^((p-1)/2)
    if chi == 1 or d == 0:
        r0, r1 = x0, a1 * t % P * half % P
    else:
        r0, r1 = a1 * t % P * half % P, (-x0) % P
    cand = (r0, r1)
    return cand if fq2_sqr(cand) == (a0 % P, a1 % P) else None


def fq_lex_largest(y: int) -> bool:
    """pairing 0.14.2 ec.rs get_point_from_x: y is 'greatest' iff y > -y."""
    return y > (P - y) % P


def fq2_lex_largest(y: Fq2) -> bool:
    """pairing 0.14.2 fq2.rs Ord: compare c1 first, then c0."""
    ny = fq2_neg(y)
    if y[1] != ny[1]:
        return y[1] > ny[1]
    return y[0] > ny[0]


# --------------------------------------------------------------------------
# generic field adaptor so curve code is written once
# --------------------------------------------------------------------------
class _F1:
    zero = 0
    one = 1
    b = B_G1

    @staticmethod
    def add(a, b):
        return (a + b) % P

    @staticmethod
    def sub(a, b):
        return (a - b) % P

    @staticmethod
    def neg(a):
        return (-a) % P

    @staticmethod
    def mul(a, b):
        return a * b % P

    @staticmethod
    def sqr(a):
        return a * a % P

    @staticmethod
    def inv(a):
        return pow(a, -1, P)

    @staticmethod
    def dbl(a):
        return 2 * a % P


class _F2:
    zero = FQ2_ZERO
    one = FQ2_ONE
    b = B_G2
    add = staticmethod(fq2_add)
    sub = staticmethod(fq2_sub)
    neg = staticmethod(fq2_neg)
    mul = staticmethod(fq2_mul)
    sqr = staticmethod(fq2_sqr)
    inv = staticmethod(fq2_inv)

    @staticmethod
    def dbl(a):
        return fq2_add(a, a)


# --------------------------------------------------------------------------
# curve arithmetic.  Affine points: None = infinity, else (x, y).
# Jacobian (X, Y, Z), Z == zero <=> infinity -- formulas as in ark-ec 0.2.0
# models/short_weierstrass_jacobian.rs (dbl-2009-l, madd-2007-bl, add-2007-bl).
# None of them uses the curve coefficient b, so they define a group law on
# whichever curve y^2 = x^3 + b' the input point happens to lie on.
# --------------------------------------------------------------------------
def jac_double(F, pt):
    X, Y, Zc = pt
    if Zc == F.zero:
        return pt
    A = F.sqr(X)
    Bv = F.sqr(Y)
    C = F.sqr(Bv)
    D = F.dbl(F.sub(F.sub(F.sqr(F.add(X, Bv)), A), C))
    E = F.add(F.dbl(A), A)
    Fv = F.sqr(E)
    Z3 = F.dbl(F.mul(Zc, Y))
    X3 = F.sub(Fv, F.dbl(D))
    Y3 = F.sub(F.mul(F.sub(D, X3), E), F.dbl(F.dbl(F.dbl(C))))
    return (X3, Y3, Z3)


def jac_add_mixed(F, pt, q):
    """pt (Jacobian) + q (affine, not infinity unless None)."""
    if q is None:
        return pt
    X1, Y1, Z1 = pt
    x2, y2 = q
    if Z1 == F.zero:
        return (x2, y2, F.one)
    Z1Z1 = F.sqr(Z1)
    U2 = F.mul(x2, Z1Z1)
    S2 = F.mul(F.mul(y2, Z1), Z1Z1)
    if U2 == X1 and S2 == Y1:
        return jac_double(F, pt)
    H = F.sub(U2, X1)
    HH = F.sqr(H)
    I = F.dbl(F.dbl(HH))
    J = F.mul(H, I)
    rr = F.dbl(F.sub(S2, Y1))
    V = F.mul(X1, I)
    X3 = F.sub(F.sub(F.sqr(rr), J), F.dbl(V))
    Y3 = F.sub(F.mul(rr, F.sub(V, X3)), F.dbl(F.mul(Y1, J)))
    Z3 = F.sub(F.sub(F.sqr(F.add(Z1, H)), Z1Z1), HH)
    return (X3, Y3, Z3)


def jac_add(F, p1, p2):
    X1, Y1, Z1 = p1
    X2, Y2, Z2 = p2
    if Z1 == F.zero:
        return p2
    if Z2 == F.zero:
        return p1
    Z1Z1 = F.sqr(Z1)
    Z2Z2 = F.sqr(Z2)
    U1 = F.mul(X1, Z2Z2)
    U2 = F.mul(X2, Z1Z1)
    S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
    S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
    if U1 == U2 and S1 == S2:
        return jac_double(F, p1)
    H = F.sub(U2, U1)
    I = F.sqr(F.dbl(H))
    J = F.mul(H, I)
    rr = F.dbl(F.sub(S2, S1))
    V = F.mul(U1, I)
    X3 = F.sub(F.sub(F.sqr(rr), J), F.dbl(V))
    Y3 = F.sub(F.mul(rr, F.sub(V, X3)), F.dbl(F.mul(S1, J)))
    Z3 = F.mul(F.sub(F.sub(F.sqr(F.add(Z1, Z2)), Z1Z1), Z2Z2), H)
    return (X3, Y3, Z3)


def jac_to_affine(F, pt):
    X, Y, Zc = pt
    if Zc == F.zero:
        return None
    zi = F.inv(Zc)
    zi2 = F.sqr(zi)
    return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))


def jac_is_zero(F, pt) -> bool:
    return pt[2] == F.zero


def mul_bits_be(F, q, k: int):
    """ark-ec 0.2.0 GroupAffine::mul_bits(BitIteratorBE): res=0; for each bit:
    double, then mixed-add the base if the bit is set.  q affine, k >= 0."""
    res = (F.zero, F.one, F.zero)
    for i in range(k.bit_length() - 1, -1, -1):
        res = jac_double(F, res)
        if (k >> i) & 1:
            res = jac_add_mixed(F, res, q)
    return res


def scalar_mul(F, q, k: int):
    """[k]q as an affine point (None = infinity); k may be negative."""
    if q is None or k == 0:
        return None
    if k < 0:
        q = (q[0], F.neg(q[1]))
        k = -k
    return jac_to_affine(F, mul_bits_be(F, q, k))


def g1_mul(q, k):
    return scalar_mul(_F1, q, k)


def g2_mul(q, k):
    return scalar_mul(_F2, q, k)


def affine_add(F, a, b):
    if a is None:
        return b
    if b is None:
        return a
    return jac_to_affine(F, jac_add_mixed(F, (a[0], a[1], F.one), b))


def on_curve(F, q) -> bool:
    if q is None:
        return True
    x, y = q
    return F.sqr(y) == F.add(F.mul(F.sqr(x), x), F.b)


def g1_on_curve(q):
    return on_curve(_F1, q)


def g2_on_curve(q):
    return on_curve(_F2, q)


# ---- subgroup predicates ---------------------------------------------------
def in_subgroup_rmul(F, q) -> bool:
    """REFERENCE SEMANTICS.  ark-ec 0.2.0 is_in_correct_subgroup_assuming_on_curve:
    self.mul_bits(BitIteratorBE::new(r)).is_zero().  No on-curve check."""
    return jac_is_zero(F, mul_bits_be(F, q, R_ORDER))


def g1_in_subgroup_rmul(q):
    return in_subgroup_rmul(_F1, q)


def g2_in_subgroup_rmul(q):
    return in_subgroup_rmul(_F2, q)


def g1_in_subgroup_glv(q) -> bool:
    """KERNEL SEMANTICS for G1: phi(P) == [-z^2]P, phi(x,y) = (BETA x, y).
    phi^2+phi+1 = 0 on every j=0 curve, so phi(P) = lambda P implies
    (lambda^2+lambda+1) P = r P = O: the test never accepts a point outside the
    r-torsion, and on G1 phi acts as -z^2, so it accepts all of G1."""
    x, y = q
    t = mul_bits_be(_F1, q, Z_ABS)
    ta = jac_to_affine(_F1, t)
    t2 = jac_to_affine(_F1, mul_bits_be(_F1, ta, Z_ABS)) if ta is not None else None
    if t2 is None:
        return False
    # [z^2]P == -phi(P)
    return t2 == (BETA * x % P, (-y) % P)


# psi(x,y) = (conj(x) * PSI_CX, conj(y) * PSI_CY) on the twist
PSI_CX: Fq2 = fq2_inv(fq2_pow((1, 1), (P - 1) // 3))
PSI_CY: Fq2 = fq2_inv(fq2_pow((1, 1), (P - 1) // 2))


def g2_psi(q):
    x, y = q
    return (fq2_mul(fq2_conj(x), PSI_CX), fq2_mul(fq2_conj(y), PSI_CY))


def g2_in_subgroup_psi(q) -> bool:
    """KERNEL SEMANTICS for G2 (on-curve points only): psi(P) == [z]P."""
    x, y = q
    t = jac_to_affine(_F2, mul_bits_be(_F2, q, Z_ABS))
    if t is None:
        return False
    t = (t[0], fq2_neg(t[1]))  # z is negative
    return t == g2_psi(q)


# --------------------------------------------------------------------------
# encodings
# --------------------------------------------------------------------------
def _be48(v: int) -> bytes:
    return v.to_bytes(48, "big")


def _le48(v: int) -> bytes:
    return v.to_bytes(48, "little")


class DecodeError(Exception):
    """Mirrors pairing::GroupDecodingError / ark SerializationError::InvalidData."""

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind


# kinds shared with include/ptau_b200.h (PTAU_BAD_*)
BAD_NON_CANONICAL = "NON_CANONICAL"
BAD_FLAGS = "BAD_FLAGS"
BAD_INFINITY = "INFINITY"
BAD_NOT_ON_CURVE = "NOT_ON_CURVE"
BAD_NOT_IN_SUBGROUP = "NOT_IN_SUBGROUP"


# ---- zcash (pairing 0.14.2 src/bls12_381/ec.rs) ----------------------------
def zcash_g1_uncompressed_encode(q) -> bytes:
    if q is None:
        return bytes([0x40]) + bytes(95)
    return _be48(q[0]) + _be48(q[1])


def zcash_g1_compressed_encode(q) -> bytes:
    if q is None:
        return bytes([0xC0]) + bytes(47)
    b = bytearray(_be48(q[0]))
    b[0] |= 0x80
    if fq_lex_largest(q[1]):
        b[0] |= 0x20
    return bytes(b)


def zcash_g2_uncompressed_encode(q) -> bytes:
    if q is None:
        return bytes([0x40]) + bytes(191)
    (x0, x1), (y0, y1) = q
    return _be48(x1) + _be48(x0) + _be48(y1) + _be48(y0)


def zcash_g2_compressed_encode(q) -> bytes:
    if q is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), y = q
    b = bytearray(_be48(x1) + _be48(x0))
    b[0] |= 0x80
    if fq2_lex_largest(y):
        b[0] |= 0x20
    return bytes(b)


def _zcash_fq(b: bytes) -> int:
    v = int.from_bytes(b, "big")
    if v >= P:
        raise DecodeError(BAD_NON_CANONICAL)
    return v


def zcash_g1_uncompressed_decode(buf: bytes):
    """G1Uncompressed::into_affine_unchecked."""
    assert len(buf) == 96
    b = bytearray(buf)
    if b[0] & 0x80:
        raise DecodeError(BAD_FLAGS)  # UnexpectedCompressionMode
    if b[0] & 0x40:
        if all(v == 0 for v in b[1:]) and b[0] == 0x40:
            return None
        raise DecodeError(BAD_FLAGS)  # UnexpectedInformation
    if b[0] & 0x20:
        raise DecodeError(BAD_FLAGS)  # sort bit on an uncompressed point
    x = _zcash_fq(bytes(b[:48]))
    y = _zcash_fq(bytes(b[48:]))
    return (x, y)


def zcash_g1_compressed_decode(buf: bytes):
    """G1Compressed::into_affine_unchecked: sqrt decompression, on curve by
    construction, NO subgroup check (CheckForCorrectness::No path used at
    preprocess-kgz.rs:105-109)."""
    assert len(buf) == 48
    b = bytearray(buf)
    if not (b[0] & 0x80):
        raise DecodeError(BAD_FLAGS)
    if b[0] & 0x40:
        if all(v == 0 for v in b[1:]) and b[0] == 0xC0:
            return None
        raise DecodeError(BAD_FLAGS)
    greatest = bool(b[0] & 0x20)
    b[0] &= 0x1F
    x = _zcash_fq(bytes(b))
    y = fq_sqrt((x * x % P * x + B_G1) % P)
    if y is None:
        raise DecodeError(BAD_NOT_ON_CURVE)
    if fq_lex_largest(y) != greatest:
        y = (-y) % P
    return (x, y)


def zcash_g2_uncompressed_decode(buf: bytes):
    assert len(buf) == 192
    b = bytearray(buf)
    if b[0] & 0x80:
        raise DecodeError(BAD_FLAGS)
    if b[0] & 0x40:
        if all(v == 0 for v in b[1:]) and b[0] == 0x40:
            return None
        raise DecodeError(BAD_FLAGS)
    if b[0] & 0x20:
        raise DecodeError(BAD_FLAGS)
    x1 = _zcash_fq(bytes(b[0:48]))
    x0 = _zcash_fq(bytes(b[48:96]))
    y1 = _zcash_fq(bytes(b[96:144]))
    y0 = _zcash_fq(bytes(b[144:192]))
    return ((x0, x1), (y0, y1))


def zcash_g2_compressed_decode(buf: bytes, sqrt=fq2_sqrt_alg9):
    assert len(buf) == 96
    b = bytearray(buf)
    if not (b[0] & 0x80):
        raise DecodeError(BAD_FLAGS)
    if b[0] & 0x40:
        if all(v == 0 for v in b[1:]) and b[0] == 0xC0:
            return None
        raise DecodeError(BAD_FLAGS)
    greatest = bool(b[0] & 0x20)
    b[0] &= 0x1F
    x1 = _zcash_fq(bytes(b[0:48]))
    x0 = _zcash_fq(bytes(b[48:96]))
    x = (x0, x1)
    y = sqrt(fq2_add(fq2_mul(fq2_sqr(x), x), B_G2))
    if y is None:
        raise DecodeError(BAD_NOT_ON_CURVE)
    if fq2_lex_largest(y) != greatest:
        y = fq2_neg(y)
    return (x, y)


# ---- arkworks 0.2.0 (ark-serialize flags.rs, ark-ff fields, ark-ec SW) ------
ARK_FLAG_INFINITY = 0x40  # SWFlags::Infinity      -> bit 6 of the last byte
ARK_FLAG_POSITIVE = 0x80  # SWFlags::PositiveY     -> bit 7 of the last byte


def ark_g1_serialize_uncompressed(q) -> bytes:
    """GroupAffine::serialize_uncompressed: x LE, then y LE with SWFlags of
    Infinity or (for finite points) no flag bits.  zero() = (0, 1, inf)."""
    if q is None:
        b = bytearray(_le48(0) + _le48(1))
        b[95] |= ARK_FLAG_INFINITY
        return bytes(b)
    return _le48(q[0]) + _le48(q[1])


def ark_g2_serialize_uncompressed(q) -> bytes:
    if q is None:
        b = bytearray(_le48(0) * 2 + _le48(1) + _le48(0))
        b[191] |= ARK_FLAG_INFINITY
        return bytes(b)
    (x0, x1), (y0, y1) = q
    return _le48(x0) + _le48(x1) + _le48(y0) + _le48(y1)


def _ark_fq_noflags(b: bytes) -> int:
    """Fp384::deserialize (EmptyFlags): 48 bytes LE, value must be < p.
    No flag bits are masked, so any set top bit makes the value >= p."""
    v = int.from_bytes(b, "little")
    if v >= P:
        raise DecodeError(BAD_NON_CANONICAL)
    return v


def _ark_fq_swflags(b: bytes) -> Tuple[int, bool]:
    """Fp384::deserialize_with_flags::<SWFlags>: top two bits of the last byte
    are flags; (1,1) is invalid, (1,0)=PositiveY is accepted and stripped,
    (0,1)=Infinity.  Returns (value, infinity)."""
    bb = bytearray(b)
    fl = bb[47] & 0xC0
    if fl == 0xC0:
        raise DecodeError(BAD_FLAGS)
    bb[47] &= 0x3F
    v = int.from_bytes(bytes(bb), "little")
    if v >= P:
        raise DecodeError(BAD_NON_CANONICAL)
    return v, fl == ARK_FLAG_INFINITY


def ark_g1_deserialize_unchecked(buf: bytes):
    """GroupAffine::deserialize_unchecked (loader primitive, src/lib.rs:180).
    Returns (x, y, infinity) exactly as read -- x,y kept even if infinity."""
    assert len(buf) == 96
    x = _ark_fq_noflags(buf[:48])
    y, inf = _ark_fq_swflags(buf[48:])
    return (x, y, inf)


def ark_g2_deserialize_unchecked(buf: bytes):
    """QuadExtField deserialization: c0 without flags, c1 carrying the flags."""
    assert len(buf) == 192
    x0 = _ark_fq_noflags(buf[0:48])
    x1 = _ark_fq_noflags(buf[48:96])  # x has EmptyFlags: top bits must be 0
    y0 = _ark_fq_noflags(buf[96:144])
    y1, inf = _ark_fq_swflags(buf[144:192])
    return ((x0, x1), (y0, y1), inf)


def ark_g1_deserialize_uncompressed(buf: bytes, strict_on_curve: bool = False):
    """GroupAffine::deserialize_uncompressed = unchecked + r-multiplication
    subgroup check (src/lib.rs:52).  strict_on_curve adds the on-curve test the
    CUDA path performs by default (a stricter superset; identical on every
    on-curve input)."""
    x, y, inf = ark_g1_deserialize_unchecked(buf)
    if inf:
        # mul_bits of the stored (x,y) with infinity=true: ark returns zero for
        # an infinity base -> passes.  Cannot be reached from read_g1 because the
        # zcash infinity bit lands in x (src/lib.rs:49-52) -> NON_CANONICAL.
        return None
    q = (x, y)
    if strict_on_curve and not g1_on_curve(q):
        raise DecodeError(BAD_NOT_ON_CURVE)
    if not g1_in_subgroup_rmul(q):
        raise DecodeError(BAD_NOT_IN_SUBGROUP)
    return q


def ark_g2_deserialize_uncompressed(buf: bytes, strict_on_curve: bool = False):
    x, y, inf = ark_g2_deserialize_unchecked(buf)
    if inf:
        return None
    q = (x, y)
    if strict_on_curve and not g2_on_curve(q):
        raise DecodeError(BAD_NOT_ON_CURVE)
    if not g2_in_subgroup_rmul(q):
        raise DecodeError(BAD_NOT_IN_SUBGROUP)
    return q


# --------------------------------------------------------------------------
# the crate's own functions
# --------------------------------------------------------------------------
def read_g1_bytes(buf: bytes) -> bytes:
    """src/lib.rs:45-50: the byte shuffle only (zcash uncompressed -> ark LE)."""
    assert len(buf) == 96
    return buf[0:48][::-1] + buf[48:96][::-1]


def read_g2_bytes(buf: bytes) -> bytes:
    """src/lib.rs:60-76."""
    assert len(buf) == 192
    return buf[48:96][::-1] + buf[0:48][::-1] + buf[144:192][::-1] + buf[96:144][::-1]


def read_g1(buf: bytes, strict_on_curve: bool = False):
    """src/lib.rs:41-54."""
    return ark_g1_deserialize_uncompressed(read_g1_bytes(buf), strict_on_curve)


def read_g2(buf: bytes, strict_on_curve: bool = False):
    """src/lib.rs:56-80."""
    return ark_g2_deserialize_uncompressed(read_g2_bytes(buf), strict_on_curve)


# ---- Montgomery limbs (what ARK_MONT_LIMBS output carries) ------------------
def fq_to_mont_limbs(v: int) -> bytes:
    """ark-ff Fp384 in-memory form: 6 x u64 little-endian limbs of v*R mod p."""
    return (v * MONT_R % P).to_bytes(48, "little")


def g1_mont_record(x: int, y: int, inf: bool) -> bytes:
    """104-byte record: x mont(48) | y mont(48) | infinity u8 | 7 pad."""
    return fq_to_mont_limbs(x) + fq_to_mont_limbs(y) + bytes([1 if inf else 0]) + bytes(7)


def g2_mont_record(x: Fq2, y: Fq2, inf: bool) -> bytes:
    """200-byte record: x.c0 | x.c1 | y.c0 | y.c1 (mont) | infinity u8 | 7 pad."""
    return (
        fq_to_mont_limbs(x[0])
        + fq_to_mont_limbs(x[1])
        + fq_to_mont_limbs(y[0])
        + fq_to_mont_limbs(y[1])
        + bytes([1 if inf else 0])
        + bytes(7)
    )


# --------------------------------------------------------------------------
# file layouts
# --------------------------------------------------------------------------
HASH_SIZE = 64
PUBKEY_SIZE = 3 * 192 + 6 * 96  # powersoftau PUBLIC_KEY_SIZE (uncompressed)


def response_size(n: int) -> int:
    """powersoftau CONTRIBUTION_BYTE_SIZE for n = TAU_POWERS_LENGTH."""
    return HASH_SIZE + (2 * n - 1) * 48 + n * 96 + n * 48 + n * 48 + 96 + PUBKEY_SIZE


def uncompressed_size(n: int) -> int:
    return (2 * n - 1) * 96 + n * 192 + n * 96 + n * 96 + 192


def kgz_size(n: int) -> int:
    return (2 * n - 1) * 96 + n * 96 + 96 + 96 + 192 + 192


def fastkgz_size(n: int) -> int:
    return (2 * n - 1) * 96 + n * 96 + 192 + 192 + n * 192


def splitmix64(state: int) -> Tuple[int, int]:
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    zz = state
    zz = ((zz ^ (zz >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    zz = ((zz ^ (zz >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, zz ^ (zz >> 31)


def derive_scalars(seed: int) -> Tuple[int, int, int]:
    """seed -> (tau, alpha, beta) in Fr \\ {0}; same derivation as oracle/cpu_ref.c
    and the GPU generator: 8 SplitMix64 words per scalar, big-endian concatenated,
    reduced mod r, 0 mapped to 1."""
    st = seed & 0xFFFFFFFFFFFFFFFF
    out = []
    for _ in range(3):
        v = 0
        for _ in range(8):
            st, w = splitmix64(st)
            v = (v << 64) | w
        v %= R_ORDER
        out.append(v or 1)
    return out[0], out[1], out[2]


class Accumulator:
    """powersoftau::Accumulator fields in serialization order."""

    def __init__(self, tau_g1, tau_g2, alpha_g1, beta_g1, beta_g2):
        self.tau_g1 = tau_g1
        self.tau_g2 = tau_g2
        self.alpha_g1 = alpha_g1
        self.beta_g1 = beta_g1
        self.beta_g2 = beta_g2


def make_accumulator(n: int, seed: int) -> Accumulator:
    tau, alpha, beta = derive_scalars(seed)
    tau_g1, tau_g2, alpha_g1, beta_g1 = [], [], [], []
    t = 1
    for i in range(2 * n - 1):
        tau_g1.append(g1_mul(G1_GEN, t))
        if i < n:
            tau_g2.append(g2_mul(G2_GEN, t))
            alpha_g1.append(g1_mul(G1_GEN, alpha * t % R_ORDER))
            beta_g1.append(g1_mul(G1_GEN, beta * t % R_ORDER))
        t = t * tau % R_ORDER
    return Accumulator(tau_g1, tau_g2, alpha_g1, beta_g1, g2_mul(G2_GEN, beta))


def filler_bytes(seed: int, count: int, tag: bytes) -> bytes:
    """Deterministic bytes for the 64-byte challenge hash and the 1152-byte
    public key, neither of which the reference reads (preprocess-kgz.rs:96-110)."""
    out = b""
    ctr = 0
    while len(out) < count:
        out += hashlib.blake2b(tag + struct.pack("<QQ", seed & (2**64 - 1), ctr)).digest()
        ctr += 1
    return out[:count]


def serialize_response(acc: Accumulator, seed: int = 0) -> bytes:
    """`powersoftau` response file: 64-byte hash | compressed accumulator | pubkey."""
    parts = [filler_bytes(seed, HASH_SIZE, b"hash")]
    parts += [zcash_g1_compressed_encode(q) for q in acc.tau_g1]
    parts += [zcash_g2_compressed_encode(q) for q in acc.tau_g2]
    parts += [zcash_g1_compressed_encode(q) for q in acc.alpha_g1]
    parts += [zcash_g1_compressed_encode(q) for q in acc.beta_g1]
    parts.append(zcash_g2_compressed_encode(acc.beta_g2))
    parts.append(filler_bytes(seed, PUBKEY_SIZE, b"pubkey"))
    return b"".join(parts)


def powersoftau_uncompress(response: bytes, n: int) -> bytes:
    """preprocess-kgz.rs:69-126: size check, skip 64-byte hash,
    Accumulator::deserialize(Compressed, CheckForCorrectness::No), then
    Accumulator::serialize(Uncompressed).  The trailing public key is unread."""
    if len(response) != response_size(n):
        raise ValueError(
            "The size of `powersoftau` should be %d, but it's %d, so something isn't right."
            % (response_size(n), len(response))
        )
    off = HASH_SIZE
    out = []

    def g1s(cnt):
        nonlocal off
        for _ in range(cnt):
            out.append(zcash_g1_uncompressed_encode(zcash_g1_compressed_decode(response[off : off + 48])))
            off += 48

    def g2s(cnt):
        nonlocal off
        for _ in range(cnt):
            out.append(zcash_g2_uncompressed_encode(zcash_g2_compressed_decode(response[off : off + 96])))
            off += 96

    g1s(2 * n - 1)
    g2s(n)
    g1s(n)
    g1s(n)
    g2s(1)
    return b"".join(out)


def preprocess_kgz(uncompressed: bytes, n: int, strict_on_curve: bool = False) -> bytes:
    """preprocess-kgz.rs:128-200 from `powersoftau_uncompressed` to `kzg_setup`."""
    off = 0
    tau_g1 = []
    for _ in range(2 * n - 1):
        tau_g1.append(read_g1(uncompressed[off : off + 96], strict_on_curve))
        off += 96
    tau_g2 = []
    for _ in range(n):
        tau_g2.append(read_g2(uncompressed[off : off + 192], strict_on_curve))
        off += 192
    alpha_g1 = []
    for _ in range(n):
        alpha_g1.append(read_g1(uncompressed[off : off + 96], strict_on_curve))
        off += 96
    out = [ark_g1_serialize_uncompressed(q) for q in tau_g1]
    out += [ark_g1_serialize_uncompressed(q) for q in alpha_g1]
    # VerifierKey::serialize_uncompressed: g, gamma_g, h, beta_h
    out.append(ark_g1_serialize_uncompressed(tau_g1[0]))
    out.append(ark_g1_serialize_uncompressed(alpha_g1[0]))
    out.append(ark_g2_serialize_uncompressed(tau_g2[0]))
    out.append(ark_g2_serialize_uncompressed(tau_g2[1]))
    return b"".join(out)


def preprocess_fastkgz(uncompressed: bytes, n: int, strict_on_curve: bool = False) -> bytes:
    """preprocess-fastkgz.rs:129-214.  beta_tau_powers_g1 is read, checked and
    dropped (:156-159)."""
    off = 0
    tau_g1 = []
    for _ in range(2 * n - 1):
        tau_g1.append(read_g1(uncompressed[off : off + 96], strict_on_curve))
        off += 96
    tau_g2 = []
    for _ in range(n):
        tau_g2.append(read_g2(uncompressed[off : off + 192], strict_on_curve))
        off += 192
    alpha_g1 = []
    for _ in range(n):
        alpha_g1.append(read_g1(uncompressed[off : off + 96], strict_on_curve))
        off += 96
    for _ in range(n):
        read_g1(uncompressed[off : off + 96], strict_on_curve)
        off += 96
    out = [ark_g1_serialize_uncompressed(q) for q in tau_g1]
    out += [ark_g1_serialize_uncompressed(q) for q in alpha_g1]
    out.append(ark_g2_serialize_uncompressed(tau_g2[0]))
    out.append(ark_g2_serialize_uncompressed(tau_g2[1]))
    out += [ark_g2_serialize_uncompressed(q) for q in tau_g2]
    return b"".join(out)


def load_kzg_setup(data: bytes, n: int):
    """src/lib.rs:174-195.  Returns (powers_of_g, powers_of_gamma_g, vk) where
    every point is the (x, y, infinity) triple deserialize_unchecked yields and
    vk = (g, gamma_g, h, beta_h)."""
    off = 0
    pg = []
    for _ in range(2 * n - 1):
        pg.append(ark_g1_deserialize_unchecked(data[off : off + 96]))
        off += 96
    pgg = []
    for _ in range(n):
        pgg.append(ark_g1_deserialize_unchecked(data[off : off + 96]))
        off += 96
    g = ark_g1_deserialize_unchecked(data[off : off + 96])
    off += 96
    gamma_g = ark_g1_deserialize_unchecked(data[off : off + 96])
    off += 96
    h = ark_g2_deserialize_unchecked(data[off : off + 192])
    off += 192
    beta_h = ark_g2_deserialize_unchecked(data[off : off + 192])
    return pg, pgg, (g, gamma_g, h, beta_h)


def load_fastkzg_setup(data: bytes, n: int):
    """src/lib.rs:197-228.  Returns (powers_of_g, powers_of_gamma_g (index
    order), h, beta_h_from_powers, beta_h_from_file, powers_of_h)."""
    off = 0
    pg = []
    for _ in range(2 * n - 1):
        pg.append(ark_g1_deserialize_unchecked(data[off : off + 96]))
        off += 96
    pgg = []
    for _ in range(n):
        pgg.append(ark_g1_deserialize_unchecked(data[off : off + 96]))
        off += 96
    h = ark_g2_deserialize_unchecked(data[off : off + 192])
    off += 192
    beta_h_file = ark_g2_deserialize_unchecked(data[off : off + 192])
    off += 192
    ph = []
    for _ in range(n):
        ph.append(ark_g2_deserialize_unchecked(data[off : off + 192]))
        off += 192
    return pg, pgg, h, ph[1], beta_h_file, ph


def load_phase1(data: bytes, m: int, strict_on_curve: bool = False):
    """src/lib.rs:82-121: alpha, beta_g1, beta_g2, m G1, m G2, m G1, m G1."""
    off = 0

    def g1():
        nonlocal off
        q = read_g1(data[off : off + 96], strict_on_curve)
        off += 96
        return q

    def g2():
        nonlocal off
        q = read_g2(data[off : off + 192], strict_on_curve)
        off += 192
        return q

    alpha = g1()
    beta_g1 = g1()
    beta_g2 = g2()
    coeffs_g1 = [g1() for _ in range(m)]
    coeffs_g2 = [g2() for _ in range(m)]
    alpha_coeffs_g1 = [g1() for _ in range(m)]
    beta_coeffs_g1 = [g1() for _ in range(m)]
    return alpha, beta_g1, beta_g2, coeffs_g1, coeffs_g2, alpha_coeffs_g1, beta_coeffs_g1


def blake2b_hex(data: bytes) -> str:
    """blake2b_simd::State::new().update(data).finalize().to_hex()
    (src/lib.rs:128-131, preprocess-kgz.rs:33-36): unkeyed BLAKE2b-512."""
    return hashlib.blake2b(data).hexdigest()


# --------------------------------------------------------------------------
# KZG10 sanity check with the known tau (config 4; no pairing needed):
# commit(p) = sum p_i [tau^i]G  must equal [p(tau)]G.
# --------------------------------------------------------------------------
def kzg_commit(powers_of_g: Sequence, coeffs: Sequence[int]):
    acc = None
    for c, g in zip(coeffs, powers_of_g):
        acc = affine_add(_F1, acc, g1_mul(g, c % R_ORDER))
    return acc


def kzg_open(powers_of_g: Sequence, coeffs: Sequence[int], zpt: int):
    """witness polynomial (p(X) - p(z)) / (X - z) committed; returns (value, proof)."""
    n = len(coeffs)
    q = [0] * (n - 1)
    carry = 0
    for i in range(n - 1, 0, -1):
        carry = (coeffs[i] + carry * zpt) % R_ORDER
        q[i - 1] = carry
    value = (coeffs[0] + carry * zpt) % R_ORDER
    return value, kzg_commit(powers_of_g, q)
