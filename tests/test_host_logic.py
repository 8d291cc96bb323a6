"""CPU-only checks of the product's host side: the limb-level arithmetic the CUDA
kernels execute (compiled for the host with an emulated carry flag), the C ABI
surface, and the N>1 sharding logic (gloo, world_size 2)."""
import csv
import ctypes
import json
import os
import random
import re
import subprocess
import sys

import pytest

import ptau_oracle as o
from conftest import GOLDEN, ROOT, golden

SZ = {1: {1: 96, 2: 48, 3: 96, 4: 104}, 2: {1: 192, 2: 96, 3: 192, 4: 200}}


def _conv(lib, group, in_fmt, data, out_fmt, checks):
    n = len(data) // SZ[group][in_fmt]
    out = (ctypes.c_uint8 * (n * SZ[group][out_fmt]))()
    st = (ctypes.c_uint32 * max(n, 1))()
    assert lib.hostemul_convert(group, in_fmt, bytes(data), out_fmt, out, ctypes.c_size_t(n), checks, st) == 0
    return bytes(out), list(st)[:n]


def test_limb_field_ops_match_bigints(hostemul):
    P = o.P
    rinv = pow(o.MONT_R, -1, P)

    def limbs(v):
        return (ctypes.c_uint32 * 12)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)])

    def op(k, a, b):
        out = (ctypes.c_uint32 * 12)()
        hostemul.hostemul_fq_op(k, limbs(a), limbs(b), out)
        return sum(int(out[i]) << (32 * i) for i in range(12))

    rnd = random.Random(7)
    edge = [0, 1, 2, P - 1, P - 2, 1 << 380, o.MONT_R, (1 << 352) - 1, 0xFFFFFFFF, (P - 1) // 2, (P + 1) // 2]
    cases = [(a, b) for a in edge for b in edge] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(1500)]
    for a, b in cases:
        assert op(0, a, b) == a * b * rinv % P
        assert op(1, a, b) == (a + b) % P
        assert op(2, a, b) == (a - b) % P
        assert op(3, a, b) == (-a) % P
        assert op(4, a, b) == a * a * rinv % P
    for _ in range(3):
        a = rnd.randrange(P)
        assert op(5, a * o.MONT_R % P, 0) == pow(a, (P - 3) // 4, P) * o.MONT_R % P


def test_limb_field_ops_carry_stress(hostemul):
    """Values built from extreme 32-bit limbs (0, 1, 2^31, 2^32-1, ...) exercise every
    carry-propagation path of the even/odd multiplier rows and of the dedicated squaring."""
    P = o.P
    rinv = pow(o.MONT_R, -1, P)

    def limbs(v):
        return (ctypes.c_uint32 * 12)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)])

    def op(k, a, b):
        out = (ctypes.c_uint32 * 12)()
        hostemul.hostemul_fq_op(k, limbs(a), limbs(b), out)
        return sum(int(out[i]) << (32 * i) for i in range(12))

    rnd = random.Random(99)
    pats = [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF, 0xFFFF0000, 0x0000FFFF, 0xAAAAAAAA, 0x55555555]
    vals = []
    for _ in range(400):
        v = 0
        for i in range(12):
            v |= rnd.choice(pats) << (32 * i)
        v &= (1 << 381) - 1
        vals.append(v % P)
    # neighbours of multiples of p-structure and of limb boundaries
    for k in range(1, 12):
        vals += [(1 << (32 * k)) - 1, 1 << (32 * k), (1 << (32 * k)) + 1, P - (1 << (32 * k)), P - (1 << (32 * k)) - 1]
    vals = [v % P for v in vals]
    for i, a in enumerate(vals):
        b = vals[(7 * i + 3) % len(vals)]
        assert op(0, a, b) == a * b * rinv % P
        assert op(4, a, 0) == a * a * rinv % P
        assert op(1, a, b) == (a + b) % P
        assert op(2, a, b) == (a - b) % P


def test_limb_wide_products_and_lazy_fq2(hostemul):
    """csrc/fqw.cuh on the host: unreduced products (operands up to 2p, as the Karatsuba sums are), the
    Montgomery reduction of a 24-limb value (every T < p 2^384, incl. the extremes), the sign fix of a difference of
    products, and fq2_mul / fq2_sqr built from them, against big-int arithmetic."""
    P = o.P
    R = 1 << 384
    rinv = pow(R, -1, P)

    def limbs(v, n=12):
        return (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])

    def op(k, a, b, na, nb, nout):
        out = (ctypes.c_uint32 * nout)()
        hostemul.hostemul_fqw_op(k, limbs(a, na), limbs(b, nb), out)
        return sum(int(out[i]) << (32 * i) for i in range(nout))

    rnd = random.Random(2024)
    pats = [0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF, 0xFFFF0000, 0x0000FFFF]

    def patterned(bound):
        v = 0
        for i in range(12):
            v |= rnd.choice(pats) << (32 * i)
        return v % bound

    ops = [0, 1, 2, P - 1, P, P + 1, 2 * P - 2, 2 * P - 1, (1 << 381) - 1, (1 << 382) - 1]
    ops += [patterned(2 * P) for _ in range(150)] + [rnd.randrange(2 * P) for _ in range(300)]
    # full-width operands too: the Karatsuba half sums carry out exactly when the halves are large
    ops += [(1 << 384) - 1, (1 << 384) - (1 << 192), (1 << 192) - 1, ((1 << 192) - 1) << 192, (1 << 383) + (1 << 191)]
    ops += [patterned(1 << 384) for _ in range(150)] + [rnd.randrange(1 << 384) for _ in range(300)]
    for i, a in enumerate(ops):
        b = ops[(5 * i + 1) % len(ops)]
        assert op(0, a, b, 12, 12, 24) == a * b
        assert op(6, a, b, 12, 12, 24) == a * b
        assert op(7, a, b, 12, 12, 24) == a * b
        if a < (1 << 383):
            assert op(1, a, 0, 12, 12, 24) == a * a
    # reduction: extremes of the admissible range and random values
    ts = [0, 1, R - 1, R, P * R - 1, P * R - P, (P - 1) * (P - 1), 4 * P * P - 1 if 4 * P * P < P * R else P * R - 2]
    ts += [rnd.randrange(P * R) for _ in range(400)]
    for k in range(24):
        ts += [(0xFFFFFFFF << (32 * k)) % (P * R), ((1 << (32 * k)) - 1) % (P * R)]
    for t in ts:
        assert op(2, t, 0, 24, 12, 12) == t * rinv % P, hex(t)
    # a - b (+ p R when negative) over products
    for _ in range(200):
        a, b = rnd.randrange(P) * rnd.randrange(P), rnd.randrange(P) * rnd.randrange(P)
        want = a - b if a >= b else a - b + P * R
        assert op(5, a, b, 24, 24, 24) == want
    assert op(5, 0, (P - 1) ** 2, 24, 24, 24) == P * R - (P - 1) ** 2
    # the two conditional-free helpers: a - b + p for reduced a, b (never negative, below 2p) and the halving
    # (a + (a odd ? p : 0)) >> 1, which is division by 2 in Montgomery form as well
    half = pow(2, -1, P)
    singles = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2] + [patterned(P) for _ in range(100)] + \
              [rnd.randrange(P) for _ in range(200)]
    for i, a in enumerate(singles):
        b = singles[(7 * i + 3) % len(singles)]
        assert op(8, a, b, 12, 12, 12) == a - b + P
        assert op(9, a, 0, 12, 12, 12) == a * half % P
    # Fq2: values in Montgomery form, c0 | c1
    def f2(v):
        return v[0] | (v[1] << 384)

    def unf2(x):
        return (x & ((1 << 384) - 1), x >> 384)

    edge = [0, 1, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, R % P, (1 << 380)]
    vals = [(a, b) for a in edge for b in edge] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(400)]
    vals += [(patterned(P), patterned(P)) for _ in range(200)]
    for i, a in enumerate(vals):
        b = vals[(3 * i + 7) % len(vals)]
        got = unf2(op(3, f2(a), f2(b), 24, 24, 24))
        assert got == ((a[0] * b[0] - a[1] * b[1]) * rinv % P, (a[0] * b[1] + a[1] * b[0]) * rinv % P)
        got = unf2(op(4, f2(a), 0, 24, 24, 24))
        assert got == ((a[0] * a[0] - a[1] * a[1]) * rinv % P, 2 * a[0] * a[1] * rinv % P)


def test_limb_g2_ladder_doubling_with_unreduced_c(hostemul):
    """csrc/curve.cuh jac_dbl_lazyc (the doubling of the G2 subgroup ladder: C = Y^4 kept as two unreduced 768-bit
    products, coordinates scaled by 1/2) against dbl-2009-l as the rest of the code runs it: for ANY field values
    X, Y, Z (not only curve points) the outputs must satisfy X' = X3/4, Y' = Y3/8, Z' = Z3/2 exactly and be reduced.
    Also the helper a - b + 2p on unreduced sums (operands below 2p)."""
    P = o.P
    R = 1 << 384
    rinv = pow(R, -1, P)

    def limbs(vals):
        w = []
        for v in vals:
            w += [(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
        return (ctypes.c_uint32 * len(w))(*w)

    def dbl(variant, vals):
        out = (ctypes.c_uint32 * 72)()
        hostemul.hostemul_g2_dbl(variant, limbs(vals), out)
        v = [sum(int(out[12 * k + i]) << (32 * i) for i in range(12)) for k in range(6)]
        assert all(x < P for x in v), "unreduced coordinate"
        return [x * rinv % P for x in v]

    rnd = random.Random(5)
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, R % P, 1 << 380]
    pats = [0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF]

    def patterned():
        v = 0
        for i in range(12):
            v |= rnd.choice(pats) << (32 * i)
        return v % P

    cases = [[rnd.choice(edge) for _ in range(6)] for _ in range(250)]
    cases += [[patterned() for _ in range(6)] for _ in range(250)]
    cases += [[rnd.randrange(P) for _ in range(6)] for _ in range(700)]
    for vals in cases:
        a, b = dbl(0, vals), dbl(1, vals)
        for k, scale in ((0, 4), (1, 4), (2, 8), (3, 8), (4, 2), (5, 2)):
            assert b[k] * scale % P == a[k], (k, vals)
    sums = [0, 1, P - 1, P, P + 1, 2 * P - 2, 2 * P - 1] + [rnd.randrange(2 * P) for _ in range(300)]
    for i, a in enumerate(sums):
        b = sums[(5 * i + 2) % len(sums)]
        out = (ctypes.c_uint32 * 12)()
        hostemul.hostemul_fq_sub_plus_2p(limbs([a]), limbs([b]), out)
        assert sum(int(out[j]) << (32 * j) for j in range(12)) == a - b + 2 * P


def test_limb_code_on_golden_files(hostemul):
    n = 8
    body = golden("n8_powersoftau.bin")[64:]
    unc = golden("n8_powersoftau_uncompressed.bin")
    kgz = golden("n8_kzg_setup_kgz.bin")
    # fused compressed -> ark of tau_g1, and decompress-only -> zcash uncompressed
    cnt = 2 * n - 1
    out, st = _conv(hostemul, 1, 2, body[:cnt * 48], 3, 14)
    assert not any(st) and out == kgz[:cnt * 96]
    out, st = _conv(hostemul, 1, 2, body[:cnt * 48], 1, 0)
    assert not any(st) and out == unc[:cnt * 96]
    g2c = body[cnt * 48:cnt * 48 + n * 96]
    out, st = _conv(hostemul, 2, 2, g2c, 1, 0)
    assert not any(st) and out == unc[cnt * 96:cnt * 96 + n * 192]
    out, st = _conv(hostemul, 2, 1, out, 3, 14)
    assert not any(st) and out == golden("n8_kzg_setup_fastkgz.bin")[(3 * n - 1) * 96 + 384:]
    out, st = _conv(hostemul, 1, 3, kgz[:(3 * n - 1) * 96], 4, 0)
    assert not any(st) and out == golden("n8_load_kgz_g1.bin")[:(3 * n - 1) * 104]
    out, st = _conv(hostemul, 2, 3, kgz[-384:], 4, 14)
    assert not any(st) and out == golden("n8_load_kgz_g2.bin")


def test_limb_code_on_edge_cases(hostemul):
    meta = json.load(open(os.path.join(GOLDEN, "edge_cases.json")))
    for cs in meta["cases"]:
        rec = bytes.fromhex(cs["rec"])
        for mode, checks in (("strict", 14), ("nocheck", 0), ("read", 4)):
            if mode not in cs:
                continue
            _, st = _conv(hostemul, cs["group"], cs["in_fmt"], rec, 3, checks)
            assert st[0] == cs[mode], (cs["desc"], mode, st[0])


def test_limb_code_vs_c_oracle_random(hostemul, cref):
    rnd = random.Random(21)
    tau = rnd.randrange(1, o.R_ORDER)
    zc1 = cref.generate(1, 2, 1, tau, 0, 40, 2)
    zc2 = cref.generate(2, 2, 1, tau, 0, 16, 2)
    for g, data in ((1, zc1), (2, zc2)):
        for out_fmt, checks in ((3, 14), (1, 0), (4, 14)):
            got, st = _conv(hostemul, g, 2, data, out_fmt, checks)
            want, st2 = cref.convert(g, 2, data, out_fmt, checks & 4)
            assert got == want and st == st2 == [0] * len(st)


def test_mont_limb_input_round_trip(hostemul, cref):
    """ARK_MONT_LIMBS as *input* (the serialize direction, preprocess-kgz.rs:188-194):
    load -> serialize reproduces the file bytes; non-canonical limbs are rejected."""
    n = 8
    kgz = golden("n8_kzg_setup_kgz.bin")
    g1_ml = golden("n8_load_kgz_g1.bin")
    g2_ml = golden("n8_load_kgz_g2.bin")
    for conv in (lambda g, i, d, of, c: _conv(hostemul, g, i, d, of, c), lambda g, i, d, of, c: cref.convert(g, i, d, of, c)):
        out, st = conv(1, 4, g1_ml, 3, 14)
        assert not any(st) and out == kgz[:(3 * n - 1) * 96] + kgz[-576:-384]
        out, st = conv(2, 4, g2_ml, 3, 14)
        assert not any(st) and out == kgz[-384:]
        out, st = conv(2, 4, g2_ml, 1, 0)
        assert not any(st) and o.read_g2_bytes(out[:192]) == kgz[-384:-192]
        out, st = conv(1, 4, g1_ml[:104], 4, 0)
        assert out == g1_ml[:104]
        bad = bytearray(g1_ml[:208])
        bad[104:152] = o.P.to_bytes(48, "little")  # x limbs = p: not a valid Fp384
        out, st = conv(1, 4, bytes(bad), 3, 0)
        assert list(st) == [0, 1]
        inf = bytearray(g1_ml[:104])
        inf[96] = 1
        out, st = conv(1, 4, bytes(inf), 3, 4)
        assert list(st) == [0] and out[95] & 0x40 and out[:95] == kgz[:95]
        out, st = conv(1, 4, bytes(inf), 3, 14)
        assert list(st) == [3]


def test_differential_fuzz_c_oracle_vs_limb_code(hostemul, cref):
    """Random and semi-valid records through both implementations, every input format and
    check mode: status and (when accepted) output bytes must agree."""
    rnd = random.Random(4242)
    P = o.P
    valid1 = [o.g1_mul(o.G1_GEN, rnd.randrange(1, o.R_ORDER)) for _ in range(6)]
    valid2 = [o.g2_mul(o.G2_GEN, rnd.randrange(1, o.R_ORDER)) for _ in range(4)]

    def mutate(rec):
        b = bytearray(rec)
        k = rnd.randrange(5)
        if k == 0:
            b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        elif k == 1:
            b[0] ^= rnd.choice([0x80, 0x40, 0x20, 0xC0, 0xE0])
        elif k == 2:
            pos = rnd.choice(range(0, len(b), 48))
            b[pos:pos + 48] = rnd.choice([P, P - 1, P + 1, 0, 1, (1 << 381) - 1, (1 << 384) - 1]).to_bytes(48, "big")
        elif k == 3:
            b[-1] ^= rnd.choice([0x80, 0x40, 0xC0])
        else:
            b[48 * rnd.randrange(len(b) // 48)] ^= rnd.choice([0x80, 0x40, 0xC0])
        return bytes(b)

    cases = []
    for q in valid1:
        cases += [(1, 1, o.zcash_g1_uncompressed_encode(q)), (1, 2, o.zcash_g1_compressed_encode(q)),
                  (1, 3, o.ark_g1_serialize_uncompressed(q)), (1, 4, o.g1_mont_record(q[0], q[1], False))]
    for q in valid2:
        cases += [(2, 1, o.zcash_g2_uncompressed_encode(q)), (2, 2, o.zcash_g2_compressed_encode(q)),
                  (2, 3, o.ark_g2_serialize_uncompressed(q)), (2, 4, o.g2_mont_record(q[0], q[1], False))]
    fuzz = []
    for g, f, rec in cases:
        fuzz.append((g, f, rec))
        for _ in range(6):
            fuzz.append((g, f, mutate(rec)))
        fuzz.append((g, f, bytes(rnd.randrange(256) for _ in range(len(rec)))))
    n_bad = 0
    for g, f, rec in fuzz:
        for checks in (0, 4, 14, 8):
            for of in (1, 3, 4):
                a, sa = _conv(hostemul, g, f, rec, of, checks)
                b, sb = cref.convert(g, f, rec, of, checks)
                assert sa == sb, (g, f, checks, rec.hex())
                if sa[0] == 0:
                    assert a == b, (g, f, checks, of, rec.hex())
                else:
                    n_bad += 1
    assert n_bad > 200  # the mutations really produce rejected records


def test_limb_ladders_reject_random_curve_points(hostemul, cref):
    """Random x coordinates with valid compressed flags: about half decompress to a point of the curve (twist), and such
    a point is outside the prime-order subgroup with overwhelming probability.  The product's ladders (GLV test in G1,
    psi test with the unreduced-C doubling in G2) must give the verdict of the reference's multiplication by r
    (oracle/cpu_ref.c) on every one of them, and the same bytes when the checks are off."""
    rnd = random.Random(977)
    for g, size, n in ((1, 48, 60), (2, 96, 40)):
        recs = []
        for _ in range(n):
            b = bytearray(rnd.randrange(o.P).to_bytes(48, "big") if g == 1 else
                          rnd.randrange(o.P).to_bytes(48, "big") + rnd.randrange(o.P).to_bytes(48, "big"))
            b[0] = (b[0] & 0x1F) | 0x80 | rnd.choice([0, 0x20])
            recs.append(bytes(b))
        data = b"".join(recs)
        a, sa = _conv(hostemul, g, 2, data, 3, 14)
        b_, sb = cref.convert(g, 2, data, 3, 14)
        assert sa == sb
        assert sum(1 for x in sa if x == 5) >= n // 4   # PTAU_BAD_NOT_IN_SUBGROUP: the ladders really ran and said no
        assert sum(1 for x in sa if x == 4) >= n // 4     # PTAU_BAD_NOT_ON_CURVE: the others have no square root
        a, sa = _conv(hostemul, g, 2, data, 3, 0)
        b_, sb = cref.convert(g, 2, data, 3, 0)
        assert sa == sb
        rs = len(a) // n
        for i in range(n):
            if sa[i] == 0:
                assert a[i * rs:(i + 1) * rs] == b_[i * rs:(i + 1) * rs]


# ---- C ABI surface ---------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ptau_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptau_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from kzg_setup_powersoftau_b200 import _ffi

    lib = _ffi.lib()  # raises if the .so was not built
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libptau_b200.so does not export %s" % s
    assert sorted(_ffi.EXPORTED_SYMBOLS) == syms


def test_ffi_constants_match_header():
    from kzg_setup_powersoftau_b200 import _ffi

    text = open(os.path.join(ROOT, "include", "ptau_b200.h")).read()
    defs = dict(re.findall(r"#define\s+PTAU_([A-Z0-9_]+)\s+\(?(-?\d+)u?\)?\s", text))
    for name, val in defs.items():
        if hasattr(_ffi, name):
            assert getattr(_ffi, name) == int(val), name
    assert _ffi.CHECKS_STRICT == 14 and _ffi.CHECKS_READ == 4


def test_layout_helpers_and_loud_failure_without_gpu():
    import torch

    import kzg_setup_powersoftau_b200 as kz
    from kzg_setup_powersoftau_b200 import _ffi

    L = _ffi.lib()
    for k in (3, 16, 21, 26):
        n = 1 << k
        assert L.ptau_response_size(n) == o.response_size(n)
        assert L.ptau_uncompressed_size(n) == o.uncompressed_size(n)
        assert L.ptau_setup_size(kz.VARIANT_KGZ, n) == o.kgz_size(n)
        assert L.ptau_setup_size(kz.VARIANT_FASTKGZ, n) == o.fastkgz_size(n)
    for g in (1, 2):
        for f in (1, 2, 3, 4):
            assert L.ptau_record_size(g, f) == SZ[g][f]
    assert kz._n_from_setup_size(kz.VARIANT_KGZ, o.kgz_size(1 << 21)) == 1 << 21
    assert kz._n_from_setup_size(kz.VARIANT_FASTKGZ, o.fastkgz_size(8)) == 8
    if not torch.cuda.is_available():
        # no CPU fallback: creating a context must fail, loudly
        with pytest.raises(kz.PtauError) as e:
            kz.Context(1)
        assert e.value.code == _ffi.ERR_CUDA


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "kzg_setup_powersoftau_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "cpu_ref" not in text and "ptau_oracle" not in text.replace("oracle/ptau_oracle.py constants", "")
    out = subprocess.run(["nm", "-D", os.path.join(pkg, "libptau_b200.so")], capture_output=True, text=True).stdout
    assert "oracle_" not in out


# ---- N > 1 host logic (gloo, world_size 2) ------------------------------------------
_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle"))
import torch, torch.distributed as dist
from kzg_setup_powersoftau_b200 import sharding
import cpu_ref
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
data = open(os.path.join(sys.argv[1], "tests", "golden", "n8_powersoftau_uncompressed.bin"), "rb").read()[:15 * 96]
data = bytearray(data * 3)[: n * 96]
bad = [11, 29]
for b in bad:
    data[b * 96 + 95] ^= 1          # corrupt y: off curve
lo, hi = sharding.shard_range(n, rank, world)
# stand-in for the GPU leg on this CPU-only box: the checker computes the shard's status
out, st = cpu_ref.convert(1, 1, bytes(data[lo * 96:hi * 96]), 3, 14)
local = sharding.STATUS_NONE
for i, s in enumerate(st):
    if s:
        local = ((lo + i) << 8) | s
        break
status = sharding.reduce_status(local)
ms = sharding.reduce_max_ms(10.0 * (rank + 1))
total = sharding.reduce_sum(hi - lo)
if rank == 0:
    print("RESULT", status >> 8, status & 0xff, ms, total, flush=True)
dist.destroy_process_group()
"""


def test_sharding_ranges():
    from kzg_setup_powersoftau_b200 import sharding

    for n in (0, 1, 7, 8, 1000, (1 << 28) - 1):
        for world in (1, 2, 4, 8):
            rs = [sharding.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0]
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


def test_world_size_2_gloo(cref, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29531", str(script), ROOT],
        capture_output=True, text=True, timeout=240, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert int(line[1]) == 11 and int(line[2]) == 4  # lowest bad index wins, NOT_ON_CURVE
    assert float(line[3]) == 20.0 and int(line[4]) == 37


def test_limb_pairing_code_vs_oracle(hostemul):
    """csrc/pairing.cuh compiled for the host (same limb arithmetic, same formulas as the kernels): GT bytes of a
    pairing product against the independent CPU pairing, and KZG10::check accepting / rejecting."""
    import pairing_oracle as po

    def g1r(q):
        return o.g1_mont_record(0, 1, True) if q is None else o.g1_mont_record(q[0], q[1], False)

    def g2r(q):
        return o.g2_mont_record((0, 0), (1, 0), True) if q is None else o.g2_mont_record(q[0], q[1], False)

    P1, Q1 = o.g1_mul(o.G1_GEN, 0xABCDEF), o.g2_mul(o.G2_GEN, 0x123457)
    P2, Q2 = o.g1_mul(o.G1_GEN, 77), o.g2_mul(o.G2_GEN, 99)
    items = [((P1, Q1), (P2, Q2)), ((P1, Q1), (None, Q2))]
    g1 = b"".join(g1r(p) for it in items for p, _ in it)
    g2 = b"".join(g2r(q) for it in items for _, q in it)
    gt = ctypes.create_string_buffer(576 * len(items))
    one = ctypes.create_string_buffer(len(items))
    hostemul.hostemul_pairing_product2(g1, g2, ctypes.c_size_t(len(items)), gt, one)
    assert one.raw == b"\x00\x00"

    def flat(rec):
        return po.f12_from_tower([int.from_bytes(rec[48 * i:48 * i + 48], "little") for i in range(12)])

    e1 = po.pairing(P1, Q1)
    assert flat(gt.raw[576:]) == e1
    assert flat(gt.raw[:576]) == po.f12_mul(e1, po.pairing(P2, Q2))
    # KZG10::check with a known tau
    R = o.R_ORDER
    tau, alpha = 0x1234567, 0x7654321
    coeffs, blind, zpt = [5, 7, 11, 13], [17, 19], 99
    ev = lambda c, x: sum(v * pow(x, i, R) for i, v in enumerate(c)) % R  # noqa: E731
    quot = lambda c, x: [sum(c[j] * pow(x, j - i - 1, R) for j in range(i + 1, len(c))) % R for i in range(len(c) - 1)]  # noqa: E731
    comm = o.g1_mul(o.G1_GEN, (ev(coeffs, tau) + alpha * ev(blind, tau)) % R)
    w = o.g1_mul(o.G1_GEN, (ev(quot(coeffs, zpt), tau) + alpha * ev(quot(blind, zpt), tau)) % R)
    vk1 = g1r(o.G1_GEN) + g1r(o.g1_mul(o.G1_GEN, alpha))
    vk2 = g2r(o.G2_GEN) + g2r(o.g2_mul(o.G2_GEN, tau))
    le = lambda v: (v % R).to_bytes(32, "little")  # noqa: E731
    ok = ctypes.create_string_buffer(2)
    comm0 = o.g1_mul(o.G1_GEN, ev(coeffs, tau))
    w0 = o.g1_mul(o.G1_GEN, ev(quot(coeffs, zpt), tau))
    big = R - 12345  # a full-width evaluation point: every window of the fixed-base tables is used
    wb = o.g1_mul(o.G1_GEN, ev(quot(coeffs, big), tau))
    for use_tables in (0, 1, 2):  # plain double-and-add, the 4-bit window tables, the 8-bit tables the library uses
        hostemul.hostemul_kzg_check(vk1, vk2, g1r(comm) * 2, le(zpt) * 2, le(ev(coeffs, zpt)) + le(ev(coeffs, zpt) + 1),
                                    g1r(w) * 2, le(ev(blind, zpt)) * 2, ctypes.c_size_t(2), ok, use_tables)
        assert ok.raw == b"\x01\x00"
        # without hiding
        hostemul.hostemul_kzg_check(vk1, vk2, g1r(comm0) * 2, le(zpt) + le(zpt + 1), le(ev(coeffs, zpt)) * 2, g1r(w0) * 2, None,
                                    ctypes.c_size_t(2), ok, use_tables)
        assert ok.raw == b"\x01\x00"
        hostemul.hostemul_kzg_check(vk1, vk2, g1r(comm0) * 2, le(big) * 2, le(ev(coeffs, big)) + le(ev(coeffs, big) - 1),
                                    g1r(wb) * 2, None, ctypes.c_size_t(2), ok, use_tables)
        assert ok.raw == b"\x01\x00"
        # degenerate openings (the two affine conversions share one inversion): a constant polynomial has the witness at
        # infinity and C - [v]g at infinity (accepted; a wrong value leaves C - [v]g finite and is rejected), and the
        # evaluation point tau itself makes beta_h - [z]h infinity (accepted only when C - [v]g is infinity too)
        cc = o.g1_mul(o.G1_GEN, 42)
        hostemul.hostemul_kzg_check(vk1, vk2, g1r(cc) * 2, le(zpt) * 2, le(42) + le(43), g1r(None) * 2, None,
                                    ctypes.c_size_t(2), ok, use_tables)
        assert ok.raw == b"\x01\x00"
        wt = o.g1_mul(o.G1_GEN, ev(quot(coeffs, tau), tau))
        hostemul.hostemul_kzg_check(vk1, vk2, g1r(comm0) * 2, le(tau) * 2, le(ev(coeffs, tau)) + le(ev(coeffs, tau) + 1),
                                    g1r(wt) * 2, None, ctypes.c_size_t(2), ok, use_tables)
        assert ok.raw == b"\x01\x00"


def test_msm_signed_window_recoding(hostemul):
    """The bucket MSM's window geometry and signed-digit recoding (csrc/msm_digits.cuh, the code the kernels run):
    widths add up to 256, digits stay within [-(2^(cw-1) - 1), 2^(cw-1)], the top digit is non-negative with no carry
    out, bucket bases are contiguous, and the digits recompose the scalar -- for every window width and for scalars
    that stress the carries."""
    R = o.R_ORDER
    rnd = random.Random(9)
    widths = set()
    for n in [1, 2, 100, 200, 300, 600, 2047, 4096, 5000, 10_000, 32768, 40_000, 65537, 131073, 262149, 1 << 20, 1 << 24]:
        geom = (ctypes.c_int * 5)()
        digits = (ctypes.c_int * 128)()
        bitoff = (ctypes.c_int * 128)()
        bases = (ctypes.c_uint32 * 128)()
        scalars = [0, 1, R - 1, (1 << 255) - 1, (1 << 254) + 12345] + [rnd.randrange(R) for _ in range(20)]
        for k in scalars:
            kw = (ctypes.c_uint32 * 8)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
            carry = hostemul.hostemul_msm_recode(ctypes.c_uint64(n), kw, geom, digits, bitoff, bases)
            c, W, a, lgL, NB = list(geom)
            assert carry == 0, (n, hex(k), carry)
            assert 3 <= c <= 16 and W == -(-256 // c) and 0 <= a <= W and a * c + (W - a) * (c - 1) == 256
            assert (NB >> 1) % (1 << lgL) == 0
            widths.add(c)
            base = 0
            for w in range(W):
                cw = c if w < a else c - 1
                assert -(1 << (cw - 1)) < digits[w] <= (1 << (cw - 1)), (n, w, digits[w])
                assert bases[w] == base
                base += 1 << (cw - 1)
            assert digits[W - 1] >= 0
            assert sum(digits[w] << bitoff[w] for w in range(W)) == k, (n, hex(k))
        # boundary digits: exactly 2^(cw-1) stays positive, one more goes negative with a carry
        for delta in (0, 1):
            k = sum(((1 << ((c if w < a else c - 1) - 1)) + delta) << bitoff[w] for w in range(W - 2))
            kw = (ctypes.c_uint32 * 8)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)])
            assert hostemul.hostemul_msm_recode(ctypes.c_uint64(n), kw, geom, digits, bitoff, bases) == 0
            assert sum(digits[w] << bitoff[w] for w in range(W)) == k
            if delta == 0:
                assert all(digits[w] == 1 << ((c if w < a else c - 1) - 1) for w in range(W - 2))
            else:
                assert digits[0] < 0
    assert widths == set(range(3, 17))


def test_limb_msm_code_vs_known_tau(hostemul):
    """The bucket MSM's per-item code (csrc/msm.cuh: bucket sums with complete additions, running-sum window
    reduction, window weights, final inversion) run serially on the host: sum c_i [tau^i]G == [sum c_i tau^i]G,
    plus repeated points, cancellation, zero scalars and infinity records."""
    R = o.R_ORDER
    tau = 0xABCDEF123
    rnd = random.Random(13)

    def rec(q):
        return o.g1_mont_record(0, 1, True) if q is None else o.g1_mont_record(q[0], q[1], False)

    def msm(points, scalars):
        out = ctypes.create_string_buffer(104)
        hostemul.hostemul_msm_g1(b"".join(points), b"".join(int(s).to_bytes(32, "little") for s in scalars),
                                 ctypes.c_size_t(len(scalars)), out)
        return out.raw

    n = 300   # c = 5: 52 windows of 5 and 4 bits
    pts, q = [], o.G1_GEN
    tp = [1]
    for i in range(n):
        pts.append(rec(q))
        q = o.g1_mul(q, tau)
        tp.append(tp[-1] * tau % R)
    for m in (1, 7, 130, n):   # c = 3, 3, 4, 5
        sc = [rnd.randrange(R) for _ in range(m)]
        assert msm(pts[:m], sc) == rec(o.g1_mul(o.G1_GEN, sum(c * t for c, t in zip(sc, tp)) % R)), m
    assert msm(pts[:50], [R - 1] * 50) == rec(o.g1_mul(o.G1_GEN, (-sum(tp[:50])) % R))
    assert msm(pts[:50], [0] * 50) == rec(None)
    assert msm([pts[3]] * 40, [7] * 40) == rec(o.g1_mul(o.G1_GEN, 280 * tp[3] % R))          # P + P inside a bucket
    assert msm([pts[3]] * 40, [9, R - 9] * 20) == rec(None)                                  # P + (-P)
    holes = [p if i % 3 else p[:96] + b"\x01" + p[97:] for i, p in enumerate(pts[:60])]
    sc = [rnd.randrange(R) for _ in range(60)]
    assert msm(holes, sc) == rec(o.g1_mul(o.G1_GEN, sum(c * t for i, (c, t) in enumerate(zip(sc, tp)) if i % 3) % R))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU: one JSON line with
    the contract's keys, alone and under torchrun (rank 0 prints, the other rank exits 0 without work)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmds = [[sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
            [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
             "--master-port", "29655", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "3"]]
    for n, cmd in ((1, cmds[0]), (2, cmds[1])):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["n_gpus"] == n and d["steps"] == 1 and d["warmup"] == 3
        assert d["unit"] == "points/s" and d["higher_is_better"] is True and d["value"] > 0
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
        assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert "workload" in d["config"] and d["metric"].startswith("G1+G2 points/sec") and d["scaling"] == "strong" and "configs[2]" in d["config"]["workload"]


def test_bench_reads_roofline_figures_from_the_committed_ncu_summary(tmp_path, monkeypatch):
    """bench.py takes `roofline.traffic` / pipe-busy from a named profiles/*.csv instead of literals: the last capture of
    the kernel whose counters are complete is used (ncu leaves "-nan" where a replay pass failed), bytes are scaled by
    the unit row, and a missing file or kernel gives None (the line then says traffic: null)."""
    sys.path.insert(0, ROOT)
    import bench

    named = bench.NCU_SUMMARY
    k = "void convert_kernel<2, 2, 3, 1>(const unsigned i"
    rows = [["metric", "unit", "void convert_kernel<1, 2, 3, 1>(const unsigned i", k, k],
            ["gpu__time_duration.sum", "ms", "47.0", "77.5", "77.2"],
            ["launch__registers_per_thread", "register/thread", "220", "255", "255"],
            ["launch__grid_size", "", "8192", "8192", "8192"], ["launch__block_size", "", "128", "128", "128"],
            ["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "%", "86.1", "81.5", "-nan"],
            ["dram__bytes_read.sum", "Mbyte", "53.5", "109.0", "-nan"], ["dram__bytes_write.sum", "Mbyte", "73.3", "252.0", "-nan"],
            ["l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "sector", "1", "2", "-nan"],
            ["l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sector", "3", "4", "-nan"]]
    (tmp_path / "profiles").mkdir()
    with open(tmp_path / "profiles" / "x.csv", "w", newline="") as f:
        csv.writer(f).writerows(rows)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "NCU_SUMMARY", os.path.join("profiles", "x.csv"))
    m = bench.ncu_summary_metrics(bench.KERNEL_G2C)
    assert m["points"] == 8192 * 128 and m["dram_bytes"] == 361.0e6 and m["fmaheavy_pct"] == 81.5 and m["registers"] == 255
    assert bench.ncu_summary_metrics(bench.KERNEL_G1C)["dram_bytes"] == pytest.approx(126.8e6)
    assert bench.ncu_summary_metrics("convert_kernel<9, 9, 9, 9>") is None
    monkeypatch.setattr(bench, "NCU_SUMMARY", os.path.join("profiles", "absent.csv"))
    assert bench.ncu_summary_metrics(bench.KERNEL_G2C) is None
    # and the file the line names is committed
    assert os.path.exists(os.path.join(ROOT, named))


def _pairing_vectors():
    with open(os.path.join(GOLDEN, "pairing_vectors.json")) as f:
        return json.load(f)


def _g1r(q):
    return o.g1_mont_record(0, 1, True) if q is None else o.g1_mont_record(q[0], q[1], False)


def _g2r(q):
    return o.g2_mont_record((0, 0), (1, 0), True) if q is None else o.g2_mont_record(q[0], q[1], False)


def test_limb_g2_prepared_coefficients(hostemul):
    """csrc/pairing.cuh g2_prepare_item (the G2Prepared of ark-ec 0.2: prepared_h / prepared_beta_h of src/lib.rs:223-224)
    on the host against oracle/pairing_oracle.py g2_prepared_coeffs, and the oracle's coefficients against the pairing
    itself: a Miller loop evaluated from them must give the same GT element as the independent construction."""
    import pairing_oracle as po

    pts = [o.g2_mul(o.G2_GEN, k) for k in (1, 7, 0xB200B200)] + [None]
    recs = b"".join(_g2r(q) for q in pts)
    out = ctypes.create_string_buffer(len(pts) * 68 * 288)
    inf = ctypes.create_string_buffer(len(pts))
    hostemul.hostemul_g2_prepare(recs, ctypes.c_size_t(len(pts)), out, inf)
    assert inf.raw == b"\0\0\0\1"
    rinv = pow(1 << 384, -1, o.P)
    for i, q in enumerate(pts):
        want, is_inf = po.g2_prepared_coeffs(q)
        blob = out.raw[i * 68 * 288:(i + 1) * 68 * 288]
        if is_inf:
            assert blob == bytes(68 * 288) and want == []
            continue
        assert len(want) == 68
        got = []
        for t in range(68):
            vals = [int.from_bytes(blob[t * 288 + 48 * k:t * 288 + 48 * k + 48], "little") * rinv % o.P for k in range(6)]
            got.append(((vals[0], vals[1]), (vals[2], vals[3]), (vals[4], vals[5])))
        assert got == want
    # the coefficients define the pairing: e(P, Q) from prepared(Q) equals the oracle's own e(P, Q) (z < 0: inverse)
    p1, q2 = o.g1_mul(o.G1_GEN, 5), pts[1]
    f = po.miller_from_prepared(po.g2_prepared_coeffs(q2)[0], p1)
    assert po.f12_pow(po.f12_inv(f), po.FINAL_EXP) == po.pairing(p1, q2)


def test_limb_pairing_code_on_golden_vectors(hostemul):
    """tests/golden/pairing_vectors.json (made by tools/make_golden_pairing.py from the independent CPU pairing)
    against csrc/pairing.cuh compiled for the host: GT values of pairing products and KZG10::check booleans."""
    import pairing_oracle as po

    vec = _pairing_vectors()
    for pv in vec["products"]:
        g1 = b"".join(_g1r(o.g1_mul(o.G1_GEN, int(s, 16)) if int(s, 16) else None) for s in pv["g1_scalars"])
        g2 = b"".join(_g2r(o.g2_mul(o.G2_GEN, int(s, 16)) if int(s, 16) else None) for s in pv["g2_scalars"])
        gt = ctypes.create_string_buffer(576)
        one = ctypes.create_string_buffer(1)
        hostemul.hostemul_pairing_product2(g1, g2, ctypes.c_size_t(1), gt, one)
        flat = po.f12_from_tower([int.from_bytes(gt.raw[48 * i:48 * i + 48], "little") for i in range(12)])
        assert flat == [int(v, 16) for v in pv["gt_flat"]]
        assert (one.raw == b"\x01") == pv["is_one"]
    tau, alpha, _ = o.derive_scalars(0xB200)
    vk1 = _g1r(o.G1_GEN) + _g1r(o.g1_mul(o.G1_GEN, alpha))
    vk2 = _g2r(o.G2_GEN) + _g2r(o.g2_mul(o.G2_GEN, tau))
    le = lambda v: int(v, 16).to_bytes(32, "little")  # noqa: E731
    for kc in vec["kzg_checks"]:
        comm = o.g1_mul(o.G1_GEN, int(kc["commitment_scalar"], 16))
        ws = int(kc["proof_scalar"], 16)
        ok = ctypes.create_string_buffer(1)
        hostemul.hostemul_kzg_check(vk1, vk2, _g1r(comm), le(kc["point"]), le(kc["value"]),
                                    _g1r(o.g1_mul(o.G1_GEN, ws) if ws else None),
                                    le(kc["random_v"]) if kc["random_v"] else None, ctypes.c_size_t(1), ok, 1)
        assert (ok.raw == b"\x01") == kc["expect"], kc


def test_kzg_quotient_host_routine():
    """ptau_kzg_quotient (the polynomial side of KZG10::open) against plain big-integer synthetic division:
    p(X) - p(z) == (X - z) q(X), including the empty, constant and linear polynomials and non-canonical input."""
    import numpy as np
    import kzg_setup_powersoftau_b200 as kz
    from kzg_setup_powersoftau_b200 import _ffi

    R = o.R_ORDER
    rnd = random.Random(17)
    L = _ffi.lib()

    def run(coeffs, z):
        n = len(coeffs)
        cb = b"".join(int(c).to_bytes(32, "little") for c in coeffs)
        q = ctypes.create_string_buffer(max(n - 1, 1) * 32)
        v = ctypes.create_string_buffer(32)
        rc = L.ptau_kzg_quotient(cb, n, int(z).to_bytes(32, "little"), q, v)
        return rc, [int.from_bytes(q.raw[32 * i:32 * i + 32], "little") for i in range(max(n - 1, 0))], int.from_bytes(v.raw, "little")

    for n in (0, 1, 2, 3, 17, 1000):
        coeffs = [rnd.randrange(R) for _ in range(n)]
        for z in (0, 1, R - 1, rnd.randrange(R)):
            rc, q, v = run(coeffs, z)
            assert rc == 0
            assert v == sum(c * pow(z, i, R) for i, c in enumerate(coeffs)) % R
            # (X - z) q(X) + v == p(X), coefficient by coefficient
            for i in range(n):
                lhs = ((q[i - 1] if 1 <= i <= n - 1 else 0) - z * (q[i] if i < n - 1 else 0) + (v if i == 0 else 0)) % R
                assert lhs == coeffs[i], (n, i)
    assert run([R, 1], 5)[0] == _ffi.ERR_ARG and run([1, 2], R)[0] == _ffi.ERR_ARG
    # the Python mirror switches to it for long polynomials: same answer as its own loop
    coeffs = [rnd.randrange(R) for _ in range(300)]
    v1, q1 = kz.KZG10._quotient(coeffs, 12345)
    v2, q2 = kz.KZG10._quotient(coeffs[:200], 12345)
    assert v1 == sum(c * pow(12345, i, R) for i, c in enumerate(coeffs)) % R and len(q1) == 299 and len(q2) == 199
