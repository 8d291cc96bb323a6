// Per-point hot path: parse -> canonical/flag checks -> (sqrt decompression) ->
// on-curve -> subgroup -> re-encode.  One point per thread.
//
// Reference behaviour being replaced (see include/ptau_b200.h for the table):
//   * zcash-uncompressed input  == read_g1 / read_g2, /root/reference/src/lib.rs:41-80:
//     reverse each 48-byte coordinate (G2: also swap c1|c0 -> c0|c1) and hand the
//     bytes to ark GroupAffine::deserialize_uncompressed.  So this format is parsed
//     with *arkworks* flag semantics on the byte-reversed record: x must be < p with
//     no flag bits; the top two bits of y (G2: y.c1) are SWFlags (11 invalid,
//     10 "positive" accepted and stripped, 01 infinity).
//   * ark-uncompressed input    == GroupAffine::deserialize_unchecked, src/lib.rs:180-214.
//   * zcash-compressed input    == pairing 0.14.2 G{1,2}Compressed::into_affine_unchecked
//     as driven by powersoftau Accumulator::deserialize(Compressed, CheckForCorrectness::No),
//     /root/reference/src/bin/preprocess-kgz.rs:105-109.
//   * ark-uncompressed output   == GroupAffine::serialize_uncompressed,
//     preprocess-kgz.rs:188-194.
//   * zcash-uncompressed output == Accumulator::serialize(UseCompression::No), :122-124.
//
// Records are handled as arrays of u32 words in memory order (word k = bytes
// 4k..4k+3 loaded little-endian).
#pragma once
#include "curve.cuh"
#include "../../include/ptau_b200.h"

namespace ptau {

PTAU_HD uint32_t bswap32(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __byte_perm(v, 0, 0x0123);
#else
  return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}

// 48 big-endian bytes -> plain limbs
PTAU_HD Fq fq_from_be_words(const uint32_t* w) {
  Fq r;
#pragma unroll
  for (int j = 0; j < 12; j++) r.l[j] = bswap32(w[11 - j]);
  return r;
}
PTAU_HD void fq_to_be_words(const Fq& a, uint32_t* w) {
#pragma unroll
  for (int j = 0; j < 12; j++) w[11 - j] = bswap32(a.l[j]);
}
PTAU_HD Fq fq_from_le_words(const uint32_t* w) {
  Fq r;
#pragma unroll
  for (int j = 0; j < 12; j++) r.l[j] = w[j];
  return r;
}
PTAU_HD void fq_to_le_words(const Fq& a, uint32_t* w) {
#pragma unroll
  for (int j = 0; j < 12; j++) w[j] = a.l[j];
}

// ---- fixed-exponent chains ----------------------------------------------------
#ifdef __CUDACC__
__constant__ uint8_t K_P34_CHAIN_D[PTAU_P34_STEPS] = PTAU_P34_CHAIN_INIT;
#endif
static const uint8_t K_P34_CHAIN_H[PTAU_P34_STEPS] = PTAU_P34_CHAIN_INIT;

// a^((p-3)/4): 376 squarings + 85 multiplications
PTAU_HD_NOINLINE Fq fq_pow_p34(const Fq& a) {
#ifdef __CUDA_ARCH__
  const uint8_t* chain = K_P34_CHAIN_D;
#else
  const uint8_t* chain = K_P34_CHAIN_H;
#endif
  Fq tbl[8];  // a^1, a^3, ..., a^15 (runtime-indexed: lives in local memory)
  tbl[0] = a;
  Fq a2 = fq_sqr(a);
#pragma unroll 1
  for (int i = 1; i < 8; i++) tbl[i] = fq_mul(tbl[i - 1], a2);
  Fq acc = tbl[(chain[0] & 15) >> 1];
#pragma unroll 1
  for (int s = 1; s < PTAU_P34_STEPS; s++) {
    int c = chain[s];
    int n = c >> 4;
#if defined(__CUDA_ARCH__) && !defined(PTAU_POW_CALLS)
    // the runs of squarings expanded in place: no call per squaring (G1 compressed 47.1 -> 45.9 ms, G2 compressed
    // 73.6 -> 72.3 ms per 2^20 points, profiles/r02c_ab_*.log)
#pragma unroll 1
    for (int k = 0; k < n; k++) acc = fq_sqr_inl(acc);
#else
#pragma unroll 1
    for (int k = 0; k < n; k++) acc = fq_sqr(acc);
#endif
    if (c & 15) acc = fq_mul(acc, tbl[(c & 15) >> 1]);
  }
  return acc;
}

// sqrt in Fq (p = 3 mod 4): s = a^((p+1)/4); ok iff s^2 == a.
// pairing 0.14.2 fq.rs sqrt computes the same value.
PTAU_HD Fq fq_sqrt(const Fq& a, bool& ok) {
  Fq s = fq_mul(fq_pow_p34(a), a);
  ok = fq_eq(fq_sqr(s), a);
  return s;
}

// sqrt in Fq2 by the norm method (two Fq exponentiations).  pairing 0.14.2 uses
// Algorithm 9 of eprint 2012/685 (two Fq2 exponentiations); the two agree up to
// sign, and the caller fixes the sign from the encoding's "largest" flag, so the
// produced bytes are identical (tests/test_oracle_pins.py::test_fq2_sqrt_methods).
PTAU_HD_NOINLINE Fq2 fq2_sqrt(const Fq2& a, bool& ok) {
  Fq n = fq_add(fq_sqr(a.c0), fq_sqr(a.c1));
  Fq s = fq_mul(fq_pow_p34(n), n);  // sqrt(norm) if it exists; verified at the end
  Fq half = k_half_mont();
  Fq d = fq_mul(fq_add(a.c0, s), half);
  if (fq_is_zero(d)) d = fq_mul(fq_sub(a.c0, s), half);  // only when a.c1 == 0
  Fq t = fq_pow_p34(d);
  Fq x0 = fq_mul(d, t);    // d^((p+1)/4)
  Fq chi = fq_mul(x0, t);  // d^((p-1)/2) = +-1 (0 if d == 0)
  Fq w = fq_mul(fq_mul(a.c1, t), half);
  bool qr = fq_eq(chi, fq_one()) || fq_is_zero(d);
  Fq2 r;
  Fq nx0 = fq_neg(x0);
#pragma unroll
  for (int i = 0; i < 12; i++) {
    r.c0.l[i] = qr ? x0.l[i] : w.l[i];
    r.c1.l[i] = qr ? w.l[i] : nx0.l[i];
  }
  ok = fq2_eq(fq2_sqr(r), a);
  return r;
}

// plain value > (p-1)/2 ?
PTAU_HD bool fq_plain_is_largest(const Fq& yp) { return fq_gt_plain(yp, k_pm1_half_plain()); }

// p - a for plain a in (0, p); 0 -> 0
PTAU_HD Fq fq_plain_neg(const Fq& a) {
  const uint32_t pl[12] = PTAU_P_LIMBS;
  Fq r;
  uint32_t bf = 0;
  bool z = fq_is_zero(a);
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)pl[i] - a.l[i] - bf;
    bf = (uint32_t)(t >> 63);
    r.l[i] = z ? 0u : (uint32_t)t;
  }
  return r;
}

// ---- record sizes (bytes) -----------------------------------------------------
PTAU_HD constexpr int record_bytes(int group, int fmt) {
  return fmt == PTAU_FMT_ZCASH_COMPRESSED    ? (group == PTAU_G1 ? 48 : 96)
         : fmt == PTAU_FMT_ARK_MONT_LIMBS    ? (group == PTAU_G1 ? 104 : 200)
                                             : (group == PTAU_G1 ? 96 : 192);
}

// =============================================================================
// G1
// =============================================================================
// HEAVY = false compiles the curve checks out (used when the caller asked for none
// of them on an uncompressed input): a small-register kernel that runs at HBM speed.
template <int INFMT, bool HEAVY = true>
PTAU_HD uint32_t g1_process(const uint32_t* in, int out_fmt, uint32_t* out, uint32_t checks) {
  if (!HEAVY) checks &= PTAU_CHECK_REJECT_INFINITY;
  Fq xp, yp, xm, ym;
  uint32_t st = PTAU_OK;
  bool inf = false;
  bool have_mont = false;

  if (INFMT == PTAU_FMT_ZCASH_COMPRESSED) {
    xp = fq_from_be_words(in);
    uint32_t fl = xp.l[11] >> 29;
    xp.l[11] &= 0x1fffffffu;
    bool greatest = fl & 1u;
    yp = fq_zero();
    if (!(fl & 4u)) {
      st = PTAU_BAD_FLAGS;
    } else if (fl & 2u) {
      if ((fl & 1u) || !fq_is_zero(xp)) st = PTAU_BAD_FLAGS;
      inf = true;
      yp.l[0] = 1;  // ark zero() = (0, 1, infinity)
    } else if (fq_plain_ge_p(xp)) {
      st = PTAU_BAD_NON_CANONICAL;
    }
    if (st == PTAU_OK && !inf) {
      xm = fq_to_mont(xp);
      Fq rhs = fq_add(fq_mul(fq_sqr(xm), xm), k_b1_mont());
      bool ok;
      ym = fq_sqrt(rhs, ok);
      if (!ok) st = PTAU_BAD_NOT_ON_CURVE;
      yp = fq_from_mont(ym);
      if (fq_plain_is_largest(yp) != greatest) {
        ym = fq_neg(ym);
        yp = fq_plain_neg(yp);
      }
      have_mont = true;
    }
  } else if (INFMT == PTAU_FMT_ARK_MONT_LIMBS) {
    // in-memory GroupAffine (ark-ff Montgomery limbs + infinity byte) -> the serialize
    // direction, preprocess-kgz.rs:188-194.  Limbs must be < p (ark's invariant).
    xm = fq_from_le_words(in);
    ym = fq_from_le_words(in + 12);
    inf = (in[24] & 0xffu) != 0;
    if (fq_plain_ge_p(xm) || fq_plain_ge_p(ym)) {
      st = PTAU_BAD_NON_CANONICAL;
      xm = fq_zero();
      ym = fq_zero();
    }
    have_mont = true;
    xp = fq_from_mont(xm);
    yp = fq_from_mont(ym);
    if (HEAVY && st == PTAU_OK && !inf && (checks & PTAU_CHECK_ON_CURVE) && !g1_on_curve(xm, ym)) st = PTAU_BAD_NOT_ON_CURVE;
  } else {
    if (INFMT == PTAU_FMT_ZCASH_UNCOMPRESSED) {
      xp = fq_from_be_words(in);
      yp = fq_from_be_words(in + 12);
    } else {
      xp = fq_from_le_words(in);
      yp = fq_from_le_words(in + 12);
    }
    uint32_t fl = yp.l[11] >> 30;
    yp.l[11] &= 0x3fffffffu;
    if (fq_plain_ge_p(xp)) {
      st = PTAU_BAD_NON_CANONICAL;
    } else if (fl == 3u) {
      st = PTAU_BAD_FLAGS;
    } else if (fq_plain_ge_p(yp)) {
      st = PTAU_BAD_NON_CANONICAL;
    }
    inf = (fl == 1u);
    bool need = (checks & (PTAU_CHECK_ON_CURVE | PTAU_CHECK_SUBGROUP)) && !inf;
    if (st == PTAU_OK && (need || out_fmt == PTAU_FMT_ARK_MONT_LIMBS)) {
      xm = fq_to_mont(xp);
      ym = fq_to_mont(yp);
      have_mont = true;
      if (HEAVY && need && (checks & PTAU_CHECK_ON_CURVE) && !g1_on_curve(xm, ym)) st = PTAU_BAD_NOT_ON_CURVE;
    }
  }
  if (st == PTAU_OK && inf && (checks & PTAU_CHECK_REJECT_INFINITY)) st = PTAU_BAD_INFINITY;
  if (HEAVY && st == PTAU_OK && !inf && (checks & PTAU_CHECK_SUBGROUP)) {
    if (!g1_in_subgroup(xm, ym)) st = PTAU_BAD_NOT_IN_SUBGROUP;
  }

  // ---- encode ----
  if (out_fmt == PTAU_FMT_ARK_UNCOMPRESSED) {
    fq_to_le_words(xp, out);
    fq_to_le_words(yp, out + 12);
    if (inf) out[23] |= 0x40000000u;
  } else if (out_fmt == PTAU_FMT_ZCASH_UNCOMPRESSED) {
    if (inf) {
#pragma unroll
      for (int i = 0; i < 24; i++) out[i] = 0;
      out[0] = 0x40u;
    } else {
      fq_to_be_words(xp, out);
      fq_to_be_words(yp, out + 12);
    }
  } else {  // PTAU_FMT_ARK_MONT_LIMBS
    if (st == PTAU_OK && !have_mont) {
      xm = fq_to_mont(xp);
      ym = fq_to_mont(yp);
      have_mont = true;
    }
    if (!have_mont) {
      xm = fq_zero();
      ym = fq_zero();
    }
    fq_to_le_words(xm, out);
    fq_to_le_words(ym, out + 12);
    out[24] = inf ? 1u : 0u;
    out[25] = 0;
  }
  return st;
}

// =============================================================================
// G2
// =============================================================================
// `out` may point to shared memory: the record is encoded BEFORE the subgroup ladder so that nothing but the
// ladder's own state is live across it.  `pk` is the ladder's operand file (curve.cuh: Park).
template <int INFMT, bool HEAVY, int PSTRIDE>
PTAU_HD uint32_t g2_process(const uint32_t* in, int out_fmt, uint32_t* out, uint32_t checks, Park<PSTRIDE> pk) {
  if (!HEAVY) checks &= PTAU_CHECK_REJECT_INFINITY;
  Fq2 xp, yp, xm, ym;
  uint32_t st = PTAU_OK;
  bool inf = false;
  bool have_mont = false;
  bool off_curve = false;

  if (INFMT == PTAU_FMT_ZCASH_COMPRESSED) {
    xp.c1 = fq_from_be_words(in);
    xp.c0 = fq_from_be_words(in + 12);
    uint32_t fl = xp.c1.l[11] >> 29;
    xp.c1.l[11] &= 0x1fffffffu;
    bool greatest = fl & 1u;
    yp = fq2_zero();
    if (!(fl & 4u)) {
      st = PTAU_BAD_FLAGS;
    } else if (fl & 2u) {
      if ((fl & 1u) || !fq2_is_zero(xp)) st = PTAU_BAD_FLAGS;
      inf = true;
      yp.c0.l[0] = 1;
    } else if (fq_plain_ge_p(xp.c1) || fq_plain_ge_p(xp.c0)) {
      st = PTAU_BAD_NON_CANONICAL;
    }
    if (st == PTAU_OK && !inf) {
      xm.c0 = fq_to_mont(xp.c0);
      xm.c1 = fq_to_mont(xp.c1);
      Fq2 b;
      b.c0 = k_b1_mont();
      b.c1 = b.c0;
      Fq2 rhs = fq2_add(fq2_mul(fq2_sqr(xm), xm), b);
      bool ok;
      ym = fq2_sqrt(rhs, ok);
      if (!ok) st = PTAU_BAD_NOT_ON_CURVE;
      yp.c0 = fq_from_mont(ym.c0);
      yp.c1 = fq_from_mont(ym.c1);
      // pairing Fq2 ordering: c1 first, then c0
      bool largest = fq_is_zero(yp.c1) ? fq_plain_is_largest(yp.c0) : fq_plain_is_largest(yp.c1);
      if (largest != greatest) {
        ym = fq2_neg(ym);
        yp.c0 = fq_plain_neg(yp.c0);
        yp.c1 = fq_plain_neg(yp.c1);
      }
      have_mont = true;
    }
  } else if (INFMT == PTAU_FMT_ARK_MONT_LIMBS) {
    xm.c0 = fq_from_le_words(in);
    xm.c1 = fq_from_le_words(in + 12);
    ym.c0 = fq_from_le_words(in + 24);
    ym.c1 = fq_from_le_words(in + 36);
    inf = (in[48] & 0xffu) != 0;
    if (fq_plain_ge_p(xm.c0) || fq_plain_ge_p(xm.c1) || fq_plain_ge_p(ym.c0) || fq_plain_ge_p(ym.c1)) {
      st = PTAU_BAD_NON_CANONICAL;
      xm = fq2_zero();
      ym = fq2_zero();
    }
    have_mont = true;
    xp.c0 = fq_from_mont(xm.c0);
    xp.c1 = fq_from_mont(xm.c1);
    yp.c0 = fq_from_mont(ym.c0);
    yp.c1 = fq_from_mont(ym.c1);
    if (HEAVY && st == PTAU_OK && !inf && (checks & (PTAU_CHECK_ON_CURVE | PTAU_CHECK_SUBGROUP)) && !g2_on_curve(xm, ym)) {
      if (checks & PTAU_CHECK_ON_CURVE) {
        st = PTAU_BAD_NOT_ON_CURVE;
      } else {
        off_curve = true;
        if (!g2_rmul_is_zero(xm, ym)) st = PTAU_BAD_NOT_IN_SUBGROUP;
      }
    }
  } else {
    if (INFMT == PTAU_FMT_ZCASH_UNCOMPRESSED) {
      xp.c1 = fq_from_be_words(in);
      xp.c0 = fq_from_be_words(in + 12);
      yp.c1 = fq_from_be_words(in + 24);
      yp.c0 = fq_from_be_words(in + 36);
    } else {
      xp.c0 = fq_from_le_words(in);
      xp.c1 = fq_from_le_words(in + 12);
      yp.c0 = fq_from_le_words(in + 24);
      yp.c1 = fq_from_le_words(in + 36);
    }
    uint32_t fl = yp.c1.l[11] >> 30;
    yp.c1.l[11] &= 0x3fffffffu;
    if (fq_plain_ge_p(xp.c0) || fq_plain_ge_p(xp.c1) || fq_plain_ge_p(yp.c0)) {
      st = PTAU_BAD_NON_CANONICAL;
    } else if (fl == 3u) {
      st = PTAU_BAD_FLAGS;
    } else if (fq_plain_ge_p(yp.c1)) {
      st = PTAU_BAD_NON_CANONICAL;
    }
    inf = (fl == 1u);
    // psi is an endomorphism of the twist only, so the psi test is used on points of the
    // twist; an off-curve point is an error under PTAU_CHECK_ON_CURVE and otherwise takes
    // the reference's own multiplication by r (curve.cuh: g2_rmul_is_zero).
    bool need = (checks & (PTAU_CHECK_ON_CURVE | PTAU_CHECK_SUBGROUP)) && !inf;
    if (st == PTAU_OK && (need || out_fmt == PTAU_FMT_ARK_MONT_LIMBS)) {
      xm.c0 = fq_to_mont(xp.c0);
      xm.c1 = fq_to_mont(xp.c1);
      ym.c0 = fq_to_mont(yp.c0);
      ym.c1 = fq_to_mont(yp.c1);
      have_mont = true;
      if (HEAVY && need && !g2_on_curve(xm, ym)) {
        if (checks & PTAU_CHECK_ON_CURVE) {
          st = PTAU_BAD_NOT_ON_CURVE;
        } else {
          off_curve = true;
          if (!g2_rmul_is_zero(xm, ym)) st = PTAU_BAD_NOT_IN_SUBGROUP;
        }
      }
    }
  }
  if (st == PTAU_OK && inf && (checks & PTAU_CHECK_REJECT_INFINITY)) st = PTAU_BAD_INFINITY;
  const bool ladder = HEAVY && st == PTAU_OK && !inf && !off_curve && (checks & PTAU_CHECK_SUBGROUP);
  if (ladder) pk.store_g2(xm, ym);

  // ---- encode (the subgroup result below only changes the status) ----
  if (out_fmt == PTAU_FMT_ARK_UNCOMPRESSED) {
    fq_to_le_words(xp.c0, out);
    fq_to_le_words(xp.c1, out + 12);
    fq_to_le_words(yp.c0, out + 24);
    fq_to_le_words(yp.c1, out + 36);
    if (inf) out[47] |= 0x40000000u;
  } else if (out_fmt == PTAU_FMT_ZCASH_UNCOMPRESSED) {
    if (inf) {
#pragma unroll
      for (int i = 0; i < 48; i++) out[i] = 0;
      out[0] = 0x40u;
    } else {
      fq_to_be_words(xp.c1, out);
      fq_to_be_words(xp.c0, out + 12);
      fq_to_be_words(yp.c1, out + 24);
      fq_to_be_words(yp.c0, out + 36);
    }
  } else {
    if (st == PTAU_OK && !have_mont) {
      xm.c0 = fq_to_mont(xp.c0);
      xm.c1 = fq_to_mont(xp.c1);
      ym.c0 = fq_to_mont(yp.c0);
      ym.c1 = fq_to_mont(yp.c1);
      have_mont = true;
    }
    if (!have_mont) {
      xm = fq2_zero();
      ym = fq2_zero();
    }
    fq_to_le_words(xm.c0, out);
    fq_to_le_words(xm.c1, out + 12);
    fq_to_le_words(ym.c0, out + 24);
    fq_to_le_words(ym.c1, out + 36);
    out[48] = inf ? 1u : 0u;
    out[49] = 0;
  }
  if (ladder && !g2_in_subgroup(pk)) st = PTAU_BAD_NOT_IN_SUBGROUP;
  return st;
}

#ifndef __CUDA_ARCH__
// host build (tests/host_emul): the operand file is a local array
template <int INFMT, bool HEAVY = true>
inline uint32_t g2_process(const uint32_t* in, int out_fmt, uint32_t* out, uint32_t checks) {
  uint32_t file[PTAU_PARK_WORDS];
  Park<1> pk;
  pk.p = file;
  return g2_process<INFMT, HEAVY, 1>(in, out_fmt, out, checks, pk);
}
#endif

}  // namespace ptau
