// Unreduced ("wide") products in Fq and the Montgomery reduction of a 24-limb value, in the same
// even/odd IMAD.WIDE scheme as fq.cuh -- the building blocks of lazy reduction in Fq2:
//
//   fq2_mul:  3 wide products (Karatsuba) + 2 reductions   = 3*144 + 2*156 = 744 MADs (fq.cuh: 3*300 = 900)
//   fq2_sqr:  3 wide squarings            + 2 reductions   = 3*78  + 2*156 = 546 MADs (fq.cuh: 2*300 = 600)
//
// ark-ff 0.2 / pairing 0.14.2 reduce after every Fq multiplication; the values produced here are
// the same field elements (the reduction is linear), checked limb for limb against the oracle in
// tests/test_host_logic.py (host build of this file) and on the GPU through ptau_selftest_fq_op.
//
// A wide value T is 24 x u32 limbs, T < p * 2^384 whenever it is handed to fq_redc.
#pragma once
#include "fq.cuh"

namespace ptau {

// a + b without the conditional subtraction (result < 2p < 2^382 for reduced inputs)
PTAU_HD Fq fq_add_nored(const Fq& a, const Fq& b) {
  Fq r;
  PX_DECL;
  PX_ADD_CC(r.l[0], a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) PX_ADDC_CC(r.l[i], a.l[i], b.l[i]);
  PX_ADDC(r.l[11], a.l[11], b.l[11]);
  return r;
}

// ---- single-instruction steps used below (fq.cuh has the others) --------------------------------
#ifdef __CUDA_ARCH__
#define PX_SUB_CC(r, a, b) asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define PX_SUBC_CC(r, a, b) asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define PX_SUBC(r, a, b) asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#else
#define PX_SUB_CC(r, a, b) do { px_cf = 0; r = emu::subc(a, b, px_cf); } while (0)
#define PX_SUBC_CC(r, a, b) r = emu::subc(a, b, px_cf)
#define PX_SUBC(r, a, b) do { r = emu::subc(a, b, px_cf); px_cf = 0; } while (0)
#endif

// a - b + p without any conditional step: in (0, 2p) for reduced inputs, for differences that only feed a multiplication
PTAU_HD Fq fq_sub_plus_p(const Fq& a, const Fq& b) {
  const uint32_t pl[12] = PTAU_P_LIMBS;
  Fq t, r;
  PX_DECL;
  PX_ADD_CC(t.l[0], a.l[0], pl[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) PX_ADDC_CC(t.l[i], a.l[i], pl[i]);
  PX_ADDC(t.l[11], a.l[11], pl[11]);
  PX_SUB_CC(r.l[0], t.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) PX_SUBC_CC(r.l[i], t.l[i], b.l[i]);
  PX_SUBC(r.l[11], t.l[11], b.l[11]);
  return r;
}

// a / 2: (a + (a odd ? p : 0)) >> 1, the same in Montgomery form
PTAU_HD Fq fq_half(const Fq& a) {
  const uint32_t pl[12] = PTAU_P_LIMBS;
  const uint32_t mask = 0u - (a.l[0] & 1u);
  uint32_t t[13];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    c += (uint64_t)a.l[i] + (pl[i] & mask);
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  t[12] = (uint32_t)c;
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
  return r;
}

// 2p, little-endian limbs
#define PTAU_2P_LIMBS                                                                      \
  {0xffff5556u, 0x73fdffffu, 0x62a7ffffu, 0x3d57fffdu, 0xed61ec48u, 0xce61a541u,           \
   0xe70a257eu, 0xc8ee9709u, 0x869759aeu, 0x96374f6cu, 0x72ffcd34u, 0x340223d4u}
// a - b + 2p for unreduced sums a, b < 2p: in (0, 4p), still below 2^384
PTAU_HD Fq fq_sub_plus_2p(const Fq& a, const Fq& b) {
  const uint32_t pl[12] = PTAU_2P_LIMBS;
  Fq t, r;
  PX_DECL;
  PX_ADD_CC(t.l[0], a.l[0], pl[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) PX_ADDC_CC(t.l[i], a.l[i], pl[i]);
  PX_ADDC(t.l[11], a.l[11], pl[11]);
  PX_SUB_CC(r.l[0], t.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < 11; i++) PX_SUBC_CC(r.l[i], t.l[i], b.l[i]);
  PX_SUBC(r.l[11], t.l[11], b.l[11]);
  return r;
}

// t[0..23] = a * b.  Same rows as fq_mul_inl without the reduction rows: after row i the lowest limb of the
// even-aligned accumulator is limb i of the product.
PTAU_HD void fq_mul_wide_plain(uint32_t* t, const Fq& a, const Fq& b) {
  uint32_t ev[12], od[12];
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    uint64_t t0 = (uint64_t)a.l[j] * b.l[0];
    uint64_t t1 = (uint64_t)a.l[j + 1] * b.l[0];
    ev[j] = (uint32_t)t0;
    ev[j + 1] = (uint32_t)(t0 >> 32);
    od[j] = (uint32_t)t1;
    od[j + 1] = (uint32_t)(t1 >> 32);
  }
  t[0] = ev[0];
#pragma unroll
  for (int i = 1; i < 12; i++) {
    uint32_t* E = (i & 1) ? od : ev;  // even-aligned accumulator of this row
    uint32_t* X = (i & 1) ? ev : od;  // previous even accumulator: its limb 0 is already in t[i-1]
    row_mac_odd_shift(X, E[0], a.l, b.l[i]);
    row_mac_even(E, X[11], a.l, b.l[i]);
    t[i] = E[0];
  }
  // last row had E = od, X = ev: high half = ev + (od >> 32)
  {
    PX_DECL;
    PX_ADD_CC(t[12], ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 11; k++) PX_ADDC_CC(t[12 + k], ev[k], od[k + 1]);
    PX_ADDC(t[23], ev[11], 0u);
  }
}

// ---- one level of Karatsuba on the 12 x 12 limb product -------------------------------------------------------------
// a = aL + aH W^6, b = bL + bH W^6:  a b = PL + (PM - PL - PH) W^6 + PH W^12 with PL = aL bL, PH = aH bH and
// PM = (aL + aH)(bL + bH): three 6 x 6 products = 108 wide MADs instead of 144.  The kernels are bound by the
// multiplier pipe (a wide MAD occupies it for 4 cycles per warp, an add issues in 1), so ~100 extra additions for 36
// fewer MADs is a gain.  The half sums are 6 limbs + a carry bit; the carry bits are folded in as masked additions.
//
// t[0..11] = x * y for 6-limb x, y: the even/odd rows of fq_mul_wide at half width.
PTAU_HD void mul6_wide(uint32_t* t, const uint32_t* x, const uint32_t* y) {
  uint32_t ev[6], od[6];
#pragma unroll
  for (int j = 0; j < 6; j += 2) {
    uint64_t t0 = (uint64_t)x[j] * y[0];
    uint64_t t1 = (uint64_t)x[j + 1] * y[0];
    ev[j] = (uint32_t)t0;
    ev[j + 1] = (uint32_t)(t0 >> 32);
    od[j] = (uint32_t)t1;
    od[j + 1] = (uint32_t)(t1 >> 32);
  }
  t[0] = ev[0];
#pragma unroll
  for (int i = 1; i < 6; i++) {
    uint32_t* E = (i & 1) ? od : ev;
    uint32_t* X = (i & 1) ? ev : od;
    const uint32_t bi = y[i];
    PX_DECL;
    PX_ADD_CC(E[0], E[0], X[1]);
    PX_MADC_LO_CC(X[0], x[1], bi, X[2]);
    PX_MADC_HI_CC(X[1], x[1], bi, X[3]);
    PX_MADC_LO_CC(X[2], x[3], bi, X[4]);
    PX_MADC_HI_CC(X[3], x[3], bi, X[5]);
    PX_MADC_LO_CC(X[4], x[5], bi, 0u);
    PX_MADC_HI_CC(X[5], x[5], bi, 0u);  // the window value stays below W^7: nothing leaves X[5]
    PX_MAD_LO_CC(E[0], x[0], bi, E[0]);
    PX_MADC_HI_CC(E[1], x[0], bi, E[1]);
    PX_MADC_LO_CC(E[2], x[2], bi, E[2]);
    PX_MADC_HI_CC(E[3], x[2], bi, E[3]);
    PX_MADC_LO_CC(E[4], x[4], bi, E[4]);
    PX_MADC_HI_CC(E[5], x[4], bi, E[5]);
    PX_ADDC(X[5], X[5], 0u);
    t[i] = E[0];
  }
  {  // last row had E = od, X = ev: high half = ev + (od >> 32)
    PX_DECL;
    PX_ADD_CC(t[6], ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 5; k++) PX_ADDC_CC(t[6 + k], ev[k], od[k + 1]);
    PX_ADDC(t[11], ev[5], 0u);
  }
}

// t[0..23] = a * b by one level of Karatsuba (a, b any 384-bit values)
PTAU_HD void fq_mul_wide_k(uint32_t* t, const Fq& a, const Fq& b) {
  uint32_t sa[6], sb[6], ca, cb;
  {
    PX_DECL;
    PX_ADD_CC(sa[0], a.l[0], a.l[6]);
#pragma unroll
    for (int i = 1; i < 6; i++) PX_ADDC_CC(sa[i], a.l[i], a.l[6 + i]);
    PX_ADDC(ca, 0u, 0u);
  }
  {
    PX_DECL;
    PX_ADD_CC(sb[0], b.l[0], b.l[6]);
#pragma unroll
    for (int i = 1; i < 6; i++) PX_ADDC_CC(sb[i], b.l[i], b.l[6 + i]);
    PX_ADDC(cb, 0u, 0u);
  }
  uint32_t m[13];
  mul6_wide(t, a.l, b.l);            // PL -> t[0..11]
  mul6_wide(t + 12, a.l + 6, b.l + 6);  // PH -> t[12..23]
  mul6_wide(m, sa, sb);              // low 12 limbs of PM
  {  // + (ca ? sb : 0) W^6 + (cb ? sa : 0) W^6 + (ca & cb) W^12
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    uint32_t c1, c2;
    {
      PX_DECL;
      PX_ADD_CC(m[6], m[6], sb[0] & ma);
#pragma unroll
      for (int i = 1; i < 6; i++) PX_ADDC_CC(m[6 + i], m[6 + i], sb[i] & ma);
      PX_ADDC(c1, 0u, 0u);
    }
    {
      PX_DECL;
      PX_ADD_CC(m[6], m[6], sa[0] & mb);
#pragma unroll
      for (int i = 1; i < 6; i++) PX_ADDC_CC(m[6 + i], m[6 + i], sa[i] & mb);
      PX_ADDC(c2, 0u, 0u);
    }
    m[12] = c1 + c2 + (ca & cb);
  }
  {  // PM - PL - PH >= 0 (13 limbs)
    PX_DECL;
    PX_SUB_CC(m[0], m[0], t[0]);
#pragma unroll
    for (int i = 1; i < 12; i++) PX_SUBC_CC(m[i], m[i], t[i]);
    PX_SUBC(m[12], m[12], 0u);
  }
  {
    PX_DECL;
    PX_SUB_CC(m[0], m[0], t[12]);
#pragma unroll
    for (int i = 1; i < 12; i++) PX_SUBC_CC(m[i], m[i], t[12 + i]);
    PX_SUBC(m[12], m[12], 0u);
  }
  {  // t += middle * W^6 (no carry leaves limb 23: the total is a * b < 2^768)
    PX_DECL;
    PX_ADD_CC(t[6], t[6], m[0]);
#pragma unroll
    for (int i = 1; i < 13; i++) PX_ADDC_CC(t[6 + i], t[6 + i], m[i]);
#pragma unroll
    for (int i = 19; i < 23; i++) PX_ADDC_CC(t[i], t[i], 0u);
    PX_ADDC(t[23], t[23], 0u);
  }
}

// t[0..23] = a * a, a < 2^383 (so that the doubled multiplicand fits 384 bits): the rows of fq_sqr_inl
// (78 product MADs) without the reduction rows.
PTAU_HD void fq_sqr_wide(uint32_t* t, const Fq& a) {
  uint32_t ev[12], od[12];
  uint32_t c2[12], d1[12];
  c2[0] = a.l[0];
  d1[0] = a.l[0];
#pragma unroll
  for (int j = 1; j < 12; j++) {
    d1[j] = a.l[j] << 1;
    c2[j] = (a.l[j] << 1) | (a.l[j - 1] >> 31);
  }
#define SQ_M(i, j) ((j) == (i) ? a.l[j] : ((j) == (i) + 1 ? d1[j] : c2[j]))
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    uint64_t t0 = (uint64_t)SQ_M(0, j) * a.l[0];
    uint64_t t1 = (uint64_t)SQ_M(0, j + 1) * a.l[0];
    ev[j] = (uint32_t)t0;
    ev[j + 1] = (uint32_t)(t0 >> 32);
    od[j] = (uint32_t)t1;
    od[j + 1] = (uint32_t)(t1 >> 32);
  }
  t[0] = ev[0];
#pragma unroll
  for (int i = 1; i < 12; i++) {
    uint32_t* E = (i & 1) ? od : ev;
    uint32_t* X = (i & 1) ? ev : od;
    const uint32_t bi = a.l[i];
    PX_DECL;
    PX_ADD_CC(E[0], E[0], X[1]);
#pragma unroll
    for (int k = 0; k < 10; k += 2) {
      if (k + 1 >= i) {
        PX_MADC_LO_CC(X[k], SQ_M(i, k + 1), bi, X[k + 2]);
        PX_MADC_HI_CC(X[k + 1], SQ_M(i, k + 1), bi, X[k + 3]);
      } else {
        PX_ADDC_CC(X[k], X[k + 2], 0u);
        PX_ADDC_CC(X[k + 1], X[k + 3], 0u);
      }
    }
    PX_MADC_LO_CC(X[10], SQ_M(i, 11), bi, 0u);
    PX_MADC_HI_CC(X[11], SQ_M(i, 11), bi, 0u);
    const int j0 = (i & 1) ? i + 1 : i;
    if (j0 <= 10) {
      PX_MAD_LO_CC(E[j0], SQ_M(i, j0), bi, E[j0]);
      PX_MADC_HI_CC(E[j0 + 1], SQ_M(i, j0), bi, E[j0 + 1]);
#pragma unroll
      for (int j = j0 + 2; j < 12; j += 2) {
        PX_MADC_LO_CC(E[j], SQ_M(i, j), bi, E[j]);
        PX_MADC_HI_CC(E[j + 1], SQ_M(i, j), bi, E[j + 1]);
      }
      PX_ADDC(X[11], X[11], 0u);
    }
    t[i] = E[0];
  }
#undef SQ_M
  {
    PX_DECL;
    PX_ADD_CC(t[12], ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 11; k++) PX_ADDC_CC(t[12 + k], ev[k], od[k + 1]);
    PX_ADDC(t[23], ev[11], 0u);
  }
}

// the wide product the rest of the code uses
PTAU_HD void fq_mul_wide(uint32_t* t, const Fq& a, const Fq& b) {
  // Measured on B200 (profiles/r02_ab_karatsuba.log): the Karatsuba product saves 25 % of the MADs of a wide product
  // and makes every kernel slower or no faster (G2 compressed 77.6 -> 79.8 ms per 2^20 points) -- at two warps per
  // scheduler the extra dependent carry chains cost more latency than the multiplier pipe gains.  Kept (and
  // tested on the host) as an opt-in.
#ifdef PTAU_KARATSUBA
  fq_mul_wide_k(t, a, b);
#else
  fq_mul_wide_plain(t, a, b);
#endif
}

// a += b, a -= b over 24 limbs (callers guarantee no carry / borrow out, except fqw_sub_fix)
PTAU_HD void fqw_add(uint32_t* a, const uint32_t* b) {
  PX_DECL;
  PX_ADD_CC(a[0], a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 23; i++) PX_ADDC_CC(a[i], a[i], b[i]);
  PX_ADDC(a[23], a[23], b[23]);
}
PTAU_HD void fqw_sub(uint32_t* a, const uint32_t* b) {
  PX_DECL;
  PX_SUB_CC(a[0], a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 23; i++) PX_SUBC_CC(a[i], a[i], b[i]);
  PX_SUBC(a[23], a[23], b[23]);
}
// a <<= 3 over 24 limbs (the caller guarantees a < 2^765)
PTAU_HD void fqw_shl3(uint32_t* a) {
#pragma unroll
  for (int i = 23; i > 0; i--) a[i] = (a[i] << 3) | (a[i - 1] >> 29);
  a[0] <<= 3;
}
// a = a - b + (a < b ? p * 2^384 : 0): the difference of two products, made non-negative (< p * 2^384)
PTAU_HD void fqw_sub_fix(uint32_t* a, const uint32_t* b) {
  const uint32_t pl[12] = PTAU_P_LIMBS;
  uint32_t mask;
  {
    PX_DECL;
    PX_SUB_CC(a[0], a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 24; i++) PX_SUBC_CC(a[i], a[i], b[i]);
    PX_SUBC(mask, 0u, 0u);  // 0xffffffff on borrow
  }
  {
    PX_DECL;
    PX_ADD_CC(a[12], a[12], pl[0] & mask);
#pragma unroll
    for (int i = 1; i < 11; i++) PX_ADDC_CC(a[12 + i], a[12 + i], pl[i] & mask);
    PX_ADDC(a[23], a[23], pl[11] & mask);
  }
}

// Montgomery reduction: T / 2^384 mod p for T = t[0..23] < p * 2^384; result < p.
// Window of 12 (+1) limbs sliding up one limb per row; limb 12+i-1... of T enters at the top of the window in row i.
PTAU_HD Fq fq_redc(const uint32_t* t) {
  uint32_t ev[12], od[12];
#pragma unroll
  for (int j = 0; j < 12; j++) {
    ev[j] = t[j];
    od[j] = 0;
  }
  {
    uint32_t m = ev[0] * PTAU_M0;
    row_red_odd(od, m);
    row_red_even(ev, od[11], m);
  }
#pragma unroll
  for (int i = 1; i < 12; i++) {
    uint32_t* E = (i & 1) ? od : ev;
    uint32_t* X = (i & 1) ? ev : od;
    uint32_t m;
    // limb 11+i of T enters at window position 11 (= X[10]) as the addend of the top product; by the bound
    // T + sum m p W^i < 2 p W^12 nothing leaves X[11]
    row_red_odd_shift(X, E[0], m, t[11 + i]);
    row_red_even(E, X[11], m);
  }
  // last row had E = od, X = ev: result = ev + (od >> 32) + t[23] * W^11
  {
    PX_DECL;
    PX_ADD_CC(ev[0], ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 11; k++) PX_ADDC_CC(ev[k], ev[k], od[k + 1]);
    PX_ADDC(ev[11], ev[11], t[23]);
  }
  fq_cond_sub_p(ev);
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = ev[i];
  return r;
}

// Montgomery product as Karatsuba wide product + separate reduction (264 MADs instead of the 300 of the interleaved
// fq_mul_inl, ~90 more additions): PTAU_MUL_K_EAGER routes the reduced multiplications through it.
PTAU_HD Fq fq_mul_kr(const Fq& a, const Fq& b) {
  uint32_t t[24];
  fq_mul_wide_k(t, a, b);
  return fq_redc(t);
}

}  // namespace ptau
