// Fq2 = Fq[u]/(u^2+1) on top of fq.cuh, plus the field-generic spellings
// (fadd/fsub/fmul/...) that curve.cuh is written against.
//
// On the device fq_mul / fq_sqr / fq2_mul / fq2_sqr are real (non-inlined)
// functions: one Montgomery multiplication is ~330 SASS instructions, and the
// double-and-add loops of the subgroup checks must stay inside the instruction
// cache.  Arguments and results travel in registers (checked with -Xptxas -v:
// no stack frame).
#pragma once
#include "fq.cuh"
#include "fqw.cuh"

namespace ptau {

#ifdef PTAU_MUL_K_EAGER
#define PTAU_FQ_MUL_IMPL(a, b) fq_mul_kr(a, b)
#else
#define PTAU_FQ_MUL_IMPL(a, b) fq_mul_inl(a, b)
#endif
#ifdef __CUDA_ARCH__
__device__ __noinline__ Fq fq_mul(Fq a, Fq b) { return PTAU_FQ_MUL_IMPL(a, b); }
__device__ __noinline__ Fq fq_sqr(Fq a) { return fq_sqr_inl(a); }
#else
inline Fq fq_mul(const Fq& a, const Fq& b) { return PTAU_FQ_MUL_IMPL(a, b); }
inline Fq fq_sqr(const Fq& a) { return fq_sqr_inl(a); }
#endif

struct Fq2 {
  Fq c0, c1;
};

PTAU_HD Fq2 fq2_zero() {
  Fq2 r;
  r.c0 = fq_zero();
  r.c1 = fq_zero();
  return r;
}
PTAU_HD Fq2 fq2_one() {
  Fq2 r;
  r.c0 = fq_one();
  r.c1 = fq_zero();
  return r;
}
PTAU_HD Fq2 fq2_add(const Fq2& a, const Fq2& b) {
  Fq2 r;
  r.c0 = fq_add(a.c0, b.c0);
  r.c1 = fq_add(a.c1, b.c1);
  return r;
}
PTAU_HD Fq2 fq2_sub(const Fq2& a, const Fq2& b) {
  Fq2 r;
  r.c0 = fq_sub(a.c0, b.c0);
  r.c1 = fq_sub(a.c1, b.c1);
  return r;
}
PTAU_HD Fq2 fq2_neg(const Fq2& a) {
  Fq2 r;
  r.c0 = fq_neg(a.c0);
  r.c1 = fq_neg(a.c1);
  return r;
}
PTAU_HD Fq2 fq2_dbl(const Fq2& a) {
  Fq2 r;
  r.c0 = fq_dbl(a.c0);
  r.c1 = fq_dbl(a.c1);
  return r;
}
PTAU_HD Fq2 fq2_conj(const Fq2& a) {
  Fq2 r;
  r.c0 = a.c0;
  r.c1 = fq_neg(a.c1);
  return r;
}
PTAU_HD bool fq2_is_zero(const Fq2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
PTAU_HD bool fq2_eq(const Fq2& a, const Fq2& b) { return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1); }

// PTAU_FQ2_LEAF (device): expand the Fq multiplications inside fq2_mul / fq2_sqr so that
// they are leaf functions (one call level, no nested argument shuffles).
#if defined(__CUDA_ARCH__) && defined(PTAU_FQ2_LEAF)
#define PTAU_FQ2_M(a, b) PTAU_FQ_MUL_IMPL(a, b)
#else
#define PTAU_FQ2_M(a, b) fq_mul(a, b)
#endif

#ifndef PTAU_FQ2_EAGER
// Lazy reduction (fqw.cuh): Karatsuba on unreduced products, one Montgomery reduction per coefficient.
//   c1 = (a0+a1)(b0+b1) - a0 b0 - a1 b1   (sums left unreduced: < 2p, products < 4p^2 < 2^768)
//   c0 = a0 b0 - a1 b1 (+ p 2^384 when negative)
// 3 x 144 + 2 x 156 = 744 MADs instead of 3 x 300.
PTAU_HD Fq2 fq2_mul_inl(const Fq2& a, const Fq2& b) {
  uint32_t S[24], V0[24], V1[24];
  {
    Fq sa = fq_add_nored(a.c0, a.c1);
    Fq sb = fq_add_nored(b.c0, b.c1);
    fq_mul_wide(S, sa, sb);
  }
  fq_mul_wide(V0, a.c0, b.c0);
  fq_mul_wide(V1, a.c1, b.c1);
  fqw_sub(S, V0);
  fqw_sub(S, V1);
  Fq2 r;
  r.c1 = fq_redc(S);
  fqw_sub_fix(V0, V1);
  r.c0 = fq_redc(V0);
  return r;
}
#else
// Karatsuba: 3 Fq multiplications
PTAU_HD Fq2 fq2_mul_inl(const Fq2& a, const Fq2& b) {
  Fq v0 = PTAU_FQ2_M(a.c0, b.c0);
  Fq v1 = PTAU_FQ2_M(a.c1, b.c1);
  Fq s = PTAU_FQ2_M(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1));
  Fq2 r;
  r.c0 = fq_sub(v0, v1);
  r.c1 = fq_sub(fq_sub(s, v0), v1);
  return r;
}
#endif

#ifdef PTAU_FQ2_SQR3
// c0 = a0^2 - a1^2, c1 = (a0+a1)^2 - a0^2 - a1^2: three dedicated wide squarings (78 MADs each) + 2 reductions
// = 546 MADs instead of the 600 of the complex-squaring formula with two full multiplications.
PTAU_HD Fq2 fq2_sqr_inl(const Fq2& a) {
  uint32_t S[24], A[24], B[24];
  {
    Fq s = fq_add_nored(a.c0, a.c1);
    fq_sqr_wide(S, s);
  }
  fq_sqr_wide(A, a.c0);
  fq_sqr_wide(B, a.c1);
  fqw_sub(S, A);
  fqw_sub(S, B);
  Fq2 r;
  r.c1 = fq_redc(S);
  fqw_sub_fix(A, B);
  r.c0 = fq_redc(A);
  return r;
}
#else
// complex squaring: 2 Fq multiplications.  (The three-squarings form above has fewer MADs, 546 against 600, but ~400 more
// ALU instructions; the squaring then becomes issue-bound instead of FMA-pipe-bound -- A/B in profiles/.)
PTAU_HD Fq2 fq2_sqr_inl(const Fq2& a) {
  Fq t = PTAU_FQ2_M(a.c0, a.c1);
  Fq2 r;
  // sum and difference only feed the multiplication: a0 + a1 and a0 - a1 + p, both < 2p with no conditional step, and
  // 2p * 2p < p 2^384 keeps the product inside the multiplier's range (it accepts operands up to 3p)
  r.c0 = PTAU_FQ2_M(fq_add_nored(a.c0, a.c1), fq_sub_plus_p(a.c0, a.c1));
  r.c1 = fq_dbl(t);
  return r;
}
#endif

#ifdef __CUDA_ARCH__
__device__ __noinline__ Fq2 fq2_mul(Fq2 a, Fq2 b) { return fq2_mul_inl(a, b); }
__device__ __noinline__ Fq2 fq2_sqr(Fq2 a) { return fq2_sqr_inl(a); }
#else
inline Fq2 fq2_mul(const Fq2& a, const Fq2& b) { return fq2_mul_inl(a, b); }
inline Fq2 fq2_sqr(const Fq2& a) { return fq2_sqr_inl(a); }
#endif

PTAU_HD Fq2 fq2_mul_fq(const Fq2& a, const Fq& k) {
  Fq2 r;
  r.c0 = fq_mul(a.c0, k);
  r.c1 = fq_mul(a.c1, k);
  return r;
}

// ---- field-generic spellings ------------------------------------------------
PTAU_HD Fq fadd(const Fq& a, const Fq& b) { return fq_add(a, b); }
PTAU_HD Fq fsub(const Fq& a, const Fq& b) { return fq_sub(a, b); }
PTAU_HD Fq fneg(const Fq& a) { return fq_neg(a); }
PTAU_HD Fq fdbl(const Fq& a) { return fq_dbl(a); }
PTAU_HD Fq fmul(const Fq& a, const Fq& b) { return fq_mul(a, b); }
PTAU_HD Fq fsqr(const Fq& a) { return fq_sqr(a); }
PTAU_HD bool fis_zero(const Fq& a) { return fq_is_zero(a); }
PTAU_HD bool feq(const Fq& a, const Fq& b) { return fq_eq(a, b); }

PTAU_HD Fq2 fadd(const Fq2& a, const Fq2& b) { return fq2_add(a, b); }
PTAU_HD Fq2 fsub(const Fq2& a, const Fq2& b) { return fq2_sub(a, b); }
PTAU_HD Fq2 fneg(const Fq2& a) { return fq2_neg(a); }
PTAU_HD Fq2 fdbl(const Fq2& a) { return fq2_dbl(a); }
PTAU_HD Fq2 fmul(const Fq2& a, const Fq2& b) { return fq2_mul(a, b); }
PTAU_HD Fq2 fsqr(const Fq2& a) { return fq2_sqr(a); }
PTAU_HD bool fis_zero(const Fq2& a) { return fq2_is_zero(a); }
PTAU_HD bool feq(const Fq2& a, const Fq2& b) { return fq2_eq(a, b); }

// to / from Montgomery form
PTAU_HD Fq fq_to_mont(const Fq& plain) { return fq_mul(plain, fq_r2()); }
PTAU_HD Fq fq_from_mont(const Fq& m) {
  Fq one = fq_zero();
  one.l[0] = 1;
  return fq_mul(m, one);
}

}  // namespace ptau
