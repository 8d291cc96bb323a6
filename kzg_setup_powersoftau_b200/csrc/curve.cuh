// Short-Weierstrass (a = 0) Jacobian arithmetic, generic over F in {Fq, Fq2},
// and the two endomorphism-based subgroup predicates.
//
// What this replaces: ark-ec 0.2.0 is_in_correct_subgroup_assuming_on_curve,
// reached from /root/reference/src/lib.rs:52 (G1) and :78 (G2), which multiplies
// by r (254 doublings + 133 mixed additions).  Here:
//   G1: phi(P) == -[z^2]P   with phi(x,y) = (beta x, y)        (2 x 64-bit ladders)
//   G2: psi(P) == [z]P      with psi = twist o Frobenius o untwist (1 x 64-bit ladder)
// Both return the same boolean as the r-multiplication for every on-curve point:
// phi^2+phi+1 = 0 gives (z^4-z^2+1)P = rP = O; psi^2-t psi+p = 0 gives
// (p-z)P = h1 r P = O and gcd(h1, h2) = 1 (checked in tests/test_oracle_pins.py).
//
// Exceptional cases: the addition formulas used here give Z3 = 0 whenever the two
// inputs share an x coordinate (H = 0), and Z = 0 is sticky through every later
// doubling/addition (Z3 is a multiple of Z1).  A point of prime order r never
// meets such a case on these ladders (all partial scalars are < 2^128 and
// != +-1 mod r), so "some Z became 0" implies "not in the subgroup", and the
// final comparison demands Z != 0.  Hence no branches on exceptional cases.
#pragma once
#include "fq2.cuh"

namespace ptau {

template <class F>
struct Jac {
  F X, Y, Z;
};

// dbl-2009-l  (2M + 5S).  Statement order chosen for register liveness: at most
// four field elements are live across any multiplication.
template <class F>
PTAU_HD void jac_dbl(Jac<F>& p) {
  F B = fsqr(p.Y);
  p.Z = fdbl(fmul(p.Z, p.Y));  // Y is dead from here
  F C = fsqr(B);
  F t = fadd(p.X, B);          // B dead
  F A = fsqr(p.X);             // X dead
  F D = fsqr(t);
  D = fsub(fsub(D, A), C);
  D = fdbl(D);
  F E = fadd(fdbl(A), A);      // A dead
  F Fv = fsqr(E);
  p.X = fsub(Fv, fdbl(D));
  C = fdbl(fdbl(fdbl(C)));
  p.Y = fsub(fmul(fsub(D, p.X), E), C);
}

// madd-2007-bl  (7M + 4S): p += (x2, y2)
template <class F>
PTAU_HD void jac_madd(Jac<F>& p, const F& x2, const F& y2) {
  F Z1Z1 = fsqr(p.Z);
  F U2 = fmul(x2, Z1Z1);
  F S2 = fmul(fmul(y2, p.Z), Z1Z1);
  F H = fsub(U2, p.X);
  F HH = fsqr(H);
  F I = fdbl(fdbl(HH));
  F J = fmul(H, I);
  F rr = fdbl(fsub(S2, p.Y));
  F V = fmul(p.X, I);
  F X3 = fsub(fsub(fsqr(rr), J), fdbl(V));
  F Y3 = fsub(fmul(rr, fsub(V, X3)), fdbl(fmul(p.Y, J)));
  // Z3 = (Z1+H)^2 - Z1Z1 - HH = 2 Z1 H
  p.Z = fdbl(fmul(p.Z, H));
  p.X = X3;
  p.Y = Y3;
}

// add-2007-bl  (11M + 5S): p += q
template <class F>
PTAU_HD void jac_add(Jac<F>& p, const Jac<F>& q) {
  F Z1Z1 = fsqr(p.Z);
  F Z2Z2 = fsqr(q.Z);
  F U1 = fmul(p.X, Z2Z2);
  F U2 = fmul(q.X, Z1Z1);
  F S1 = fmul(fmul(p.Y, q.Z), Z2Z2);
  F S2 = fmul(fmul(q.Y, p.Z), Z1Z1);
  F H = fsub(U2, U1);
  F I = fsqr(fdbl(H));
  F J = fmul(H, I);
  F rr = fdbl(fsub(S2, S1));
  F V = fmul(U1, I);
  F X3 = fsub(fsub(fsqr(rr), J), fdbl(V));
  F Y3 = fsub(fmul(rr, fsub(V, X3)), fdbl(fmul(S1, J)));
  // Z3 = ((Z1+Z2)^2 - Z1Z1 - Z2Z2) H = 2 Z1 Z2 H
  p.Z = fmul(fdbl(fmul(p.Z, q.Z)), H);
  p.X = X3;
  p.Y = Y3;
}

// |z| = 0xd201000000010000 = 2^63 + 2^62 + 2^60 + 2^57 + 2^48 + 2^16
#define PTAU_Z_ABS 0xd201000000010000ull

// The doubling of the G1 ladders.  PTAU_G1_DBL_INLINE (device) expands the seven field multiplications in place
// instead of calling fq_mul / fq_sqr: no argument shuffles, at the price of a ~45 KB loop body.
//
// Sums that only feed multiplications stay unreduced (fq_add_nored): the Montgomery product accepts operands up to
// 3p (9 p^2 < p 2^384) and the dedicated squaring operands below 2^383, so X + B < 2p, 3A < 3p and Z3 = 2YZ < 2p need
// no conditional subtraction.  Z stays in [0, 2p) across the ladder; it is only ever multiplied, squared or tested for
// zero (2YZ = p is impossible for YZ < p, p odd).  Saves 100 of ~2900 instructions per doubling: the ladders are
// bound by instruction issue as much as by the multiplier.  The host build runs the same formula (tests/host_emul).
#if defined(__CUDA_ARCH__) && defined(PTAU_G1_DBL_INLINE)
#define PTAU_LAD_MUL(a, b) PTAU_FQ_MUL_IMPL(a, b)
#define PTAU_LAD_SQR(a) fq_sqr_inl(a)
#else
#define PTAU_LAD_MUL(a, b) fq_mul(a, b)
#define PTAU_LAD_SQR(a) fq_sqr(a)
#endif
PTAU_HD void jac_dbl_ladder(Jac<Fq>& p) {
#if !defined(PTAU_DBL_EAGER_ADDS) && !defined(PTAU_DBL_NO_SHARED_RED)
  // dbl-2009-l with two of its seven Montgomery reductions shared (fqw.cuh): C = Y^4 is only ever used inside the
  // sums D = 2((X+B)^2 - A - C) and Y3 = E (D - X3) - 8C, so it stays an unreduced 768-bit square and each sum is
  // reduced once:  7 products + 6 reductions instead of 7 + 7 (1614 MADs instead of 1770).
  //   (X+B)^2 - X^2 - B^2 = 2 X B exactly (the sum X + B is left unreduced), so that difference needs no sign fix;
  //   E (D - X3) - 8C may be negative: fqw_sub_fix adds p 2^384, keeping the value below p 2^384 (E < 3p, 8C < 8p^2).
  Fq B = PTAU_LAD_SQR(p.Y);
  Fq zy = PTAU_LAD_MUL(p.Z, p.Y);
  p.Z = fq_add_nored(zy, zy);
  uint32_t wC[24], wA[24], wS[24];
  fq_sqr_wide(wC, B);
  Fq t = fq_add_nored(p.X, B);
  fq_sqr_wide(wA, p.X);
  fq_sqr_wide(wS, t);
  fqw_sub(wS, wA);
  fqw_sub(wS, wC);
  Fq A = fq_redc(wA);
  Fq D = fq_redc(wS);  // S - A - C
  D = fq_dbl(D);
  Fq E = fq_add_nored(fq_add_nored(A, A), A);
  Fq Fv = PTAU_LAD_SQR(E);
  p.X = fq_sub(Fv, fq_dbl(D));
  fq_mul_wide(wS, fq_sub(D, p.X), E);
  fqw_shl3(wC);
  fqw_sub_fix(wS, wC);
  p.Y = fq_redc(wS);
#elif !defined(PTAU_DBL_EAGER_ADDS)
  Fq B = PTAU_LAD_SQR(p.Y);
  Fq zy = PTAU_LAD_MUL(p.Z, p.Y);
  p.Z = fq_add_nored(zy, zy);
  Fq C = PTAU_LAD_SQR(B);
  Fq t = fq_add_nored(p.X, B);
  Fq A = PTAU_LAD_SQR(p.X);
  Fq D = PTAU_LAD_SQR(t);
  D = fq_sub(fq_sub(D, A), C);
  D = fq_dbl(D);
  Fq E = fq_add_nored(fq_add_nored(A, A), A);
  Fq Fv = PTAU_LAD_SQR(E);
  p.X = fq_sub(Fv, fq_dbl(D));
  C = fq_dbl(fq_dbl(fq_dbl(C)));
  p.Y = fq_sub(PTAU_LAD_MUL(fq_sub(D, p.X), E), C);
#else
  Fq B = PTAU_LAD_SQR(p.Y);
  p.Z = fq_dbl(PTAU_LAD_MUL(p.Z, p.Y));
  Fq C = PTAU_LAD_SQR(B);
  Fq t = fq_add(p.X, B);
  Fq A = PTAU_LAD_SQR(p.X);
  Fq D = PTAU_LAD_SQR(t);
  D = fq_sub(fq_sub(D, A), C);
  D = fq_dbl(D);
  Fq E = fq_add(fq_dbl(A), A);
  Fq Fv = PTAU_LAD_SQR(E);
  p.X = fq_sub(Fv, fq_dbl(D));
  C = fq_dbl(fq_dbl(fq_dbl(C)));
  p.Y = fq_sub(PTAU_LAD_MUL(fq_sub(D, p.X), E), C);
#endif
}
#if defined(__CUDA_ARCH__) && defined(PTAU_G2_DBL_INLINE)
// experiment: the G2 doubling with its seven Fq2 multiplications expanded in place (A/B knob)
PTAU_HD void jac_dbl_ladder(Jac<Fq2>& p) {
  Fq2 B = fq2_sqr_inl(p.Y);
  p.Z = fq2_dbl(fq2_mul_inl(p.Z, p.Y));
  Fq2 C = fq2_sqr_inl(B);
  Fq2 t = fq2_add(p.X, B);
  Fq2 A = fq2_sqr_inl(p.X);
  Fq2 D = fq2_sqr_inl(t);
  D = fq2_sub(fq2_sub(D, A), C);
  D = fq2_dbl(D);
  Fq2 E = fq2_add(fq2_dbl(A), A);
  Fq2 Fv = fq2_sqr_inl(E);
  p.X = fq2_sub(Fv, fq2_dbl(D));
  C = fq2_dbl(fq2_dbl(fq2_dbl(C)));
  p.Y = fq2_sub(fq2_mul_inl(fq2_sub(D, p.X), E), C);
}
#else
PTAU_HD void jac_dbl_ladder(Jac<Fq2>& p) { jac_dbl(p); }
#endif

// acc = [|z|] (x, y), affine base, mixed additions
template <class F>
PTAU_HD Jac<F> mul_zabs_affine(const F& x, const F& y, const F& one) {
  Jac<F> acc;
  acc.X = x;
  acc.Y = y;
  acc.Z = one;
#pragma unroll 1
  for (int i = 62; i >= 0; --i) {
    jac_dbl_ladder(acc);
    if ((PTAU_Z_ABS >> i) & 1ull) jac_madd(acc, x, y);
  }
  return acc;
}

// acc = [|z|] q, Jacobian base, full additions
template <class F>
PTAU_HD Jac<F> mul_zabs_jac(const Jac<F>& q) {
  Jac<F> acc = q;
#pragma unroll 1
  for (int i = 62; i >= 0; --i) {
    jac_dbl_ladder(acc);
    if ((PTAU_Z_ABS >> i) & 1ull) jac_add(acc, q);
  }
  return acc;
}

// Jacobian p equals the finite affine point (x, y)?
template <class F>
PTAU_HD bool jac_eq_affine(const Jac<F>& p, const F& x, const F& y) {
  F ZZ = fsqr(p.Z);
  F ZZZ = fmul(ZZ, p.Z);
  bool ok = !fis_zero(p.Z);
  ok = ok && feq(p.X, fmul(x, ZZ));
  ok = ok && feq(p.Y, fmul(y, ZZZ));
  return ok;
}

// ---- constants (Montgomery form) --------------------------------------------
#include "consts.inc"

// ---- Fermat inversion a^(p-2), host- and device-callable -------------------------------------------------
#define PTAU_PM2_INIT                                                                                       \
  {0xffffaaa9u, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, \
   0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
#ifdef __CUDACC__
__constant__ uint32_t K_PM2_PAIR_D[12] = PTAU_PM2_INIT;
#endif
static const uint32_t K_PM2_PAIR_H[12] = PTAU_PM2_INIT;
PTAU_HD_NOINLINE Fq fq_inv_fermat(Fq a) {
#ifdef __CUDA_ARCH__
  const uint32_t* e = K_PM2_PAIR_D;
#else
  const uint32_t* e = K_PM2_PAIR_H;
#endif
  Fq acc = a;
#pragma unroll 1
  for (int i = 379; i >= 0; --i) {
    acc = fq_sqr(acc);
    if ((e[i >> 5] >> (i & 31)) & 1u) acc = fq_mul(acc, a);
  }
  return acc;
}


// y^2 == x^3 + 4
PTAU_HD_NOINLINE bool g1_on_curve(const Fq& x, const Fq& y) {
  Fq rhs = fq_add(fq_mul(fq_sqr(x), x), k_b1_mont());
  return fq_eq(fq_sqr(y), rhs);
}

// y^2 == x^3 + 4(1+u)
PTAU_HD_NOINLINE bool g2_on_curve(const Fq2& x, const Fq2& y) {
  Fq2 b;
  b.c0 = k_b1_mont();
  b.c1 = b.c0;
  Fq2 rhs = fq2_add(fq2_mul(fq2_sqr(x), x), b);
  return fq2_eq(fq2_sqr(y), rhs);
}

// phi(P) == -[z^2]P   <=>   [|z|]([|z|]P) == (beta x, -y)
PTAU_HD_NOINLINE bool g1_in_subgroup(const Fq& x, const Fq& y) {
  Jac<Fq> q = mul_zabs_affine(x, y, fq_one());
  Jac<Fq> q2 = mul_zabs_jac(q);
  return jac_eq_affine(q2, fq_mul(x, k_beta_mont()), fq_neg(y));
}

// Operand file of the G2 ladder: the base point (x, y) is needed only at the five additions and in the final
// comparison, so it is parked outside the register file and the 63 doublings run with the accumulator and the
// temporaries of the field operations only.  On the device the file is shared memory, transposed (word j of a
// thread at byte address base + j * 4 * STRIDE, STRIDE = block size: conflict-free) and addressed in the shared
// window directly (ld.shared / st.shared); on the host it is a plain array.
// Layout: x.c0 | x.c1 | y.c0 | y.c1, 12 words each.
template <int STRIDE>
struct Park {
#ifdef __CUDA_ARCH__
  uint32_t base;  // shared-window byte address of this thread's word 0
  __device__ __forceinline__ uint32_t ld(int j) const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + (uint32_t)(j * 4 * STRIDE)) : "memory");
    return v;
  }
  __device__ __forceinline__ void st(int j, uint32_t v) const {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + (uint32_t)(j * 4 * STRIDE)), "r"(v) : "memory");
  }
#else
  uint32_t* p;
  uint32_t ld(int j) const { return p[j * STRIDE]; }
  void st(int j, uint32_t v) const { p[j * STRIDE] = v; }
#endif
  PTAU_HD void store_fq(int at, const Fq& a) const {
#pragma unroll
    for (int j = 0; j < 12; j++) st(at + j, a.l[j]);
  }
  PTAU_HD Fq load_fq(int at) const {
    Fq r;
#pragma unroll
    for (int j = 0; j < 12; j++) r.l[j] = ld(at + j);
    return r;
  }
  PTAU_HD Fq2 load_fq2(int at) const {
    Fq2 r;
    r.c0 = load_fq(at);
    r.c1 = load_fq(at + 12);
    return r;
  }
  PTAU_HD void store_g2(const Fq2& x, const Fq2& y) const {
    store_fq(0, x.c0);
    store_fq(12, x.c1);
    store_fq(24, y.c0);
    store_fq(36, y.c1);
  }
  // 24-limb (unreduced) values, words 48..95 of the file: only with PTAU_G2_LAZYC_PARK
  PTAU_HD void store_wide(int at, const uint32_t* t) const {
#pragma unroll
    for (int j = 0; j < 24; j++) st(at + j, t[j]);
  }
  PTAU_HD void load_wide(int at, uint32_t* t) const {
#pragma unroll
    for (int j = 0; j < 24; j++) t[j] = ld(at + j);
  }
};
// words per thread of the operand file
#ifdef PTAU_G2_LAZYC_PARK
#define PTAU_PARK_WORDS 96
#else
#define PTAU_PARK_WORDS 48
#endif

// The doubling of the G2 ladder with C = Y^4 never reduced (the Fq2 counterpart of the G1 doubling above), in
// coordinates scaled by lambda = 1/2 -- (X3/4, Y3/8, Z3/2) is the same Jacobian point as dbl-2009-l's (X3, Y3, Z3):
//     B = Y^2, C = B^2, A = X^2, D = (X+B)^2 - A - C = 2XB, M = 3A/2, S = D/2
//     X3 = M^2 - D,  Y3 = M (S - X3) - C,  Z3 = Y Z
// so C enters Y3 with coefficient 1 and every wide difference below needs at most one conditional + p 2^384.
// An Fq2 square is taken as (a0+a1)(a0-a1+p) + (2 a0) a1 u with unreduced operands; the two wide values of C stay in
// registers (48 limbs; A/B: parked in the operand file) and are used twice, inside the sums that are reduced instead of C:
//     D.c1 = redc(2 t0 t1 - 2 x0 x1 - 2 b0 b1)                        = 2 (x0 b1 + x1 b0)            in [0, 4p^2)
//     D.c0 = redc((t0+t1)(t0-t1+2p) - (x0+x1)(x0-x1+p) - (b0+b1)(b0-b1+p))
//          = 2 x0 b0 - 2 x1 b1 + p (x0+x1+b0+b1)                                                     in [0, 6p^2)
// with t = X + B the unreduced sum (t_i = x_i + b_i as integers, which makes both identities exact: no sign fix), and
//     Y3.c1 = redc(m0 u1 + m1 u0 - C.c1)   in (-2p^2, 2p^2) + fix;   Y3.c0 = redc(m0 u0 - (m1 u1 + C.c0))  in (-5p^2, p^2) + fix
// 16 wide products + 12 reductions = 4176 MADs instead of 16 + 14 = 4488, and 4 halvings + 9 field additions
// instead of 26 field additions.  All bounds are < p 2^384 = 9.84 p^2 (fq_redc) and every operand is < 2^384.
template <int STRIDE>
PTAU_HD void jac_dbl_lazyc(Jac<Fq2>& p, const Park<STRIDE>& pk) {
#if defined(__CUDA_ARCH__) && defined(PTAU_G2_LAZYC_CALLS)  // A/B: the three plain Fq2 products as calls (smaller loop body)
#define PTAU_LZ_SQR(a) fq2_sqr(a)
#define PTAU_LZ_MUL(a, b) fq2_mul(a, b)
#else
#define PTAU_LZ_SQR(a) fq2_sqr_inl(a)
#define PTAU_LZ_MUL(a, b) fq2_mul_inl(a, b)
#endif
  Fq2 B = PTAU_LZ_SQR(p.Y);
  p.Z = PTAU_LZ_MUL(p.Z, p.Y);
  uint32_t w0[24], w1[24];
  // C (never reduced)
#ifndef PTAU_G2_LAZYC_PARK  // the wide C stays in registers (no spills at 255 registers: -Xptxas -v)
  uint32_t wc0[24], wc1[24];
  fq_mul_wide(wc0, fq_add_nored(B.c0, B.c1), fq_sub_plus_p(B.c0, B.c1));
  fq_mul_wide(wc1, fq_add_nored(B.c0, B.c0), B.c1);
#define PTAU_LOAD_C0(dst) do { for (int j_ = 0; j_ < 24; j_++) (dst)[j_] = wc0[j_]; } while (0)
#define PTAU_LOAD_C1(dst) do { for (int j_ = 0; j_ < 24; j_++) (dst)[j_] = wc1[j_]; } while (0)
#else  // A/B: C parked in the operand file (words 48..95) -- measured slower, profiles/r02c_ab_*.log
  fq_mul_wide(w0, fq_add_nored(B.c0, B.c1), fq_sub_plus_p(B.c0, B.c1));
  pk.store_wide(48, w0);
  fq_mul_wide(w0, fq_add_nored(B.c0, B.c0), B.c1);
  pk.store_wide(72, w0);
#define PTAU_LOAD_C0(dst) pk.load_wide(48, dst)
#define PTAU_LOAD_C1(dst) pk.load_wide(72, dst)
#endif
  Fq2 t;
  t.c0 = fq_add_nored(p.X.c0, B.c0);
  t.c1 = fq_add_nored(p.X.c1, B.c1);
  Fq2 A, D;
  // imaginary parts: A.c1 = 2 x0 x1, D.c1
  fq_mul_wide(w0, fq_add_nored(p.X.c0, p.X.c0), p.X.c1);
  fq_mul_wide(w1, fq_add_nored(t.c0, t.c0), t.c1);
  fqw_sub(w1, w0);
  A.c1 = fq_redc(w0);
  PTAU_LOAD_C1(w0);
  fqw_sub(w1, w0);
  D.c1 = fq_redc(w1);
  // real parts
  fq_mul_wide(w0, fq_add_nored(p.X.c0, p.X.c1), fq_sub_plus_p(p.X.c0, p.X.c1));
  fq_mul_wide(w1, fq_add_nored(t.c0, t.c1), fq_sub_plus_2p(t.c0, t.c1));
  fqw_sub(w1, w0);
  A.c0 = fq_redc(w0);
  PTAU_LOAD_C0(w0);
  fqw_sub(w1, w0);
  D.c0 = fq_redc(w1);
  // M = 3A/2 = A + A/2
  Fq2 M;
  M.c0 = fq_add(A.c0, fq_half(A.c0));
  M.c1 = fq_add(A.c1, fq_half(A.c1));
  Fq2 Fv = PTAU_LZ_SQR(M);
  p.X = fq2_sub(Fv, D);
  Fq2 U;
  U.c0 = fq_sub(fq_half(D.c0), p.X.c0);
  U.c1 = fq_sub(fq_half(D.c1), p.X.c1);
  // Y3 = M U - C: the Karatsuba rows of fq2_mul_inl with C subtracted before the two reductions
  uint32_t w2[24];
  fq_mul_wide(w2, fq_add_nored(M.c0, M.c1), fq_add_nored(U.c0, U.c1));
  fq_mul_wide(w0, M.c0, U.c0);
  fq_mul_wide(w1, M.c1, U.c1);
  fqw_sub(w2, w0);
  fqw_sub(w2, w1);
  {
    uint32_t c[24];
    PTAU_LOAD_C1(c);
    fqw_sub_fix(w2, c);
    p.Y.c1 = fq_redc(w2);
    PTAU_LOAD_C0(c);
    fqw_add(w1, c);
  }
  fqw_sub_fix(w0, w1);
  p.Y.c0 = fq_redc(w0);
#undef PTAU_LZ_SQR
#undef PTAU_LZ_MUL
#undef PTAU_LOAD_C0
#undef PTAU_LOAD_C1
}

// psi(P) == [z]P   <=>   [|z|]P == -psi(P) = (psi_x, -psi_y); P = (x, y) parked in `pk`
template <int STRIDE>
PTAU_HD_NOINLINE bool g2_in_subgroup(Park<STRIDE> pk) {
  Jac<Fq2> q;
  q.X = pk.load_fq2(0);
  q.Y = pk.load_fq2(24);
  q.Z = fq2_one();
#pragma unroll 1
  for (int i = 62; i >= 0; --i) {
#ifndef PTAU_G2_DBL_EAGERC
    jac_dbl_lazyc(q, pk);
#else  // A/B: dbl-2009-l with every product reduced (the doubling of rounds 1-2)
    jac_dbl_ladder(q);
#endif
    if ((PTAU_Z_ABS >> i) & 1ull) jac_madd(q, pk.load_fq2(0), pk.load_fq2(24));
  }
  // psi_x = conj(x) * (0, cx1) = (x1*cx1, x0*cx1)
  Fq cx1 = k_psi_cx1_mont();
  Fq2 px;
  px.c0 = fq_mul(pk.load_fq(12), cx1);
  px.c1 = fq_mul(pk.load_fq(0), cx1);
  Fq2 cy;
  cy.c0 = k_psi_cy0_mont();
  cy.c1 = k_psi_cy1_mont();
  Fq2 py = fq2_mul(fq2_conj(pk.load_fq2(24)), cy);
  return jac_eq_affine(q, px, fq2_neg(py));
}

// ---------------------------------------------------------------------------
// Reference-exact fallback for G2 points that are NOT on the twist (only reachable
// with PTAU_CHECKS_READ, i.e. when the caller asks for the reference's behaviour and
// not for the on-curve test).  ark-ec 0.2.0 does not test the curve equation in
// deserialize_uncompressed; it multiplies by r with formulas that never use b, so an
// off-curve (x, y) is accepted iff it is r-torsion on the curve y^2 = x^3 + b' it
// happens to lie on.  psi is not an endomorphism of those curves, so the only way to
// reproduce that boolean is the multiplication by r itself:
//   res = 0; for each bit of r, MSB first: res = 2 res; if bit: res = res + P
// with the special cases of add_assign_mixed (res == 0 -> P; res == P -> double).
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
__constant__ uint32_t K_R_ORDER_D[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                        0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
#endif
static const uint32_t K_R_ORDER_H[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                        0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};

PTAU_HD_NOINLINE bool g2_rmul_is_zero(const Fq2& x, const Fq2& y) {
#ifdef __CUDA_ARCH__
  const uint32_t* r = K_R_ORDER_D;
#else
  const uint32_t* r = K_R_ORDER_H;
#endif
  Jac<Fq2> acc;
  acc.X = fq2_zero();
  acc.Y = fq2_one();
  acc.Z = fq2_zero();
#pragma unroll 1
  for (int i = 254; i >= 0; --i) {
    if (!fq2_is_zero(acc.Z)) jac_dbl(acc);  // doubling of zero is zero; a point with Y = 0 doubles to Z = 0
    if ((r[i >> 5] >> (i & 31)) & 1u) {
      if (fq2_is_zero(acc.Z)) {
        acc.X = x;
        acc.Y = y;
        acc.Z = fq2_one();
      } else {
        Fq2 zz = fq2_sqr(acc.Z);
        Fq2 u2 = fq2_mul(x, zz);
        Fq2 s2 = fq2_mul(fq2_mul(y, acc.Z), zz);
        if (fq2_eq(u2, acc.X) && fq2_eq(s2, acc.Y))
          jac_dbl(acc);
        else
          jac_madd(acc, x, y);  // opposite points: H = 0 gives Z3 = 0, i.e. zero
      }
    }
  }
  return fq2_is_zero(acc.Z);
}

}  // namespace ptau
