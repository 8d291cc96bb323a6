// BLS12-381 optimal ate pairing on the device: the tower Fq2[v]/(v^3 - (1+u)) = Fq6, Fq6[w]/(w^2 - v) = Fq12,
// the Miller loop with projective line coefficients, and the final exponentiation.
//
// What this restates: ark-ec 0.2.0 models/bls12 (G2Prepared doubling_step / addition_step, ell with the
// M-type twist, miller_loop over BitIterator(|z|) with the final conjugation for z < 0, final_exponentiation)
// as reached from ark-poly-commit 0.2 KZG10::check / batch_check, which is what the reference's consumer
// code calls (/root/reference/src/lib.rs:276-286: KZG10::check(&vk, &comm, point, value, &proof)).
// The hard part of the final exponentiation uses
//     (p^4 - p^2 + 1) / r = ((z-1)^2 / 3) (z + p) (z^2 + p^2 - 1) + 1
// (checked numerically in tests/test_oracle_pins.py), i.e. the exact exponent, not its multiple by 3, so the
// value in GT is the same as a plain f^((p^12-1)/r).
//
// One pairing product per thread; an Fq12 is 144 registers' worth of limbs, so these functions work on
// references to thread-local values and are not inlined.  This is a latency-tolerant side path (two pairings
// per KZG opening), data-parallel over openings, not a throughput kernel like the point path.
#pragma once
#include "curve.cuh"

namespace ptau {

// A block-wide barrier at the top of every iteration of the long loops keeps the warps of a block (and so most of an SM)
// in the same part of the ~450 KB instruction stream: measured 1.15 -> 1.23 M openings/s at 128 threads per block
// (profiles/r02_kzg_ab2.log).  The kernels keep control flow uniform per block (a thread past the end repeats the last
// item).  -DPTAU_PAIR_NOSYNC removes the barriers.
#if defined(__CUDA_ARCH__) && !defined(PTAU_PAIR_NOSYNC)
#define PTAU_TOWER_SYNC() __syncthreads()
#else
#define PTAU_TOWER_SYNC() ((void)0)
#endif

struct Fq6 {
  Fq2 c0, c1, c2;
};
struct Fq12 {
  Fq6 c0, c1;
};

// Like the rest of csrc/, everything here also compiles for the host (tests/host_emul) so that the CPU-only
// suite runs the exact limb-level code against the oracle.
// (z-1)^2 / 3 = 0x396c8c005555e1568c00aaab0000aaab (126 bits)
#define PTAU_H1_INIT {0x0000aaabu, 0x8c00aaabu, 0x5555e156u, 0x396c8c00u}
#ifdef __CUDACC__
__constant__ uint32_t K_H1_D[4] = PTAU_H1_INIT;
#endif
static const uint32_t K_H1_H[4] = PTAU_H1_INIT;

PTAU_HD_NOINLINE Fq2 fq2_inv(const Fq2& a) {
  Fq n = fq_inv_fermat(fq_add(fq_sqr(a.c0), fq_sqr(a.c1)));
  Fq2 r;
  r.c0 = fq_mul(a.c0, n);
  r.c1 = fq_neg(fq_mul(a.c1, n));
  return r;
}
PTAU_HD Fq finv(const Fq& a) { return fq_inv_fermat(a); }
PTAU_HD Fq2 finv(const Fq2& a) { return fq2_inv(a); }

// a * (1 + u)
PTAU_HD Fq2 fq2_mul_xi(const Fq2& a) {
  Fq2 r;
  r.c0 = fq_sub(a.c0, a.c1);
  r.c1 = fq_add(a.c0, a.c1);
  return r;
}

// ---- Fq6 ----------------------------------------------------------------------------------------
PTAU_HD void fq6_add(Fq6& r, const Fq6& a, const Fq6& b) {
  r.c0 = fq2_add(a.c0, b.c0);
  r.c1 = fq2_add(a.c1, b.c1);
  r.c2 = fq2_add(a.c2, b.c2);
}
PTAU_HD void fq6_sub(Fq6& r, const Fq6& a, const Fq6& b) {
  r.c0 = fq2_sub(a.c0, b.c0);
  r.c1 = fq2_sub(a.c1, b.c1);
  r.c2 = fq2_sub(a.c2, b.c2);
}
PTAU_HD void fq6_neg(Fq6& r, const Fq6& a) {
  r.c0 = fq2_neg(a.c0);
  r.c1 = fq2_neg(a.c1);
  r.c2 = fq2_neg(a.c2);
}
// Karatsuba, 6 Fq2 multiplications; r may alias a or b
PTAU_HD_NOINLINE void fq6_mul(Fq6& r, const Fq6& a, const Fq6& b) {
  Fq2 v0 = fq2_mul(a.c0, b.c0), v1 = fq2_mul(a.c1, b.c1), v2 = fq2_mul(a.c2, b.c2);
  Fq2 t0 = fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c1, a.c2), fq2_add(b.c1, b.c2)), v1), v2);
  Fq2 t1 = fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c0, a.c1), fq2_add(b.c0, b.c1)), v0), v1);
  Fq2 t2 = fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c0, a.c2), fq2_add(b.c0, b.c2)), v0), v2);
  r.c0 = fq2_add(v0, fq2_mul_xi(t0));
  r.c1 = fq2_add(t1, fq2_mul_xi(v2));
  r.c2 = fq2_add(t2, v1);
}
// a * v
PTAU_HD void fq6_mul_v(Fq6& r, const Fq6& a) {
  Fq2 t = fq2_mul_xi(a.c2);
  r.c2 = a.c1;
  r.c1 = a.c0;
  r.c0 = t;
}
PTAU_HD_NOINLINE void fq6_inv(Fq6& r, const Fq6& a) {
  Fq2 t0 = fq2_sub(fq2_sqr(a.c0), fq2_mul_xi(fq2_mul(a.c1, a.c2)));
  Fq2 t1 = fq2_sub(fq2_mul_xi(fq2_sqr(a.c2)), fq2_mul(a.c0, a.c1));
  Fq2 t2 = fq2_sub(fq2_sqr(a.c1), fq2_mul(a.c0, a.c2));
  Fq2 d = fq2_add(fq2_mul(a.c0, t0), fq2_mul_xi(fq2_add(fq2_mul(a.c2, t1), fq2_mul(a.c1, t2))));
  Fq2 di = fq2_inv(d);
  r.c0 = fq2_mul(t0, di);
  r.c1 = fq2_mul(t1, di);
  r.c2 = fq2_mul(t2, di);
}
// Frobenius: conjugate the Fq2 coefficients, times xi^((p-1)/3), xi^(2(p-1)/3)
PTAU_HD_NOINLINE void fq6_frob(Fq6& r, const Fq6& a) {
  Fq2 g1, g2;
  g1.c0 = k_frob6_1_c0_mont();
  g1.c1 = k_frob6_1_c1_mont();
  g2.c0 = k_frob6_2_c0_mont();
  g2.c1 = k_frob6_2_c1_mont();
  r.c0 = fq2_conj(a.c0);
  r.c1 = fq2_mul(fq2_conj(a.c1), g1);
  r.c2 = fq2_mul(fq2_conj(a.c2), g2);
}

// ---- Fq12 ---------------------------------------------------------------------------------------
PTAU_HD void fq12_one(Fq12& r) {
  r.c0.c0 = fq2_one();
  r.c0.c1 = fq2_zero();
  r.c0.c2 = fq2_zero();
  r.c1.c0 = fq2_zero();
  r.c1.c1 = fq2_zero();
  r.c1.c2 = fq2_zero();
}
// r = a * b (3 Fq6 multiplications); r may alias a or b
PTAU_HD_NOINLINE void fq12_mul(Fq12& r, const Fq12& a, const Fq12& b) {
  Fq6 s, v0, v1;  // three temporaries: the sum of b's halves lives in v0 until the cross product is formed
  fq6_add(s, a.c0, a.c1);
  fq6_add(v0, b.c0, b.c1);
  fq6_mul(s, s, v0);
  fq6_mul(v0, a.c0, b.c0);
  fq6_mul(v1, a.c1, b.c1);
  fq6_sub(s, s, v0);
  fq6_sub(r.c1, s, v1);
  fq6_mul_v(v1, v1);
  fq6_add(r.c0, v0, v1);
}
// r = a^2 by the complex method (2 Fq6 multiplications instead of 3): with t = a0 a1,
// c0 = (a0 + a1)(a0 + v a1) - t - v t, c1 = 2t; r may alias a
PTAU_HD_NOINLINE void fq12_sqr(Fq12& r, const Fq12& a) {
  Fq6 t, s, u;
  fq6_mul(t, a.c0, a.c1);
  fq6_add(s, a.c0, a.c1);
  fq6_mul_v(u, a.c1);
  fq6_add(u, u, a.c0);
  fq6_mul(s, s, u);
  fq6_sub(s, s, t);
  fq6_mul_v(u, t);
  fq6_sub(r.c0, s, u);
  fq6_add(r.c1, t, t);
}
PTAU_HD void fq12_conj(Fq12& r, const Fq12& a) {
  r.c0 = a.c0;
  fq6_neg(r.c1, a.c1);
}
PTAU_HD_NOINLINE void fq12_inv(Fq12& r, const Fq12& a) {
  Fq6 t, u;
  fq6_mul(t, a.c0, a.c0);
  fq6_mul(u, a.c1, a.c1);
  fq6_mul_v(u, u);
  fq6_sub(t, t, u);
  fq6_inv(t, t);
  fq6_mul(u, a.c1, t);
  fq6_mul(r.c0, a.c0, t);
  fq6_neg(r.c1, u);
}
PTAU_HD_NOINLINE void fq12_frob(Fq12& r, const Fq12& a) {
  Fq2 g;
  g.c0 = k_frob12_c0_mont();
  g.c1 = k_frob12_c1_mont();
  fq6_frob(r.c0, a.c0);
  Fq6 t;
  fq6_frob(t, a.c1);
  r.c1.c0 = fq2_mul(t.c0, g);
  r.c1.c1 = fq2_mul(t.c1, g);
  r.c1.c2 = fq2_mul(t.c2, g);
}
PTAU_HD_NOINLINE bool fq12_is_one(const Fq12& a) {
  bool ok = fq2_eq(a.c0.c0, fq2_one());
  ok = ok && fq2_is_zero(a.c0.c1) && fq2_is_zero(a.c0.c2);
  ok = ok && fq2_is_zero(a.c1.c0) && fq2_is_zero(a.c1.c1) && fq2_is_zero(a.c1.c2);
  return ok;
}
// a^2 for a in the cyclotomic subgroup (Granger-Scott: three Fq4 squarings, 6 Fq2 multiplications instead of 18);
// r may alias a
PTAU_HD_NOINLINE void fq12_cyclotomic_sqr(Fq12& r, const Fq12& a) {
  Fq2 z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2;
  // (x + y s)^2 in Fq4 = Fq2[s]/(s^2 - xi): t_even = x^2 + xi y^2 = (x + y)(x + xi y) - xy - xi xy, t_odd = 2xy
  Fq2 tmp = fq2_mul(z0, z1);
  Fq2 t0 = fq2_sub(fq2_sub(fq2_mul(fq2_add(z0, z1), fq2_add(z0, fq2_mul_xi(z1))), tmp), fq2_mul_xi(tmp));
  Fq2 t1 = fq2_dbl(tmp);
  tmp = fq2_mul(z2, z3);
  Fq2 t2 = fq2_sub(fq2_sub(fq2_mul(fq2_add(z2, z3), fq2_add(z2, fq2_mul_xi(z3))), tmp), fq2_mul_xi(tmp));
  Fq2 t3 = fq2_dbl(tmp);
  tmp = fq2_mul(z4, z5);
  Fq2 t4 = fq2_sub(fq2_sub(fq2_mul(fq2_add(z4, z5), fq2_add(z4, fq2_mul_xi(z5))), tmp), fq2_mul_xi(tmp));
  Fq2 t5 = fq2_dbl(tmp);
  z0 = fq2_sub(t0, z0);  // 3 t0 - 2 z0
  z0 = fq2_add(fq2_dbl(z0), t0);
  z1 = fq2_add(t1, z1);  // 3 t1 + 2 z1
  z1 = fq2_add(fq2_dbl(z1), t1);
  tmp = fq2_mul_xi(t5);  // 3 xi t5 + 2 z2
  z2 = fq2_add(tmp, z2);
  z2 = fq2_add(fq2_dbl(z2), tmp);
  z3 = fq2_sub(t4, z3);  // 3 t4 - 2 z3
  z3 = fq2_add(fq2_dbl(z3), t4);
  z4 = fq2_sub(t2, z4);  // 3 t2 - 2 z4
  z4 = fq2_add(fq2_dbl(z4), t2);
  z5 = fq2_add(t3, z5);  // 3 t3 + 2 z5
  z5 = fq2_add(fq2_dbl(z5), t3);
  r.c0.c0 = z0;
  r.c0.c1 = z4;
  r.c0.c2 = z3;
  r.c1.c0 = z2;
  r.c1.c1 = z1;
  r.c1.c2 = z5;
}
// r = a^e for a in the cyclotomic subgroup; e = nbits-bit exponent in 32-bit words, top bit set.  r must not alias a
// (the accumulator is r itself: every Fq12 temporary is 576 bytes of local memory per thread, and the working set of
// a full wave is what decides whether it stays in L2)
PTAU_HD_NOINLINE void fq12_pow_cyclotomic(Fq12& r, const Fq12& a, const uint32_t* e, int nbits) {
  r = a;
#pragma unroll 1
  for (int i = nbits - 2; i >= 0; --i) {
    PTAU_TOWER_SYNC();
    fq12_cyclotomic_sqr(r, r);
    if ((e[i >> 5] >> (i & 31)) & 1u) fq12_mul(r, r, a);
  }
}
// a^z for a in the cyclotomic subgroup (z < 0: inverse = conjugate); r must not alias a
PTAU_HD_NOINLINE void fq12_exp_z(Fq12& r, const Fq12& a) {
  const uint32_t za[2] = {0x00010000u, 0xd2010000u};
  fq12_pow_cyclotomic(r, a, za, 64);
  fq12_conj(r, r);
}

// f^((p^6 - 1)(p^2 + 1)): the result is in the cyclotomic subgroup
PTAU_HD_NOINLINE void final_exp_easy(Fq12& m, const Fq12& f) {
  Fq12 t;
  fq12_conj(t, f);
  fq12_inv(m, f);
  fq12_mul(m, t, m);  // f^(p^6 - 1)
  fq12_frob(t, m);
  fq12_frob(t, t);
  fq12_mul(m, t, m);  // ^(p^2 + 1)
}
// r = a^((z + p)(z^2 + p^2 - 1)) * last for a in the cyclotomic subgroup; `a` is used as scratch (destroyed); r, a, last
// are three different objects
PTAU_HD_NOINLINE void final_exp_tail(Fq12& r, Fq12& a, const Fq12& last) {
  Fq12 b;
  fq12_exp_z(b, a);
  fq12_frob(r, a);
  fq12_mul(b, b, r);  // b = a^(z + p); a is free from here
  fq12_exp_z(a, b);
  fq12_exp_z(r, a);   // b^(z^2)
  fq12_frob(a, b);
  fq12_frob(a, a);
  fq12_mul(r, r, a);  // * b^(p^2)
  fq12_conj(a, b);
  fq12_mul(r, r, a);  // * b^-1
  fq12_mul(r, r, last);
}
// f^((p^12 - 1) / r)
PTAU_HD_NOINLINE void final_exponentiation(Fq12& r, const Fq12& f) {
  Fq12 m, a;
  final_exp_easy(m, f);
  {
    uint32_t h1[4];
#ifdef __CUDA_ARCH__
    const uint32_t* hc = K_H1_D;
#else
    const uint32_t* hc = K_H1_H;
#endif
#pragma unroll
    for (int i = 0; i < 4; i++) h1[i] = hc[i];
    fq12_pow_cyclotomic(a, m, h1, 126);  // ^((z-1)^2 / 3)
  }
  final_exp_tail(r, a, m);  // ^((z + p)(z^2 + p^2 - 1)), + 1 (r may be f itself: f is not read any more)
}
// f^((p^12 - 1) / r) == 1, decided on the cube: 3 (p^4 - p^2 + 1) / r = (z-1)^2 (z + p)(z^2 + p^2 - 1) + 3, and
// x -> x^3 is a bijection of the order-r group the value lies in (r is a prime != 3).  (z-1)^2 costs two
// exponentiations by the sparse z instead of one by the dense 126-bit (z-1)^2 / 3: 37 fewer Fq12 multiplications.
// Only the boolean is the same as final_exponentiation's; KZG10::check needs nothing else.  f is used as scratch.
PTAU_HD_NOINLINE bool final_exp_is_one(Fq12& f) {
  Fq12 m, a, t;
  final_exp_easy(m, f);
  fq12_exp_z(a, m);
  fq12_conj(t, m);
  fq12_mul(a, a, t);  // m^(z - 1)
  fq12_exp_z(t, a);
  fq12_conj(a, a);
  fq12_mul(a, t, a);  // m^((z - 1)^2)
  fq12_cyclotomic_sqr(t, m);
  fq12_mul(t, t, m);  // m^3
  final_exp_tail(f, a, t);
  return fq12_is_one(f);
}

// ---- Miller loop (ark-ec 0.2 bls12, TwistType::M) -------------------------------------------------
struct G2Hom {
  Fq2 x, y, z;
};
struct EllCoeff {
  Fq2 c0, c1, c2;
};

// fq_half: fqw.cuh
PTAU_HD Fq2 fq2_half(const Fq2& a) {
  Fq2 r;
  r.c0 = fq_half(a.c0);
  r.c1 = fq_half(a.c1);
  return r;
}
// ark-ec 0.2 doubling_step.  Two of its multiplications are by constants and are done without the multiplier, with the
// same field values: `* two_inv` is a halving, and `* b'` (the twist coefficient 4(1 + u) = 4 xi) is two doublings and
// a multiplication by xi.
PTAU_HD_NOINLINE void doubling_step(G2Hom& r, EllCoeff& co) {
  Fq2 a = fq2_half(fq2_mul(r.x, r.y));
  Fq2 b = fq2_sqr(r.y);
  Fq2 c = fq2_sqr(r.z);
  Fq2 e = fq2_mul_xi(fq2_dbl(fq2_dbl(fq2_add(fq2_dbl(c), c))));  // b' * 3c
  Fq2 f = fq2_add(fq2_dbl(e), e);
  Fq2 g = fq2_half(fq2_add(b, f));
  Fq2 h = fq2_sub(fq2_sqr(fq2_add(r.y, r.z)), fq2_add(b, c));
  Fq2 i = fq2_sub(e, b);
  Fq2 j = fq2_sqr(r.x);
  Fq2 e2 = fq2_sqr(e);
  r.x = fq2_mul(a, fq2_sub(b, f));
  r.y = fq2_sub(fq2_sqr(g), fq2_add(fq2_dbl(e2), e2));
  r.z = fq2_mul(b, h);
  co.c0 = i;
  co.c1 = fq2_add(fq2_dbl(j), j);
  co.c2 = fq2_neg(h);
}
PTAU_HD_NOINLINE void addition_step(G2Hom& r, const Fq2& qx, const Fq2& qy, EllCoeff& co) {
  Fq2 theta = fq2_sub(r.y, fq2_mul(qy, r.z));
  Fq2 lambda = fq2_sub(r.x, fq2_mul(qx, r.z));
  Fq2 c = fq2_sqr(theta);
  Fq2 d = fq2_sqr(lambda);
  Fq2 e = fq2_mul(lambda, d);
  Fq2 f = fq2_mul(r.z, c);
  Fq2 g = fq2_mul(r.x, d);
  Fq2 h = fq2_sub(fq2_add(e, f), fq2_dbl(g));
  r.x = fq2_mul(lambda, h);
  r.y = fq2_sub(fq2_mul(theta, fq2_sub(g, h)), fq2_mul(e, r.y));
  r.z = fq2_mul(r.z, e);
  co.c0 = fq2_sub(fq2_mul(theta, qx), fq2_mul(lambda, qy));
  co.c1 = fq2_neg(theta);
  co.c2 = lambda;
}
// G2Prepared of ark-ec 0.2 (`prepared_h`, `prepared_beta_h`: /root/reference/src/lib.rs:223-224,
// src/bin/preprocess-kgz.rs:177-184): the 68 line-coefficient triples of the loop above, as ark keeps them in memory
// (Montgomery limbs; triple t at out + 72 t words: c0 | c1 | c2, each c0-part then c1-part).  rec: one ARK_MONT_LIMBS G2
// record.  A point at infinity has no coefficients (ark: empty vector, infinity = true): the slot is zero-filled.
#define PTAU_G2PREP_COEFFS 68
PTAU_HD_NOINLINE bool g2_prepare_item(const uint32_t* rec, uint32_t* out) {
  G2Hom r;
  Fq2 qx, qy;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    qx.c0.l[i] = rec[i];
    qx.c1.l[i] = rec[12 + i];
    qy.c0.l[i] = rec[24 + i];
    qy.c1.l[i] = rec[36 + i];
  }
  if ((rec[48] & 0xffu) != 0) {
#pragma unroll 1
    for (int i = 0; i < PTAU_G2PREP_COEFFS * 72; i++) out[i] = 0;
    return true;
  }
  r.x = qx;
  r.y = qy;
  r.z = fq2_one();
  EllCoeff co;
  int t = 0;
#pragma unroll 1
  for (int i = 62; i >= 0; --i) {
#pragma unroll 1
    for (int step = 0; step < 2; step++) {
      if (step == 0)
        doubling_step(r, co);
      else if ((PTAU_Z_ABS >> i) & 1ull)
        addition_step(r, qx, qy, co);
      else
        break;
      const Fq* f[6] = {&co.c0.c0, &co.c0.c1, &co.c1.c0, &co.c1.c1, &co.c2.c0, &co.c2.c1};
#pragma unroll 1
      for (int k = 0; k < 6; k++)
#pragma unroll
        for (int j = 0; j < 12; j++) out[t * 72 + k * 12 + j] = f[k]->l[j];
      t++;
    }
  }
  return false;
}

// Fq6 times the sparse b0 + b1 v (5 Fq2 multiplications); r may alias a
PTAU_HD_NOINLINE void fq6_mul_by_01(Fq6& r, const Fq6& a, const Fq2& b0, const Fq2& b1) {
  Fq2 v0 = fq2_mul(a.c0, b0), v1 = fq2_mul(a.c1, b1);
  Fq2 t0 = fq2_add(v0, fq2_mul_xi(fq2_sub(fq2_mul(fq2_add(a.c1, a.c2), b1), v1)));
  Fq2 t1 = fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c0, a.c1), fq2_add(b0, b1)), v0), v1);
  Fq2 t2 = fq2_add(fq2_sub(fq2_mul(fq2_add(a.c0, a.c2), b0), v0), v1);
  r.c0 = t0;
  r.c1 = t1;
  r.c2 = t2;
}
// Fq6 times b1 v (3 Fq2 multiplications); r may alias a
PTAU_HD void fq6_mul_by_1(Fq6& r, const Fq6& a, const Fq2& b1) {
  Fq2 t0 = fq2_mul_xi(fq2_mul(a.c2, b1)), t1 = fq2_mul(a.c0, b1), t2 = fq2_mul(a.c1, b1);
  r.c0 = t0;
  r.c1 = t1;
  r.c2 = t2;
}
// f *= (c0 + c1 px v) + (c2 py v) w     (ark: mul_by_014; 13 Fq2 multiplications instead of 18)
// One Fq6 temporary: aa and bb are formed in place in f.c0 and f.c1 (every Fq6 a thread keeps is 288 bytes of local
// memory, and the working set of a full wave sits right at the size of L2).
PTAU_HD_NOINLINE void ell(Fq12& f, const EllCoeff& co, const Fq& px, const Fq& py) {
  const Fq2 c1 = fq2_mul_fq(co.c1, px), c4 = fq2_mul_fq(co.c2, py);
  Fq6 s;
  fq6_add(s, f.c1, f.c0);
  fq6_mul_by_01(f.c0, f.c0, co.c0, c1);  // aa
  fq6_mul_by_1(f.c1, f.c1, c4);          // bb
  fq6_mul_by_01(s, s, co.c0, fq2_add(c1, c4));
  fq6_sub(s, s, f.c0);
  fq6_sub(s, s, f.c1);   // new c1 = (f0 + f1)(c0 + (c1 + c4) v) - aa - bb
  fq6_mul_v(f.c1, f.c1);
  fq6_add(f.c0, f.c1, f.c0);  // new c0 = v bb + aa
  f.c1 = s;
}

// Miller value of up to two pairs (P_k affine in G1, Q_k affine on the twist); a pair with use[k] == false
// (P or Q at infinity) contributes 1, like ark's filter in miller_loop.
PTAU_HD_NOINLINE void miller_loop2(Fq12& f, const Fq* px, const Fq* py, const Fq2* qx, const Fq2* qy,
                                                 const bool* use) {
  fq12_one(f);
  G2Hom r[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    r[k].x = qx[k];
    r[k].y = qy[k];
    r[k].z = fq2_one();
  }
  EllCoeff co;
#pragma unroll 1
  for (int i = 62; i >= 0; --i) {
    PTAU_TOWER_SYNC();
    fq12_sqr(f, f);
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
      if (!use[k]) continue;
      doubling_step(r[k], co);
      ell(f, co, px[k], py[k]);
    }
    if ((PTAU_Z_ABS >> i) & 1ull) {
#pragma unroll 1
      for (int k = 0; k < 2; k++) {
        if (!use[k]) continue;
        addition_step(r[k], qx[k], qy[k], co);
        ell(f, co, px[k], py[k]);
      }
    }
  }
  fq12_conj(f, f);  // z < 0
}

// The same value when the line coefficients of Q_0 are already known (g2_prepare_item's layout, 68 triples of 72 words):
// KZG10::check pairs every opening with the verifier key's h, so its doubling / addition steps are done once per key.
PTAU_HD void load_ell_coeff(EllCoeff& co, const uint32_t* p) {
#pragma unroll
  for (int j = 0; j < 12; j++) {
    co.c0.c0.l[j] = p[j];
    co.c0.c1.l[j] = p[12 + j];
    co.c1.c0.l[j] = p[24 + j];
    co.c1.c1.l[j] = p[36 + j];
    co.c2.c0.l[j] = p[48 + j];
    co.c2.c1.l[j] = p[60 + j];
  }
}
PTAU_HD_NOINLINE void miller_loop2_prep0(Fq12& f, const Fq* px, const Fq* py, const uint32_t* prep0, const Fq2& qx1,
                                         const Fq2& qy1, const bool* use) {
  fq12_one(f);
  G2Hom r;
  r.x = qx1;
  r.y = qy1;
  r.z = fq2_one();
  EllCoeff co;
  int t = 0;
#pragma unroll 1
  for (int i = 62; i >= 0; --i) {
    PTAU_TOWER_SYNC();
    if (i != 62) fq12_sqr(f, f);  // 1^2
#pragma unroll 1
    for (int step = 0; step < 2; step++) {
      if (step == 1 && !((PTAU_Z_ABS >> i) & 1ull)) break;
      if (use[0]) {
        load_ell_coeff(co, prep0 + t * 72);
        ell(f, co, px[0], py[0]);
      }
      t++;
      if (use[1]) {
        if (step == 0)
          doubling_step(r, co);
        else
          addition_step(r, qx1, qy1, co);
        ell(f, co, px[1], py[1]);
      }
    }
  }
  fq12_conj(f, f);  // z < 0
}

// ---- group operations with every special case, generic over the field -----------------------------------
template <class F>
PTAU_HD_NOINLINE void jac_madd_complete_t(Jac<F>& acc, const F& x, const F& y, const F& one) {
  if (fis_zero(acc.Z)) {
    acc.X = x;
    acc.Y = y;
    acc.Z = one;
    return;
  }
  F zz = fsqr(acc.Z);
  F u2 = fmul(x, zz);
  F s2 = fmul(fmul(y, acc.Z), zz);
  if (feq(u2, acc.X)) {
    if (feq(s2, acc.Y)) {
      jac_dbl(acc);
    } else {
      acc.Z = fsub(acc.Z, acc.Z);  // P + (-P)
    }
    return;
  }
  F H = fsub(u2, acc.X);
  F I = fsqr(fdbl(H));
  F J = fmul(H, I);
  F rr = fdbl(fsub(s2, acc.Y));
  F V = fmul(acc.X, I);
  F X3 = fsub(fsub(fsqr(rr), J), fdbl(V));
  acc.Y = fsub(fmul(rr, fsub(V, X3)), fdbl(fmul(acc.Y, J)));
  acc.Z = fdbl(fmul(acc.Z, H));
  acc.X = X3;
}
// acc = [k] (x, y), k = 8 little-endian words (< 2^255); double-and-add
template <class F>
PTAU_HD_NOINLINE void jac_scalar_mul_t(Jac<F>& acc, const F& x, const F& y, const uint32_t* k, const F& one) {
  acc.X = fsub(one, one);
  acc.Y = one;
  acc.Z = acc.X;
#pragma unroll 1
  for (int i = 254; i >= 0; --i) {
    if (!fis_zero(acc.Z)) jac_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1u) jac_madd_complete_t(acc, x, y, one);
  }
}
// Jacobian -> affine; returns false for the point at infinity
template <class F>
PTAU_HD_NOINLINE bool jac_to_affine_t(const Jac<F>& p, F& x, F& y) {
  if (fis_zero(p.Z)) return false;
  F zi = finv(p.Z);
  F zi2 = fsqr(zi);
  x = fmul(p.X, zi2);
  y = fmul(p.Y, fmul(zi2, zi));
  return true;
}
// Both conversions of one KZG10 opening with ONE field inversion (Montgomery's trick across Fq and Fq2:
// 1 / Z2 = conj(Z2) / N(Z2) with the norm N in Fq, and 1/Z1, 1/N come from 1 / (Z1 N)).  Same affine values as two
// jac_to_affine_t calls; returns through pok / qok whether each point is finite.
PTAU_HD_NOINLINE void jac_to_affine_pair(const Jac<Fq>& p, const Jac<Fq2>& q, Fq& px, Fq& py, Fq2& qx, Fq2& qy, bool& pok,
                                         bool& qok) {
  pok = !fq_is_zero(p.Z);
  qok = !fq2_is_zero(q.Z);
  const Fq z1 = pok ? p.Z : fq_one();
  // N = c0^2 + c1^2 is zero only for Z2 = 0 (-1 is not a square mod p)
  const Fq n = qok ? fq_add(fq_sqr(q.Z.c0), fq_sqr(q.Z.c1)) : fq_one();
  const Fq ti = fq_inv_fermat(fq_mul(z1, n));
  if (pok) {
    Fq zi = fq_mul(ti, n);
    Fq zi2 = fq_sqr(zi);
    px = fq_mul(p.X, zi2);
    py = fq_mul(p.Y, fq_mul(zi2, zi));
  }
  if (qok) {
    Fq ni = fq_mul(ti, z1);
    Fq2 zi;
    zi.c0 = fq_mul(q.Z.c0, ni);
    zi.c1 = fq_neg(fq_mul(q.Z.c1, ni));
    Fq2 zi2 = fq2_sqr(zi);
    qx = fq2_mul(q.X, zi2);
    qy = fq2_mul(q.Y, fq2_mul(zi2, zi));
  }
}
// p += q, both Jacobian, every special case
template <class F>
PTAU_HD_NOINLINE void jac_add_complete_t(Jac<F>& p, const Jac<F>& q) {
  if (fis_zero(q.Z)) return;
  if (fis_zero(p.Z)) {
    p = q;
    return;
  }
  F z1z1 = fsqr(p.Z), z2z2 = fsqr(q.Z);
  F u1 = fmul(p.X, z2z2), u2 = fmul(q.X, z1z1);
  F s1 = fmul(fmul(p.Y, q.Z), z2z2), s2 = fmul(fmul(q.Y, p.Z), z1z1);
  if (feq(u1, u2)) {
    if (feq(s1, s2)) {
      jac_dbl(p);
    } else {
      p.Z = fsub(p.Z, p.Z);
    }
    return;
  }
  F H = fsub(u2, u1);
  F I = fsqr(fdbl(H));
  F J = fmul(H, I);
  F rr = fdbl(fsub(s2, s1));
  F V = fmul(u1, I);
  F X3 = fsub(fsub(fsqr(rr), J), fdbl(V));
  p.Y = fsub(fmul(rr, fsub(V, X3)), fdbl(fmul(s1, J)));
  p.Z = fmul(fdbl(fmul(p.Z, q.Z)), H);
  p.X = X3;
}

// ---- ARK_MONT_LIMBS records ----------------------------------------------------------------------
PTAU_HD void load_g1_rec(const uint32_t* rec, Fq& x, Fq& y, bool& inf) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x.l[i] = rec[i];
    y.l[i] = rec[12 + i];
  }
  inf = (rec[24] & 0xffu) != 0;
}
PTAU_HD void load_g2_rec(const uint32_t* rec, Fq2& x, Fq2& y, bool& inf) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x.c0.l[i] = rec[i];
    x.c1.l[i] = rec[12 + i];
    y.c0.l[i] = rec[24 + i];
    y.c1.l[i] = rec[36 + i];
  }
  inf = (rec[48] & 0xffu) != 0;
}
// GT element -> 12 x 48 canonical little-endian bytes in arkworks' Fq12 order (c0.c0.c0, c0.c0.c1, c0.c1.c0, ...)
PTAU_HD_NOINLINE void store_gt(uint32_t* out, const Fq12& f) {
  const Fq2* c[6] = {&f.c0.c0, &f.c0.c1, &f.c0.c2, &f.c1.c0, &f.c1.c1, &f.c1.c2};
#pragma unroll 1
  for (int k = 0; k < 6; k++) {
    Fq a = fq_from_mont(c[k]->c0), b = fq_from_mont(c[k]->c1);
#pragma unroll
    for (int i = 0; i < 12; i++) {
      out[k * 24 + i] = a.l[i];
      out[k * 24 + 12 + i] = b.l[i];
    }
  }
}

// ---- fixed-base windowed multiplication -----------------------------------------------------------------
// g, gamma_g and h are the same for every opening of a call (and, in practice, of a context), so [v]g, [rv]gamma_g
// and [z]h use a table T[w][j-1] = [j 16^w] B, w < 64, j = 1..15 (ARK_MONT_LIMBS records): 64 mixed additions instead
// of 255 doublings + ~127 additions.  One window per thread builds it: 4w doublings, 14 additions, and one inversion
// shared by the 15 points (Montgomery's trick).
#define PTAU_FB_WINDOWS 64
#define PTAU_FB_ENTRIES (PTAU_FB_WINDOWS * 15)
PTAU_HD void store_rec(uint32_t* rec, const Fq& x, const Fq& y, bool inf) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    rec[i] = inf ? 0u : x.l[i];
    rec[12 + i] = inf ? 0u : y.l[i];
  }
  rec[24] = inf ? 1u : 0u;
  rec[25] = 0;
}
PTAU_HD void store_rec(uint32_t* rec, const Fq2& x, const Fq2& y, bool inf) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    rec[i] = inf ? 0u : x.c0.l[i];
    rec[12 + i] = inf ? 0u : x.c1.l[i];
    rec[24 + i] = inf ? 0u : y.c0.l[i];
    rec[36 + i] = inf ? 0u : y.c1.l[i];
  }
  rec[48] = inf ? 1u : 0u;
  rec[49] = 0;
}
PTAU_HD void load_rec(const uint32_t* rec, Fq& x, Fq& y, bool& inf) { load_g1_rec(rec, x, y, inf); }
PTAU_HD void load_rec(const uint32_t* rec, Fq2& x, Fq2& y, bool& inf) { load_g2_rec(rec, x, y, inf); }
template <class F>
struct RecWords;
template <>
struct RecWords<Fq> {
  static constexpr int value = 26;
};
template <>
struct RecWords<Fq2> {
  static constexpr int value = 50;
};

// window w of the table of the base given as a record
template <class F>
PTAU_HD_NOINLINE void fixed_base_window(uint32_t* tbl, const uint32_t* base_rec, int w, const F& one) {
  constexpr int REC = RecWords<F>::value;
  uint32_t* out = tbl + (size_t)w * 15 * REC;
  F bx, by;
  bool binf;
  load_rec(base_rec, bx, by, binf);
  Jac<F> m[15];
  m[0].X = bx;
  m[0].Y = by;
  m[0].Z = binf ? fsub(one, one) : one;
#pragma unroll 1
  for (int i = 0; i < 4 * w; i++)
    if (!fis_zero(m[0].Z)) jac_dbl(m[0]);
#pragma unroll 1
  for (int j = 1; j < 15; j++) {
    m[j] = m[j - 1];
    jac_add_complete_t(m[j], m[0]);
  }
  // one inversion for the 15 Z coordinates
  F pre[15];
  F acc = one;
#pragma unroll 1
  for (int j = 0; j < 15; j++) {
    pre[j] = acc;
    if (!fis_zero(m[j].Z)) acc = fmul(acc, m[j].Z);
  }
  F inv = finv(acc);
#pragma unroll 1
  for (int j = 14; j >= 0; --j) {
    if (fis_zero(m[j].Z)) {
      store_rec(out + j * REC, bx, by, true);
      continue;
    }
    F zi = fmul(inv, pre[j]);
    inv = fmul(inv, m[j].Z);
    F zi2 = fsqr(zi);
    store_rec(out + j * REC, fmul(m[j].X, zi2), fmul(m[j].Y, fmul(zi2, zi)), false);
  }
}
// Second level: T8[w][e-1] = [e 256^w] B, w < 32, e = 1..255, from the 4-bit table: entry e = 16 hi + lo is
// T[2w+1][hi] + T[2w][lo].  One thread per (w, hi) builds its 16 entries with one inversion.  32 mixed additions per
// scalar instead of 64 (KZG10::check does three such multiplications per opening; the tables are built once per key).
#define PTAU_FB8_WINDOWS 32
#define PTAU_FB8_ENTRIES (PTAU_FB8_WINDOWS * 255)
template <class F>
PTAU_HD_NOINLINE void fixed_base_window8(uint32_t* tbl8, const uint32_t* tbl4, int w, int hi, const F& one) {
  constexpr int REC = RecWords<F>::value;
  uint32_t* out = tbl8 + ((size_t)w * 255 + hi * 16) * REC;  // entry e = 16 hi + lo lives at out + (lo - 1) REC
  const F zero = fsub(one, one);
  F bx = zero, by = one;
  bool binf = true;
  if (hi) load_rec(tbl4 + ((size_t)(2 * w + 1) * 15 + hi - 1) * REC, bx, by, binf);
  Jac<F> m[16];
#pragma unroll 1
  for (int lo = 0; lo < 16; lo++) {
    m[lo].X = bx;
    m[lo].Y = by;
    m[lo].Z = binf ? zero : one;
    if (lo) {
      F x, y;
      bool inf;
      load_rec(tbl4 + ((size_t)(2 * w) * 15 + lo - 1) * REC, x, y, inf);
      if (!inf) jac_madd_complete_t(m[lo], x, y, one);
    }
  }
  F pre[16];
  F acc = one;
#pragma unroll 1
  for (int j = 0; j < 16; j++) {
    pre[j] = acc;
    if (!fis_zero(m[j].Z)) acc = fmul(acc, m[j].Z);
  }
  F inv = finv(acc);
#pragma unroll 1
  for (int j = 15; j >= 0; --j) {
    if (hi == 0 && j == 0) continue;  // e = 0 has no entry
    if (fis_zero(m[j].Z)) {
      store_rec(out + (j - 1) * REC, bx, by, true);
      continue;
    }
    F zi = fmul(inv, pre[j]);
    inv = fmul(inv, m[j].Z);
    F zi2 = fsqr(zi);
    store_rec(out + (j - 1) * REC, fmul(m[j].X, zi2), fmul(m[j].Y, fmul(zi2, zi)), false);
  }
}
// acc = [k] B from B's table (w8: the 8-bit second-level table, else the 4-bit one); k = 8 little-endian words
template <class F>
PTAU_HD_NOINLINE void fixed_base_mul(Jac<F>& acc, const uint32_t* tbl, const uint32_t* k, const F& one, bool w8 = false) {
  constexpr int REC = RecWords<F>::value;
  acc.X = fsub(one, one);
  acc.Y = one;
  acc.Z = acc.X;
  const int bits = w8 ? 8 : 4;
  const uint32_t per = w8 ? 255u : 15u;
#pragma unroll 1
  for (int w = 0; w < 256 / bits; w++) {
    const uint32_t d = (k[(w * bits) >> 5] >> ((w * bits) & 31)) & per;
    if (!d) continue;
    F x, y;
    bool inf;
    load_rec(tbl + ((size_t)w * per + d - 1) * REC, x, y, inf);
    if (!inf) jac_madd_complete_t(acc, x, y, one);
  }
}

// ---- one item of the two public operations --------------------------------------------------------
// prod_{k<2} e(P_k, Q_k): g1 = 2 G1 records (26 words each), g2 = 2 G2 records (50 words each);
// gt_out (144 words) may be null.  Returns whether the product is 1.
PTAU_HD_NOINLINE bool pairing_product2_item(const uint32_t* g1, const uint32_t* g2, uint32_t* gt_out) {
  Fq px[2], py[2];
  Fq2 qx[2], qy[2];
  bool use[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    bool pi, qi;
    load_g1_rec(g1 + k * 26, px[k], py[k], pi);
    load_g2_rec(g2 + k * 50, qx[k], qy[k], qi);
    use[k] = !pi && !qi;
  }
  Fq12 f;
  miller_loop2(f, px, py, qx, qy, use);
  final_exponentiation(f, f);
  if (gt_out) store_gt(gt_out, f);
  return fq12_is_one(f);
}

// KZG10::check of one opening: e(C - [v]g - [rv]gamma_g, h) == e(w, beta_h - [z]h), evaluated as
// e(inner, h) * e(-w, beta_h - [z]h) == 1.  Scalars: 8 little-endian words, < r; random_v may be null.
// tbl_g / tbl_gg / tbl_h: fixed-base tables of g, gamma_g, h (fixed_base_window, or with w8 the second-level tables of
// fixed_base_window8), or null for plain double-and-add;
// prep_h: h's line coefficients (g2_prepare_item), or null to run its doubling / addition steps per opening.
PTAU_HD_NOINLINE bool kzg_check_item(const uint32_t* vk_g1, const uint32_t* vk_g2, const uint32_t* comm, const uint32_t* point,
                                     const uint32_t* value, const uint32_t* proof_w, const uint32_t* random_v,
                                     const uint32_t* tbl_g = nullptr, const uint32_t* tbl_gg = nullptr,
                                     const uint32_t* tbl_h = nullptr, const uint32_t* prep_h = nullptr, bool w8 = false) {
  Fq px[2], py[2];
  Fq2 qx[2], qy[2];
  bool use[2];
  uint32_t k[8];
  bool hinf;
  load_g2_rec(vk_g2, qx[0], qy[0], hinf);
  Jac<Fq> inner;
  {  // inner = C - [v] g - [rv] gamma_g
    Fq gx, gy;
    bool ginf;
    load_g1_rec(vk_g1, gx, gy, ginf);
    Jac<Fq> acc;
#pragma unroll
    for (int w = 0; w < 8; w++) k[w] = ginf ? 0u : value[w];
    if (tbl_g)
      fixed_base_mul(acc, tbl_g, k, fq_one(), w8);
    else
      jac_scalar_mul_t(acc, gx, gy, k, fq_one());
    if (random_v) {
      load_g1_rec(vk_g1 + 26, gx, gy, ginf);
#pragma unroll
      for (int w = 0; w < 8; w++) k[w] = ginf ? 0u : random_v[w];
      Jac<Fq> t;
      if (tbl_gg)
        fixed_base_mul(t, tbl_gg, k, fq_one(), w8);
      else
        jac_scalar_mul_t(t, gx, gy, k, fq_one());
      jac_add_complete_t(acc, t);
    }
    acc.Y = fq_neg(acc.Y);
    Fq cx, cy;
    bool cinf;
    load_g1_rec(comm, cx, cy, cinf);
    if (!cinf) jac_madd_complete_t(acc, cx, cy, fq_one());
    inner = acc;
  }
  {  // Q = beta_h - [z] h
    Jac<Fq2> acc;
#pragma unroll
    for (int w = 0; w < 8; w++) k[w] = hinf ? 0u : point[w];
    if (tbl_h)
      fixed_base_mul(acc, tbl_h, k, fq2_one(), w8);
    else
      jac_scalar_mul_t(acc, qx[0], qy[0], k, fq2_one());
    acc.Y = fq2_neg(acc.Y);
    Fq2 bx, by;
    bool binf;
    load_g2_rec(vk_g2 + 50, bx, by, binf);
    if (!binf) jac_madd_complete_t(acc, bx, by, fq2_one());
    bool winf;
    load_g1_rec(proof_w, px[1], py[1], winf);
    py[1] = fq_neg(py[1]);
    bool pok, qok;
    jac_to_affine_pair(inner, acc, px[0], py[0], qx[1], qy[1], pok, qok);
    use[0] = pok && !hinf;
    use[1] = qok && !winf;
  }
  Fq12 f;
  if (prep_h)
    miller_loop2_prep0(f, px, py, prep_h, qx[1], qy[1], use);
  else
    miller_loop2(f, px, py, qx, qy, use);
  return final_exp_is_one(f);
}

}  // namespace ptau
