// BLS12-381 base field Fq, 12 x u32 limbs, Montgomery form (R = 2^384).
//
// Restates what pairing 0.14.2 `Fq` / ark-ff 0.2.0 `Fp384` compute (6 x u64
// Montgomery; call sites /root/reference/src/lib.rs:52,78,180) as 32-bit
// carry-chained multiply-adds so that ptxas emits IMAD.WIDE.U32(.X) chains on
// sm_100a.  The multiplier keeps two accumulators, one holding the products of
// the even limbs of `a` and one the odd limbs, so that every 32x32->64 product
// lands on a 64-bit aligned register pair of its accumulator and a whole row is
// one carry chain of 6 wide MADs.
//
// fq_sqr_inl is the dedicated squaring (78 product MADs instead of 144).  Measured on B200:
// IMAD 18.4 T/s, IMAD.WIDE.X carry chains 8.9 T/s (a wide carry MAD = two IMAD issue
// slots), one multiplication = 581 slots, one squaring = 446.
//
// Every function is also compiled for the host (plain C emulation of the same
// instruction sequences, including the carry flag) so that tests/host_emul can
// check the exact limb algorithm against the oracle without a GPU.  The host
// build is a test vehicle only; the library never calls it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PTAU_HD __host__ __device__ __forceinline__
#define PTAU_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define PTAU_HD inline
#define PTAU_HD_NOINLINE static inline
#endif

namespace ptau {

struct Fq {
  uint32_t l[12];
};

// p, little-endian 32-bit limbs (also spelled as immediates in the asm below).
#define PTAU_P_LIMBS                                                                       \
  {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,           \
   0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau}
#define PTAU_M0 0xfffcfffdu  // -p^-1 mod 2^32

#define P0S "0xffffaaab"
#define P1S "0xb9feffff"
#define P2S "0xb153ffff"
#define P3S "0x1eabfffe"
#define P4S "0xf6b0f624"
#define P5S "0x6730d2a0"
#define P6S "0xf38512bf"
#define P7S "0x64774b84"
#define P8S "0x434bacd7"
#define P9S "0x4b1ba7b6"
#define P10S "0x397fe69a"
#define P11S "0x1a0111ea"

PTAU_HD Fq fq_zero() {
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = 0;
  return r;
}

// R mod p  (Montgomery one)
PTAU_HD Fq fq_one() {
  const uint32_t c[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                          0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = c[i];
  return r;
}

// R^2 mod p
PTAU_HD Fq fq_r2() {
  const uint32_t c[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                          0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = c[i];
  return r;
}

PTAU_HD bool fq_is_zero(const Fq& a) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) t |= a.l[i];
  return t == 0;
}

PTAU_HD bool fq_eq(const Fq& a, const Fq& b) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) t |= a.l[i] ^ b.l[i];
  return t == 0;
}

// ---------------------------------------------------------------------------
// host emulation helpers (carry flag made explicit)
// ---------------------------------------------------------------------------
#ifndef __CUDA_ARCH__
namespace emu {
static inline uint32_t addc(uint32_t a, uint32_t b, uint32_t& cf) {
  uint64_t t = (uint64_t)a + b + cf;
  cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
static inline uint32_t subc(uint32_t a, uint32_t b, uint32_t& bf) {
  uint64_t t = (uint64_t)a - b - bf;
  bf = (uint32_t)(t >> 63);
  return (uint32_t)t;
}
static inline uint32_t madlo(uint32_t a, uint32_t b, uint32_t c, uint32_t& cf) {
  uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + cf;
  cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
static inline uint32_t madhi(uint32_t a, uint32_t b, uint32_t c, uint32_t& cf) {
  uint64_t t = (((uint64_t)a * b) >> 32) + c + cf;
  cf = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
static const uint32_t PL[12] = PTAU_P_LIMBS;
}  // namespace emu
#endif

// ---------------------------------------------------------------------------
// carry-chain rows of the multiplier.  Absolute limb positions: accumulator E
// is "even aligned" (E[k] at position k); accumulator X is "odd aligned"
// (X[k] at position k+1).
// ---------------------------------------------------------------------------

// E += sum_{j even} a_j * bi * W^j ; carry out of position 12 goes to top (=X[11]).
PTAU_HD void row_mac_even(uint32_t* E, uint32_t& top, const uint32_t* a, uint32_t bi) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %13, %19, %0;\n\t"
      "madc.hi.cc.u32 %1, %13, %19, %1;\n\t"
      "madc.lo.cc.u32 %2, %14, %19, %2;\n\t"
      "madc.hi.cc.u32 %3, %14, %19, %3;\n\t"
      "madc.lo.cc.u32 %4, %15, %19, %4;\n\t"
      "madc.hi.cc.u32 %5, %15, %19, %5;\n\t"
      "madc.lo.cc.u32 %6, %16, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %16, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %17, %19, %8;\n\t"
      "madc.hi.cc.u32 %9, %17, %19, %9;\n\t"
      "madc.lo.cc.u32 %10, %18, %19, %10;\n\t"
      "madc.hi.cc.u32 %11, %18, %19, %11;\n\t"
      "addc.u32 %12, %12, 0;"
      : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
        "+r"(E[8]), "+r"(E[9]), "+r"(E[10]), "+r"(E[11]), "+r"(top)
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(bi));
#else
  uint32_t cf = 0;
  for (int j = 0; j < 12; j += 2) {
    E[j] = emu::madlo(a[j], bi, E[j], cf);
    E[j + 1] = emu::madhi(a[j], bi, E[j + 1], cf);
  }
  top += cf;
#endif
}

// Shift-and-accumulate: on entry X holds the previous even accumulator (its
// limb 0 is zero, limb 1 is the left-over at new position 0, limbs 2..11 move to
// odd-aligned indices 0..9).  E0 (new E[0]) absorbs the left-over and its carry
// enters the X chain at position 1.  X = shifted X + sum_{j odd} a_j * bi.
PTAU_HD void row_mac_odd_shift(uint32_t* X, uint32_t& E0, const uint32_t* a, uint32_t bi) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %12, %12, %1;\n\t"
      "madc.lo.cc.u32 %0, %13, %19, %2;\n\t"
      "madc.hi.cc.u32 %1, %13, %19, %3;\n\t"
      "madc.lo.cc.u32 %2, %14, %19, %4;\n\t"
      "madc.hi.cc.u32 %3, %14, %19, %5;\n\t"
      "madc.lo.cc.u32 %4, %15, %19, %6;\n\t"
      "madc.hi.cc.u32 %5, %15, %19, %7;\n\t"
      "madc.lo.cc.u32 %6, %16, %19, %8;\n\t"
      "madc.hi.cc.u32 %7, %16, %19, %9;\n\t"
      "madc.lo.cc.u32 %8, %17, %19, %10;\n\t"
      "madc.hi.cc.u32 %9, %17, %19, %11;\n\t"
      "madc.lo.cc.u32 %10, %18, %19, 0;\n\t"
      "madc.hi.cc.u32 %11, %18, %19, 0;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]),
        "+r"(X[8]), "+r"(X[9]), "+r"(X[10]), "+r"(X[11]), "+r"(E0)
      : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(a[9]), "r"(a[11]), "r"(bi));
#else
  uint32_t cf = 0;
  E0 = emu::addc(E0, X[1], cf);
  for (int j = 0; j < 10; j += 2) {
    X[j] = emu::madlo(a[j + 1], bi, X[j + 2], cf);
    X[j + 1] = emu::madhi(a[j + 1], bi, X[j + 3], cf);
  }
  X[10] = emu::madlo(a[11], bi, 0, cf);
  X[11] = emu::madhi(a[11], bi, 0, cf);
#endif
}

// X += sum_{j odd} p_j * m   (no carry out of the chain: X < W^12 by the bound)
PTAU_HD void row_red_odd(uint32_t* X, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %12, " P1S ", %0;\n\t"
      "madc.hi.cc.u32 %1, %12, " P1S ", %1;\n\t"
      "madc.lo.cc.u32 %2, %12, " P3S ", %2;\n\t"
      "madc.hi.cc.u32 %3, %12, " P3S ", %3;\n\t"
      "madc.lo.cc.u32 %4, %12, " P5S ", %4;\n\t"
      "madc.hi.cc.u32 %5, %12, " P5S ", %5;\n\t"
      "madc.lo.cc.u32 %6, %12, " P7S ", %6;\n\t"
      "madc.hi.cc.u32 %7, %12, " P7S ", %7;\n\t"
      "madc.lo.cc.u32 %8, %12, " P9S ", %8;\n\t"
      "madc.hi.cc.u32 %9, %12, " P9S ", %9;\n\t"
      "madc.lo.cc.u32 %10, %12, " P11S ", %10;\n\t"
      "madc.hi.cc.u32 %11, %12, " P11S ", %11;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]),
        "+r"(X[8]), "+r"(X[9]), "+r"(X[10]), "+r"(X[11])
      : "r"(m));
#else
  uint32_t cf = 0;
  for (int j = 0; j < 12; j += 2) {
    X[j] = emu::madlo(m, emu::PL[j + 1], X[j], cf);
    X[j + 1] = emu::madhi(m, emu::PL[j + 1], X[j + 1], cf);
  }
#endif
}

// E += sum_{j even} p_j * m ; carry out goes to top (=X[11]).
PTAU_HD void row_red_even(uint32_t* E, uint32_t& top, uint32_t m) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %13, " P0S ", %0;\n\t"
      "madc.hi.cc.u32 %1, %13, " P0S ", %1;\n\t"
      "madc.lo.cc.u32 %2, %13, " P2S ", %2;\n\t"
      "madc.hi.cc.u32 %3, %13, " P2S ", %3;\n\t"
      "madc.lo.cc.u32 %4, %13, " P4S ", %4;\n\t"
      "madc.hi.cc.u32 %5, %13, " P4S ", %5;\n\t"
      "madc.lo.cc.u32 %6, %13, " P6S ", %6;\n\t"
      "madc.hi.cc.u32 %7, %13, " P6S ", %7;\n\t"
      "madc.lo.cc.u32 %8, %13, " P8S ", %8;\n\t"
      "madc.hi.cc.u32 %9, %13, " P8S ", %9;\n\t"
      "madc.lo.cc.u32 %10, %13, " P10S ", %10;\n\t"
      "madc.hi.cc.u32 %11, %13, " P10S ", %11;\n\t"
      "addc.u32 %12, %12, 0;"
      : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
        "+r"(E[8]), "+r"(E[9]), "+r"(E[10]), "+r"(E[11]), "+r"(top)
      : "r"(m));
#else
  uint32_t cf = 0;
  for (int j = 0; j < 12; j += 2) {
    E[j] = emu::madlo(m, emu::PL[j], E[j], cf);
    E[j + 1] = emu::madhi(m, emu::PL[j], E[j + 1], cf);
  }
  top += cf;
#endif
}

// r = (a >= p) ? a - p : a     (a < 2p)
PTAU_HD void fq_cond_sub_p(uint32_t* a) {
  uint32_t t[12];
  uint32_t borrow;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %13, " P0S ";\n\t"
      "subc.cc.u32 %1, %14, " P1S ";\n\t"
      "subc.cc.u32 %2, %15, " P2S ";\n\t"
      "subc.cc.u32 %3, %16, " P3S ";\n\t"
      "subc.cc.u32 %4, %17, " P4S ";\n\t"
      "subc.cc.u32 %5, %18, " P5S ";\n\t"
      "subc.cc.u32 %6, %19, " P6S ";\n\t"
      "subc.cc.u32 %7, %20, " P7S ";\n\t"
      "subc.cc.u32 %8, %21, " P8S ";\n\t"
      "subc.cc.u32 %9, %22, " P9S ";\n\t"
      "subc.cc.u32 %10, %23, " P10S ";\n\t"
      "subc.cc.u32 %11, %24, " P11S ";\n\t"
      "subc.u32 %12, 0, 0;"
      : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]),
        "=&r"(t[8]), "=&r"(t[9]), "=&r"(t[10]), "=&r"(t[11]), "=&r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
        "r"(a[9]), "r"(a[10]), "r"(a[11]));
#else
  uint32_t bf = 0;
  for (int i = 0; i < 12; i++) t[i] = emu::subc(a[i], emu::PL[i], bf);
  borrow = bf ? 0xffffffffu : 0u;
#endif
#pragma unroll
  for (int i = 0; i < 12; i++) a[i] = borrow ? a[i] : t[i];
}

// Montgomery product a*b/R mod p, inputs < p, output < p.
PTAU_HD Fq fq_mul_inl(const Fq& a, const Fq& b) {
  uint32_t ev[12], od[12];
  // row 0: plain products
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    uint64_t t0 = (uint64_t)a.l[j] * b.l[0];
    uint64_t t1 = (uint64_t)a.l[j + 1] * b.l[0];
    ev[j] = (uint32_t)t0;
    ev[j + 1] = (uint32_t)(t0 >> 32);
    od[j] = (uint32_t)t1;
    od[j + 1] = (uint32_t)(t1 >> 32);
  }
  {
    uint32_t m = ev[0] * PTAU_M0;
    row_red_odd(od, m);
    row_red_even(ev, od[11], m);
  }
#pragma unroll
  for (int i = 1; i < 12; i += 2) {
    {  // roles: E = od, X = ev
      row_mac_odd_shift(ev, od[0], a.l, b.l[i]);
      row_mac_even(od, ev[11], a.l, b.l[i]);
      uint32_t m = od[0] * PTAU_M0;
      row_red_odd(ev, m);
      row_red_even(od, ev[11], m);
    }
    if (i + 1 < 12) {  // roles: E = ev, X = od
      row_mac_odd_shift(od, ev[0], a.l, b.l[i + 1]);
      row_mac_even(ev, od[11], a.l, b.l[i + 1]);
      uint32_t m = ev[0] * PTAU_M0;
      row_red_odd(od, m);
      row_red_even(ev, od[11], m);
    }
  }
  // after 12 rows (last row had E = od, X = ev): result = ev + (od >> 32)
  Fq r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %0, %12;\n\t"
      "addc.cc.u32 %1, %1, %13;\n\t"
      "addc.cc.u32 %2, %2, %14;\n\t"
      "addc.cc.u32 %3, %3, %15;\n\t"
      "addc.cc.u32 %4, %4, %16;\n\t"
      "addc.cc.u32 %5, %5, %17;\n\t"
      "addc.cc.u32 %6, %6, %18;\n\t"
      "addc.cc.u32 %7, %7, %19;\n\t"
      "addc.cc.u32 %8, %8, %20;\n\t"
      "addc.cc.u32 %9, %9, %21;\n\t"
      "addc.cc.u32 %10, %10, %22;\n\t"
      "addc.u32 %11, %11, 0;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]),
        "+r"(ev[7]), "+r"(ev[8]), "+r"(ev[9]), "+r"(ev[10]), "+r"(ev[11])
      : "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]),
        "r"(od[9]), "r"(od[10]), "r"(od[11]));
#else
  {
    uint32_t cf = 0;
    for (int k = 0; k < 11; k++) ev[k] = emu::addc(ev[k], od[k + 1], cf);
    ev[11] = emu::addc(ev[11], 0, cf);
  }
#endif
  fq_cond_sub_p(ev);
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = ev[i];
  return r;
}

// ---------------------------------------------------------------------------
// single-instruction carry-chain steps (device: one PTX instruction each, the
// carry lives in CC.CF between them; host: explicit flag `px_cf` in scope)
// ---------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define PX_DECL
#define PX_ADD_CC(r, a, b) asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define PX_ADDC_CC(r, a, b) asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define PX_ADDC(r, a, b) asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define PX_MAD_LO_CC(r, a, b, c) asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c))
#define PX_MADC_LO_CC(r, a, b, c) asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c))
#define PX_MADC_HI_CC(r, a, b, c) asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c))
#define PX_MADC_HI(r, a, b, c) asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c))
#else
#define PX_DECL uint32_t px_cf = 0; (void)px_cf
#define PX_ADD_CC(r, a, b) do { px_cf = 0; r = emu::addc(a, b, px_cf); } while (0)
#define PX_ADDC_CC(r, a, b) r = emu::addc(a, b, px_cf)
#define PX_ADDC(r, a, b) do { r = emu::addc(a, b, px_cf); px_cf = 0; } while (0)
#define PX_MAD_LO_CC(r, a, b, c) do { px_cf = 0; r = emu::madlo(a, b, c, px_cf); } while (0)
#define PX_MADC_LO_CC(r, a, b, c) r = emu::madlo(a, b, c, px_cf)
#define PX_MADC_HI_CC(r, a, b, c) r = emu::madhi(a, b, c, px_cf)
#define PX_MADC_HI(r, a, b, c) do { r = emu::madhi(a, b, c, px_cf); px_cf = 0; } while (0)
#endif

// One reduction row with the window shift fused in.  On entry X is the previous even-aligned accumulator whose limb 0
// has just been zeroed (limb 1 is the left-over at the new position 0); E0 is limb 0 of the new even-aligned
// accumulator.  E0 += left-over; m = E0 * (-p^-1); X = (X >> 64) + m * (odd limbs of p) + tin W^10, carry chained from E0.
PTAU_HD void row_red_odd_shift(uint32_t* X, uint32_t& E0, uint32_t& m, uint32_t tin) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %12, %12, %1;\n\t"
      "mul.lo.u32 %13, %12, 0xfffcfffd;\n\t"
      "madc.lo.cc.u32 %0, %13, " P1S ", %2;\n\t"
      "madc.hi.cc.u32 %1, %13, " P1S ", %3;\n\t"
      "madc.lo.cc.u32 %2, %13, " P3S ", %4;\n\t"
      "madc.hi.cc.u32 %3, %13, " P3S ", %5;\n\t"
      "madc.lo.cc.u32 %4, %13, " P5S ", %6;\n\t"
      "madc.hi.cc.u32 %5, %13, " P5S ", %7;\n\t"
      "madc.lo.cc.u32 %6, %13, " P7S ", %8;\n\t"
      "madc.hi.cc.u32 %7, %13, " P7S ", %9;\n\t"
      "madc.lo.cc.u32 %8, %13, " P9S ", %10;\n\t"
      "madc.hi.cc.u32 %9, %13, " P9S ", %11;\n\t"
      "madc.lo.cc.u32 %10, %13, " P11S ", %14;\n\t"
      "madc.hi.cc.u32 %11, %13, " P11S ", 0;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]),
        "+r"(X[8]), "+r"(X[9]), "+r"(X[10]), "+r"(X[11]), "+r"(E0), "=&r"(m)
      : "r"(tin));
#else
  uint32_t cf = 0;
  E0 = emu::addc(E0, X[1], cf);
  m = E0 * PTAU_M0;
  for (int j = 0; j < 10; j += 2) {
    X[j] = emu::madlo(m, emu::PL[j + 1], X[j + 2], cf);
    X[j + 1] = emu::madhi(m, emu::PL[j + 1], X[j + 3], cf);
  }
  X[10] = emu::madlo(m, emu::PL[11], tin, cf);
  X[11] = emu::madhi(m, emu::PL[11], 0, cf);
#endif
}

// Montgomery square a*a/R mod p.  Row i only forms the products with limbs j >= i of the multiplicand
//     c(i) = a_i W^i + 2 * sum_{j>i} a_j W^j        (a^2 = sum_i a_i W^i c(i)),
// i.e. 78 wide MADs for the product instead of 144; the 12 reduction rows are unchanged.  Limbs of c(i):
// j == i -> a_i ; j == i+1 -> a_j << 1 ; j > i+1 -> (a_j << 1) | (a_{j-1} >> 31).  a < 2^383 so nothing is shifted out
// of limb 11.
// Row order: the products of row i >= 1 sit at window positions >= i, so they do not touch position 0 and the
// reduction of the row can run FIRST, fused with the one-limb shift of the window (row_red_odd_shift): every
// position gets its shift from a reduction MAD and the product chains simply start at their first position.
// (With the products first, the skipped low positions needed 60 add-with-carry instructions per squaring just to
// move the window; the ladders are issue-bound, not multiplier-bound, so those were pure cost.)
PTAU_HD Fq fq_sqr_inl(const Fq& a) {
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
  // row 0: all 12 products, plain
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    uint64_t t0 = (uint64_t)SQ_M(0, j) * a.l[0];
    uint64_t t1 = (uint64_t)SQ_M(0, j + 1) * a.l[0];
    ev[j] = (uint32_t)t0;
    ev[j + 1] = (uint32_t)(t0 >> 32);
    od[j] = (uint32_t)t1;
    od[j + 1] = (uint32_t)(t1 >> 32);
  }
  {
    uint32_t m = ev[0] * PTAU_M0;
    row_red_odd(od, m);
    row_red_even(ev, od[11], m);
  }
#pragma unroll
  for (int i = 1; i < 12; i++) {
    uint32_t* E = (i & 1) ? od : ev;  // even-aligned accumulator of this row
    uint32_t* X = (i & 1) ? ev : od;  // odd-aligned (the previous even accumulator: limb 0 is zero, limb 1 the left-over)
    const uint32_t bi = a.l[i];
    {  // reduction of this row, with the window shift
      uint32_t m;
      row_red_odd_shift(X, E[0], m, 0u);
      row_red_even(E, X[11], m);
    }
    PX_DECL;
    // odd-j products (j >= i) into X[j-1], X[j]; the chain starts at the first such j
    const int jo = (i & 1) ? i : i + 1;
    if (jo <= 11) {
      PX_MAD_LO_CC(X[jo - 1], SQ_M(i, jo), bi, X[jo - 1]);
      PX_MADC_HI_CC(X[jo], SQ_M(i, jo), bi, X[jo]);
#pragma unroll
      for (int j = jo + 2; j < 12; j += 2) {
        PX_MADC_LO_CC(X[j - 1], SQ_M(i, j), bi, X[j - 1]);
        PX_MADC_HI_CC(X[j], SQ_M(i, j), bi, X[j]);  // no carry leaves X[11] (bound on the window)
      }
    }
    // even-j products (j >= i) into E[j], E[j+1]; the carry of the chain goes to window position 12 = X[11]
    const int je = (i & 1) ? i + 1 : i;
    if (je <= 10) {
      PX_MAD_LO_CC(E[je], SQ_M(i, je), bi, E[je]);
      PX_MADC_HI_CC(E[je + 1], SQ_M(i, je), bi, E[je + 1]);
#pragma unroll
      for (int j = je + 2; j < 12; j += 2) {
        PX_MADC_LO_CC(E[j], SQ_M(i, j), bi, E[j]);
        PX_MADC_HI_CC(E[j + 1], SQ_M(i, j), bi, E[j + 1]);
      }
      PX_ADDC(X[11], X[11], 0u);
    }
  }
#undef SQ_M
  // last row (i = 11) had E = od, X = ev: result = ev + (od >> 32)
  {
    PX_DECL;
    PX_ADD_CC(ev[0], ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 11; k++) PX_ADDC_CC(ev[k], ev[k], od[k + 1]);
    PX_ADDC(ev[11], ev[11], 0u);
  }
  fq_cond_sub_p(ev);
  Fq r;
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = ev[i];
  return r;
}

// a + b mod p
PTAU_HD Fq fq_add(const Fq& a, const Fq& b) {
  Fq r;
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %12, %24;\n\t"
      "addc.cc.u32 %1, %13, %25;\n\t"
      "addc.cc.u32 %2, %14, %26;\n\t"
      "addc.cc.u32 %3, %15, %27;\n\t"
      "addc.cc.u32 %4, %16, %28;\n\t"
      "addc.cc.u32 %5, %17, %29;\n\t"
      "addc.cc.u32 %6, %18, %30;\n\t"
      "addc.cc.u32 %7, %19, %31;\n\t"
      "addc.cc.u32 %8, %20, %32;\n\t"
      "addc.cc.u32 %9, %21, %33;\n\t"
      "addc.cc.u32 %10, %22, %34;\n\t"
      "addc.u32 %11, %23, %35;"
      : "=&r"(r.l[0]), "=&r"(r.l[1]), "=&r"(r.l[2]), "=&r"(r.l[3]), "=&r"(r.l[4]), "=&r"(r.l[5]), "=&r"(r.l[6]),
        "=&r"(r.l[7]), "=&r"(r.l[8]), "=&r"(r.l[9]), "=&r"(r.l[10]), "=&r"(r.l[11])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]),
        "r"(a.l[7]), "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]), "r"(b.l[0]), "r"(b.l[1]),
        "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]), "r"(b.l[8]),
        "r"(b.l[9]), "r"(b.l[10]), "r"(b.l[11]));
#else
  uint32_t cf = 0;
  for (int i = 0; i < 12; i++) r.l[i] = emu::addc(a.l[i], b.l[i], cf);
#endif
  fq_cond_sub_p(r.l);
  return r;
}

// a - b mod p
PTAU_HD Fq fq_sub(const Fq& a, const Fq& b) {
  Fq r;
  uint32_t mask;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %13, %25;\n\t"
      "subc.cc.u32 %1, %14, %26;\n\t"
      "subc.cc.u32 %2, %15, %27;\n\t"
      "subc.cc.u32 %3, %16, %28;\n\t"
      "subc.cc.u32 %4, %17, %29;\n\t"
      "subc.cc.u32 %5, %18, %30;\n\t"
      "subc.cc.u32 %6, %19, %31;\n\t"
      "subc.cc.u32 %7, %20, %32;\n\t"
      "subc.cc.u32 %8, %21, %33;\n\t"
      "subc.cc.u32 %9, %22, %34;\n\t"
      "subc.cc.u32 %10, %23, %35;\n\t"
      "subc.cc.u32 %11, %24, %36;\n\t"
      "subc.u32 %12, 0, 0;"
      : "=&r"(r.l[0]), "=&r"(r.l[1]), "=&r"(r.l[2]), "=&r"(r.l[3]), "=&r"(r.l[4]), "=&r"(r.l[5]), "=&r"(r.l[6]),
        "=&r"(r.l[7]), "=&r"(r.l[8]), "=&r"(r.l[9]), "=&r"(r.l[10]), "=&r"(r.l[11]), "=&r"(mask)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]),
        "r"(a.l[7]), "r"(a.l[8]), "r"(a.l[9]), "r"(a.l[10]), "r"(a.l[11]), "r"(b.l[0]), "r"(b.l[1]),
        "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]), "r"(b.l[8]),
        "r"(b.l[9]), "r"(b.l[10]), "r"(b.l[11]));
  asm("{\n\t"
      ".reg .u32 t;\n\t"
      "and.b32 t, %12, " P0S ";\n\t add.cc.u32 %0, %0, t;\n\t"
      "and.b32 t, %12, " P1S ";\n\t addc.cc.u32 %1, %1, t;\n\t"
      "and.b32 t, %12, " P2S ";\n\t addc.cc.u32 %2, %2, t;\n\t"
      "and.b32 t, %12, " P3S ";\n\t addc.cc.u32 %3, %3, t;\n\t"
      "and.b32 t, %12, " P4S ";\n\t addc.cc.u32 %4, %4, t;\n\t"
      "and.b32 t, %12, " P5S ";\n\t addc.cc.u32 %5, %5, t;\n\t"
      "and.b32 t, %12, " P6S ";\n\t addc.cc.u32 %6, %6, t;\n\t"
      "and.b32 t, %12, " P7S ";\n\t addc.cc.u32 %7, %7, t;\n\t"
      "and.b32 t, %12, " P8S ";\n\t addc.cc.u32 %8, %8, t;\n\t"
      "and.b32 t, %12, " P9S ";\n\t addc.cc.u32 %9, %9, t;\n\t"
      "and.b32 t, %12, " P10S ";\n\t addc.cc.u32 %10, %10, t;\n\t"
      "and.b32 t, %12, " P11S ";\n\t addc.u32 %11, %11, t;\n\t"
      "}"
      : "+r"(r.l[0]), "+r"(r.l[1]), "+r"(r.l[2]), "+r"(r.l[3]), "+r"(r.l[4]), "+r"(r.l[5]), "+r"(r.l[6]),
        "+r"(r.l[7]), "+r"(r.l[8]), "+r"(r.l[9]), "+r"(r.l[10]), "+r"(r.l[11])
      : "r"(mask));
#else
  uint32_t bf = 0;
  for (int i = 0; i < 12; i++) r.l[i] = emu::subc(a.l[i], b.l[i], bf);
  mask = bf ? 0xffffffffu : 0u;
  uint32_t cf = 0;
  for (int i = 0; i < 12; i++) r.l[i] = emu::addc(r.l[i], emu::PL[i] & mask, cf);
#endif
  return r;
}

PTAU_HD Fq fq_neg(const Fq& a) {
  // p - a, with 0 -> 0
  Fq z = fq_zero();
  return fq_sub(z, a);
}

PTAU_HD Fq fq_dbl(const Fq& a) { return fq_add(a, a); }

// lexicographic compare of canonical (non-Montgomery) values: a > b
PTAU_HD bool fq_gt_plain(const Fq& a, const Fq& b) {
  // b - a borrows  <=>  a > b
  uint32_t bf = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)b.l[i] - a.l[i] - bf;
    bf = (uint32_t)(t >> 63);
  }
  return bf != 0;
}

// is the plain (non-Montgomery) 384-bit value >= p ?
PTAU_HD bool fq_plain_ge_p(const Fq& a) {
  const uint32_t pl[12] = PTAU_P_LIMBS;
  uint32_t bf = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)a.l[i] - pl[i] - bf;
    bf = (uint32_t)(t >> 63);
  }
  return bf == 0;
}

}  // namespace ptau
