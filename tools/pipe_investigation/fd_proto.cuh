// Prototype: BLS12-381 Fq on the FP64 pipe.  8 limbs of 48 bits held in doubles, Montgomery R = 2^384
// (same R as the 12x32 integer code, so conversions are pure re-packing).
// Values are kept in [0, 2p); limbs normalised to [0, 2^48).
#pragma once
#include <stdint.h>
#include "../../kzg_setup_powersoftau_b200/csrc/fq.cuh"

namespace ptau {

struct Fd { double l[8]; };

#define FD_C1 1267650600228229401496703205376.0   /* 2^100 */
#define FD_2M48 3.552713678800501e-15              /* 2^-48 */
#define FD_2P52 4503599627370496.0                 /* 2^52 */

// p limbs (48-bit) and -p^-1 mod 2^48, as exact doubles
__device__ __forceinline__ double fd_p(int j) {
  switch (j) {
    case 0: return (double)0xffffffffaaabull;
    case 1: return (double)0xb153ffffb9feull;
    case 2: return (double)0xf6241eabfffeull;
    case 3: return (double)0x6730d2a0f6b0ull;
    case 4: return (double)0x4b84f38512bfull;
    case 5: return (double)0x434bacd76477ull;
    case 6: return (double)0xe69a4b1ba7b6ull;
    default: return (double)0x1a0111ea397full;
  }
}
#define FD_PINV ((double)0xfffcfffcfffdull) /* patched by gen: -p^-1 mod 2^48 */

// one product a*b split at 2^48: high part accumulated in the chain H (ulp 2^48), low part into L
__device__ __forceinline__ void fd_prod(double a, double b, double& H, double& L) {
  double hn = __fma_rz(a, b, H);
  double d = __dsub_rn(H, hn);
  double lo = __fma_rn(a, b, d);
  L = __dadd_rn(L, lo);
  H = hn;
}

__device__ __forceinline__ Fd fd_mul(const Fd& a, const Fd& b) {
  double H[16], L[16];
#pragma unroll
  for (int k = 0; k < 16; k++) { H[k] = FD_C1; L[k] = 0.0; }
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) fd_prod(a.l[i], b.l[j], H[i + j], L[i + j]);
  double carry = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    double V = L[i];
    if (i > 0) {
      V = __dadd_rn(V, __fma_rn(H[i - 1], FD_2M48, -FD_2P52));
      V = __fma_rn(carry, FD_2M48, V);
    }
    double hq = __dadd_rz(V, FD_C1);
    double qv = __dsub_rn(hq, FD_C1);
    double t = __dsub_rn(V, qv);
    double hm = __fma_rz(t, FD_PINV, FD_C1);
    double dm = __dsub_rn(FD_C1, hm);
    double m = __fma_rn(t, FD_PINV, dm);
    {
      double hn = __fma_rz(m, fd_p(0), H[i]);
      double d = __dsub_rn(H[i], hn);
      double lo = __fma_rn(m, fd_p(0), d);
      H[i] = hn;
      V = __dadd_rn(V, lo);
    }
#pragma unroll
    for (int j = 1; j < 8; j++) fd_prod(m, fd_p(j), H[i + j], L[i + j]);
    carry = V;
  }
  Fd r;
#pragma unroll
  for (int k = 8; k < 16; k++) {
    double V = __dadd_rn(L[k], __fma_rn(H[k - 1], FD_2M48, -FD_2P52));
    V = __fma_rn(carry, FD_2M48, V);
    if (k < 15) {
      double hq = __dadd_rz(V, FD_C1);
      double qv = __dsub_rn(hq, FD_C1);
      r.l[k - 8] = __dsub_rn(V, qv);
      carry = qv;
    } else {
      r.l[7] = V;
    }
  }
  return r;
}

// 12 x u32 (value < 2^384) -> 8 x 48-bit doubles
__device__ __forceinline__ Fd fd_from_fq(const Fq& a) {
  Fd r;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    uint32_t w0 = a.l[3 * m], w1 = a.l[3 * m + 1], w2 = a.l[3 * m + 2];
    uint32_t lo0 = w0, hi0 = 0x43300000u | (w1 & 0xffffu);
    uint32_t lo1 = (w1 >> 16) | (w2 << 16), hi1 = 0x43300000u | (w2 >> 16);
    r.l[2 * m] = __dsub_rn(__hiloint2double(hi0, lo0), FD_2P52);
    r.l[2 * m + 1] = __dsub_rn(__hiloint2double(hi1, lo1), FD_2P52);
  }
  return r;
}
// normalised limbs -> 12 x u32
__device__ __forceinline__ Fq fd_to_fq(const Fd& a) {
  Fq r;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    double e = __dadd_rn(a.l[2 * m], FD_2P52), o = __dadd_rn(a.l[2 * m + 1], FD_2P52);
    uint32_t lo0 = __double2loint(e), hi0 = __double2hiint(e) & 0xffffu;
    uint32_t lo1 = __double2loint(o), hi1 = __double2hiint(o) & 0xffffu;
    r.l[3 * m] = lo0;
    r.l[3 * m + 1] = hi0 | (lo1 << 16);
    r.l[3 * m + 2] = (lo1 >> 16) | (hi1 << 16);
  }
  return r;
}

}  // namespace ptau
