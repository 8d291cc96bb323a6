// Window geometry and signed-digit recoding of the bucket MSM (kernels.cu), host- and device-callable so that
// the CPU-only suite can check the recoding (tests/host_emul): sum_w d_w 2^bitoff(w) == k for every scalar.
#pragma once
#include <stdint.h>
#include "fq.cuh"  // PTAU_HD

namespace ptau {

// windows [0, a) are c bits wide with NB = 2^(c-1) buckets, windows [a, W) are c - 1 bits wide with NB / 2 buckets;
// a c + (W - a)(c - 1) = 256, so no window is partial and the top window ends at bit 255 (0 for every scalar < r)
struct MsmGeom {
  int c, W, a, lgL;
  uint32_t NB;
};
PTAU_HD MsmGeom msm_geometry(uint64_t n) {
  int lg = 0;
  while (lg < 63 && (1ull << lg) < n) lg++;
  int c = lg - 4;  // about 32 points per bucket
  c = c < 3 ? 3 : c > 16 ? 16 : c;
  MsmGeom g;
  g.c = c;
  g.W = (256 + c - 1) / c;
  g.a = 256 - (c - 1) * g.W;
  g.NB = 1u << (c - 1);
  g.lgL = c - 2 < 4 ? c - 2 : 4;  // runs of 2^lgL buckets in the window reduction; 2^lgL divides NB / 2
  return g;
}
PTAU_HD int msm_bitoff(const MsmGeom& g, int w) { return w < g.a ? w * g.c : g.a * g.c + (w - g.a) * (g.c - 1); }
PTAU_HD uint32_t msm_bucket_base(const MsmGeom& g, int w) {
  return w < g.a ? (uint32_t)w * g.NB : (uint32_t)g.a * g.NB + (uint32_t)(w - g.a) * (g.NB >> 1);
}

// signed digit of the cw-bit window (cw <= 16) starting at `bit` of the 256-bit little-endian scalar k, with the
// carry of the windows below.  v in [0, 2^cw]; v > 2^(cw-1) becomes v - 2^cw with a carry into the next window.
PTAU_HD int msm_digit(const uint32_t* k, int bit, int cw, uint32_t& carry) {
  const int wi = bit >> 5, sh = bit & 31;
  uint64_t t = k[wi];
  if (wi + 1 < 8) t |= (uint64_t)k[wi + 1] << 32;
  uint32_t v = ((uint32_t)(t >> sh) & ((1u << cw) - 1u)) + carry;
  if (v > (1u << (cw - 1))) {
    carry = 1;
    return (int)v - (1 << cw);
  }
  carry = 0;
  return (int)v;
}

}  // namespace ptau
