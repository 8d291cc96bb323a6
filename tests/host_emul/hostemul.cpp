// Host build of the *product's* per-point code (csrc/point.cuh compiled as plain
// C++ with the carry-flag emulation in fq.cuh).  TEST VEHICLE ONLY: lets the
// CPU-only test-suite check the exact limb-level algorithm the CUDA kernels run
// against the oracle.  libptau_b200.so never links or calls this.
#include <cstring>
#include "../../kzg_setup_powersoftau_b200/csrc/point.cuh"
#include "../../kzg_setup_powersoftau_b200/csrc/pairing.cuh"
#include "../../kzg_setup_powersoftau_b200/csrc/msm.cuh"
#include <vector>

using namespace ptau;

template <int G, int INFMT>
static uint32_t one(const uint8_t* in, int out_fmt, uint8_t* out, unsigned checks) {
  uint32_t win[50], wout[50];
  std::memcpy(win, in, record_bytes(G, INFMT));
  std::memset(wout, 0, sizeof(wout));
  uint32_t st = (G == PTAU_G1) ? g1_process<INFMT>(win, out_fmt, wout, checks)
                               : g2_process<INFMT>(win, out_fmt, wout, checks);
  std::memcpy(out, wout, record_bytes(G, out_fmt));
  return st;
}

extern "C" int hostemul_convert(int group, int in_fmt, const uint8_t* in, int out_fmt, uint8_t* out,
                                size_t n, unsigned checks, uint32_t* status) {
  size_t ri = record_bytes(group, in_fmt), ro = record_bytes(group, out_fmt);
  for (size_t i = 0; i < n; i++) {
    uint32_t st;
    const uint8_t* p = in + i * ri;
    uint8_t* q = out + i * ro;
    if (group == PTAU_G1) {
      if (in_fmt == PTAU_FMT_ZCASH_UNCOMPRESSED) st = one<PTAU_G1, PTAU_FMT_ZCASH_UNCOMPRESSED>(p, out_fmt, q, checks);
      else if (in_fmt == PTAU_FMT_ZCASH_COMPRESSED) st = one<PTAU_G1, PTAU_FMT_ZCASH_COMPRESSED>(p, out_fmt, q, checks);
      else if (in_fmt == PTAU_FMT_ARK_UNCOMPRESSED) st = one<PTAU_G1, PTAU_FMT_ARK_UNCOMPRESSED>(p, out_fmt, q, checks);
      else if (in_fmt == PTAU_FMT_ARK_MONT_LIMBS) st = one<PTAU_G1, PTAU_FMT_ARK_MONT_LIMBS>(p, out_fmt, q, checks);
      else return -2;
    } else {
      if (in_fmt == PTAU_FMT_ZCASH_UNCOMPRESSED) st = one<PTAU_G2, PTAU_FMT_ZCASH_UNCOMPRESSED>(p, out_fmt, q, checks);
      else if (in_fmt == PTAU_FMT_ZCASH_COMPRESSED) st = one<PTAU_G2, PTAU_FMT_ZCASH_COMPRESSED>(p, out_fmt, q, checks);
      else if (in_fmt == PTAU_FMT_ARK_UNCOMPRESSED) st = one<PTAU_G2, PTAU_FMT_ARK_UNCOMPRESSED>(p, out_fmt, q, checks);
      else if (in_fmt == PTAU_FMT_ARK_MONT_LIMBS) st = one<PTAU_G2, PTAU_FMT_ARK_MONT_LIMBS>(p, out_fmt, q, checks);
      else return -2;
    }
    status[i] = st;
  }
  return 0;
}

// raw field ops for limb-level tests: a,b,out are 12 x u32 Montgomery limbs
extern "C" void hostemul_fq_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fq x, y, r;
  std::memcpy(x.l, a, 48);
  std::memcpy(y.l, b, 48);
  switch (op) {
    case 0: r = fq_mul(x, y); break;
    case 1: r = fq_add(x, y); break;
    case 2: r = fq_sub(x, y); break;
    case 3: r = fq_neg(x); break;
    case 4: r = fq_sqr(x); break;
    case 5: r = fq_pow_p34(x); break;
    default: r = fq_zero();
  }
  std::memcpy(out, r.l, 48);
}

// wide (lazy-reduction) building blocks of csrc/fqw.cuh and the Fq2 operations built from them.
//   op 0: t[24] = a * b            (a, b: 12 limbs, any value < 2^384 whose product rows fit, i.e. < 2p)
//   op 1: t[24] = a * a            (a < 2^383)
//   op 2: out[12] = redc(t[24])    (t < p 2^384; passed in `a` as 24 limbs)
//   op 3: out[24] = fq2_mul(a, b)  (a, b: c0 | c1, 24 limbs, reduced)
//   op 4: out[24] = fq2_sqr(a)
//   op 5: out[24] = a - b + (a < b ? p 2^384 : 0) over 24 limbs (fqw_sub_fix)
//   op 8: out[12] = a - b + p (fq_sub_plus_p);  op 9: out[12] = a / 2 mod p (fq_half)
extern "C" void hostemul_fqw_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fq x, y;
  Fq2 u, v, w;
  switch (op) {
    case 0:
      std::memcpy(x.l, a, 48);
      std::memcpy(y.l, b, 48);
      fq_mul_wide(out, x, y);
      break;
    case 1:
      std::memcpy(x.l, a, 48);
      fq_sqr_wide(out, x);
      break;
    case 2:
      x = fq_redc(a);
      std::memcpy(out, x.l, 48);
      break;
    case 3:
      std::memcpy(&u, a, 96);
      std::memcpy(&v, b, 96);
      w = fq2_mul(u, v);
      std::memcpy(out, &w, 96);
      break;
    case 4:
      std::memcpy(&u, a, 96);
      w = fq2_sqr(u);
      std::memcpy(out, &w, 96);
      break;
    case 5: {
      uint32_t t[24];
      std::memcpy(t, a, 96);
      fqw_sub_fix(t, b);
      std::memcpy(out, t, 96);
      break;
    }
    case 6:  // the row-by-row 12 x 12 product
      std::memcpy(x.l, a, 48);
      std::memcpy(y.l, b, 48);
      fq_mul_wide_plain(out, x, y);
      break;
    case 7:  // one level of Karatsuba
      std::memcpy(x.l, a, 48);
      std::memcpy(y.l, b, 48);
      fq_mul_wide_k(out, x, y);
      break;
    case 8:  // a - b + p without a conditional step (fqw.cuh), 12 limbs
      std::memcpy(x.l, a, 48);
      std::memcpy(y.l, b, 48);
      x = fq_sub_plus_p(x, y);
      std::memcpy(out, x.l, 48);
      break;
    case 9:  // a / 2 mod p (pairing.cuh), 12 limbs
      std::memcpy(x.l, a, 48);
      x = fq_half(x);
      std::memcpy(out, x.l, 48);
      break;
  }
}

// one Jacobian doubling in G2 (csrc/curve.cuh): variant 0 = dbl-2009-l as the other code uses it, 1 = the ladder's
// doubling with C never reduced (jac_dbl_lazyc; coordinates scaled by 1/2).  in / out: X | Y | Z, 72 Montgomery limbs.
extern "C" void hostemul_g2_dbl(int variant, const uint32_t* in, uint32_t* out) {
  Jac<Fq2> q;
  std::memcpy(&q, in, 288);
  if (variant == 0) {
    jac_dbl(q);
  } else {
    uint32_t file[96];
    Park<1> pk;
    pk.p = file;
    jac_dbl_lazyc(q, pk);
  }
  std::memcpy(out, &q, 288);
}
// a - b + 2p (fqw.cuh), 12 limbs
extern "C" void hostemul_fq_sub_plus_2p(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fq x, y;
  std::memcpy(x.l, a, 48);
  std::memcpy(y.l, b, 48);
  x = fq_sub_plus_2p(x, y);
  std::memcpy(out, x.l, 48);
}

// the product's pairing code (csrc/pairing.cuh) on the host: same limb arithmetic, same formulas
extern "C" void hostemul_pairing_product2(const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* gt_out, uint8_t* is_one) {
  for (size_t i = 0; i < n; i++) {
    uint32_t a[52], b[100], gt[144];
    std::memcpy(a, g1 + i * 208, 208);
    std::memcpy(b, g2 + i * 400, 400);
    is_one[i] = pairing_product2_item(a, b, gt) ? 1 : 0;
    std::memcpy(gt_out + i * 576, gt, 576);
  }
}
// G2Prepared line coefficients (csrc/pairing.cuh: g2_prepare_item) on the host
extern "C" void hostemul_g2_prepare(const uint8_t* g2, size_t n, uint8_t* coeffs_out, uint8_t* infinity) {
  for (size_t i = 0; i < n; i++) {
    uint32_t rec[50];
    static uint32_t out[PTAU_G2PREP_COEFFS * 72];
    std::memcpy(rec, g2 + i * 200, 200);
    infinity[i] = g2_prepare_item(rec, out) ? 1 : 0;
    std::memcpy(coeffs_out + i * sizeof(out), out, sizeof(out));
  }
}
// use_tables != 0: fixed-base window tables for g, gamma_g, h (the path the library takes), else double-and-add
extern "C" void hostemul_kzg_check(const uint8_t* vk_g1, const uint8_t* vk_g2, const uint8_t* comms, const uint8_t* points,
                                   const uint8_t* values, const uint8_t* proofs, const uint8_t* random_v, size_t n, uint8_t* ok,
                                   int use_tables) {
  uint32_t v1[52], v2[100];
  std::memcpy(v1, vk_g1, 208);
  std::memcpy(v2, vk_g2, 400);
  static uint32_t tg[PTAU_FB_ENTRIES * 26], tgg[PTAU_FB_ENTRIES * 26], th[PTAU_FB_ENTRIES * 50], ph[PTAU_G2PREP_COEFFS * 72];
  static uint32_t tg8[PTAU_FB8_ENTRIES * 26], tgg8[PTAU_FB8_ENTRIES * 26], th8[PTAU_FB8_ENTRIES * 50];
  if (use_tables) {  // 1: 4-bit tables; 2: the second-level 8-bit tables the library's check kernel uses
    g2_prepare_item(v2, ph);
    for (int w = 0; w < PTAU_FB_WINDOWS; w++) {
      fixed_base_window<Fq>(tg, v1, w, fq_one());
      fixed_base_window<Fq>(tgg, v1 + 26, w, fq_one());
      fixed_base_window<Fq2>(th, v2, w, fq2_one());
    }
    if (use_tables == 2)
      for (int w = 0; w < PTAU_FB8_WINDOWS; w++)
        for (int hi = 0; hi < 16; hi++) {
          fixed_base_window8<Fq>(tg8, tg, w, hi, fq_one());
          fixed_base_window8<Fq>(tgg8, tgg, w, hi, fq_one());
          fixed_base_window8<Fq2>(th8, th, w, hi, fq2_one());
        }
  }
  for (size_t i = 0; i < n; i++) {
    uint32_t c[26], w[26], z[8], v[8], rv[8];
    std::memcpy(c, comms + i * 104, 104);
    std::memcpy(w, proofs + i * 104, 104);
    std::memcpy(z, points + i * 32, 32);
    std::memcpy(v, values + i * 32, 32);
    if (random_v) std::memcpy(rv, random_v + i * 32, 32);
    ok[i] = (use_tables == 2   ? kzg_check_item(v1, v2, c, z, v, w, random_v ? rv : nullptr, tg8, tgg8, th8, ph, true)
             : use_tables == 1 ? kzg_check_item(v1, v2, c, z, v, w, random_v ? rv : nullptr, tg, tgg, th, ph)
                               : kzg_check_item(v1, v2, c, z, v, w, random_v ? rv : nullptr))
                ? 1
                : 0;
  }
}

// bucket-MSM recoding: geometry for n terms and the signed digits of one scalar (digits[w], bit offsets, bucket bases);
// returns the final carry (must be 0 for scalars < 2^255)
extern "C" int hostemul_msm_recode(uint64_t n, const uint32_t* k, int* geom /*c, W, a, lgL, NB*/, int* digits, int* bitoff,
                                   uint32_t* bases) {
  MsmGeom g = msm_geometry(n);
  geom[0] = g.c;
  geom[1] = g.W;
  geom[2] = g.a;
  geom[3] = g.lgL;
  geom[4] = (int)g.NB;
  uint32_t carry = 0;
  int bit = 0;
  for (int w = 0; w < g.W; w++) {
    const int cw = w < g.a ? g.c : g.c - 1;
    digits[w] = msm_digit(k, bit, cw, carry);
    bitoff[w] = msm_bitoff(g, w);
    bases[w] = msm_bucket_base(g, w);
    if (bitoff[w] != bit) return -1;
    bit += cw;
  }
  return bit == 256 ? (int)carry : -2;
}

// the whole bucket MSM with the product's per-item code (csrc/msm.cuh), the kernels' loops run serially
extern "C" void hostemul_msm_g1(const uint8_t* pts8, const uint8_t* sc8, size_t n, uint8_t* out104) {
  std::vector<uint32_t> pts(n * 26 + 1), sc(n * 8 + 1);
  std::memcpy(pts.data(), pts8, n * 104);
  std::memcpy(sc.data(), sc8, n * 32);
  const MsmGeom g = msm_geometry(n);
  const uint32_t m = (uint32_t)g.a * g.NB + (uint32_t)(g.W - g.a) * (g.NB >> 1);
  std::vector<uint32_t> cnt(m + 1, 0), off(m + 1, 0), cur(m + 1, 0), entries(n * (size_t)g.W + 1);
  auto digits = [&](bool scatter) {
    for (size_t i = 0; i < n; i++) {
      if (pts[i * 26 + 24] & 0xffu) continue;
      uint32_t carry = 0, base = 0;
      int bit = 0;
      for (int w = 0; w < g.W; w++) {
        const int cw = w < g.a ? g.c : g.c - 1;
        int d = msm_digit(&sc[i * 8], bit, cw, carry);
        if (d != 0) {
          uint32_t b = base + (uint32_t)(d < 0 ? -d : d) - 1u;
          if (scatter) entries[cur[b]++] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
          else cnt[b]++;
        }
        bit += cw;
        base += 1u << (cw - 1);
      }
    }
  };
  digits(false);
  for (uint32_t b = 0; b < m; b++) off[b + 1] = off[b] + cnt[b];
  cur = off;
  digits(true);
  std::vector<uint32_t> buckets((size_t)m * 36), segs(((size_t)m >> g.lgL) * 36), wsum((size_t)g.W * 36);
  for (uint32_t b = 0; b < m; b++) msm_bucket_item(pts.data(), entries.data(), off.data(), b, buckets.data());
  const uint32_t nseg = m >> g.lgL;
  for (uint32_t t = 0; t < nseg; t++) msm_segment_item(buckets.data(), g, t, segs.data());
  for (int w = 0; w < g.W; w++) {
    const uint32_t first = msm_bucket_base(g, w) >> g.lgL, count = (w < g.a ? g.NB : g.NB >> 1) >> g.lgL;
    Jac<Fq> acc = jac_infinity();
    for (uint32_t i = 0; i < count; i++) {
      Jac<Fq> q = jac_load(segs.data() + (size_t)(first + i) * 36);
      g1_add_complete(acc, q);
    }
    msm_window_weight(acc, g, w);
    jac_store(wsum.data() + (size_t)w * 36, acc);
  }
  uint32_t out[26];
  msm_finish_item(wsum.data(), g.W, out);
  std::memcpy(out104, out, 104);
}
