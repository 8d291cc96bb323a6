// Per-item code of the bucket MSM (kernels.cu holds the kernels, which are thin loops over these), host- and
// device-callable so that tests/host_emul can run the whole algorithm on the CPU.
#pragma once
#include "curve.cuh"
#include "msm_digits.cuh"

namespace ptau {

// acc += (x, y) with every special case of the group law (madd-2007-bl, 7M + 4S; the quantities that
// decide the special cases are the ones the formula needs anyway)
PTAU_HD_NOINLINE void g1_madd_complete(Jac<Fq>& acc, const Fq& x, const Fq& y) {
  if (fq_is_zero(acc.Z)) {
    acc.X = x;
    acc.Y = y;
    acc.Z = fq_one();
    return;
  }
  Fq zz = fq_sqr(acc.Z);
  Fq u2 = fq_mul(x, zz);
  Fq s2 = fq_mul(fq_mul(y, acc.Z), zz);
  if (fq_eq(u2, acc.X)) {
    if (fq_eq(s2, acc.Y)) {
      jac_dbl(acc);
    } else {
      acc.Z = fq_zero();  // P + (-P)
    }
    return;
  }
  Fq H = fq_sub(u2, acc.X);
  Fq I = fq_sqr(fq_dbl(H));
  Fq J = fq_mul(H, I);
  Fq rr = fq_dbl(fq_sub(s2, acc.Y));
  Fq V = fq_mul(acc.X, I);
  Fq X3 = fq_sub(fq_sub(fq_sqr(rr), J), fq_dbl(V));
  acc.Y = fq_sub(fq_mul(rr, fq_sub(V, X3)), fq_dbl(fq_mul(acc.Y, J)));
  acc.Z = fq_dbl(fq_mul(acc.Z, H));
  acc.X = X3;
}
// p += q, both Jacobian (add-2007-bl, 11M + 5S), every special case
PTAU_HD_NOINLINE void g1_add_complete(Jac<Fq>& p, const Jac<Fq>& q) {
  if (fq_is_zero(q.Z)) return;
  if (fq_is_zero(p.Z)) {
    p = q;
    return;
  }
  Fq z1z1 = fq_sqr(p.Z), z2z2 = fq_sqr(q.Z);
  Fq u1 = fq_mul(p.X, z2z2), u2 = fq_mul(q.X, z1z1);
  Fq s1 = fq_mul(fq_mul(p.Y, q.Z), z2z2), s2 = fq_mul(fq_mul(q.Y, p.Z), z1z1);
  if (fq_eq(u1, u2)) {
    if (fq_eq(s1, s2)) {
      jac_dbl(p);
    } else {
      p.Z = fq_zero();
    }
    return;
  }
  Fq H = fq_sub(u2, u1);
  Fq I = fq_sqr(fq_dbl(H));
  Fq J = fq_mul(H, I);
  Fq rr = fq_dbl(fq_sub(s2, s1));
  Fq V = fq_mul(u1, I);
  Fq X3 = fq_sub(fq_sub(fq_sqr(rr), J), fq_dbl(V));
  p.Y = fq_sub(fq_mul(rr, fq_sub(V, X3)), fq_dbl(fq_mul(s1, J)));
  p.Z = fq_mul(fq_dbl(fq_mul(p.Z, q.Z)), H);
  p.X = X3;
}
PTAU_HD Jac<Fq> jac_infinity() {
  Jac<Fq> a;
  a.X = fq_zero();
  a.Y = fq_one();
  a.Z = fq_zero();
  return a;
}
PTAU_HD void jac_store(uint32_t* o, const Jac<Fq>& a) {
#pragma unroll
  for (int w = 0; w < 12; w++) {
    o[w] = a.X.l[w];
    o[12 + w] = a.Y.l[w];
    o[24 + w] = a.Z.l[w];
  }
}
PTAU_HD Jac<Fq> jac_load(const uint32_t* s) {
  Jac<Fq> q;
#pragma unroll
  for (int w = 0; w < 12; w++) {
    q.X.l[w] = s[w];
    q.Y.l[w] = s[12 + w];
    q.Z.l[w] = s[24 + w];
  }
  return q;
}

// one 48-byte field element of an ARK_MONT_LIMBS record (8-byte aligned)
PTAU_HD Fq msm_load_fq(const uint32_t* p) {
  Fq r;
#ifdef __CUDA_ARCH__
  const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint2 v = __ldg(q + i);
    r.l[2 * i] = v.x;
    r.l[2 * i + 1] = v.y;
  }
#else
  for (int i = 0; i < 12; i++) r.l[i] = p[i];
#endif
  return r;
}

// bucket b: sum of its points (entries[off[b] .. off[b+1]): point index, sign in bit 31)
PTAU_HD void msm_bucket_item(const uint32_t* pts, const uint32_t* entries, const uint32_t* off, uint32_t b, uint32_t* buckets) {
  Jac<Fq> acc = jac_infinity();
  const uint32_t e1 = off[b + 1];
#pragma unroll 1
  for (uint32_t e = off[b]; e < e1; e++) {
    const uint32_t v = entries[e];
    const uint32_t* rec = pts + (uint64_t)(v & 0x7fffffffu) * 26;
    Fq x = msm_load_fq(rec), y = msm_load_fq(rec + 12);
    if (v >> 31) y = fq_neg(y);
    g1_madd_complete(acc, x, y);
  }
  jac_store(buckets + (uint64_t)b * 36, acc);
}

// run t_id of L = 2^lgL consecutive buckets of one window: sum_{j in run} j * B_j, where bucket index j0 (0-based)
// holds the points of digit magnitude j0 + 1.  Running sums from the top give sum (j0 - lo + 1) B_j0 and
// t = sum B_j0; the run's offset adds [lo] t.
PTAU_HD void msm_segment_item(const uint32_t* buckets, const MsmGeom& g, uint32_t t_id, uint32_t* seg) {
  // runs are laid out exactly like the buckets, L buckets per run
  const uint32_t lo_abs = t_id << g.lgL;
  const uint32_t wide = (uint32_t)g.a * g.NB;
  const uint32_t in_window = lo_abs < wide ? lo_abs % g.NB : (lo_abs - wide) % (g.NB >> 1);
  const uint32_t sidx = in_window >> g.lgL;
  const uint32_t* base = buckets + (uint64_t)lo_abs * 36;
  Jac<Fq> t = jac_infinity(), sacc = jac_infinity();
#pragma unroll 1
  for (int j = (1 << g.lgL) - 1; j >= 0; --j) {
    Jac<Fq> q = jac_load(base + (uint64_t)j * 36);
    g1_add_complete(t, q);
    g1_add_complete(sacc, t);
  }
  if (sidx) {  // sacc += [lo] t = [2^lgL] [sidx] t
    Jac<Fq> m = jac_infinity();
    int top = 31;
    while (!((sidx >> top) & 1u)) top--;
#pragma unroll 1
    for (int bit = top; bit >= 0; --bit) {
      if (!fq_is_zero(m.Z)) jac_dbl(m);
      if ((sidx >> bit) & 1u) g1_add_complete(m, t);
    }
#pragma unroll 1
    for (int k = 0; k < g.lgL; k++)
      if (!fq_is_zero(m.Z)) jac_dbl(m);
    g1_add_complete(sacc, m);
  }
  jac_store(seg + (uint64_t)t_id * 36, sacc);
}

// dbl-2009-l with the multiplications expanded in place: the weight of a window is a chain of up to 240 dependent
// doublings in ONE thread, so what counts is latency -- inlined, the three independent products of every level
// (Y^2, ZY, X^2; then B^2, (X+B)^2, E^2) overlap in the pipe instead of running one call after the other
PTAU_HD void jac_dbl_latency(Jac<Fq>& p) {
  Fq B = fq_sqr_inl(p.Y);
  Fq ZY = fq_mul_inl(p.Z, p.Y);
  Fq A = fq_sqr_inl(p.X);
  Fq C = fq_sqr_inl(B);
  Fq D = fq_sqr_inl(fq_add(p.X, B));
  Fq E = fq_add(fq_dbl(A), A);
  Fq Fv = fq_sqr_inl(E);
  D = fq_dbl(fq_sub(fq_sub(D, A), C));
  p.Z = fq_dbl(ZY);
  p.X = fq_sub(Fv, fq_dbl(D));
  C = fq_dbl(fq_dbl(fq_dbl(C)));
  p.Y = fq_sub(fq_mul_inl(fq_sub(D, p.X), E), C);
}
// the window's weight 2^bitoff(w)
PTAU_HD void msm_window_weight(Jac<Fq>& acc, const MsmGeom& g, int w) {
  const int nd = msm_bitoff(g, w);
  if (fq_is_zero(acc.Z)) return;  // later on Z3 = 2 Y Z keeps an infinity (Z = 0) an infinity, whatever X and Y become
#pragma unroll 1
  for (int k = 0; k < nd; k++) jac_dbl_latency(acc);
}

// sum of the weighted window sums, then one ARK_MONT_LIMBS record (affine, or ark zero() = (0, 1, infinity))
PTAU_HD void msm_finish_item(const uint32_t* wsum, int W, uint32_t* out) {
  Jac<Fq> acc = jac_infinity();
#pragma unroll 1
  for (int w = 0; w < W; w++) {
    Jac<Fq> q = jac_load(wsum + (uint64_t)w * 36);
    g1_add_complete(acc, q);
  }
  if (fq_is_zero(acc.Z)) {
    Fq one = fq_one();
#pragma unroll
    for (int w = 0; w < 12; w++) {
      out[w] = 0;
      out[12 + w] = one.l[w];
    }
    out[24] = 1;
    out[25] = 0;
  } else {
    Fq zi = fq_inv_fermat(acc.Z);
    Fq zi2 = fq_sqr(zi);
    Fq x = fq_mul(acc.X, zi2), y = fq_mul(acc.Y, fq_mul(zi2, zi));
#pragma unroll
    for (int w = 0; w < 12; w++) {
      out[w] = x.l[w];
      out[12 + w] = y.l[w];
    }
    out[24] = 0;
    out[25] = 0;
  }
}

}  // namespace ptau
