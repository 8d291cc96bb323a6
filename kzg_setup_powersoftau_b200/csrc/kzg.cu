// KZG10::check on the GPU: the pairing side of ark-poly-commit 0.2 kzg10, data-parallel over
// openings -- one opening (two Miller loops sharing the squarings, one final exponentiation) per thread.
// Reference usage: /root/reference/src/lib.rs:276-286 (KZG10::check(&vk, &comm, point, value, &proof)).
#include "pairing.cuh"
#include "kernels.h"

namespace ptau {

#define PTAU_PAIR_BLOCK 64

static __device__ __forceinline__ void load_g1_rec(const uint32_t* rec, Fq& x, Fq& y, bool& inf) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x.l[i] = rec[i];
    y.l[i] = rec[12 + i];
  }
  inf = (rec[24] & 0xffu) != 0;
}
static __device__ __forceinline__ void load_g2_rec(const uint32_t* rec, Fq2& x, Fq2& y, bool& inf) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    x.c0.l[i] = rec[i];
    x.c1.l[i] = rec[12 + i];
    y.c0.l[i] = rec[24 + i];
    y.c1.l[i] = rec[36 + i];
  }
  inf = (rec[48] & 0xffu) != 0;
}

// p += q, both Jacobian, every special case
template <class F>
static __device__ __noinline__ void jac_add_complete_t(Jac<F>& p, const Jac<F>& q) {
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

// GT element -> 12 x 48 canonical little-endian bytes in arkworks' Fq12 order (c0.c0.c0, c0.c0.c1, c0.c1.c0, ...)
static __device__ __noinline__ void store_gt(uint32_t* out, const Fq12& f) {
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

// prod_{k<2} e(P_ik, Q_ik) for n items.  g1: n x 2 ARK_MONT_LIMBS G1 records, g2: n x 2 G2 records.
__global__ void __launch_bounds__(PTAU_PAIR_BLOCK) pairing_product2_kernel(const uint32_t* __restrict__ g1, const uint32_t* __restrict__ g2,
                                                                           uint64_t n, uint32_t* __restrict__ gt_out,
                                                                           uint8_t* __restrict__ is_one) {
  const uint64_t i = (uint64_t)blockIdx.x * PTAU_PAIR_BLOCK + threadIdx.x;
  if (i >= n) return;
  Fq px[2], py[2];
  Fq2 qx[2], qy[2];
  bool use[2];
#pragma unroll
  for (int k = 0; k < 2; k++) {
    bool pi, qi;
    load_g1_rec(g1 + (i * 2 + k) * 26, px[k], py[k], pi);
    load_g2_rec(g2 + (i * 2 + k) * 50, qx[k], qy[k], qi);
    use[k] = !pi && !qi;
  }
  Fq12 f;
  miller_loop2(f, px, py, qx, qy, use);
  final_exponentiation(f, f);
  if (gt_out) store_gt(gt_out + i * 144, f);
  if (is_one) is_one[i] = fq12_is_one(f) ? 1 : 0;
}

// KZG10::check for n openings: e(C - [v]g - [rv]gamma_g, h) == e(w, beta_h - [z]h), evaluated as
// e(inner, h) * e(-w, beta_h - [z]h) == 1.
__global__ void __launch_bounds__(PTAU_PAIR_BLOCK) kzg_check_kernel(const uint32_t* __restrict__ vk_g1, const uint32_t* __restrict__ vk_g2,
                                                                    const uint32_t* __restrict__ comms, const uint32_t* __restrict__ points,
                                                                    const uint32_t* __restrict__ values, const uint32_t* __restrict__ proofs,
                                                                    const uint32_t* __restrict__ random_v, uint64_t n,
                                                                    uint8_t* __restrict__ ok) {
  const uint64_t i = (uint64_t)blockIdx.x * PTAU_PAIR_BLOCK + threadIdx.x;
  if (i >= n) return;
  Fq px[2], py[2];
  Fq2 qx[2], qy[2];
  bool use[2];
  uint32_t k[8];
  bool hinf;
  load_g2_rec(vk_g2, qx[0], qy[0], hinf);
  {  // inner = C - [v] g - [rv] gamma_g
    Fq gx, gy;
    bool ginf;
    load_g1_rec(vk_g1, gx, gy, ginf);
    Jac<Fq> acc;
#pragma unroll
    for (int w = 0; w < 8; w++) k[w] = ginf ? 0u : values[i * 8 + w];
    jac_scalar_mul_t(acc, gx, gy, k, fq_one());
    if (random_v) {
      load_g1_rec(vk_g1 + 26, gx, gy, ginf);
#pragma unroll
      for (int w = 0; w < 8; w++) k[w] = ginf ? 0u : random_v[i * 8 + w];
      Jac<Fq> t;
      jac_scalar_mul_t(t, gx, gy, k, fq_one());
      jac_add_complete_t(acc, t);
    }
    acc.Y = fq_neg(acc.Y);
    Fq cx, cy;
    bool cinf;
    load_g1_rec(comms + i * 26, cx, cy, cinf);
    if (!cinf) jac_madd_complete_t(acc, cx, cy, fq_one());
    use[0] = jac_to_affine_t(acc, px[0], py[0]) && !hinf;
  }
  {  // Q = beta_h - [z] h
    Jac<Fq2> acc;
#pragma unroll
    for (int w = 0; w < 8; w++) k[w] = hinf ? 0u : points[i * 8 + w];
    jac_scalar_mul_t(acc, qx[0], qy[0], k, fq2_one());
    acc.Y = fq2_neg(acc.Y);
    Fq2 bx, by;
    bool binf;
    load_g2_rec(vk_g2 + 50, bx, by, binf);
    if (!binf) jac_madd_complete_t(acc, bx, by, fq2_one());
    bool winf;
    load_g1_rec(proofs + i * 26, px[1], py[1], winf);
    py[1] = fq_neg(py[1]);
    use[1] = jac_to_affine_t(acc, qx[1], qy[1]) && !winf;
  }
  Fq12 f;
  miller_loop2(f, px, py, qx, qy, use);
  final_exponentiation(f, f);
  ok[i] = fq12_is_one(f) ? 1 : 0;
}

cudaError_t launch_pairing_product2(const void* d_g1, const void* d_g2, uint64_t n, void* d_gt, void* d_is_one,
                                    cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  pairing_product2_kernel<<<(unsigned)((n + PTAU_PAIR_BLOCK - 1) / PTAU_PAIR_BLOCK), PTAU_PAIR_BLOCK, 0, stream>>>(
      (const uint32_t*)d_g1, (const uint32_t*)d_g2, n, (uint32_t*)d_gt, (uint8_t*)d_is_one);
  return cudaGetLastError();
}

cudaError_t launch_kzg_check(const void* d_vk_g1, const void* d_vk_g2, const void* d_comms, const void* d_points,
                             const void* d_values, const void* d_proofs, const void* d_random_v, uint64_t n, void* d_ok,
                             cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  kzg_check_kernel<<<(unsigned)((n + PTAU_PAIR_BLOCK - 1) / PTAU_PAIR_BLOCK), PTAU_PAIR_BLOCK, 0, stream>>>(
      (const uint32_t*)d_vk_g1, (const uint32_t*)d_vk_g2, (const uint32_t*)d_comms, (const uint32_t*)d_points,
      (const uint32_t*)d_values, (const uint32_t*)d_proofs, (const uint32_t*)d_random_v, n, (uint8_t*)d_ok);
  return cudaGetLastError();
}

}  // namespace ptau
