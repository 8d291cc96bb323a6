// KZG10::check on the GPU: the pairing side of ark-poly-commit 0.2 kzg10, data-parallel over
// openings -- one opening (two Miller loops sharing the squarings, one final exponentiation) per thread.
// Reference usage: /root/reference/src/lib.rs:276-286 (KZG10::check(&vk, &comm, point, value, &proof)).
#include "pairing.cuh"
#include "kernels.h"

namespace ptau {

#ifndef PTAU_PAIR_BLOCK
#define PTAU_PAIR_BLOCK 128  // 255 registers: two blocks per SM, two warps per scheduler
#endif

// prod_{k<2} e(P_ik, Q_ik) for n items.  g1: n x 2 ARK_MONT_LIMBS G1 records, g2: n x 2 G2 records.
__global__ void __launch_bounds__(PTAU_PAIR_BLOCK) pairing_product2_kernel(const uint32_t* __restrict__ g1, const uint32_t* __restrict__ g2,
                                                                           uint64_t n, uint32_t* __restrict__ gt_out,
                                                                           uint8_t* __restrict__ is_one) {
  uint64_t i = (uint64_t)blockIdx.x * PTAU_PAIR_BLOCK + threadIdx.x;
  const bool live = i < n;  // a thread past the end repeats the last item: one instruction stream per block
  if (!live) i = n - 1;
  uint32_t* gt = gt_out ? gt_out + i * 144 : nullptr;
  bool one = pairing_product2_item(g1 + i * 52, g2 + i * 100, live ? gt : nullptr);
  if (is_one && live) is_one[i] = one ? 1 : 0;
}

// fixed-base tables of g, gamma_g (blocks 0, 1) and h (block 2): one window per thread; block 3: the line coefficients
// of h (ark's G2Prepared), which every opening's first pairing shares.  tbl = [g | gamma_g | h | prepared h]
#define PTAU_KZG_TBL_PREP_OFF ((size_t)PTAU_FB_ENTRIES * (26 + 26 + 50))
// then the 8-bit second-level tables [g | gamma_g | h], which are what the check kernel reads
#define PTAU_KZG_TBL8_OFF (PTAU_KZG_TBL_PREP_OFF + (size_t)PTAU_G2PREP_COEFFS * 72)
#define PTAU_KZG_TBL_WORDS (PTAU_KZG_TBL8_OFF + (size_t)PTAU_FB8_ENTRIES * (26 + 26 + 50))
__global__ void __launch_bounds__(PTAU_FB_WINDOWS) kzg_tables_kernel(const uint32_t* __restrict__ vk_g1, const uint32_t* __restrict__ vk_g2,
                                                                     uint32_t* __restrict__ tbl) {
  const int w = threadIdx.x;
  if (blockIdx.x < 2)
    fixed_base_window<Fq>(tbl + (size_t)blockIdx.x * PTAU_FB_ENTRIES * 26, vk_g1 + blockIdx.x * 26, w, fq_one());
  else if (blockIdx.x == 2)
    fixed_base_window<Fq2>(tbl + (size_t)2 * PTAU_FB_ENTRIES * 26, vk_g2, w, fq2_one());
  else if (w == 0)
    g2_prepare_item(vk_g2, tbl + PTAU_KZG_TBL_PREP_OFF);
}

// second-level tables: thread = (window, high nibble), blockIdx.y = base
__global__ void __launch_bounds__(64) kzg_tables8_kernel(uint32_t* __restrict__ tbl) {
  const int idx = blockIdx.x * 64 + threadIdx.x, w = idx >> 4, hi = idx & 15;
  if (w >= PTAU_FB8_WINDOWS) return;
  uint32_t* t8 = tbl + PTAU_KZG_TBL8_OFF;
  if (blockIdx.y < 2)
    fixed_base_window8<Fq>(t8 + (size_t)blockIdx.y * PTAU_FB8_ENTRIES * 26, tbl + (size_t)blockIdx.y * PTAU_FB_ENTRIES * 26, w, hi, fq_one());
  else
    fixed_base_window8<Fq2>(t8 + (size_t)2 * PTAU_FB8_ENTRIES * 26, tbl + (size_t)2 * PTAU_FB_ENTRIES * 26, w, hi, fq2_one());
}

// KZG10::check for n openings, one per thread
__global__ void __launch_bounds__(PTAU_PAIR_BLOCK) kzg_check_kernel(const uint32_t* __restrict__ vk_g1, const uint32_t* __restrict__ vk_g2,
                                                                    const uint32_t* __restrict__ comms, const uint32_t* __restrict__ points,
                                                                    const uint32_t* __restrict__ values, const uint32_t* __restrict__ proofs,
                                                                    const uint32_t* __restrict__ random_v, const uint32_t* __restrict__ tbl,
                                                                    uint64_t n, uint8_t* __restrict__ ok) {
  uint64_t i = (uint64_t)blockIdx.x * PTAU_PAIR_BLOCK + threadIdx.x;
  const bool live = i < n;  // a thread past the end repeats the last opening: one instruction stream per block
  if (!live) i = n - 1;
  const bool good = kzg_check_item(vk_g1, vk_g2, comms + i * 26, points + i * 8, values + i * 8, proofs + i * 26,
                                   random_v ? random_v + i * 8 : nullptr, tbl + PTAU_KZG_TBL8_OFF,
                                   tbl + PTAU_KZG_TBL8_OFF + (size_t)PTAU_FB8_ENTRIES * 26,
                                   tbl + PTAU_KZG_TBL8_OFF + (size_t)2 * PTAU_FB8_ENTRIES * 26, tbl + PTAU_KZG_TBL_PREP_OFF, true);
  if (live) ok[i] = good ? 1 : 0;
}

// G2Prepared line coefficients, one point per thread
__global__ void __launch_bounds__(PTAU_PAIR_BLOCK) g2_prepare_kernel(const uint32_t* __restrict__ g2, uint64_t n, uint32_t* __restrict__ coeffs,
                                                                     uint8_t* __restrict__ infinity) {
  const uint64_t i = (uint64_t)blockIdx.x * PTAU_PAIR_BLOCK + threadIdx.x;
  if (i >= n) return;
  infinity[i] = g2_prepare_item(g2 + i * 50, coeffs + i * (uint64_t)(PTAU_G2PREP_COEFFS * 72)) ? 1 : 0;
}

cudaError_t launch_g2_prepare(const void* d_g2, uint64_t n, void* d_coeffs, void* d_infinity, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  g2_prepare_kernel<<<(unsigned)((n + PTAU_PAIR_BLOCK - 1) / PTAU_PAIR_BLOCK), PTAU_PAIR_BLOCK, 0, stream>>>(
      (const uint32_t*)d_g2, n, (uint32_t*)d_coeffs, (uint8_t*)d_infinity);
  return cudaGetLastError();
}

cudaError_t launch_pairing_product2(const void* d_g1, const void* d_g2, uint64_t n, void* d_gt, void* d_is_one,
                                    cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  pairing_product2_kernel<<<(unsigned)((n + PTAU_PAIR_BLOCK - 1) / PTAU_PAIR_BLOCK), PTAU_PAIR_BLOCK, 0, stream>>>(
      (const uint32_t*)d_g1, (const uint32_t*)d_g2, n, (uint32_t*)d_gt, (uint8_t*)d_is_one);
  return cudaGetLastError();
}

size_t kzg_tables_bytes() { return PTAU_KZG_TBL_WORDS * 4; }

cudaError_t launch_kzg_tables(const void* d_vk_g1, const void* d_vk_g2, void* d_tbl, cudaStream_t stream) {
  kzg_tables_kernel<<<4, PTAU_FB_WINDOWS, 0, stream>>>((const uint32_t*)d_vk_g1, (const uint32_t*)d_vk_g2, (uint32_t*)d_tbl);
  kzg_tables8_kernel<<<dim3(PTAU_FB8_WINDOWS * 16 / 64, 3), 64, 0, stream>>>((uint32_t*)d_tbl);
  return cudaGetLastError();
}

cudaError_t launch_kzg_check(const void* d_vk_g1, const void* d_vk_g2, const void* d_comms, const void* d_points,
                             const void* d_values, const void* d_proofs, const void* d_random_v, const void* d_tbl,
                             uint64_t n, void* d_ok, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  kzg_check_kernel<<<(unsigned)((n + PTAU_PAIR_BLOCK - 1) / PTAU_PAIR_BLOCK), PTAU_PAIR_BLOCK, 0, stream>>>(
      (const uint32_t*)d_vk_g1, (const uint32_t*)d_vk_g2, (const uint32_t*)d_comms, (const uint32_t*)d_points,
      (const uint32_t*)d_values, (const uint32_t*)d_proofs, (const uint32_t*)d_random_v, (const uint32_t*)d_tbl, n,
      (uint8_t*)d_ok);
  return cudaGetLastError();
}

}  // namespace ptau
