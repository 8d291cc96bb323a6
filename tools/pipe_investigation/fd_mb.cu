// Microbenchmark: FP64-pipe field multiplication vs the IMAD one, alone and co-resident.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "fd_proto.cuh"
using namespace ptau;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
  double m = 1.0000001, c = 1e-9;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      a0 = __fma_rz(a0, m, c); a1 = __fma_rz(a1, m, c); a2 = __fma_rz(a2, m, c); a3 = __fma_rz(a3, m, c);
      a4 = __fma_rz(a4, m, c); a5 = __fma_rz(a5, m, c); a6 = __fma_rz(a6, m, c); a7 = __fma_rz(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void __launch_bounds__(256) k_dadd(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
  double c = 1e-9;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      a0 = __dadd_rz(a0, c); a1 = __dadd_rz(a1, c); a2 = __dadd_rz(a2, c); a3 = __dadd_rz(a3, c);
      a4 = __dadd_rz(a4, c); a5 = __dadd_rz(a5, c); a6 = __dadd_rz(a6, c); a7 = __dadd_rz(a7, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) { return fq_mul_inl(a, b); }
__device__ __noinline__ Fd fd_mul_call(Fd a, Fd b) { return fd_mul(a, b); }

// mode 0: every warp runs the IMAD chain; 1: every warp the FP64 chain; 2: even warps IMAD, odd warps FP64
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_mix(uint32_t* out, const Fq* xs, int itI, int itD, int mode) {
  int wid = threadIdx.x >> 5;
  bool imad = mode == 0 ? true : mode == 1 ? false : ((wid & 1) == 0);
  Fq x = xs[threadIdx.x & 31], y = xs[32 + (threadIdx.x & 31)];
  uint32_t r = 0;
  if (imad) {
#pragma unroll 1
    for (int i = 0; i < itI; i++) { x = fq_mul_call(x, y); y = fq_mul_call(y, x); }
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= x.l[i] ^ y.l[i];
  } else {
    Fd a = fd_from_fq(x), b = fd_from_fq(y);
#pragma unroll 1
    for (int i = 0; i < itD; i++) { a = fd_mul_call(a, b); b = fd_mul_call(b, a); }
    Fq u = fd_to_fq(a), v = fd_to_fq(b);
#pragma unroll
    for (int i = 0; i < 12; i++) r ^= u.l[i] ^ v.l[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_check(const Fq* xs, const Fq* ys, int n, int chain, unsigned* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x = xs[i], y = ys[i];
  Fd a = fd_from_fq(x), b = fd_from_fq(y);
  for (int k = 0; k < chain; k++) {
    x = fq_mul_call(x, y); y = fq_mul_call(y, x);
    a = fd_mul_call(a, b); b = fd_mul_call(b, a);
  }
  Fq u = fd_to_fq(a), v = fd_to_fq(b);
  fq_cond_sub_p(u.l); fq_cond_sub_p(v.l);
  if (!fq_eq(u, x) || !fq_eq(v, y)) atomicAdd(bad, 1u);
}

template <class F>
static float timeit(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int sm = pr.multiProcessorCount; printf("%s SMs=%d\n", pr.name, sm);
  void* dout; CK(cudaMalloc(&dout, 64 << 20));
  // inputs
  const int N = 1 << 16;
  std::vector<Fq> hx(N), hy(N);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
  for (int i = 0; i < N; i++) for (int k = 0; k < 12; k++) { hx[i].l[k] = rnd(); hy[i].l[k] = rnd(); if (k == 11) { hx[i].l[k] %= 0x1a000000u; hy[i].l[k] %= 0x1a000000u; } }
  Fq *dx, *dy; CK(cudaMalloc(&dx, N * sizeof(Fq))); CK(cudaMalloc(&dy, N * sizeof(Fq)));
  CK(cudaMemcpy(dx, hx.data(), N * sizeof(Fq), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, hy.data(), N * sizeof(Fq), cudaMemcpyHostToDevice));
  unsigned* dbad; CK(cudaMalloc(&dbad, 4)); CK(cudaMemset(dbad, 0, 4));
  k_check<<<N / 128, 128>>>(dx, dy, N, 1, dbad);
  k_check<<<N / 128, 128>>>(dx, dy, N, 25, dbad);
  unsigned bad; CK(cudaMemcpy(&bad, dbad, 4, cudaMemcpyDeviceToHost)); CK(cudaDeviceSynchronize());
  printf("check: %u mismatches of %d\n", bad, 2 * N);

  {
    int it = 4000; int grid = sm * 8;
    float ms = timeit([&] { k_dfma<<<grid, 256>>>((double*)dout, it); });
    printf("DFMA.RZ : %.3f ms  %.2f T/s\n", ms, (double)grid * 256 * it * 64 / ms / 1e9);
    ms = timeit([&] { k_dadd<<<grid, 256>>>((double*)dout, it); });
    printf("DADD.RZ : %.3f ms  %.2f T/s\n", ms, (double)grid * 256 * it * 64 / ms / 1e9);
  }
  auto run = [&](int minb, int mode, int itI, int itD) {
    int grid = sm * minb;
    float ms;
    if (minb == 1) ms = timeit([&] { k_mix<1><<<grid, 128>>>((uint32_t*)dout, dx, itI, itD, mode); });
    else if (minb == 2) ms = timeit([&] { k_mix<2><<<grid, 128>>>((uint32_t*)dout, dx, itI, itD, mode); });
    else if (minb == 3) ms = timeit([&] { k_mix<3><<<grid, 128>>>((uint32_t*)dout, dx, itI, itD, mode); });
    else ms = timeit([&] { k_mix<4><<<grid, 128>>>((uint32_t*)dout, dx, itI, itD, mode); });
    double thr = (double)grid * 128;
    double nI = mode == 0 ? thr * itI * 2 : mode == 1 ? 0 : thr / 2 * itI * 2;
    double nD = mode == 1 ? thr * itD * 2 : mode == 0 ? 0 : thr / 2 * itD * 2;
    printf("blocks/SM=%d mode=%d itI=%d itD=%d : %.3f ms  imad %.2f Gmul/s  fp64 %.2f Gmul/s  total %.2f Gmul/s\n", minb, mode, itI, itD, ms,
           nI / ms / 1e6, nD / ms / 1e6, (nI + nD) / ms / 1e6);
  };
  for (int minb = 1; minb <= 4; minb++) {
    run(minb, 0, 2000, 0);
    run(minb, 1, 0, 2000);
    run(minb, 2, 2000, 0);     // only the even warps work (IMAD), odd idle
    run(minb, 2, 0, 2000);     // only odd warps work (FP64)
    run(minb, 2, 2000, 2000);
    run(minb, 2, 2000, 1500);
    run(minb, 2, 2000, 1000);
  }
  return 0;
}
