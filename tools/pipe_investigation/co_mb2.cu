// What limits IMAD.WIDE: the carry, or operand (register-file) bandwidth?
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#define WIDE(acc, a, b) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b))
#define WIDE_CO(lo, hi, cnt, a, b) \
  asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo), "+r"(hi), "+r"(cnt) : "r"(a), "r"(b))

// KIND: 0 same a,b | 1 four a's, one b | 2 eight a's, two b's | 3 = 0 + carry-out | 4 = 1 + carry-out | 5 = 2 + carry-out
// 6: eight a's one b, 7: one a, 8 accumulators but b varies per MAD (8 b's)
template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, uint32_t seed) {
  unsigned long long acc[8];
  uint32_t lo[8], hi[8], cnt[8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    acc[i] = seed + i + threadIdx.x; lo[i] = seed + i + threadIdx.x; hi[i] = seed * 3 + i; cnt[i] = 0;
    a[i] = seed * 7 + i * 5 + threadIdx.x; b[i] = (seed * 11 + i * 3 + threadIdx.x) | 1u;
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] += (uint32_t)acc[i] ^ lo[i]; b[i] += (uint32_t)acc[i] ^ hi[i]; }
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (KIND == 0) WIDE(acc[i], a[i], b[u]);
        if (KIND == 1) WIDE_CO(lo[i], hi[i], cnt[i], a[i], b[u]);
        if (KIND == 2) WIDE(acc[i], a[(i + u) & 7], b[i]);
        if (KIND == 3) WIDE_CO(lo[i], hi[i], cnt[i], a[(i + u) & 7], b[i]);
        if (KIND == 4) asm volatile("{.reg .u32 l, h;\n\tmov.b64 {l, h}, %0;\n\tmul.wide.u32 %0, l, h;}" : "+l"(acc[i]));
        if (KIND == 5) asm volatile("add.u32 %0, %0, %2;\n\tadd.u32 %1, %1, %2;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(b[u]));
        if (KIND == 6) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(b[u]));
        if (KIND == 7) asm volatile("lop3.b32 %0, %0, %2, %3, 0x96;\n\tlop3.b32 %1, %1, %3, %2, 0x96;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a[i]), "r"(b[u]));
      }
    }
  }
  unsigned long long r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= acc[i] ^ lo[i] ^ ((unsigned long long)hi[i] << 32) ^ cnt[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)r ^ (uint32_t)(r >> 32);
}

template <class F>
static float timeit(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
}
template <int KIND> void run(uint32_t* dout, int sm, const char* name) {
  int it = 4000, grid = sm * 8;
  float ms = timeit([&] { k<KIND><<<grid, 256>>>(dout, it, 12345u); });
  printf("%-50s: %.3f ms  %.2f T wideMAD/s\n", name, ms, (double)grid * 256 * it * 64 / ms / 1e9);
}
int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  int sm = pr.multiProcessorCount;
  uint32_t* dout; CK(cudaMalloc(&dout, 64 << 20));
  run<0>(dout, sm, "plain, a[i], b[u] (row-constant b)");
  run<1>(dout, sm, "carry-out, a[i], b[u] (row-constant b)");
  run<2>(dout, sm, "plain, a and b vary per MAD");
  run<3>(dout, sm, "carry-out, a and b vary per MAD");
  run<4>(dout, sm, "mul.wide chain (IMAD.WIDE, RZ addend) [x1 op]");
  run<5>(dout, sm, "add.u32 pairs [x2 ALU ops per count]");
  run<6>(dout, sm, "add.cc/addc pairs [x2 ALU ops per count]");
  run<7>(dout, sm, "lop3 pairs [x2 ALU ops per count]");
  return 0;
}
