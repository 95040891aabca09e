// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tf_bench scripts/rng_microbench_threefry.cu   (run on the GPU box)
// microbenchmark: threefry2x32 x4 lockstep, rotate variants (pipe balance experiment)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t rotl_shf(uint32_t x, int r) { return __funnelshift_l(x, x, r); }
template <int MODE>
__device__ __forceinline__ void round4(uint32_t (&x0)[4], uint32_t (&x1)[4], int r, bool alt) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x0[j] += x1[j];
    if (MODE == 1 && alt) {
      // (x1 << r) on the FMA pipe (IMAD.SHL), (x1 >> (32 - r)) via IMAD.HI, one LOP3 for (a | b) ^ x0
      const uint32_t lo = x1[j] * (1u << r);
      const uint32_t hi = __umulhi(x1[j], 1u << r);
      x1[j] = (lo | hi) ^ x0[j];
    } else {
      x1[j] = rotl_shf(x1[j], r) ^ x0[j];
    }
  }
}
template <int MODE>
__global__ void bench(uint32_t* out, int iters, uint32_t k0, uint32_t k1) {
  uint32_t x0[4], x1[4];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = 0; j < 4; ++j) { x0[j] = t * 4 + j; x1[j] = t * 4 + j + 12345u; }
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  const int R[8] = {13, 15, 26, 6, 17, 29, 16, 24};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < 5; ++g) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = (g & 1) * 4 + q;
        round4<MODE>(x0, x1, R[idx], (g * 4 + q) % 3 == 1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { x0[j] += (g % 3 == 0 ? k1 : (g % 3 == 1 ? k2 : k0)); x1[j] += (g % 3 == 0 ? k2 : (g % 3 == 1 ? k0 : k1)) + g + 1; }
    }
  }
  uint32_t acc = 0;
  for (int j = 0; j < 4; ++j) acc ^= x0[j] ^ x1[j];
  out[t] = acc;
}
int main() {
  uint32_t* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 2000;
  uint32_t h[2][4];
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) bench<0><<<148 * 8, 256>>>(out, iters, 1u, 2u); else bench<1><<<148 * 8, 256>>>(out, iters, 1u, 2u);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("mode %d: %.3f ms  %.1f G blocks/s\n", mode, ms, 148.0 * 8 * 256 * 4 * iters / ms / 1e6);
    }
    cudaMemcpy(h[mode], out, 16, cudaMemcpyDeviceToHost);
  }
  printf("same results: %d\n", h[0][0] == h[1][0] && h[0][1] == h[1][1]);
  return 0;
}
