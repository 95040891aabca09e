// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/nz_bench scripts/rng_microbench_normals.cu   (run on the GPU box)
// microbenchmark: normals per second of the in-kernel generator (threefry2x32 x4 lockstep + XLA's erf_inv polynomial)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../fbs_b200/csrc/fbs_rng.cuh"
using namespace fbs;
template <int MODE>
__global__ void bench(float* out, int iters, uint32_t k0, uint32_t k1) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    const uint32_t b = (t * iters + it) * 4u;
    uint32_t x0[4] = {b, b + 1u, b + 2u, b + 3u};
    uint32_t x1[4] = {b + 77u, b + 78u, b + 79u, b + 80u};
    threefry2x32_x4(k0, k1, x0, x1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (MODE == 0) { acc += bits_to_normal(x0[c]) + bits_to_normal(x1[c]); }
      else { acc += bits_to_unit(x0[c]) + bits_to_unit(x1[c]); }
    }
  }
  out[t] = acc;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1000;
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) bench<0><<<148 * 8, 256>>>(out, iters, 1u, 2u); else bench<1><<<148 * 8, 256>>>(out, iters, 1u, 2u);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("%s: %.3f ms  %.1f G draws/s\n", mode == 0 ? "normals" : "uniforms", ms, 148.0 * 8 * 256 * 8 * iters / ms / 1e6);
    }
  return 0;
}
