// Non-GEMM pieces of the score network (fbs/nn/unet.py) and the closures around it (experiments/imgs/inpainting.py:
// 94-147): normalisations fused with their activations / residuals, the two attention flavours, the time-embedding
// MLP, the tiny first / last convolutions, the space-to-depth copy in front of the stride-2 convolutions, and the
// image assembly + Euler--Maruyama / log-weight step.  All activations are NHWC; "P" = H * W pixels.
#include <cuda_bf16.h>
#include <math.h>
#include "fbs_common.cuh"
#include "fbs_rng.cuh"

namespace fbs {
namespace nnops {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum (blockDim.x <= 1024, multiple of 32); red: shared float[32]
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = lane < nw ? red[lane] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = lane < nw ? red[lane] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}
__device__ __forceinline__ float swishf(float x) { return x / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------
// GroupNorm (+ time-embedding scale / shift) + swish (+ residual): unet.py:144-155,159-160,172.
// One CTA per (sample, group); the slice (P x C / groups values, L2 / L1 resident) is read three times.
// ---------------------------------------------------------------------------------------------------------
template <bool STAGED>
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, int P, int C, int groups,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ tss, const float* __restrict__ residual,
                                                       float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, float eps) {
  extern __shared__ float4 slice[];  // STAGED: the (sample, group) slice, read from global memory exactly once
  __shared__ float red[32];
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups, q4 = cpg / 4;  // float4 per pixel of this group
  const size_t base = (size_t)b * P * C + (size_t)g * cpg;
  const int n4 = P * q4;
  float s = 0.f;
  for (int e = threadIdx.x; e < n4; e += blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(x + base + (size_t)(e / q4) * C + 4 * (e % q4));
    if (STAGED) slice[e] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = block_sum(s, red) / (float)(P * cpg);
  float ss = 0.f;
  for (int e = threadIdx.x; e < n4; e += blockDim.x) {
    const float4 v = STAGED ? slice[e] : *reinterpret_cast<const float4*>(x + base + (size_t)(e / q4) * C + 4 * (e % q4));
    const float a = v.x - mean, bq = v.y - mean, c = v.z - mean, d = v.w - mean;
    ss += (a * a + bq * bq) + (c * c + d * d);
  }
  const float rstd = rsqrtf(block_sum(ss, red) / (float)(P * cpg) + eps);
  for (int e = threadIdx.x; e < n4; e += blockDim.x) {
    const int c0 = g * cpg + 4 * (e % q4);
    const size_t off = (size_t)b * P * C + (size_t)(e / q4) * C + c0;
    const float4 v = STAGED ? slice[e] : *reinterpret_cast<const float4*>(x + off);
    float y[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = (y[j] - mean) * rstd * __ldg(gamma + c0 + j) + __ldg(beta + c0 + j);
      if (tss) t = t * (1.0f + __ldg(tss + c0 + j)) + __ldg(tss + C + c0 + j);  // h * (1 + scale) + shift
      y[j] = swishf(t);
    }
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + off);
      y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_bf16) {
      __align__(8) __nv_bfloat162 o[2] = {__floats2bfloat162_rn(y[0], y[1]), __floats2bfloat162_rn(y[2], y[3])};
      *reinterpret_cast<uint2*>(out_bf16 + off) = *reinterpret_cast<const uint2*>(o);
    }
  }
}

// Same operation, one CTA of 1024 threads per SAMPLE (all groups): a (sample, group) slice is 32 bytes out of every 256-byte
// pixel at 64 channels / 8 groups, so the per-group kernel above spends its time on 32-byte segments; here a warp reads and
// writes whole pixels (512 contiguous bytes per instruction).  1024 % (C / 4) == 0 makes a thread's channel quad -- hence
// its group and its affine constants -- fixed; group statistics are reduced by segmented shuffles, then over the warps in a
// fixed order (deterministic).  Used for large samples when there are enough of them to fill the machine.
template <bool STAGED, int NT>
__global__ void __launch_bounds__(NT) gn_sample_kernel(const float* __restrict__ x, int P, int C, int groups,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ tss, const float* __restrict__ residual,
                                                       float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, float eps) {
  extern __shared__ float4 slice[];  // STAGED: the sample, read from global memory exactly once
  constexpr int NW = NT / 32;
  __shared__ float part[NW][17];
  __shared__ float tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q4 = C >> 2, cpg = C / groups, L = cpg >> 2;  // float4 per pixel / per group inside a pixel (powers of two)
  const int quad = tid % q4, g = quad / L;
  const int n4 = P * q4;
  const size_t base4 = (size_t)blockIdx.x * n4;
  const float4* xb = reinterpret_cast<const float4*>(x) + base4;
  const int Lw = L < 32 ? L : 32, qw = q4 < 32 ? q4 : 32;
  const int GW = qw / Lw;                                       // groups one warp holds partial sums of
  const int gf_lane = lane < NW ? ((lane * 32) % q4) / L : 0;   // first group warp `lane` holds
  auto group_total = [&](float v) -> float {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
      if (o < L || o >= q4) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane < qw && (lane % Lw) == 0) part[warp][lane / Lw] = v;
    __syncthreads();
    // warp j adds up group j (j, j + NW, ...) over the warps that hold it: lane = source warp, fixed xor tree
    for (int j = warp; j < groups; j += NW) {
      float t = (lane < NW && j >= gf_lane && j < gf_lane + GW) ? part[lane][j - gf_lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) tot[j] = t;
    }
    __syncthreads();
    return tot[g];
  };
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float s = 0.f;
  for (int e0 = tid; e0 < n4; e0 += 4 * NT) {  // four independent loads in flight per thread
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = e0 + u * NT < n4 ? xb[e0 + u * NT] : zero4;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (STAGED && e0 + u * NT < n4) slice[e0 + u * NT] = v[u];
      s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
  }
  const float inv_n = 1.0f / (float)(P * cpg);
  const float mean = group_total(s) * inv_n;
  float ss = 0.f;
  for (int e0 = tid; e0 < n4; e0 += 4 * NT) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = e0 + u * NT < n4 ? (STAGED ? slice[e0 + u * NT] : xb[e0 + u * NT]) : make_float4(mean, mean, mean, mean);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float a = v[u].x - mean, bq = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
      ss += (a * a + bq * bq) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(group_total(ss) * inv_n + eps);
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + quad), bt = __ldg(reinterpret_cast<const float4*>(beta) + quad);
  float4 sc = zero4, sh = zero4;
  if (tss) {
    sc = __ldg(reinterpret_cast<const float4*>(tss) + quad);
    sh = __ldg(reinterpret_cast<const float4*>(tss + C) + quad);
  }
  const float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, btv[4] = {bt.x, bt.y, bt.z, bt.w};
  const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
  for (int e0 = tid; e0 < n4; e0 += 4 * NT) {
    float4 v[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * NT;
      r[u] = (residual && e < n4) ? reinterpret_cast<const float4*>(residual)[base4 + e] : zero4;
      v[u] = e < n4 ? (STAGED ? slice[e] : xb[e]) : zero4;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * NT;
      if (e >= n4) continue;
      float y[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      const float rr[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = (y[j] - mean) * rstd * gmv[j] + btv[j];
        if (tss) t = t * (1.0f + scv[j]) + shv[j];  // h * (1 + scale) + shift
        y[j] = swishf(t) + rr[j];
      }
      if (out_f32) reinterpret_cast<float4*>(out_f32)[base4 + e] = make_float4(y[0], y[1], y[2], y[3]);
      if (out_bf16) {
        __align__(8) __nv_bfloat162 o[2] = {__floats2bfloat162_rn(y[0], y[1]), __floats2bfloat162_rn(y[2], y[3])};
        reinterpret_cast<uint2*>(out_bf16)[base4 + e] = *reinterpret_cast<const uint2*>(o);
      }
    }
  }
}

// GroupNorm + swish with the statistics handed over by the convolution that produced x (nn_conv.cu, `gn_partials`:
// [B][slots][C / 4] float2 = sum, sum of squares per (sample, slot, channel quad)): one streaming pass over the activations,
// CTA = (sample, pixel chunk), a thread keeps one channel quad (blockDim % (C / 4) == 0).  var = E[x^2] - E[x]^2 is flax's
// own formula (flax.linen.normalization._compute_stats, use_fast_variance = True).
// LN: the attention block that follows normalises the block output once more over the channels (LayerNorm, scale only,
// unet.py:258); a pixel's C / 4 channel quads sit in C / 4 <= 32 adjacent lanes, so that second normalisation is two shuffle
// reductions on values already in registers and one more bf16 store -- it saves a launch and a pass over the fp32 tensor.
template <bool IN16, bool LN>
__global__ void __launch_bounds__(256) gn_stats_apply_kernel(const float* __restrict__ x32, const __nv_bfloat16* __restrict__ x16,
                                                             const float2* __restrict__ partials, int slots, int P, int C, int groups,
                                                             int pchunks, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ tss,
                                                             const float* __restrict__ residual, float* __restrict__ out_f32,
                                                             __nv_bfloat16* __restrict__ out_bf16, float eps,
                                                             const float* __restrict__ ln_gamma, float ln_eps,
                                                             __nv_bfloat16* __restrict__ ln_out) {
  __shared__ float s_mean[32], s_rstd[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / pchunks, pc = blockIdx.x % pchunks;
  const int q4 = C >> 2, cpg = C / groups, L = cpg >> 2;
  // group statistics: warp w adds up groups w, w + 8, ... (lanes stride over the (slot, quad) partials; fixed xor tree)
  const float2* pb = partials + (size_t)b * slots * q4;
  for (int g = warp; g < groups; g += 8) {
    float s = 0.f, ss = 0.f;
    for (int e = lane; e < slots * L; e += 32) {
      const float2 t = __ldg(pb + (size_t)(e / L) * q4 + g * L + e % L);
      s += t.x;
      ss += t.y;
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) {
      const float inv_n = 1.0f / (float)(P * cpg);
      const float mean = s * inv_n;
      s_mean[g] = mean;
      s_rstd[g] = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + eps);
    }
  }
  __syncthreads();
  const int quad = tid % q4, g = quad / L;
  const float mean = s_mean[g], rstd = s_rstd[g];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + quad), bt = __ldg(reinterpret_cast<const float4*>(beta) + quad);
  float4 sc = zero4, sh = zero4;
  if (tss) {
    sc = __ldg(reinterpret_cast<const float4*>(tss) + quad);
    sh = __ldg(reinterpret_cast<const float4*>(tss + C) + quad);
  }
  // y = swish(x * A + Bc): the affine parts folded per channel
  float A[4], Bc[4];
  {
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w}, btv[4] = {bt.x, bt.y, bt.z, bt.w};
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a0 = rstd * gmv[j], b0 = btv[j] - mean * a0;
      A[j] = a0 * (1.0f + scv[j]);
      Bc[j] = b0 * (1.0f + scv[j]) + shv[j];
    }
  }
  const int n4 = P * q4;
  const int per = ((P + pchunks - 1) / pchunks) * q4;  // whole pixels per chunk
  const int e_begin = pc * per, e_end = min(n4, e_begin + per);
  const size_t base4 = (size_t)b * n4;
  float4 lg = zero4;
  if (LN) lg = __ldg(reinterpret_cast<const float4*>(ln_gamma) + quad);
  for (int eb = e_begin; eb < e_end; eb += 4 * 256) {  // (the same trip count for every thread: the LN shuffles are warp-wide)
    const int e0 = eb + tid;
    float4 v[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      const bool ok = e < e_end;
      if (IN16) {
        uint2 raw = make_uint2(0u, 0u);
        if (ok) raw = reinterpret_cast<const uint2*>(x16)[base4 + e];
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
        const float2 fl = __bfloat1622float2(lo), fh = __bfloat1622float2(hi);
        v[u] = make_float4(fl.x, fl.y, fh.x, fh.y);
      } else {
        v[u] = ok ? reinterpret_cast<const float4*>(x32)[base4 + e] : zero4;
      }
      r[u] = (residual && ok) ? reinterpret_cast<const float4*>(residual)[base4 + e] : zero4;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      const bool ok = e < e_end;
      if (!LN && !ok) continue;
      float4 y;
      y.x = swishf(fmaf(v[u].x, A[0], Bc[0])) + r[u].x;
      y.y = swishf(fmaf(v[u].y, A[1], Bc[1])) + r[u].y;
      y.z = swishf(fmaf(v[u].z, A[2], Bc[2])) + r[u].z;
      y.w = swishf(fmaf(v[u].w, A[3], Bc[3])) + r[u].w;
      if (ok) {
        if (out_f32) reinterpret_cast<float4*>(out_f32)[base4 + e] = y;
        if (out_bf16) {
          __align__(8) __nv_bfloat162 o[2] = {__floats2bfloat162_rn(y.x, y.y), __floats2bfloat162_rn(y.z, y.w)};
          reinterpret_cast<uint2*>(out_bf16)[base4 + e] = *reinterpret_cast<const uint2*>(o);
        }
      }
      if (LN) {
        // the q4 lanes of this pixel: mean, then the variance about it (the stand-alone LayerNorm's two passes)
        float sm = (y.x + y.y) + (y.z + y.w);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
          if (o < q4) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        const float mu = sm / (float)C;
        const float a = y.x - mu, bq = y.y - mu, c = y.z - mu, d = y.w - mu;
        float sq = (a * a + bq * bq) + (c * c + d * d);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
          if (o < q4) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rs = rsqrtf(sq / (float)C + ln_eps);
        if (ok) {
          __align__(8) __nv_bfloat162 o[2] = {__floats2bfloat162_rn(a * rs * lg.x, bq * rs * lg.y),
                                              __floats2bfloat162_rn(c * rs * lg.z, d * rs * lg.w)};
          reinterpret_cast<uint2*>(ln_out)[base4 + e] = *reinterpret_cast<const uint2*>(o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over channels, scale only (+ residual): unet.py:243,258,264.  One warp per pixel.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t R, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ residual,
                                                        float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const float* xr = x + row * C;
  float v[16];  // C <= 512
  const int per = C / 32;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < per) {
      v[j] = xr[lane + 32 * j];
      s += v[j];
    }
  const float mean = warp_sum(s) / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < per) {
      const float d = v[j] - mean;
      ss += d * d;
    }
  const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < per) {
      const int c = lane + 32 * j;
      float y = (v[j] - mean) * rstd * __ldg(gamma + c);
      if (residual) y += residual[row * C + c];
      if (out_f32) out_f32[row * C + c] = y;
      if (out_bf16) out_bf16[row * C + c] = __float2bfloat16_rn(y);
    }
}

// Same operation with 16-byte accesses: min(32, C / 4) lanes per pixel, C in {64, 128, 256, 512}.
__global__ void __launch_bounds__(256) layernorm4_kernel(const float* __restrict__ x, int64_t R, int C,
                                                         const float* __restrict__ gamma, const float* __restrict__ residual,
                                                         float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q4 = C >> 2, LPR = q4 < 32 ? q4 : 32, per = q4 / LPR, rpw = 32 / LPR;
  const int64_t row = ((int64_t)blockIdx.x * 8 + warp) * rpw + lane / LPR;
  const int l = lane % LPR;
  const bool active = row < R;
  const float4* xr = reinterpret_cast<const float4*>(x) + row * q4;
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < per) {
      v[j] = active ? xr[l + LPR * j] : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    if (o < LPR) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < per) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    if (o < LPR) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)C + eps);
  if (!active) return;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < per) {
      const int c4 = l + LPR * j;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      float4 y = make_float4((v[j].x - mean) * rstd * gm.x, (v[j].y - mean) * rstd * gm.y, (v[j].z - mean) * rstd * gm.z,
                             (v[j].w - mean) * rstd * gm.w);
      const size_t off = (size_t)row * q4 + c4;
      if (residual) {
        const float4 r = reinterpret_cast<const float4*>(residual)[off];
        y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
      }
      if (out_f32) reinterpret_cast<float4*>(out_f32)[off] = y;
      if (out_bf16) {
        __align__(8) __nv_bfloat162 o[2] = {__floats2bfloat162_rn(y.x, y.y), __floats2bfloat162_rn(y.z, y.w)};
        reinterpret_cast<uint2*>(out_bf16)[off] = *reinterpret_cast<const uint2*>(o);
      }
    }
}

// ---------------------------------------------------------------------------------------------------------
// LinearAttention core (unet.py:227-239): q softmax over the head dimension, k softmax over the pixels,
// context = k^T (v / P), out = context^T (q / sqrt(d)).  One CTA per (sample, head); dim_head = 32.
// qkv: bf16 [B, P, 3 * heads * 32] (q | k | v, each (head, d));  out: bf16 [B, P, heads * 32].
//
// Both contractions (32 x P x 32 and P x 32 x 32 per head: tiny, but 1.6 MFLOP per CTA of scalar FMAs was the whole run
// time) run on the tensor cores with warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate) -- far too small for a
// tcgen05 tile.  Global traffic is 16-byte loads; tiles of 128 pixels are staged in shared memory as bf16 rows padded to
// 80 bytes (conflict-free ldmatrix); exp(k - max) and the q softmax are applied while staging.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) linear_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int P, int heads,
                                                               __nv_bfloat16* __restrict__ out) {
  constexpr int TN = 128, LDS = 40;  // pixels per tile; shared-memory row stride in bf16 (32 + 8 pad)
  __shared__ __align__(16) unsigned char raw[2 * TN * LDS * 2];
  __nv_bfloat16 (*at)[LDS] = reinterpret_cast<__nv_bfloat16 (*)[LDS]>(raw);                 // exp(k - max) tile, then the softmaxed q tile
  __nv_bfloat16 (*vt)[LDS] = reinterpret_cast<__nv_bfloat16 (*)[LDS]>(raw + TN * LDS * 2);  // v tile
  float (*pb)[32][33] = reinterpret_cast<float (*)[32][33]>(raw);                           // 4 partial contexts (after the tiles)
  __shared__ __align__(16) __nv_bfloat16 cb[32][LDS];  // context [d][e], bf16
  __shared__ float red[64][33];
  __shared__ float kmax[32], ksum[32];
  __shared__ float ctx[32][33];
  const int b = blockIdx.x / heads, hd = blockIdx.x % heads;
  const int HD = heads * 32, ld = 3 * HD;
  const __nv_bfloat16* q = qkv + (size_t)b * P * ld + hd * 32;
  const __nv_bfloat16* k = q + HD;
  const __nv_bfloat16* v = k + HD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cgp = tid & 3, prow = tid >> 2;  // this thread stages channels [8 cgp, 8 cgp + 8) of pixels prow and prow + 64 of a tile
  auto load8 = [&](const __nv_bfloat16* base, int n, float (&f)[8]) {
    const uint4 raw = *reinterpret_cast<const uint4*>(base + (size_t)n * ld + 8 * cgp);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __bfloat1622float2(h2[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
  };
  auto store8 = [&](__nv_bfloat16 (*tile)[LDS], int row, const float (&f)[8]) {
    __align__(16) __nv_bfloat162 h2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    *reinterpret_cast<uint4*>(&tile[row][8 * cgp]) = *reinterpret_cast<const uint4*>(h2);
  };
  // ---- column max of k over the pixels
  {
    float mx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mx[j] = -INFINITY;
    for (int n = prow; n < P; n += 64) {
      float f[8];
      load8(k, n, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) mx[j] = fmaxf(mx[j], f[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[prow][8 * cgp + j] = mx[j];
    __syncthreads();
    if (tid < 32) {
      float t = red[0][tid];
      for (int r = 1; r < 64; ++r) t = fmaxf(t, red[r][tid]);
      kmax[tid] = t;
    }
    __syncthreads();
  }
  // ---- context[d][e] = sum_n exp(k[n,d] - kmax[d]) v[n,e]: A = (exp k)^T through ldmatrix.trans, B = v through ldmatrix.trans;
  //      warp w takes pixels [16 w, 16 w + 16) of every tile (one k16 step) and keeps a full 32 x 32 partial in registers
  float km[8], ks[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    km[j] = kmax[8 * cgp + j];
    ks[j] = 0.f;
  }
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
  for (int n0 = 0; n0 < P; n0 += TN) {
    __syncthreads();
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int row = prow + 64 * hlf, n = n0 + row;
      float fk[8], fv[8];
      if (n < P) {
        load8(k, n, fk);
        load8(v, n, fv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          fk[j] = __expf(fk[j] - km[j]);
          ks[j] += fk[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) fk[j] = fv[j] = 0.f;
      }
      store8(at, row, fk);
      store8(vt, row, fv);
    }
    __syncthreads();
    const int r0 = 16 * warp;
    uint32_t a[2][4], bv[2][4];
    // A (m = d, k = pixel) from at[pixel][d]: matrices (pix 0-7, d 0-7), (pix 0-7, d 8-15), (pix 8-15, d 0-7), (pix 8-15, d 8-15)
    const int arow = r0 + (lane & 7) + 8 * (lane >> 4), acol = 8 * ((lane >> 3) & 1);
    ldmatrix_x4_trans(a[0], &at[arow][acol]);
    ldmatrix_x4_trans(a[1], &at[arow][acol + 16]);
    // B (k = pixel, n = e) from vt[pixel][e]: matrices (pix 0-7, e0), (pix 8-15, e0), (pix 0-7, e0 + 8), (pix 8-15, e0 + 8)
    const int brow = r0 + (lane & 7) + 8 * ((lane >> 3) & 1), bcol = 8 * (lane >> 4);
    ldmatrix_x4_trans(bv[0], &vt[brow][bcol]);
    ldmatrix_x4_trans(bv[1], &vt[brow][bcol + 16]);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a[mt], bv[nt >> 1][2 * (nt & 1)], bv[nt >> 1][2 * (nt & 1) + 1]);
  }
  // ---- reduce the partial contexts of the 8 warps (fixed order: the result must not depend on scheduling) and the
  //      softmax denominators
  __syncthreads();  // the tiles are dead: their storage holds the partials now
#pragma unroll
  for (int j = 0; j < 8; ++j) red[prow][8 * cgp + j] = ks[j];
  {
    const int g = lane >> 2, tg = lane & 3;
    for (int round = 0; round < 2; ++round) {
      if ((warp >> 2) == round) {
        float (*dstp)[33] = pb[warp & 3];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float* r0p = &dstp[16 * mt + g][8 * nt + 2 * tg];
            float* r1p = &dstp[16 * mt + g + 8][8 * nt + 2 * tg];
            if (round == 0) {
              r0p[0] = acc[mt][nt][0]; r0p[1] = acc[mt][nt][1];
              r1p[0] = acc[mt][nt][2]; r1p[1] = acc[mt][nt][3];
            } else {
              r0p[0] += acc[mt][nt][0]; r0p[1] += acc[mt][nt][1];
              r1p[0] += acc[mt][nt][2]; r1p[1] += acc[mt][nt][3];
            }
          }
      }
      __syncthreads();
    }
  }
  if (tid < 32) {
    float t = 0.f;
    for (int r = 0; r < 64; ++r) t += red[r][tid];
    ksum[tid] = t;
  }
  for (int t = tid; t < 32 * 32; t += 256) {
    const int d = t >> 5, e = t & 31;
    ctx[d][e] = (pb[0][d][e] + pb[1][d][e]) + (pb[2][d][e] + pb[3][d][e]);
  }
  __syncthreads();
  for (int t = tid; t < 32 * 32; t += 256) {
    const int d = t >> 5, e = t & 31;
    cb[d][e] = __float2bfloat16_rn(ctx[d][e] / (ksum[d] * (float)P));  // softmax denominator and v / (H W)
  }
  __syncthreads();
  // ---- out[n][e] = sum_d softmax_d(q[n,:])[d] / sqrt(32) ctx[d][e]: A = q tile (row major), B = context through ldmatrix.trans
  const float inv_sqrt_d = 0.17677669529663687f;
  uint32_t bc[2][2][4];  // [k step (d 0-15 / 16-31)][e half][matrices (d lo, e0), (d hi, e0), (d lo, e0 + 8), (d hi, e0 + 8)]
  {
    const int brow = (lane & 7) + 8 * ((lane >> 3) & 1), bcol = 8 * (lane >> 4);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      ldmatrix_x4_trans(bc[kk][0], &cb[16 * kk + brow][bcol]);
      ldmatrix_x4_trans(bc[kk][1], &cb[16 * kk + brow][bcol + 16]);
    }
  }
  for (int n0 = 0; n0 < P; n0 += TN) {
    __syncthreads();
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int row = prow + 64 * hlf, n = n0 + row;
      float f[8];
      if (n < P) {
        load8(q, n, f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      float m = f[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) m = fmaxf(m, f[j]);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));  // the 4 threads of a pixel are adjacent lanes
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = __expf(f[j] - m);
        sum += f[j];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float sc = inv_sqrt_d / sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= sc;
      store8(at, row, f);
    }
    __syncthreads();
    const int r0 = 16 * warp;
    uint32_t a[2][4];
    // A (m = pixel, k = d), row major: matrices (pix 0-7, d 0-7), (pix 8-15, d 0-7), (pix 0-7, d 8-15), (pix 8-15, d 8-15)
    const int arow = r0 + (lane & 7) + 8 * ((lane >> 3) & 1), acol = 8 * (lane >> 4);
    ldmatrix_x4(a[0], &at[arow][acol]);
    ldmatrix_x4(a[1], &at[arow][acol + 16]);
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) o[nt][c] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) mma_bf16_16816(o[nt], a[kk], bc[kk][nt >> 1][2 * (nt & 1)], bc[kk][nt >> 1][2 * (nt & 1) + 1]);
    }
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int n = n0 + r0 + g + 8 * hh;
      if (n < P) {
        __nv_bfloat16* dst = out + ((size_t)b * P + n) * HD + hd * 32 + 2 * tg;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<__nv_bfloat162*>(dst + 8 * nt) = __floats2bfloat162_rn(o[nt][2 * hh], o[nt][2 * hh + 1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Attention core of the middle block (unet.py:192-199): q, k l2-normalised ALONG THE TOKEN AXIS (axis=1 of
// 'b (x y) h d', as written upstream), sim = 10 q k^T, softmax over keys, out = attn v.  One CTA per
// (sample, head, 128-query block); keys / values streamed through shared memory in blocks of 128, online softmax.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attention_kernel(const __nv_bfloat16* __restrict__ qkv, int P, int heads, float scale,
                                                        __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) float ks[128][36], vs[128][36];  // rows of 32 (+ 4: 16-byte aligned, conflict-free broadcasts)
  __shared__ float qn[32], kn[32];
  __shared__ float part[4][32];
  const int qblocks = (P + (int)blockDim.x - 1) / (int)blockDim.x;
  const int b = blockIdx.x / (heads * qblocks), rem = blockIdx.x % (heads * qblocks);
  const int hd = rem / qblocks, qb = rem % qblocks;
  const int HD = heads * 32, ld = 3 * HD;
  const __nv_bfloat16* q = qkv + (size_t)b * P * ld + hd * 32;
  const __nv_bfloat16* k = q + HD;
  const __nv_bfloat16* v = k + HD;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5, nthr = blockDim.x;  // 64 or 128 threads
  // column norms over the tokens
  {
    float sq = 0.f, sk = 0.f;
    for (int n = w; n < P; n += nw) {
      const float a = __bfloat162float(q[(size_t)n * ld + lane]), c = __bfloat162float(k[(size_t)n * ld + lane]);
      sq = fmaf(a, a, sq);
      sk = fmaf(c, c, sk);
    }
    part[w][lane] = sq;
    if (w + 2 < 4 && nw == 2) part[w + 2][lane] = 0.f;
    __syncthreads();
    if (w == 0) qn[lane] = 1.0f / fmaxf(sqrtf(part[0][lane] + part[1][lane] + part[2][lane] + part[3][lane]), 1e-12f);
    __syncthreads();
    part[w][lane] = sk;
    __syncthreads();
    if (w == 0) kn[lane] = 1.0f / fmaxf(sqrtf(part[0][lane] + part[1][lane] + part[2][lane] + part[3][lane]), 1e-12f);
    __syncthreads();
  }
  const int i = qb * nthr + threadIdx.x;  // this thread's query
  float qi[32], acc[32];
  auto load8 = [&](const __nv_bfloat16* src, float (&f)[8]) {  // 16-byte load of 8 bf16
    const uint4 raw = *reinterpret_cast<const uint4*>(src);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __bfloat1622float2(h2[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
  };
#pragma unroll
  for (int d8 = 0; d8 < 4; ++d8) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (i < P) load8(q + (size_t)i * ld + 8 * d8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      qi[8 * d8 + j] = f[j] * qn[8 * d8 + j] * scale;
      acc[8 * d8 + j] = 0.f;
    }
  }
  float mx = -INFINITY, den = 0.f;
  for (int j0 = 0; j0 < P; j0 += 128) {
    const int jn = min(128, P - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < jn * 4; t += nthr) {  // only the rows that exist; 8 channels per thread and load
      const int j = t >> 2, d8 = t & 3;
      float fk[8], fv[8];
      load8(k + (size_t)(j0 + j) * ld + 8 * d8, fk);
      load8(v + (size_t)(j0 + j) * ld + 8 * d8, fv);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        ks[j][8 * d8 + c] = fk[c] * kn[8 * d8 + c];
        vs[j][8 * d8 + c] = fv[c];
      }
    }
    __syncthreads();
    for (int j = 0; j < jn; ++j) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four partial sums: the 32-deep dependent chain was the key loop's latency
#pragma unroll
      for (int dd = 0; dd < 32; dd += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&ks[j][dd]);
        s0 = fmaf(qi[dd], kk.x, s0);
        s1 = fmaf(qi[dd + 1], kk.y, s1);
        s2 = fmaf(qi[dd + 2], kk.z, s2);
        s3 = fmaf(qi[dd + 3], kk.w, s3);
      }
      const float sdot = (s0 + s1) + (s2 + s3);
      const float nm = fmaxf(mx, sdot);
      const float corr = __expf(mx - nm), pj = __expf(sdot - nm);
      den = den * corr + pj;
#pragma unroll
      for (int dd = 0; dd < 32; dd += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&vs[j][dd]);
        acc[dd] = fmaf(acc[dd], corr, pj * vv.x);
        acc[dd + 1] = fmaf(acc[dd + 1], corr, pj * vv.y);
        acc[dd + 2] = fmaf(acc[dd + 2], corr, pj * vv.z);
        acc[dd + 3] = fmaf(acc[dd + 3], corr, pj * vv.w);
      }
      mx = nm;
    }
  }
  if (i < P) {
    const float inv = 1.0f / den;
#pragma unroll
    for (int dd = 0; dd < 32; ++dd) out[((size_t)b * P + i) * HD + hd * 32 + dd] = __float2bfloat16_rn(acc[dd] * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Time embedding (unet.py:293-300, base.py:44-77) and every ResnetBlock's Dense(2 dim)(swish(time_emb))
// (:148-149) in one launch: table[j] = sum_i swish(temb)[i] Wcat[i][j] + bcat[j].  Each CTA recomputes the
// shared 64 -> 4 dim -> 4 dim trunk (dim = 64) and produces 64 outputs.  `tval` holds the network time.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) time_mlp_kernel(const float* __restrict__ tval, float inv_dt, int dim,
                                                       const float* __restrict__ W0, const float* __restrict__ b0,
                                                       const float* __restrict__ W1, const float* __restrict__ b1,
                                                       const float* __restrict__ Wcat, const float* __restrict__ bcat, int nout,
                                                       float* __restrict__ table) {
  __shared__ float emb[256], h1[1024], h2[1024];
  const int D4 = 4 * dim, half = dim / 2;
  const float t = tval[0] * inv_dt;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    const int j = i < half ? i : i - half;
    const float f = expf(-logf(10000.0f) * (float)j / (float)(half - 1));
    emb[i] = i < half ? sinf(t * f) : cosf(t * f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < D4; o += blockDim.x) {
    float a = b0[o];
    for (int i = 0; i < dim; ++i) a = fmaf(emb[i], W0[(size_t)i * D4 + o], a);
    // flax nn.gelu: tanh approximation
    h1[o] = 0.5f * a * (1.0f + tanhf(0.7978845608028654f * (a + 0.044715f * a * a * a)));
  }
  __syncthreads();
  for (int o = threadIdx.x; o < D4; o += blockDim.x) {
    float a = b1[o];
    for (int i = 0; i < D4; ++i) a = fmaf(h1[i], W1[(size_t)i * D4 + o], a);
    h2[o] = a / (1.0f + expf(-a));  // swish(time_emb), the input of every block's time MLP
  }
  __syncthreads();
  // 64 outputs per CTA, 4 threads per output
  const int o = blockIdx.x * 64 + (threadIdx.x >> 2), part = threadIdx.x & 3;
  float a = 0.f;
  if (o < nout)
    for (int i = part; i < D4; i += 4) a = fmaf(h2[i], Wcat[(size_t)i * nout + o], a);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  if (o < nout && part == 0) table[o] = a + bcat[o];
}

// ---------------------------------------------------------------------------------------------------------
// First convolution: 7x7, padding 3, Cin in {1..4} -> 64 (unet.py:286-291).  CUDA cores: K = 49 Cin is tiny.
// thread = (pixel, 4 output channels); weights [7][7][Cin][Cout] in shared memory.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ x, int B, int H, int W, int Cin, int Cout,
                                                        const float* __restrict__ wgt, const float* __restrict__ bias,
                                                        float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  extern __shared__ float ws[];
  const int nw = 49 * Cin * Cout;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) ws[i] = wgt[i];
  __syncthreads();
  const int cgs = Cout / 16;  // thread = (pixel, 16 output channels)
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * W * cgs) return;
  const int cg = (int)(idx % cgs);
  const int64_t pix = idx / cgs;
  const int w = (int)(pix % W), h = (int)((pix / W) % H);
  const int64_t b = pix / ((int64_t)W * H);
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = bias[16 * cg + j];
  for (int ty = 0; ty < 7; ++ty) {
    const int hh = h + ty - 3;
    if (hh < 0 || hh >= H) continue;
    for (int tx = 0; tx < 7; ++tx) {
      const int ww = w + tx - 3;
      if (ww < 0 || ww >= W) continue;
      const float* xp = x + ((b * H + hh) * W + ww) * Cin;
      const float* wp = ws + ((ty * 7 + tx) * Cin) * Cout + 16 * cg;
      for (int c = 0; c < Cin; ++c) {
        const float xv = xp[c];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(wp + c * Cout + j);
          acc[j] = fmaf(xv, wv.x, acc[j]); acc[j + 1] = fmaf(xv, wv.y, acc[j + 1]);
          acc[j + 2] = fmaf(xv, wv.z, acc[j + 2]); acc[j + 3] = fmaf(xv, wv.w, acc[j + 3]);
        }
      }
    }
  }
  const size_t off = (size_t)pix * Cout + 16 * cg;
  if (out_f32) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(out_f32 + off + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  if (out_bf16) {
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      __align__(16) __nv_bfloat162 o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t] = __floats2bfloat162_rn(acc[j + 2 * t], acc[j + 2 * t + 1]);
      *reinterpret_cast<uint4*>(out_bf16 + off + j) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// Same convolution, thread = (4 horizontally adjacent pixels, 16 output channels): a tap's 16 weights (four 16-byte shared
// loads) serve 64 FMAs instead of 16 -- the kernel above is bound by its shared-memory loads.  W % 4 == 0.
template <int CPT>  // output channels per thread: 16, or 8 (twice the threads, half the accumulators: three CTAs per SM)
__global__ void __launch_bounds__(256, CPT == 8 ? 3 : 2) stem_conv4_kernel(const float* __restrict__ x, int B, int H, int W, int Cin, int Cout,
                                                         const float* __restrict__ wgt, const float* __restrict__ bias,
                                                         float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16) {
  extern __shared__ float ws[];
  const int nw = 49 * Cin * Cout;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) ws[i] = wgt[i];
  __syncthreads();
  const int cgs = Cout / CPT, wqs = W / 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * wqs * cgs) return;
  const int cg = (int)(idx % cgs);
  int64_t t = idx / cgs;
  const int w0 = 4 * (int)(t % wqs);
  t /= wqs;
  const int h = (int)(t % H);
  const int64_t b = t / H;
  float acc[4][CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const float bj = bias[CPT * cg + j];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][j] = bj;
  }
  for (int ty = 0; ty < 7; ++ty) {
    const int hh = h + ty - 3;
    if (hh < 0 || hh >= H) continue;
    for (int c = 0; c < Cin; ++c) {
      float xv[10];  // columns w0 - 3 .. w0 + 6 of this row / channel (zero outside the image)
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const int ww = w0 - 3 + i;
        xv[i] = (ww >= 0 && ww < W) ? __ldg(x + ((b * H + hh) * W + ww) * Cin + c) : 0.f;
      }
#pragma unroll
      for (int tx = 0; tx < 7; ++tx) {
        const float* wp = ws + ((ty * 7 + tx) * Cin + c) * Cout + CPT * cg;
#pragma unroll
        for (int j = 0; j < CPT; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(wp + j);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q][j] = fmaf(xv[q + tx], wv.x, acc[q][j]);
            acc[q][j + 1] = fmaf(xv[q + tx], wv.y, acc[q][j + 1]);
            acc[q][j + 2] = fmaf(xv[q + tx], wv.z, acc[q][j + 2]);
            acc[q][j + 3] = fmaf(xv[q + tx], wv.w, acc[q][j + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const size_t off = (((size_t)b * H + h) * W + w0 + q) * Cout + CPT * cg;
    if (out_f32) {
#pragma unroll
      for (int j = 0; j < CPT; j += 4)
        *reinterpret_cast<float4*>(out_f32 + off + j) = make_float4(acc[q][j], acc[q][j + 1], acc[q][j + 2], acc[q][j + 3]);
    }
    if (out_bf16) {
#pragma unroll
      for (int j = 0; j < CPT; j += 8) {
        __align__(16) __nv_bfloat162 o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u] = __floats2bfloat162_rn(acc[q][j + 2 * u], acc[q][j + 2 * u + 1]);
        *reinterpret_cast<uint4*>(out_bf16 + off + j) = *reinterpret_cast<const uint4*>(o);
      }
    }
  }
}

// Last convolution: 1x1, 64 -> Cimg (unet.py:363).  One warp per pixel.
__global__ void __launch_bounds__(256) head_conv_kernel(const float* __restrict__ x, int64_t R, int C, int Cimg,
                                                        const float* __restrict__ wgt, const float* __restrict__ bias,
                                                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  for (int o = 0; o < Cimg; ++o) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(x[row * C + c], __ldg(wgt + (size_t)c * Cimg + o), a);
    a = warp_sum(a);
    if (lane == 0) out[row * Cimg + o] = a + bias[o];
  }
}

// Space-to-depth in front of the 4x4 stride-2 convolution (unet.py:50): out[b, i, j, (r, s, c)] =
// in[b, 2 i - 1 + r, 2 j - 1 + s, c] (zero outside), i in [0, H/2], j in [0, W/2]; the convolution then is 2x2, stride 1.
__global__ void __launch_bounds__(256) space_to_depth_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int C,
                                                             __nv_bfloat16* __restrict__ out) {
  const int Ho = H / 2 + 1, Wo = W / 2 + 1, c8 = C / 8;
  const int64_t total = (int64_t)B * Ho * Wo * 4 * c8;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(idx % c8);
    int64_t t = idx / c8;
    const int rs = (int)(t % 4);
    t /= 4;
    const int j = (int)(t % Wo);
    t /= Wo;
    const int i = (int)(t % Ho);
    const int64_t b = t / Ho;
    const int hh = 2 * i - 1 + (rs >> 1), ww = 2 * j - 1 + (rs & 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = *reinterpret_cast<const uint4*>(in + ((b * H + hh) * W + ww) * C + 8 * cc);
    *reinterpret_cast<uint4*>(out + (((b * Ho + i) * Wo + j) * 4 + rs) * C + 8 * cc) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Closures (experiments/imgs/inpainting.py:94-147).
// assemble: image[b] = concat(us[b], v) = scatter by the mask's index lists (fbs/data/images.py:352-363).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assemble_kernel(const float* __restrict__ us, const float* __restrict__ v,
                                                       const int32_t* __restrict__ unobs, const int32_t* __restrict__ obs, int64_t B,
                                                       int p, int q, int c, float* __restrict__ img) {
  const int P = p + q;
  const int64_t total = B * (int64_t)P * c;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx % c);
    const int e = (int)((idx / c) % P);
    const int64_t b = idx / ((int64_t)c * P);
    if (e < p)
      img[(b * P + unobs[e]) * c + ch] = us[(b * p + e) * c + ch];
    else
      img[(b * P + obs[e - p]) * c + ch] = v[(int64_t)(e - p) * c + ch];
  }
}

// One Euler--Maruyama step of the reverse SDE for every particle + the Gaussian log-weight of the next observation:
//   rd = -a x + g2 score                                  (inpainting.py:102-103; linear SDE drift a x)
//   u' = u + rd_u dt + sd normal(key, (B, p, c))            (:122-128)      [skipped when us_new == nullptr]
//   lw[b] = sum logN(v_next; v_prev + rd_v dt, sd)           (:141-147)      [skipped when lw == nullptr]
// One CTA per particle.
__global__ void __launch_bounds__(256) em_step_kernel(const float* __restrict__ img, const float* __restrict__ score,
                                                      const int32_t* __restrict__ unobs, const int32_t* __restrict__ obs,
                                                      const float* __restrict__ v_next, const uint32_t* __restrict__ key, int64_t B,
                                                      int p, int q, int c, float a, float g2, float dt, float sd,
                                                      int64_t row_offset, int64_t rows_total, const int32_t* __restrict__ pin_row,
                                                      const float* __restrict__ pin_value,
                                                      float* __restrict__ us_new, float* __restrict__ mean_out, float* __restrict__ lw) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  // csmc.py:143: the reference particle's slot receives u*_{k+1} instead of a propagated particle
  const bool pinned = pin_row != nullptr && (row_offset + b) == (int64_t)pin_row[0];
  const int P = p + q;
  const float* xi = img + b * (int64_t)P * c;
  const float* si = score + b * (int64_t)P * c;
  if (us_new || mean_out) {
    const uint32_t nel = (uint32_t)(rows_total * p * c);  // the noise is normal(key, (rows_total, p, c)); this launch owns rows row_offset ..
    Key k{0u, 0u};
    if (key) k = Key{key[0], key[1]};
    for (int e = threadIdx.x; e < p * c; e += blockDim.x) {
      const int pix = unobs[e / c], ch = e % c;
      const float x = xi[pix * c + ch];
      const float rd = -a * x + g2 * si[pix * c + ch];
      const float mean = x + rd * dt;
      if (mean_out) mean_out[b * (int64_t)p * c + e] = mean;
      if (us_new) {
        const uint32_t el = (uint32_t)((row_offset + b) * p * c + e);
        us_new[b * (int64_t)p * c + e] = pinned ? pin_value[e] : mean + sd * bits_to_normal(random_bits_elem(k, nel, el));
      }
    }
  }
  if (lw) {
    float acc = 0.f;
    const float inv = 1.0f / sd, lognorm = -logf(sd) - 0.9189385332046727f;
    for (int e = threadIdx.x; e < q * c; e += blockDim.x) {
      const int pix = obs[e / c], ch = e % c;
      const float x = xi[pix * c + ch];
      const float rd = -a * x + g2 * si[pix * c + ch];
      const float z = (v_next[e] - (x + rd * dt)) * inv;
      acc += -0.5f * z * z + lognorm;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) lw[b] = acc;
  }
}

// normalise (csmc.py:273-292) of a batch of log-weight vectors: log_w = lw - logsumexp(lw), w = exp(log_w).  One CTA per
// chain; the reduction order is a function of (N, blockDim) only, so the sharded and the unsharded sweep -- which both run
// this kernel on the full weight vector -- agree bit for bit.
__global__ void __launch_bounds__(256) normalise_logw_kernel(const float* __restrict__ lw, int N, float* __restrict__ log_w,
                                                             float* __restrict__ w) {
  __shared__ float red[32];
  const float* x = lw + (int64_t)blockIdx.x * N;
  float m = -INFINITY;
  for (int q = threadIdx.x; q < N; q += blockDim.x) m = fmaxf(m, x[q]);
  m = block_max(m, red);
  if (!(fabsf(m) < INFINITY)) m = 0.f;  // jax logsumexp: a non-finite maximum is replaced by 0
  float s = 0.f;
  for (int q = threadIdx.x; q < N; q += blockDim.x) s += expf(x[q] - m);
  s = block_sum(s, red);
  const float lse = logf(s) + m;
  for (int q = threadIdx.x; q < N; q += blockDim.x) {
    const float v = x[q] - lse;
    if (log_w) log_w[(int64_t)blockIdx.x * N + q] = v;
    if (w) w[(int64_t)blockIdx.x * N + q] = expf(v);
  }
}

// One Euler--Maruyama sub-step with the drift ALREADY evaluated (a score / drift network, simulators.py:87):
//   out[b, e] = x[b, e] + drift[b, e] * ddt + gs * normal(key_b, (n,))[e],   gs = dispersion(t) * sqrt(ddt)
// One thread per threefry block (elements e and e + h of the chain's stream).  Products and sums are rounded one by
// one, as the NumPy oracle evaluates the expression.
__global__ void __launch_bounds__(256) em_drift_step_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ x,
                                                            const float* __restrict__ drift, int64_t B, uint32_t n, float ddt,
                                                            float gs, float* __restrict__ out) {
  const uint32_t h = (n + 1u) >> 1;
  const int64_t total = B * (int64_t)h;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / h;
    const uint32_t e = (uint32_t)(t - b * h);
    uint32_t y0, y1;
    random_bits_block(Key{keys[2 * b], keys[2 * b + 1]}, n, e, y0, y1);
    const int64_t o = b * (int64_t)n;
    out[o + e] = __fadd_rn(__fadd_rn(x[o + e], __fmul_rn(drift[o + e], ddt)), __fmul_rn(gs, bits_to_normal(y0)));
    if (e + h < n)
      out[o + e + h] = __fadd_rn(__fadd_rn(x[o + e + h], __fmul_rn(drift[o + e + h], ddt)), __fmul_rn(gs, bits_to_normal(y1)));
  }
}

// rows of src gathered by an index list: dst[b, :] = src[idx[b], :]  (csmc.py:140, the ancestor gather)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int64_t B,
                                                          int64_t row, int src_rows, float* __restrict__ dst) {
  const int64_t total = B * row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / row;
    dst[t] = src[(int64_t)min(max(idx[b], 0), src_rows - 1) * row + (t - b * row)];
  }
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float* __restrict__ x, int64_t n, __nv_bfloat16* __restrict__ y) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
    y[t] = __float2bfloat16_rn(x[t]);
}

static inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace nnops
}  // namespace fbs

using namespace fbs;
using namespace fbs::nnops;

extern "C" {

int fbs_nn_groupnorm_swish_f32(fbs_stream_t s, const float* x, int64_t B, int32_t P, int32_t C, int32_t groups, const float* gamma,
                               const float* beta, const float* time_scale_shift, const float* residual, float eps,
                               float* out_f32, void* out_bf16) {
  FBS_REQUIRE(x && gamma && beta && (out_f32 || out_bf16), "groupnorm: null argument");
  FBS_REQUIRE(groups > 0 && C % groups == 0 && (C / groups) % 4 == 0, "groupnorm: channels per group must be a multiple of 4");
  const int q4 = C / 4, L = (C / groups) / 4;
  const bool pow2 = (q4 & (q4 - 1)) == 0 && (L & (L - 1)) == 0;
  if (pow2 && q4 <= 256 && groups <= 32 && (q4 < 32 ? q4 : 32) / (L < 32 ? L : 32) <= 17 && 2 * B >= sm_count() &&
      (int64_t)P * q4 >= 8192) {
    // large samples, enough of them for one CTA each: whole-pixel accesses (measured: 21.9 -> 16.3 us at 101 x 784 x 64; the
    // per-group kernel stays ahead below ~8k float4 per sample, where its 8x more CTAs hide latency better)
    const size_t sample_bytes = (size_t)P * C * sizeof(float);
    const auto bf = reinterpret_cast<__nv_bfloat16*>(out_bf16);
#define FBS_GN_LAUNCH(STAGED, NT, SMEM)                                                                                     \
  do {                                                                                                                       \
    if ((SMEM) > 48 * 1024)                                                                                                  \
      cudaFuncSetAttribute(gn_sample_kernel<STAGED, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SMEM));          \
    gn_sample_kernel<STAGED, NT><<<(unsigned)B, NT, (SMEM), as_stream(s)>>>(x, P, C, groups, gamma, beta, time_scale_shift, \
                                                                            residual, out_f32, bf, eps);                    \
  } while (0)
    if (sample_bytes <= 200 * 1024) {
      FBS_GN_LAUNCH(true, 1024, sample_bytes);
    } else {
      FBS_GN_LAUNCH(false, 1024, 0);
    }
#undef FBS_GN_LAUNCH
    return check_launch("gn_sample_kernel");
  }
  const size_t slice_bytes = (size_t)P * (C / groups) * sizeof(float);
  if (slice_bytes <= 160 * 1024) {
    if (slice_bytes > 48 * 1024)
      cudaFuncSetAttribute(gn_apply_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice_bytes);
    gn_apply_kernel<true><<<(unsigned)(B * groups), 256, slice_bytes, as_stream(s)>>>(x, P, C, groups, gamma, beta, time_scale_shift, residual, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), eps);
  } else {
    gn_apply_kernel<false><<<(unsigned)(B * groups), 256, 0, as_stream(s)>>>(x, P, C, groups, gamma, beta, time_scale_shift, residual, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), eps);
  }
  return check_launch("gn_apply_kernel");
}

int fbs_nn_groupnorm_swish_stats(fbs_stream_t s, const float* x_f32, const void* x_bf16, const float* partials, int32_t slots,
                                 int64_t B, int32_t P, int32_t C, int32_t groups, const float* gamma, const float* beta,
                                 const float* time_scale_shift, const float* residual, float eps, float* out_f32, void* out_bf16,
                                 const float* ln_gamma, float ln_eps, void* ln_out_bf16) {
  FBS_REQUIRE(((x_f32 != nullptr) != (x_bf16 != nullptr)) && partials && gamma && beta && (out_f32 || out_bf16),
              "groupnorm_stats: null argument (exactly one of x_f32 / x_bf16)");
  FBS_REQUIRE(groups > 0 && groups <= 32 && C % groups == 0 && (C / groups) % 4 == 0 && slots > 0,
              "groupnorm_stats: channels per group must be a multiple of 4, at most 32 groups");
  FBS_REQUIRE(256 % (C / 4) == 0, "groupnorm_stats: C / 4 must divide 256");
  FBS_REQUIRE((ln_gamma != nullptr) == (ln_out_bf16 != nullptr), "groupnorm_stats: ln_gamma and ln_out_bf16 go together");
  FBS_REQUIRE(ln_gamma == nullptr || C <= 128, "groupnorm_stats: the fused LayerNorm needs C <= 128 (a pixel inside one warp)");
  // about five CTAs per SM's worth of (sample, pixel chunk) items, at least 16 pixels each
  int64_t pch = (5 * (int64_t)sm_count() + B - 1) / B;
  if (pch > P / 16) pch = P / 16;
  if (pch < 1) pch = 1;
  const auto bf = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  const auto lnb = reinterpret_cast<__nv_bfloat16*>(ln_out_bf16);
  const auto x16 = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
  const auto pt = reinterpret_cast<const float2*>(partials);
#define FBS_GNS_LAUNCH(IN16, LN)                                                                                                \
  gn_stats_apply_kernel<IN16, LN><<<(unsigned)(B * pch), 256, 0, as_stream(s)>>>(x_f32, x16, pt, slots, P, C, groups, (int)pch, gamma, \
                                                                                 beta, time_scale_shift, residual, out_f32, bf, eps,  \
                                                                                 ln_gamma, ln_eps, lnb)
  if (x_bf16) {
    if (ln_gamma) FBS_GNS_LAUNCH(true, true); else FBS_GNS_LAUNCH(true, false);
  } else {
    if (ln_gamma) FBS_GNS_LAUNCH(false, true); else FBS_GNS_LAUNCH(false, false);
  }
#undef FBS_GNS_LAUNCH
  return check_launch("gn_stats_apply_kernel");
}

int fbs_nn_layernorm_f32(fbs_stream_t s, const float* x, int64_t rows, int32_t C, const float* gamma, const float* residual,
                         float eps, float* out_f32, void* out_bf16) {
  FBS_REQUIRE(x && gamma && (out_f32 || out_bf16), "layernorm: null argument");
  FBS_REQUIRE(C % 32 == 0 && C <= 512, "layernorm: C must be a multiple of 32, <= 512");
  if (C == 64 || C == 128 || C == 256 || C == 512) {
    const int rows_per_cta = 8 * (32 / (C / 4 < 32 ? C / 4 : 32));
    layernorm4_kernel<<<(unsigned)((rows + rows_per_cta - 1) / rows_per_cta), 256, 0, as_stream(s)>>>(x, rows, C, gamma, residual, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), eps);
  } else {
    layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(s)>>>(x, rows, C, gamma, residual, out_f32,
                                                                           reinterpret_cast<__nv_bfloat16*>(out_bf16), eps);
  }
  return check_launch("layernorm_kernel");
}

int fbs_nn_linear_attention_bf16(fbs_stream_t s, const void* qkv, int64_t B, int32_t P, int32_t heads, int32_t dim_head,
                                 void* out_bf16) {
  FBS_REQUIRE(qkv && out_bf16, "linear_attention: null argument");
  FBS_REQUIRE(dim_head == 32 && heads >= 1 && heads <= 8, "linear_attention: dim_head must be 32");
  linear_attention_kernel<<<(unsigned)(B * heads), 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), P, heads,
                                                                           reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("linear_attention_kernel");
}

int fbs_nn_attention_bf16(fbs_stream_t s, const void* qkv, int64_t B, int32_t P, int32_t heads, int32_t dim_head, float scale,
                          void* out_bf16) {
  FBS_REQUIRE(qkv && out_bf16, "attention: null argument");
  FBS_REQUIRE(dim_head == 32 && heads >= 1, "attention: dim_head must be 32");
  const int nthr = P <= 64 ? 64 : 128;  // one query per thread: 49 tokens leave 79 of 128 threads idle in the key loop
  const int qblocks = (P + nthr - 1) / nthr;
  attention_kernel<<<(unsigned)(B * heads * qblocks), nthr, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), P, heads, scale,
                                                                             reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("attention_kernel");
}

int fbs_nn_time_mlp_f32(fbs_stream_t s, const float* tval, float dt, int32_t dim, const float* W0, const float* b0, const float* W1,
                        const float* b1, const float* Wcat, const float* bcat, int32_t nout, float* table) {
  FBS_REQUIRE(tval && W0 && b0 && W1 && b1 && Wcat && bcat && table, "time_mlp: null argument");
  FBS_REQUIRE(dim >= 4 && dim <= 256 && dim % 2 == 0, "time_mlp: 4 <= dim <= 256");
  time_mlp_kernel<<<(unsigned)((nout + 63) / 64), 256, 0, as_stream(s)>>>(tval, 1.0f / dt, dim, W0, b0, W1, b1, Wcat, bcat, nout, table);
  return check_launch("time_mlp_kernel");
}

int fbs_nn_stem_conv_f32(fbs_stream_t s, const float* x, int64_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                         const float* weight, const float* bias, float* out_f32, void* out_bf16) {
  FBS_REQUIRE(x && weight && bias && (out_f32 || out_bf16), "stem_conv: null argument");
  FBS_REQUIRE(Cout % 16 == 0 && (size_t)49 * Cin * Cout * 4 <= 96 * 1024, "stem_conv: weights must fit shared memory");
  const size_t smem = (size_t)49 * Cin * Cout * 4;
  if (W % 4 == 0) {
    // 8 channels per thread: the 16-channel variant needs 128 registers, and 101 x 28 x 28 then misses one resident wave by 4 %
    cudaFuncSetAttribute(stem_conv4_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t n4 = B * H * (W / 4) * (Cout / 8);
    stem_conv4_kernel<8><<<(unsigned)((n4 + 255) / 256), 256, smem, as_stream(s)>>>(x, (int)B, H, W, Cin, Cout, weight, bias, out_f32,
                                                                                    reinterpret_cast<__nv_bfloat16*>(out_bf16));
    return check_launch("stem_conv4_kernel");
  }
  cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t n = B * H * W * (Cout / 16);
  stem_conv_kernel<<<(unsigned)((n + 255) / 256), 256, smem, as_stream(s)>>>(x, (int)B, H, W, Cin, Cout, weight, bias, out_f32,
                                                                             reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("stem_conv_kernel");
}

int fbs_nn_head_conv_f32(fbs_stream_t s, const float* x, int64_t rows, int32_t C, int32_t Cimg, const float* weight,
                         const float* bias, float* out) {
  FBS_REQUIRE(x && weight && bias && out, "head_conv: null argument");
  head_conv_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(s)>>>(x, rows, C, Cimg, weight, bias, out);
  return check_launch("head_conv_kernel");
}

int fbs_nn_space_to_depth_bf16(fbs_stream_t s, const void* in, int64_t B, int32_t H, int32_t W, int32_t C, void* out) {
  FBS_REQUIRE(in && out, "space_to_depth: null argument");
  FBS_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "space_to_depth: need C % 8 == 0 and even H, W");
  const int64_t n = B * (H / 2 + 1) * (W / 2 + 1) * 4 * (C / 8);
  space_to_depth_kernel<<<grid_for(n), 256, 0, as_stream(s)>>>(reinterpret_cast<const __nv_bfloat16*>(in), (int)B, H, W, C,
                                                               reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("space_to_depth_kernel");
}

int fbs_nn_assemble_image_f32(fbs_stream_t s, const float* us, const float* v, const int32_t* unobs_idx, const int32_t* obs_idx,
                              int64_t B, int32_t p, int32_t q, int32_t c, float* img) {
  FBS_REQUIRE(us && v && unobs_idx && obs_idx && img, "assemble_image: null argument");
  assemble_kernel<<<grid_for(B * (int64_t)(p + q) * c), 256, 0, as_stream(s)>>>(us, v, unobs_idx, obs_idx, B, p, q, c, img);
  return check_launch("assemble_kernel");
}

int fbs_nn_em_step_f32(fbs_stream_t s, const float* img, const float* score, const int32_t* unobs_idx, const int32_t* obs_idx,
                       const float* v_next, const uint32_t* key, int64_t B, int32_t p, int32_t q, int32_t c, float a, float g2,
                       float dt, float sd, int64_t row_offset, int64_t rows_total, const int32_t* pin_row, const float* pin_value,
                       float* us_new, float* mean_out, float* lw) {
  FBS_REQUIRE(img && score && unobs_idx && obs_idx, "em_step: null argument");
  FBS_REQUIRE(lw == nullptr || v_next != nullptr, "em_step: v_next missing");
  FBS_REQUIRE(us_new == nullptr || key != nullptr, "em_step: key missing");
  FBS_REQUIRE(row_offset >= 0 && row_offset + B <= rows_total, "em_step: rows [row_offset, row_offset + B) must lie inside rows_total");
  FBS_REQUIRE((pin_row == nullptr) == (pin_value == nullptr), "em_step: pin_row and pin_value go together");
  em_step_kernel<<<(unsigned)B, 256, 0, as_stream(s)>>>(img, score, unobs_idx, obs_idx, v_next, key, B, p, q, c, a, g2, dt, sd,
                                                        row_offset, rows_total, pin_row, pin_value, us_new, mean_out, lw);
  return check_launch("em_step_kernel");
}

int fbs_normalise_logw_f32(fbs_stream_t s, const float* lw, int64_t B, int64_t N, float* log_w, float* w) {
  if (B == 0) return FBS_OK;
  FBS_REQUIRE(lw && (log_w || w), "normalise_logw: null argument");
  FBS_REQUIRE(N >= 1 && N < (1ll << 31) && B < (1ll << 31), "normalise_logw: bad sizes");
  normalise_logw_kernel<<<(unsigned)B, 256, 0, as_stream(s)>>>(lw, (int)N, log_w, w);
  return check_launch("normalise_logw_kernel");
}

int fbs_em_drift_step_f32(fbs_stream_t s, const uint32_t* keys, const float* x, const float* drift, int64_t B, int64_t n,
                          float ddt, float gs, float* out) {
  if (B == 0 || n == 0) return FBS_OK;
  FBS_REQUIRE(keys && x && drift && out, "em_drift_step: null argument");
  FBS_REQUIRE(n > 0 && n < 0xFFFFFFFFll, "em_drift_step: bad n");
  em_drift_step_kernel<<<grid_for(B * ((n + 1) / 2)), 256, 0, as_stream(s)>>>(keys, x, drift, B, (uint32_t)n, ddt, gs, out);
  return check_launch("em_drift_step_kernel");
}

int fbs_gather_rows_f32(fbs_stream_t s, const float* src, const int32_t* idx, int64_t B, int64_t row, int64_t src_rows,
                        float* dst) {
  if (B == 0) return FBS_OK;
  FBS_REQUIRE(src && idx && dst, "gather_rows: null argument");
  FBS_REQUIRE(src_rows >= 1 && src_rows < (1ll << 31), "gather_rows: bad src_rows");
  gather_rows_kernel<<<grid_for(B * row), 256, 0, as_stream(s)>>>(src, idx, B, row, (int)src_rows, dst);
  return check_launch("gather_rows_kernel");
}

int fbs_nn_f32_to_bf16(fbs_stream_t s, const float* x, int64_t n, void* y) {
  FBS_REQUIRE(x && y, "f32_to_bf16: null argument");
  f32_to_bf16_kernel<<<grid_for(n), 256, 0, as_stream(s)>>>(x, n, reinterpret_cast<__nv_bfloat16*>(y));
  return check_launch("f32_to_bf16_kernel");
}

}  // extern "C"
