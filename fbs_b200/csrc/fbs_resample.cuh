// Warp-cooperative resampling primitives shared by the standalone resampling kernels and the
// fused CSMC / pMCMC sweep kernels.  One warp works on one chain's N weights.
//
// Summation-order contract: every cumulative sum / sum that decides an index is
// SEQUENTIAL float32, c[i] = fl(c[i-1] + w[i]) -- done by lane 0 -- so that indices are
// bit-exact against the oracle on identical weights and keys.  max() is order-free.
//
// Reference: fbs/samplers/csmc/resamplings.py (conditional), fbs/samplers/resampling.py.
#pragma once
#include "fbs_rng.cuh"

namespace fbs {

// Reference-particle indices (bs_star, i, j) address shared / global memory inside the kernels.  JAX clamps an
// out-of-range gather / scatter index silently; here it is clamped where it is loaded so that a stale bs_star (reused
// after nparticles changed) cannot corrupt memory.  The Python layer validates the range up front where that costs nothing.
__device__ __forceinline__ int clamp_index(int b, int n) { return min(max(b, 0), n - 1); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Summation-order contract for LONG rows (n >= kChunkedMinN; the oracle's jax_random.seq_cumsum / seq_sum follow the same
// definition): chunks of 8 consecutive elements,
//   local_c[t] = sequential sum of the chunk's first t + 1 elements,  P_0 = 0,  P_{c+1} = fl(P_c + local_c[last]),
//   cum[8 c + t] = fl(P_c + local_c[t]).
// The serial dependency is n / 8 additions instead of n (a 16384-particle row: 8 k instead of 65 k cycles) and the 8-element
// chunks map onto lanes; rows of the sweep kernels (N <= 128) and every golden file (N <= 257) keep the plain sequential sum.
constexpr int kChunkedMinN = 1024;

// chunked cumulative sum of a long row by the whole warp; w and cum may alias.  Returns the total to all lanes.
__device__ __forceinline__ float warp_chunked_cumsum(const float* w, float* cum, int n, int lane, bool store) {
  float running = 0.f;  // P of the tile's first chunk (uniform over the warp)
  for (int base = 0; base < n; base += 256) {
    const int c0 = base + 8 * lane;  // this lane's chunk
    float loc[8];
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float x = c0 + t < n ? w[c0 + t] : 0.f;
      acc = __fadd_rn(acc, x);
      loc[t] = acc;
    }
    float myP = 0.f;
#pragma unroll 4
    for (int s = 0; s < 32; ++s) {  // P_{c+1} = fl(P_c + total_c), chunk by chunk (the only serial part)
      const float t = __shfl_sync(0xffffffffu, acc, s);
      if (lane == s) myP = running;
      running = __fadd_rn(running, t);
    }
    if (store) {
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (c0 + t < n) cum[c0 + t] = __fadd_rn(myP, loc[t]);
    }
  }
  __syncwarp();
  return running;
}

// sequential float32 sum of a row (the oracle's seq_sum): by lane 0 for short rows, chunked for long ones.  To all lanes.
__device__ __forceinline__ float warp_seq_sum(const float* w, int n, int lane) {
  if (n >= kChunkedMinN) return warp_chunked_cumsum(w, nullptr, n, lane, false);
  float acc = 0.f;
  if (lane == 0) {
    int q = 0;
    for (; q + 8 <= n; q += 8) {
      float x[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) x[t] = w[q + t];
#pragma unroll
      for (int t = 0; t < 8; ++t) acc = __fadd_rn(acc, x[t]);
    }
    for (; q < n; ++q) acc = __fadd_rn(acc, w[q]);
  }
  return __shfl_sync(0xffffffffu, acc, 0);
}

// cum[i] = cumsum of w in the contract's order (sequential by lane 0 for n < kChunkedMinN).  w and cum may alias.
// Returns cum[n-1] to all lanes.
__device__ __forceinline__ float warp_seq_cumsum(const float* w, float* cum, int n, int lane) {
  if (n >= kChunkedMinN) return warp_chunked_cumsum(w, cum, n, lane, true);
  float acc = 0.f;
  if (lane == 0) {
    // batches of 8: independent loads first, then the dependent FADD chain (the only serial part), then stores
    int i = 0;
    if (((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(cum)) & 15) == 0) {
      // 16-byte aligned rows: two vector loads / stores per 8 elements (the lane issues 12 instead of 24 instructions)
      for (; i + 8 <= n; i += 8) {
        float4 a = *reinterpret_cast<const float4*>(w + i), b = *reinterpret_cast<const float4*>(w + i + 4);
        a.x = acc = __fadd_rn(acc, a.x);
        a.y = acc = __fadd_rn(acc, a.y);
        a.z = acc = __fadd_rn(acc, a.z);
        a.w = acc = __fadd_rn(acc, a.w);
        b.x = acc = __fadd_rn(acc, b.x);
        b.y = acc = __fadd_rn(acc, b.y);
        b.z = acc = __fadd_rn(acc, b.z);
        b.w = acc = __fadd_rn(acc, b.w);
        *reinterpret_cast<float4*>(cum + i) = a;
        *reinterpret_cast<float4*>(cum + i + 4) = b;
      }
    }
    for (; i + 8 <= n; i += 8) {
      float x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = w[i + q];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        acc = __fadd_rn(acc, x[q]);
        x[q] = acc;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) cum[i + q] = x[q];
    }
    for (; i < n; ++i) {
      acc = __fadd_rn(acc, w[i]);
      cum[i] = acc;
    }
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, acc, 0);
}

// jax.random.choice(key, n, (n_draws,), p=w) given the cumulative sums: draw e.
__device__ __forceinline__ int choice_from_cum(const float* cum, int n, float u) {
  float r = __fmul_rn(cum[n - 1], __fsub_rn(1.0f, u));
  return searchsorted_left(cum, n, r);
}

// Unconditional part of `killing` (resamplings.py:63-74 / resampling.py:93-101):
// idx[n] = killed ? choice : n.  cum must hold cumsum(w).  Writes idx for all n.
__device__ __forceinline__ void warp_killing_unconditional(Key key_1, Key key_2, const float* w, const float* cum,
                                                           float w_max, int n, int lane, int* idx) {
  const uint32_t h = ((uint32_t)n + 1u) >> 1;
  for (uint32_t b = lane; b < h; b += 32) {
    uint32_t a0, a1, c0, c1;
    random_bits_block(key_1, n, b, a0, a1);
    random_bits_block(key_2, n, b, c0, c1);
    {
      float u1 = bits_to_unit(a0);
      bool killed = __fmul_rn(u1, w_max) >= w[b];
      idx[b] = killed ? choice_from_cum(cum, n, bits_to_unit(c0)) : (int)b;
    }
    uint32_t e = b + h;
    if (e < (uint32_t)n) {
      float u1 = bits_to_unit(a1);
      bool killed = __fmul_rn(u1, w_max) >= w[e];
      idx[e] = killed ? choice_from_cum(cum, n, bits_to_unit(c1)) : (int)e;
    }
  }
  __syncwarp();
}

// Conditional killing resampling, resamplings.py:40-88.
//   w     [n] normalised weights (shared or global memory, read only)
//   cum   [n] scratch            tmp [n] int scratch          out [n] ancestor indices
// All pointers must be visible to the whole warp.  i = pinned ancestor value, j = pinned slot.
__device__ __forceinline__ void warp_cond_killing(Key key, const float* w, int n, int i, int j, bool conditional,
                                                  float* cum, int* tmp, int* out, int lane) {
  Key key_1, key_2, key_3;
  split3(key, key_1, key_2, key_3);  // :66
  i = clamp_index(i, n);
  j = clamp_index(j, n);
  float m = -INFINITY;
  for (int q = lane; q < n; q += 32) m = fmaxf(m, w[q]);
  const float w_max = warp_max(m);  // :69
  warp_seq_cumsum(w, cum, n, lane);
  int* dst = conditional ? tmp : out;
  warp_killing_unconditional(key_1, key_2, w, cum, w_max, n, lane, dst);
  if (!conditional) return;

  // J_prob = (1 - w / w_max) / N, J_prob[i] = max(1 - sum(J_prob with J_prob[i] = 0), 0)   (:79-82)
  const float fn = (float)n;
  // J_prob computed in parallel into cum[], then two passes in the contract's summation order (sum, cumulative sum)
  for (int q = lane; q < n; q += 32) cum[q] = (q == i) ? 0.f : __fdiv_rn(__fsub_rn(1.0f, __fdiv_rn(w[q], w_max)), fn);
  __syncwarp();
  const float acc = warp_seq_sum(cum, n, lane);
  if (lane == 0) cum[i] = fmaxf(__fsub_rn(1.0f, acc), 0.f);
  __syncwarp();
  warp_seq_cumsum(cum, cum, n, lane);
  // J ~ Cat(J_prob): choice(key_3, N, (), p=J_prob)   (:84) -- random_bits(key_3, 1) = block (0, 0), word 0
  uint32_t x0 = 0u, x1 = 0u;
  threefry2x32(key_3.k0, key_3.k1, x0, x1);
  const int J = choice_from_cum(cum, n, bits_to_unit(x0));
  // idx = roll(idx, j - J); idx[j] = i   (:85-86):  out[m] = tmp[(m - (j - J)) mod n]
  int shift = (j - J) % n;
  if (shift < 0) shift += n;
  for (int q = lane; q < n; q += 32) {
    int src = q - shift;
    if (src < 0) src += n;
    out[q] = (q == j) ? i : tmp[src];
  }
  __syncwarp();
}

// Conditional multinomial, resamplings.py:10-37.
__device__ __forceinline__ void warp_cond_multinomial(Key key, const float* w, int n, int i, int j, bool conditional,
                                                      float* cum, int* out, int lane) {
  i = clamp_index(i, n);
  j = clamp_index(j, n);
  warp_seq_cumsum(w, cum, n, lane);
  const uint32_t h = ((uint32_t)n + 1u) >> 1;
  for (uint32_t b = lane; b < h; b += 32) {
    uint32_t c0, c1;
    random_bits_block(key, n, b, c0, c1);
    out[b] = (conditional && (int)b == j) ? i : choice_from_cum(cum, n, bits_to_unit(c0));
    uint32_t e = b + h;
    if (e < (uint32_t)n) out[e] = (conditional && (int)e == j) ? i : choice_from_cum(cum, n, bits_to_unit(c1));
  }
  __syncwarp();
}

// systematic / stratified.  clip = true: fbs/samplers/resampling.py:43-59; clip = false with
// systematic: fbs/samplers/csmc/resamplings.py:120-125.
__device__ __forceinline__ void warp_systematic_or_stratified(Key key, const float* w, int n, bool is_systematic,
                                                              bool clip, float* cum, int* out, int lane) {
  warp_seq_cumsum(w, cum, n, lane);
  const float fn = (float)n;
  if (is_systematic) {
    uint32_t x0 = 0u, x1 = 0u;  // uniform(key, ()) = random_bits(key, 1) word 0
    threefry2x32(key.k0, key.k1, x0, x1);
    const float u = bits_to_unit(x0);
    for (int q = lane; q < n; q += 32) {
      float pt = __fdiv_rn(__fadd_rn((float)q, u), fn);
      int id = searchsorted_left(cum, n, pt);
      out[q] = clip ? min(max(id, 0), n - 1) : id;
    }
  } else {
    const uint32_t h = ((uint32_t)n + 1u) >> 1;
    for (uint32_t b = lane; b < h; b += 32) {
      uint32_t c0, c1;
      random_bits_block(key, n, b, c0, c1);
      {
        float pt = __fdiv_rn(__fadd_rn((float)b, bits_to_unit(c0)), fn);
        int id = searchsorted_left(cum, n, pt);
        out[b] = clip ? min(max(id, 0), n - 1) : id;
      }
      uint32_t e = b + h;
      if (e < (uint32_t)n) {
        float pt = __fdiv_rn(__fadd_rn((float)e, bits_to_unit(c1)), fn);
        int id = searchsorted_left(cum, n, pt);
        out[e] = clip ? min(max(id, 0), n - 1) : id;
      }
    }
  }
  __syncwarp();
}

// Sorted-uniform multinomial, fbs/samplers/resampling.py:36-40,62-68 ("Not tested." upstream).
// pts scratch [n + 1].
__device__ __forceinline__ void warp_sorted_multinomial(Key key, const float* w, int n, float* cum, float* pts, int* out,
                                                        int lane) {
  warp_seq_cumsum(w, cum, n, lane);
  const uint32_t n1 = (uint32_t)n + 1u;
  const uint32_t h = (n1 + 1u) >> 1;
  for (uint32_t b = lane; b < h; b += 32) {
    uint32_t c0, c1;
    random_bits_block(key, n1, b, c0, c1);
    pts[b] = -logf(bits_to_unit(c0));
    if (b + h < n1) pts[b + h] = -logf(bits_to_unit(c1));
  }
  __syncwarp();
  const float z_last = warp_seq_cumsum(pts, pts, (int)n1, lane);
  for (int q = lane; q < n; q += 32) {
    int id = searchsorted_left(cum, n, __fdiv_rn(pts[q], z_last));
    out[q] = min(max(id, 0), n - 1);
  }
  __syncwarp();
}

}  // namespace fbs
