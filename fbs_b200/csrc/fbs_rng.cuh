// Device-side jax.random (threefry2x32, jax 0.4.26 non-partitionable layout).
//
// The algorithm is the published Threefry-2x32-20 (Random123); the mapping from a flat
// element index to (counter pair, output word) follows jax's `threefry_2x32` wrapper:
// counters iota(n) zero-padded to even length 2h, x0 = c[:h], x1 = c[h:], outputs
// concat(y0, y1)[:n] -- so elements e and e+h share one block.
#pragma once
#include <stdint.h>

namespace fbs {

struct Key {
  uint32_t k0, k1;
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define FBS_TF_ROUND(r) \
  x0 += x1;             \
  x1 = rotl32(x1, r);   \
  x1 ^= x0;
  FBS_TF_ROUND(13) FBS_TF_ROUND(15) FBS_TF_ROUND(26) FBS_TF_ROUND(6)
  x0 += k1;
  x1 += k2 + 1u;
  FBS_TF_ROUND(17) FBS_TF_ROUND(29) FBS_TF_ROUND(16) FBS_TF_ROUND(24)
  x0 += k2;
  x1 += k0 + 2u;
  FBS_TF_ROUND(13) FBS_TF_ROUND(15) FBS_TF_ROUND(26) FBS_TF_ROUND(6)
  x0 += k0;
  x1 += k1 + 3u;
  FBS_TF_ROUND(17) FBS_TF_ROUND(29) FBS_TF_ROUND(16) FBS_TF_ROUND(24)
  x0 += k1;
  x1 += k2 + 4u;
  FBS_TF_ROUND(13) FBS_TF_ROUND(15) FBS_TF_ROUND(26) FBS_TF_ROUND(6)
  x0 += k2;
  x1 += k0 + 5u;
#undef FBS_TF_ROUND
}

// Four blocks in lockstep (independent dependency chains interleaved: the rounds are latency bound one at a time).
// ptxas splits the additions between the ALU pipe (IADD3 / VIADD) and the FMA pipe (IMAD.IADD) by itself.
__device__ __forceinline__ void threefry2x32_x4(uint32_t k0, uint32_t k1, uint32_t (&x0)[4], uint32_t (&x1)[4]) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define FBS_TF_INJ(a, b)            \
  _Pragma("unroll") for (int j = 0; j < 4; ++j) { \
    x0[j] += (a);                   \
    x1[j] += (b);                   \
  }
#define FBS_TF_ROUND4(r)            \
  _Pragma("unroll") for (int j = 0; j < 4; ++j) { \
    x0[j] += x1[j];                 \
    x1[j] = rotl32(x1[j], r);       \
    x1[j] ^= x0[j];                 \
  }
  FBS_TF_INJ(k0, k1)
  FBS_TF_ROUND4(13) FBS_TF_ROUND4(15) FBS_TF_ROUND4(26) FBS_TF_ROUND4(6)
  FBS_TF_INJ(k1, k2 + 1u)
  FBS_TF_ROUND4(17) FBS_TF_ROUND4(29) FBS_TF_ROUND4(16) FBS_TF_ROUND4(24)
  FBS_TF_INJ(k2, k0 + 2u)
  FBS_TF_ROUND4(13) FBS_TF_ROUND4(15) FBS_TF_ROUND4(26) FBS_TF_ROUND4(6)
  FBS_TF_INJ(k0, k1 + 3u)
  FBS_TF_ROUND4(17) FBS_TF_ROUND4(29) FBS_TF_ROUND4(16) FBS_TF_ROUND4(24)
  FBS_TF_INJ(k1, k2 + 4u)
  FBS_TF_ROUND4(13) FBS_TF_ROUND4(15) FBS_TF_ROUND4(26) FBS_TF_ROUND4(6)
  FBS_TF_INJ(k2, k0 + 5u)
#undef FBS_TF_INJ
#undef FBS_TF_ROUND4
}

// Block b (0 <= b < h) of random_bits(key, n): y0 is element b, y1 is element b + h (valid iff b + h < n).
__device__ __forceinline__ void random_bits_block(Key key, uint32_t n, uint32_t b, uint32_t& y0, uint32_t& y1) {
  const uint32_t h = (n + 1u) >> 1;
  uint32_t x0 = b;
  uint32_t x1 = (b + h < n) ? (b + h) : 0u;  // odd n: the single pad counter is zero
  threefry2x32(key.k0, key.k1, x0, x1);
  y0 = x0;
  y1 = x1;
}

// Element e of random_bits(key, n) (one block evaluation, one word kept).
__device__ __forceinline__ uint32_t random_bits_elem(Key key, uint32_t n, uint32_t e) {
  const uint32_t h = (n + 1u) >> 1;
  uint32_t y0, y1;
  random_bits_block(key, n, e < h ? e : e - h, y0, y1);
  return e < h ? y0 : y1;
}

// jax.random.split(key, num)[i]
__device__ __forceinline__ Key split_key(Key key, uint32_t num, uint32_t i) {
  Key out;
  out.k0 = random_bits_elem(key, 2u * num, 2u * i);
  out.k1 = random_bits_elem(key, 2u * num, 2u * i + 1u);
  return out;
}

// split(key, 2) -> both children with two block evaluations.
__device__ __forceinline__ void split2(Key key, Key& a, Key& b) {
  uint32_t y00, y01, y10, y11;
  random_bits_block(key, 4u, 0u, y00, y01);  // elements 0 and 2
  random_bits_block(key, 4u, 1u, y10, y11);  // elements 1 and 3
  a.k0 = y00; a.k1 = y10;
  b.k0 = y01; b.k1 = y11;
}

// split(key, 3)
__device__ __forceinline__ void split3(Key key, Key& a, Key& b, Key& c) {
  uint32_t e0, e3, e1, e4, e2, e5;
  random_bits_block(key, 6u, 0u, e0, e3);
  random_bits_block(key, 6u, 1u, e1, e4);
  random_bits_block(key, 6u, 2u, e2, e5);
  a.k0 = e0; a.k1 = e1;
  b.k0 = e2; b.k1 = e3;
  c.k0 = e4; c.k1 = e5;
}

__device__ __forceinline__ float bits_to_unit(uint32_t bits) {
  return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
}

// jax.random.uniform(key, shape, f32, minval, maxval) applied to one word.
__device__ __forceinline__ float bits_to_uniform(uint32_t bits, float lo, float hi) {
  return fmaxf(lo, __fadd_rn(__fmul_rn(bits_to_unit(bits), __fsub_rn(hi, lo)), lo));
}

// XLA's float32 erf_inv (Giles' single-precision polynomial), |x| < 1.
// w = -log1p(-x^2) = -ln2 * lg2(a), a = 1 - fl(x^2) in [2^-23, 1] (the subtraction is exact or 2^-24-accurate).  a is split as m * 2^e with
// m in [0.75, 1.5) so that the hardware lg2 (MUFU.LG2: absolute error 2^-22 on [0.5, 2]) is only used where it is
// accurate; w then carries ~1e-7 absolute error + its own float32 rounding, i.e. the normal stays within an ULP or
// two of the log1pf evaluation (tests/test_gpu_random.py) at a third of its instruction count.
__device__ __forceinline__ float erfinv_f32(float x) {
  const int ia = __float_as_int(__fsub_rn(1.0f, __fmul_rn(x, x)));  // x * x rounded first, as log1p(-x * x) sees it
  const int eb = (ia - 0x3F400000) & 0xFF800000;            // e << 23
  const float m = __int_as_float(ia - eb);                  // [0.75, 1.5)
  const float fe = __int_as_float((eb >> 23) + 0x4B400000) - 12582912.0f;  // (float)e without the XU-pipe I2F
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(m));
  float w = fmaf(fe, -0.693147182f, -0.693147182f * l2);
  float p;
  if (w < 5.0f) {
    w -= 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    asm volatile("");  // keep the tail (0.3 % of the draws) a real branch: if-converted it costs every draw ~15 instructions
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return p * x;
}

// jax.random.normal applied to one word: sqrt(2) * erf_inv(uniform(nextafter(-1, 0), 1)).
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0) in float32
  // hi - lo = fl(1 + 0.99999994) = 2.0f exactly (ties-to-even), as in the reference computation.
  float u = fmaxf(lo, fmaf(bits_to_unit(bits), 2.0f, lo));
  return 1.41421354f * erfinv_f32(u);
}

// First index i in [0, n) with c[i] >= r (jnp.searchsorted side='left'); n if none.
__device__ __forceinline__ int searchsorted_left(const float* c, int n, float r) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (c[mid] < r) lo = mid + 1; else hi = mid;
  }
  return lo;
}

}  // namespace fbs
