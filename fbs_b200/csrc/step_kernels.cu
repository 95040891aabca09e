// Per-timestep CSMC kernels for particle sets held in global memory (large N): the body of
// fbs/samplers/csmc/csmc.py:132-148 as three launches -- ancestors, fused transition + weight,
// normalise.  The persistent whole-sweep kernel (csmc_kernels.cu) is the fast path when the
// particle set of a chain fits in shared memory; this is the general one.
#include <stdlib.h>
#include "fbs_common.cuh"
#include "fbs_resample.cuh"

namespace fbs {

// (1) A = cond_resampling(key_resampling, exp(log_ws), b_star_prev, b_star, True): one warp per chain.
__global__ void step_ancestors_kernel(int scheme, const uint32_t* __restrict__ step_keys,
                                      const float* __restrict__ log_ws, const int32_t* __restrict__ b_prev,
                                      const int32_t* __restrict__ b_cur, int64_t B, int N, int32_t* __restrict__ A_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  float* w = smem;
  float* cum = w + N;
  int* tmp = reinterpret_cast<int*>(cum + N + 1);
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int q = lane; q < N; q += 32) w[q] = expf(log_ws[b * N + q]);
    __syncwarp();
    const Key key_k{step_keys[2 * b], step_keys[2 * b + 1]};
    Key key_res, key_tr;
    split2(key_k, key_res, key_tr);  // csmc.py:136
    if (scheme == FBS_RESAMPLE_KILLING)
      warp_cond_killing(key_res, w, N, b_prev[b], b_cur[b], true, cum, tmp, A_out + b * N, lane);
    else
      warp_cond_multinomial(key_res, w, N, b_prev[b], b_cur[b], true, cum, A_out + b * N, lane);
    __syncwarp();
  }
}

// (2) fused transition + weight: one warp per child particle.
//   parent = us_prev[A[n]];  drift = M_k [parent; v_prev] + m_k
//   us_out[n] = parent + dt drift_u + sd eps[n]  (reference particle pinned to u_star)
//   lw_out[n] = -0.5 (sum_v (v - v_prev - dt drift_v)^2 / sd^2 + lognorm)         (unnormalised)
__global__ void step_transition_kernel(int du, int dv, const float* __restrict__ MTk, const float* __restrict__ mk,
                                       const float* __restrict__ dtp, const float* __restrict__ sdp,
                                       const float* __restrict__ lnp, int k, const uint32_t* __restrict__ step_keys,
                                       int split_first, const float* __restrict__ us_prev,
                                       const int32_t* __restrict__ A, const float* __restrict__ v,
                                       const float* __restrict__ v_prev, const float* __restrict__ u_star,
                                       const int32_t* __restrict__ b_cur, const float* __restrict__ u_eval, int64_t B,
                                       int N, float* __restrict__ us_out, float* __restrict__ lw_out,
                                       float* __restrict__ tlp_out) {
  const int D = du + dv;
  const float dt = dtp[k], sd = sdp[k], lognorm = lnp[k];
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_s2 = 1.0f / (sd * sd);
  const uint32_t nel = (uint32_t)N * du;
  for (int64_t row = warp_global; row < B * N; row += nwarps) {
    const int64_t b = row / N;
    const int n = (int)(row - b * N);
    Key key_tr{0u, 0u};
    if (step_keys) {
      key_tr = Key{step_keys[2 * b], step_keys[2 * b + 1]};
      if (split_first) {
        Key key_res;
        split2(Key{key_tr.k0, key_tr.k1}, key_res, key_tr);
      }
    }
    const float* parent = us_prev + (b * N + (A ? A[b * N + n] : n)) * du;
    const float* vp = v_prev + b * dv;
    const float* vc = v ? v + b * dv : nullptr;
    const bool pinned = b_cur && (n == b_cur[b]);
    float ss = 0.f, st = 0.f;
    for (int i = lane; i < D; i += 32) {
      float acc = 0.f;
      for (int j = 0; j < du; ++j) acc = fmaf(__ldg(MTk + (size_t)j * D + i), __ldg(parent + j), acc);
      for (int j = 0; j < dv; ++j) acc = fmaf(__ldg(MTk + (size_t)(du + j) * D + i), __ldg(vp + j), acc);
      const float drift = acc + mk[i];
      if (i < du) {
        const float mean = parent[i] + dt * drift;
        if (us_out) {
          const float eps = bits_to_normal(random_bits_elem(key_tr, nel, (uint32_t)n * du + i));
          us_out[(b * N + n) * du + i] = pinned ? u_star[b * du + i] : mean + sd * eps;
        }
        if (u_eval) {
          const float resid = u_eval[b * du + i] - mean;
          st = fmaf(resid, resid, st);
        }
      } else if (vc) {
        const int q = i - du;
        const float resid = (vc[q] - vp[q]) - dt * drift;
        ss = fmaf(resid, resid, ss);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      st += __shfl_xor_sync(0xffffffffu, st, o);
    }
    if (lane == 0) {
      if (lw_out) lw_out[b * N + n] = -0.5f * (ss * inv_s2 + lognorm);
      // sum_du logN(u; mean, sd): the normaliser is du * log(2 pi sd^2) = lognorm * du / dv
      if (tlp_out) tlp_out[b * N + n] = -0.5f * (st * inv_s2 + lognorm * ((float)du / (float)dv));
    }
  }
}

// (2b) fused transition + weight for SMALL state dimensions (du, dv <= 16): one thread per particle PAIR (n, n + N/2) --
// the two particles whose noise comes from the same threefry blocks -- with the parent, the drift and the children in
// registers, the step matrix and the chain's constant drift part in shared memory (broadcast reads), 8-byte vector
// loads / stores of the particle rows.  Per particle: du (du + dv) + dv (du + dv) FMAs, du normals (~65 instructions
// each): the kernel is bound by the in-kernel RNG, not by HBM (profiles/r1_hbm_kernels.md).
template <int DUT, int DVT>
__global__ void __launch_bounds__(128) step_transition_small_kernel(
    int du, int dv, const float* __restrict__ MTk, const float* __restrict__ mk, const float* __restrict__ dtp,
    const float* __restrict__ sdp, const float* __restrict__ lnp, int k, const uint32_t* __restrict__ step_keys,
    const float* __restrict__ us_prev, const int32_t* __restrict__ A, const float* __restrict__ v,
    const float* __restrict__ v_prev, const float* __restrict__ u_star, const int32_t* __restrict__ b_cur, int64_t B, int N,
    int blocks_per_chain, float* __restrict__ us_out, float* __restrict__ lw_out) {
  constexpr int DT = DUT + DVT;
  __shared__ float Ms[DUT][DT];  // Ms[j][i] = M_k[i][j], u inputs only
  __shared__ float cs[DT];       // m_k + M_k[:, du:] v_prev  (u rows), and for v rows: (v - v_prev) - dt (m + M v_prev)
  const int D = du + dv, half = N / 2;
  const float dt = dtp[k], sd = sdp[k], lognorm = lnp[k];
  const float inv_s2 = 1.0f / (sd * sd);
  for (int64_t blk = blockIdx.x; blk < B * blocks_per_chain; blk += gridDim.x) {
    const int64_t b = blk / blocks_per_chain;
    const int pb = (int)(blk - b * blocks_per_chain);
    __syncthreads();
    for (int t = threadIdx.x; t < DUT * DT; t += blockDim.x) {
      const int j = t / DT, i = t - j * DT;
      const int ii = i < DUT ? i : du + (i - DUT);  // padded column -> model row
      const bool ok = j < du && (i < DUT ? i < du : i - DUT < dv);
      Ms[j][i] = ok ? __ldg(MTk + (size_t)j * D + ii) : 0.f;
    }
    if (threadIdx.x < DT) {
      const int i = threadIdx.x;
      const int ii = i < DUT ? i : du + (i - DUT);
      const bool ok = i < DUT ? i < du : i - DUT < dv;
      float acc = 0.f;
      if (ok) {
        acc = mk[ii];
        for (int j = 0; j < dv; ++j) acc = fmaf(__ldg(MTk + (size_t)(du + j) * D + ii), v_prev[b * dv + j], acc);
        if (i >= DUT) acc = (v[b * dv + (i - DUT)] - v_prev[b * dv + (i - DUT)]) - dt * acc;
      }
      cs[i] = acc;
    }
    __syncthreads();
    const int n = pb * blockDim.x + threadIdx.x;
    if (n >= half) continue;
    Key key_res, key_tr;
    split2(Key{step_keys[2 * b], step_keys[2 * b + 1]}, key_res, key_tr);  // csmc.py:136
    const int bc = b_cur[b];
    const uint32_t nel = (uint32_t)N * du;
    float u[2][DUT], child[2][DUT];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const int nn = n + s2 * half;
      const float* parent = us_prev + ((size_t)b * N + A[b * N + nn]) * du;
#pragma unroll
      for (int j = 0; j < DUT; j += 2) {  // rows are 8-byte aligned when du is even
        if (j + 1 < du) {
          const float2 x = *reinterpret_cast<const float2*>(parent + j);
          u[s2][j] = x.x;
          u[s2][j + 1] = x.y;
        } else {
          u[s2][j] = j < du ? parent[j] : 0.f;
          if (j + 1 < DUT) u[s2][j + 1] = 0.f;
        }
      }
    }
    // drift: u rows -> means, v rows -> residuals
    float ss[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j < DUT; ++j) {
        const float mji = Ms[j][i];
        a0 = fmaf(mji, u[0][j], a0);
        a1 = fmaf(mji, u[1][j], a1);
      }
      if (i < DUT) {
        child[0][i] = u[0][i] + dt * (a0 + cs[i]);
        child[1][i] = u[1][i] + dt * (a1 + cs[i]);
      } else {
        const float r0 = cs[i] - dt * a0, r1 = cs[i] - dt * a1;
        ss[0] = fmaf(r0, r0, ss[0]);
        ss[1] = fmaf(r1, r1, ss[1]);
      }
    }
    // noise: block e = n * du + i gives element e (particle n) and e + N du / 2 (particle n + N/2)
#pragma unroll
    for (int i = 0; i < DUT; ++i) {
      if (i < du) {
        uint32_t y0, y1;
        random_bits_block(key_tr, nel, (uint32_t)n * du + i, y0, y1);
        child[0][i] += sd * bits_to_normal(y0);
        child[1][i] += sd * bits_to_normal(y1);
      }
    }
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const int nn = n + s2 * half;
      float* dst = us_out + ((size_t)b * N + nn) * du;
      const bool pinned = nn == bc;
#pragma unroll
      for (int j = 0; j < DUT; j += 2) {
        if (j + 1 < du) {
          float2 x = make_float2(child[s2][j], child[s2][j + 1]);
          if (pinned) x = *reinterpret_cast<const float2*>(u_star + (size_t)b * du + j);
          *reinterpret_cast<float2*>(dst + j) = x;
        } else if (j < du) {
          dst[j] = pinned ? u_star[(size_t)b * du + j] : child[s2][j];
        }
      }
      lw_out[(size_t)b * N + nn] = -0.5f * (ss[s2] * inv_s2 + lognorm);
    }
  }
}

// (3) log_ws -= logsumexp(log_ws): one CTA per chain.
__global__ void step_normalise_kernel(float* __restrict__ lw, int64_t B, int N) {
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    float* x = lw + b * N;
    float m = -INFINITY;
    for (int q = tid; q < N; q += blockDim.x) m = fmaxf(m, x[q]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = (lane < nw) ? red[lane] : -INFINITY;
    m = warp_max(m);
    if (!(fabsf(m) < INFINITY)) m = 0.f;
    __syncthreads();
    float s = 0.f;
    for (int q = tid; q < N; q += blockDim.x) s += expf(x[q] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = (lane < nw) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lse = logf(s) + m;
    for (int q = tid; q < N; q += blockDim.x) x[q] -= lse;
    __syncthreads();
  }
}

}  // namespace fbs

using namespace fbs;

extern "C" int fbs_csmc_step_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, int32_t k, int scheme,
                                        const uint32_t* step_keys, const float* us_prev, const float* log_ws,
                                        const float* v, const float* v_prev, const float* u_star,
                                        const int32_t* b_star_prev, const int32_t* b_star, int64_t B, int64_t N,
                                        int32_t* A_out, float* us_out, float* log_ws_out) {
  FBS_REQUIRE(model && model->MT && model->m && model->dt && model->sd && model->lognorm, "csmc_step: bad model");
  FBS_REQUIRE(k >= 0 && k < model->K, "csmc_step: step index %d out of range [0, %d)", k, model->K);
  FBS_REQUIRE(step_keys && us_prev && log_ws && v && v_prev && u_star && b_star_prev && b_star && A_out && us_out &&
                  log_ws_out,
              "csmc_step: null pointer");
  FBS_REQUIRE(scheme == FBS_RESAMPLE_KILLING || scheme == FBS_RESAMPLE_MULTINOMIAL, "csmc_step: bad scheme %d", scheme);
  FBS_REQUIRE(B >= 0 && N >= 1 && N < (1ll << 30), "csmc_step: bad sizes");
  FBS_REQUIRE(us_prev != us_out, "csmc_step: us_out must not alias us_prev");
  if (B == 0) return FBS_OK;
  const int du = model->du, dv = model->dv, D = du + dv;
  int rc = launch_resample_tile(as_stream(s), scheme, step_keys, log_ws, b_star_prev, b_star, 1, 0, 1, 1, B, N, A_out);
  const bool tiled = rc != FBS_ERR_UNSUPPORTED;
  if (tiled && rc) return rc;
  const size_t smem = ((size_t)3 * N + 1) * sizeof(float);
  if (!tiled && smem > 200 * 1024) {
    set_error("csmc_step: N=%lld exceeds the single-warp resampling limit (multi-CTA scan not built yet)", (long long)N);
    return FBS_ERR_UNSUPPORTED;
  }
  const int64_t cap = (int64_t)sm_count() * 8;
  if (!tiled) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(step_ancestors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    step_ancestors_kernel<<<(int)(B > cap ? cap : B), 32, smem, as_stream(s)>>>(scheme, step_keys, log_ws, b_star_prev,
                                                                             b_star, B, (int)N, A_out);
    rc = check_launch("step_ancestors_kernel");
    if (rc) return rc;
  }
  const float* MTk = model->MT + (size_t)k * D * D;
  const float* mk = model->m + (size_t)k * D;
  int64_t blocks = (B * N + 7) / 8;  // 8 warps (children) per CTA
  const int64_t cap2 = (int64_t)sm_count() * 8;
  if (blocks > cap2) blocks = cap2;
  // OPT_STEP_IMPL = 1 pins the CUDA-core kernels (tests compare the two)
  rc = (debug_opt(OPT_STEP_IMPL) == 1) ? -1
                                           : launch_step_transition_tc(as_stream(s), model, k, step_keys, us_prev, A_out, v,
                                                                       v_prev, u_star, b_star, B, N, us_out, log_ws_out);
  if (rc > 0) return rc;
  if (rc == 0) {
    // done on the tensor cores
  } else if (du <= 16 && dv <= 16 && du % 2 == 0 && N % 2 == 0) {
    // small state: thread per particle pair, everything in registers
    const int bpc = (int)((N / 2 + 127) / 128);
    int64_t nblk = B * bpc;
    if (nblk > cap2 * 4) nblk = cap2 * 4;
    if (du == 10 && dv == 10)
      step_transition_small_kernel<10, 10><<<(int)nblk, 128, 0, as_stream(s)>>>(du, dv, MTk, mk, model->dt, model->sd, model->lognorm,
                                                                              k, step_keys, us_prev, A_out, v, v_prev, u_star,
                                                                              b_star, B, (int)N, bpc, us_out, log_ws_out);
    else
      step_transition_small_kernel<16, 16><<<(int)nblk, 128, 0, as_stream(s)>>>(du, dv, MTk, mk, model->dt, model->sd, model->lognorm,
                                                                              k, step_keys, us_prev, A_out, v, v_prev, u_star,
                                                                              b_star, B, (int)N, bpc, us_out, log_ws_out);
    rc = check_launch("step_transition_small_kernel");
  } else {
    step_transition_kernel<<<(int)blocks, 256, 0, as_stream(s)>>>(du, dv, MTk, mk, model->dt, model->sd, model->lognorm, k,
                                                                  step_keys, 1, us_prev, A_out, v, v_prev, u_star, b_star,
                                                                  nullptr, B, (int)N, us_out, log_ws_out, nullptr);
    rc = check_launch("step_transition_kernel");
  }
  if (rc) return rc;
  step_normalise_kernel<<<(int)(B > cap ? cap : B), 256, 0, as_stream(s)>>>(log_ws_out, B, (int)N);
  return check_launch("step_normalise_kernel");
}

// The three closures of the reference drivers evaluated on their own (experiments/toy/gp_gibbs.py:120-135):
//   us_out  = transition_sampler(us_prev, v_prev, t_k, key)      (needs tr_keys [B,2])
//   lw_out  = likelihood_logpdf(v, us_prev, v_prev, t_k)         (needs v [B,dv])
//   tlp_out = transition_logpdf(u_eval, us_prev, v_prev, t_k)    (needs u_eval [B,du])
// Any output may be NULL.  us_prev [B,N,du], v_prev [B,dv].
extern "C" int fbs_affine_eval_f32(fbs_stream_t s, const fbs_affine_model_t* model, int32_t k, const uint32_t* tr_keys,
                                   const float* us_prev, const float* v, const float* v_prev, const float* u_eval,
                                   int64_t B, int64_t N, float* us_out, float* lw_out, float* tlp_out) {
  FBS_REQUIRE(model && model->MT && model->m && model->dt && model->sd && model->lognorm, "affine_eval: bad model");
  FBS_REQUIRE(k >= 0 && k < model->K, "affine_eval: step index %d out of range [0, %d)", k, model->K);
  FBS_REQUIRE(us_prev && v_prev, "affine_eval: null input");
  FBS_REQUIRE(!us_out || tr_keys, "affine_eval: us_out needs tr_keys");
  FBS_REQUIRE(!lw_out || v, "affine_eval: lw_out needs v");
  FBS_REQUIRE(!tlp_out || u_eval, "affine_eval: tlp_out needs u_eval");
  FBS_REQUIRE(us_out != us_prev, "affine_eval: us_out must not alias us_prev");
  FBS_REQUIRE(B >= 0 && N >= 1, "affine_eval: bad sizes");
  if (B == 0) return FBS_OK;
  const int du = model->du, dv = model->dv, D = du + dv;
  int64_t blocks = (B * N + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  step_transition_kernel<<<(int)blocks, 256, 0, as_stream(s)>>>(
      du, dv, model->MT + (size_t)k * D * D, model->m + (size_t)k * D, model->dt, model->sd, model->lognorm, k, tr_keys, 0,
      us_prev, nullptr, lw_out ? v : nullptr, v_prev, nullptr, nullptr, tlp_out ? u_eval : nullptr, B, (int)N, us_out,
      lw_out, tlp_out);
  return check_launch("step_transition_kernel");
}
