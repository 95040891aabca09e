// Twisted SMC (fbs/samplers/smc.py:261-309; Algorithm 1 of arXiv 2306.17775) for an affine reverse drift with the
// Gaussian twisting function of experiments/toy/gp_twisted.py:88-129 -- the comparison sampler of SURVEY 8(f) rank 4 --
// as ONE launch for the whole K-step scan.
//
// With reverse_drift(u, t) = M_t u + m_t the closures of gp_twisted.py are all affine / quadratic:
//   denoising estimate   xhat(u, t) = u + dt (M_t u + m_t)                                        (:115)
//   twisting_logpdf      log p~(y | u, t) = sum logN(y; xhat(u, t), sqrt(obs_var))                (:114-116)
//   its gradient         (I + dt M_t)^T (y - xhat) / obs_var      (what jax.grad returns, :88-90)
//   proposal mean        u + dt (M_t u + m_t + g_t^2 grad) = xhat + dt g_t^2 grad                 (:122-124)
//   weights              transition_logpdf + log p~(new) - twisting_prop_logpdf - log p~(prev)    (smc.py:291-293)
// so a particle-step is three d x d matrix-vector products (M u_prev, M^T z, M u_new) with the SAME matrix.  One CTA per
// chain, particles in shared memory, a warp per particle; resampling by the shared warp primitives (sequential cumulative
// sums: indices bit-exact against the oracle).  The transition and proposal densities share their variance, so their
// difference is accumulated per coordinate (no cancellation of two O(d) sums).
#include "fbs_common.cuh"
#include "fbs_resample.cuh"

namespace fbs {

struct TwistedParams {
  int K, d, N, scheme;
  const float *MT, *Mr, *m, *sd, *g2;  // per time index 0..K: MT[t][j][i] = M_t[i][j], Mr[t][j][i] = M_t[j][i], m [K+1, d], sd, g2 [K+1]
  float dt, obs_var;
  const uint32_t* keys;  // [B, 2]: key_filter of smc.py:296
  const float* y;        // [B, d] or [1, d] (y_batched == 0)
  int y_batched;
  const float* x0;       // [B, N, d]: init_sampler(key_init, nparticles), smc.py:299
  int64_t B;
  float *samples, *log_ws;  // [B, N, d], [B, N]
  int32_t* inds;            // optional history [B, K, N]
  float *xs_hist, *lw_hist; // optional history [B, K, N, d], [B, K, N] (normalised)
};

__device__ __forceinline__ float tw_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void tw_normalise(float* lw, int n, int lane) {
  float m = -INFINITY;
  for (int q = lane; q < n; q += 32) m = fmaxf(m, lw[q]);
  m = warp_max(m);
  if (!(fabsf(m) < INFINITY)) m = 0.f;
  float s = 0.f;
  for (int q = lane; q < n; q += 32) s += expf(lw[q] - m);
  s = tw_wsum(s);
  const float lse = logf(s) + m;
  for (int q = lane; q < n; q += 32) lw[q] -= lse;
  __syncwarp();
}

__global__ void __launch_bounds__(256) twisted_smc_kernel(const TwistedParams p) {
  extern __shared__ __align__(16) float sm[];
  const int d = p.d, N = p.N, K = p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  float* xa = sm;                       // [N][d] current particles
  float* xb = xa + (size_t)N * d;       // [N][d] new particles
  float* lps = xb + (size_t)N * d;      // [N] log p~ of the current particles
  float* lpsn = lps + N;                // [N]
  float* lw = lpsn + N;                 // [N]
  float* w = lw + N;                    // [N]
  float* cum = w + N;                   // [N + 1]
  int* idx = reinterpret_cast<int*>(cum + N + 1);
  int* tmp = idx + N;                   // [N + 1]
  float* yv = reinterpret_cast<float*>(tmp + N + 1);  // [d]
  float* scr = yv + d + (warp * 3) * d; // per warp: mu [d], z [d], xnew [d]
  const float ov = p.obs_var, inv_ov = 1.0f / p.obs_var;
  const float lognorm_y = (float)d * logf(6.283185307179586f * ov);
  const uint32_t nel = (uint32_t)N * d;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    const Key kf{p.keys[2 * b], p.keys[2 * b + 1]};
    for (int t = tid; t < N * d; t += blockDim.x) xa[t] = p.x0[(size_t)b * N * d + t];
    for (int t = tid; t < d; t += blockDim.x) yv[t] = p.y[(p.y_batched ? (size_t)b * d : 0) + t];
    __syncthreads();
    float* xs = xa;
    float* xn = xb;
    float* lp = lps;
    float* lpn = lpsn;
    // log p~(y | u, t) for the particles in `src` at time index ti -> dst
    auto twisting = [&](const float* src, int ti, float* dst) {
      const float* MT = p.MT + (size_t)ti * d * d;
      const float* mt = p.m + (size_t)ti * d;
      for (int n = warp; n < N; n += nwarps) {
        const float* u = src + (size_t)n * d;
        float sps = 0.f;
        for (int i = lane; i < d; i += 32) {
          float acc = 0.f;
          for (int j = 0; j < d; ++j) acc = fmaf(__ldg(MT + (size_t)j * d + i), u[j], acc);
          const float res = yv[i] - (u[i] + p.dt * (acc + mt[i]));
          sps = fmaf(res, res, sps);
        }
        sps = tw_wsum(sps);
        if (lane == 0) dst[n] = -0.5f * (sps * inv_ov + lognorm_y);
      }
    };
    twisting(xs, 0, lp);  // smc.py:300
    __syncthreads();
    if (warp == 0) {
      for (int q = lane; q < N; q += 32) lw[q] = lp[q];
      __syncwarp();
      tw_normalise(lw, N, lane);  // smc.py:301
    }
    __syncthreads();

    for (int k = 0; k < K; ++k) {
      const int ti = k + 1;  // the scan walks ts[1:] (smc.py:305)
      const Key key_step = split_key(kf, (uint32_t)K, (uint32_t)k);
      Key key_res, key_prop;
      split2(key_step, key_res, key_prop);  // smc.py:280
      if (warp == 0) {
        for (int q = lane; q < N; q += 32) w[q] = expf(lw[q]);  // smc.py:283
        __syncwarp();
        if (p.scheme == FBS_RESAMPLE_KILLING)
          warp_cond_killing(key_res, w, N, 0, 0, false, cum, tmp, idx, lane);
        else if (p.scheme == FBS_RESAMPLE_MULTINOMIAL)
          warp_sorted_multinomial(key_res, w, N, cum, reinterpret_cast<float*>(tmp), idx, lane);
        else
          warp_systematic_or_stratified(key_res, w, N, p.scheme == FBS_RESAMPLE_SYSTEMATIC, true, cum, idx, lane);
        if (p.inds)
          for (int q = lane; q < N; q += 32) p.inds[((size_t)b * K + k) * N + q] = idx[q];
      }
      __syncthreads();
      const float* MT = p.MT + (size_t)ti * d * d;
      const float* Mr = p.Mr + (size_t)ti * d * d;
      const float* mt = p.m + (size_t)ti * d;
      const float sd = p.sd[ti], g2 = p.g2[ti];
      const float inv_var = 1.0f / (sd * sd);
      float* mu = scr;
      float* z = scr + d;
      float* xw = scr + 2 * d;
      for (int n = warp; n < N; n += nwarps) {
        const int a = idx[n];
        const float* up = xs + (size_t)a * d;  // smc.py:284: xs_prev = xs_prev[resampling_inds]
        for (int i = lane; i < d; i += 32) {
          float acc = 0.f;
          for (int j = 0; j < d; ++j) acc = fmaf(__ldg(MT + (size_t)j * d + i), up[j], acc);
          const float mi = up[i] + p.dt * (acc + mt[i]);  // transition mean = denoising estimate of the parent
          mu[i] = mi;
          z[i] = (yv[i] - mi) * inv_ov;
        }
        __syncwarp();
        float dlp = 0.f;
        for (int i = lane; i < d; i += 32) {
          float acc = 0.f;
          for (int j = 0; j < d; ++j) acc = fmaf(__ldg(Mr + (size_t)j * d + i), z[j], acc);  // (M^T z)_i
          const float grad = z[i] + p.dt * acc;                                               // gp_twisted.py:88-90
          const float mp = mu[i] + p.dt * g2 * grad;                                          // proposal mean, :122-123
          const float x = mp + sd * bits_to_normal(random_bits_elem(key_prop, nel, (uint32_t)n * d + i));  // :124
          xn[(size_t)n * d + i] = x;
          xw[i] = x;
          const float ra = x - mu[i], rb = x - mp;
          dlp = fmaf(-0.5f * inv_var, ra * ra - rb * rb, dlp);  // transition_logpdf - twisting_prop_logpdf, per coordinate
        }
        __syncwarp();
        float sps = 0.f;
        for (int i = lane; i < d; i += 32) {
          float acc = 0.f;
          for (int j = 0; j < d; ++j) acc = fmaf(__ldg(MT + (size_t)j * d + i), xw[j], acc);
          const float res = yv[i] - (xw[i] + p.dt * (acc + mt[i]));
          sps = fmaf(res, res, sps);
        }
        dlp = tw_wsum(dlp);
        sps = tw_wsum(sps);
        if (lane == 0) {
          const float lpnew = -0.5f * (sps * inv_ov + lognorm_y);  // smc.py:290
          lpn[n] = lpnew;
          lw[n] = dlp + lpnew - lp[a];                            // smc.py:291-293 (log_ps_prev gathered, :285)
        }
        __syncwarp();
      }
      __syncthreads();
      if (warp == 0) tw_normalise(lw, N, lane);  // smc.py:294
      {
        float* t1 = xs; xs = xn; xn = t1;
        float* t2 = lp; lp = lpn; lpn = t2;
      }
      __syncthreads();
      if (p.xs_hist)
        for (int t = tid; t < N * d; t += blockDim.x) p.xs_hist[((size_t)b * K + k) * N * d + t] = xs[t];
      if (p.lw_hist)
        for (int t = tid; t < N; t += blockDim.x) p.lw_hist[((size_t)b * K + k) * N + t] = lw[t];
    }
    for (int t = tid; t < N * d; t += blockDim.x) p.samples[(size_t)b * N * d + t] = xs[t];
    for (int t = tid; t < N; t += blockDim.x) p.log_ws[(size_t)b * N + t] = lw[t];
    __syncthreads();
  }
}

}  // namespace fbs

using namespace fbs;

extern "C" {

int fbs_twisted_smc_affine_f32(fbs_stream_t s, const float* MT, const float* Mrow, const float* m, const float* sd,
                               const float* g2, float dt, float obs_var, int64_t K, int64_t d, const uint32_t* keys,
                               const float* y, int y_batched, const float* x0, int scheme, int64_t B, int64_t N,
                               float* samples, float* log_ws, int32_t* inds, float* xs_hist, float* lw_hist) {
  if (B == 0) return FBS_OK;
  FBS_REQUIRE(MT && Mrow && m && sd && g2 && keys && y && x0 && samples && log_ws, "twisted_smc: null argument");
  FBS_REQUIRE(K >= 1 && d >= 1 && N >= 1 && N < (1 << 20) && d < (1 << 15), "twisted_smc: bad sizes");
  FBS_REQUIRE(scheme >= FBS_RESAMPLE_MULTINOMIAL && scheme <= FBS_RESAMPLE_STRATIFIED, "twisted_smc: bad scheme %d", scheme);
  FBS_REQUIRE(obs_var > 0.f && dt > 0.f, "twisted_smc: obs_var and dt must be positive");
  TwistedParams p{};
  p.K = (int)K; p.d = (int)d; p.N = (int)N; p.scheme = scheme;
  p.MT = MT; p.Mr = Mrow; p.m = m; p.sd = sd; p.g2 = g2; p.dt = dt; p.obs_var = obs_var;
  p.keys = keys; p.y = y; p.y_batched = y_batched; p.x0 = x0; p.B = B;
  p.samples = samples; p.log_ws = log_ws; p.inds = inds; p.xs_hist = xs_hist; p.lw_hist = lw_hist;
  const int nwarps = 8;
  const size_t smem = ((size_t)2 * N * d + 4 * N + (N + 1) + N + (N + 1) + d + (size_t)nwarps * 3 * d) * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("twisted_smc: N=%lld d=%lld needs %zu B of shared memory per chain", (long long)N, (long long)d, smem);
    return FBS_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(twisted_smc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("twisted_smc: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, twisted_smc_kernel, 32 * nwarps, smem);
  if (occ < 1) occ = 1;
  int64_t grid = B;
  const int64_t cap = (int64_t)sm_count() * occ;
  if (grid > cap) grid = cap;
  twisted_smc_kernel<<<(int)grid, 32 * nwarps, smem, as_stream(s)>>>(p);
  return check_launch("twisted_smc_kernel");
}

}  // extern "C"
