// Backward passes over a stored particle history for the affine-Gaussian model, as device kernels:
//   * backward_sampling_pass      fbs/samplers/csmc/csmc.py:167-227   (mode 0: weights of the forward pass given)
//   * bootstrap_backward_smoother fbs/samplers/smc.py:91-112          (mode 1: no weights, u_T drawn uniformly)
// Both walk t = K-1 .. 0 and draw B_t ~ Cat(w_t), w_t[n] prop. to p(x_{t+1} | u_t^n) (x the stored weight in mode 0); the
// transition density needs the drift of every stored particle, i.e. one du x du matrix-vector product per particle and
// step -- the same work as the forward sweep.  One CTA per chain: the step's u-u block of the matrix and the chain's
// step vector are staged in shared memory, a warp per particle evaluates log p(x | u^n), warp 0 normalises, runs the
// sequential cumulative sum of the summation-order contract and draws the index.
#include "fbs_common.cuh"
#include "fbs_resample.cuh"

namespace fbs {

struct BackwardParams {
  int K, du, dv, N, mode;
  const float *MT, *m, *dt, *sd, *lognorm;
  const uint32_t* keys;
  const float* vs;       // [B, K+1, dv]
  const float* uss;      // [B, K+1, N, du]
  const float* log_wss;  // [B, K+1, N] (mode 0) or NULL
  int64_t B;
  int shared;            // 1: vs / uss / log_wss carry no chain axis (every chain walks the same stored history)
  float* xs;             // [B, K+1, du]
  int32_t* bs;           // [B, K+1] or NULL
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// w[0..n) = exp(lw - logsumexp(lw)) in place (one warp); jax logsumexp: a non-finite maximum is replaced by 0
__device__ __forceinline__ void warp_softmax(float* lw, int n, int lane) {
  float m = -INFINITY;
  for (int q = lane; q < n; q += 32) m = fmaxf(m, lw[q]);
  m = warp_max(m);
  if (!(fabsf(m) < INFINITY)) m = 0.f;
  float s = 0.f;
  for (int q = lane; q < n; q += 32) s += expf(lw[q] - m);
  s = warp_sum_f(s);
  const float lse = logf(s) + m;
  for (int q = lane; q < n; q += 32) lw[q] = expf(lw[q] - lse);
  __syncwarp();
}

__global__ void __launch_bounds__(256) backward_sample_kernel(const BackwardParams p) {
  extern __shared__ __align__(16) float sm[];
  const int du = p.du, dv = p.dv, D = du + dv, N = p.N, K = p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  float* Muu = sm;                       // [du][du]: Muu[j * du + i] = MT[j][i] (input-major)
  float* cvec = Muu + (size_t)du * du;   // [du]: m_u + M_uv v_prev
  float* x = cvec + du;                  // [du]: the trajectory's state at t + 1
  float* w = x + du;                     // [N]
  __shared__ int s_idx;
  int S = 1;
  while (S < du && S < 32) S <<= 1;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    const Key key{p.keys[2 * b], p.keys[2 * b + 1]};
    const size_t hb = p.shared ? 0 : (size_t)b;  // chain index into the stored history
    const float* uss = p.uss + hb * (K + 1) * N * du;
    Key key_steps = key;  // mode 0: keys = split(key, K + 1) (csmc.py:194); mode 1: split(split(key)[1], K) (smc.py:108,110)
    // ---- terminal index
    if (warp == 0) {
      int BT;
      if (p.mode == 0) {
        const float* lw = p.log_wss + (hb * (K + 1) + K) * N;
        for (int q = lane; q < N; q += 32) w[q] = lw[q];
        __syncwarp();
        warp_softmax(w, N, lane);  // csmc.py:200
        warp_seq_cumsum(w, w, N, lane);
        const Key kT = split_key(key, (uint32_t)(K + 1), (uint32_t)K);  // keys[-1], csmc.py:201
        uint32_t x0 = 0u, x1 = 0u;
        threefry2x32(kT.k0, kT.k1, x0, x1);
        BT = choice_from_cum(w, N, bits_to_unit(x0));
      } else {
        // choice(key, n, ()) without p = randint(key, (), 0, n) with the UNSPLIT key (smc.py:109, as written upstream)
        Key k1, k2;
        split2(key, k1, k2);
        uint32_t h0 = 0u, h1 = 0u, l0 = 0u, l1 = 0u;
        threefry2x32(k1.k0, k1.k1, h0, h1);
        threefry2x32(k2.k0, k2.k1, l0, l1);
        const uint32_t span = (uint32_t)N;
        uint32_t mult = 65536u % span;
        mult = (mult * mult) % span;
        BT = (int)(((h0 % span) * mult + (l0 % span)) % span);
      }
      if (lane == 0) s_idx = min(BT, N - 1);
    }
    if (p.mode == 1) {
      Key a;
      split2(key, a, key_steps);  // key_last (unused upstream), key_smoother
    }
    __syncthreads();
    int Bt = s_idx;
    for (int i = tid; i < du; i += blockDim.x) {
      const float v = uss[((size_t)K * N + Bt) * du + i];
      x[i] = v;
      p.xs[((size_t)b * (K + 1) + K) * du + i] = v;
    }
    if (tid == 0 && p.bs) p.bs[b * (K + 1) + K] = Bt;
    __syncthreads();

    // ---- t = K - 1 .. 0; step q consumes keys[q] (csmc.py:217 / smc.py:110-111)
    for (int q = 0; q < K; ++q) {
      const int t = K - 1 - q;
      const float* MTk = p.MT + (size_t)t * D * D;
      const float* mk = p.m + (size_t)t * D;
      const float* vp = p.vs + (hb * (K + 1) + t) * dv;
      const float dt = p.dt[t], sd = p.sd[t];
      const float inv_s2 = 1.0f / (sd * sd);
      const float norm_u = p.lognorm[t] * ((float)du / (float)dv);  // du log(2 pi sd^2)
      for (int e = tid; e < du * du; e += blockDim.x) {
        const int j = e / du, i = e - j * du;
        Muu[e] = MTk[(size_t)j * D + i];
      }
      for (int i = tid; i < du; i += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < dv; ++j) acc = fmaf(MTk[(size_t)(du + j) * D + i], vp[j], acc);
        cvec[i] = acc + mk[i];
      }
      __syncthreads();
      // log p(x | u_t^n) for every stored particle: S lanes per particle (S = the power of two covering du, at most a
      // warp), lanes over the outputs, so that narrow states still fill the warp
      const float* ut = uss + (size_t)t * N * du;
      const int sl = lane & (S - 1), sg = lane / S, ppw = 32 / S;
      for (int n0 = warp * ppw; n0 < N; n0 += nwarps * ppw) {
        const int n = n0 + sg;
        float st = 0.f;
        if (n < N) {
          const float* parent = ut + (size_t)n * du;
          for (int i = sl; i < du; i += S) {
            float acc = 0.f;
            for (int j = 0; j < du; ++j) acc = fmaf(Muu[j * du + i], __ldg(parent + j), acc);
            const float mean = parent[i] + dt * (acc + cvec[i]);
            const float resid = x[i] - mean;
            st = fmaf(resid, resid, st);
          }
        }
        for (int o = S >> 1; o > 0; o >>= 1) st += __shfl_xor_sync(0xffffffffu, st, o);
        if (sl == 0 && n < N) w[n] = -0.5f * (st * inv_s2 + norm_u);
      }
      __syncthreads();
      if (warp == 0) {
        if (p.mode == 0) {
          float m = -INFINITY;
          for (int e = lane; e < N; e += 32) m = fmaxf(m, w[e]);
          m = warp_max(m);  // csmc.py:207
          const float* lw = p.log_wss + (hb * (K + 1) + t) * N;
          for (int e = lane; e < N; e += 32) w[e] = (w[e] - m) + lw[e];  // csmc.py:208
          __syncwarp();
        }
        warp_softmax(w, N, lane);  // csmc.py:209 / smc.py:103
        warp_seq_cumsum(w, w, N, lane);
        const Key kq = split_key(key_steps, (uint32_t)(p.mode == 0 ? K + 1 : K), (uint32_t)q);
        uint32_t x0 = 0u, x1 = 0u;
        threefry2x32(kq.k0, kq.k1, x0, x1);
        const int id = choice_from_cum(w, N, bits_to_unit(x0));  // csmc.py:210 / smc.py:104
        if (lane == 0) s_idx = min(id, N - 1);
      }
      __syncthreads();
      Bt = s_idx;
      for (int i = tid; i < du; i += blockDim.x) {
        const float v = ut[(size_t)Bt * du + i];
        x[i] = v;
        p.xs[((size_t)b * (K + 1) + t) * du + i] = v;
      }
      if (tid == 0 && p.bs) p.bs[b * (K + 1) + t] = Bt;
      __syncthreads();
    }
  }
}

}  // namespace fbs

using namespace fbs;

extern "C" {

int fbs_backward_sample_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, int mode, const uint32_t* keys,
                                   const float* vs, const float* uss, const float* log_wss, int shared_history, int64_t B,
                                   int64_t N, float* xs_star, int32_t* bs_star) {
  if (B == 0) return FBS_OK;
  FBS_REQUIRE(model && model->MT && model->m && model->dt && model->sd && model->lognorm, "backward_sample: null model");
  FBS_REQUIRE(mode == 0 || mode == 1, "backward_sample: mode must be 0 (csmc.py:167-227) or 1 (smc.py:91-112)");
  FBS_REQUIRE(keys && vs && uss && xs_star, "backward_sample: null argument");
  FBS_REQUIRE(mode == 1 || log_wss, "backward_sample: the CSMC backward sampling pass needs the forward log-weights");
  FBS_REQUIRE(N >= 1 && N < (1 << 24), "backward_sample: bad N");
  BackwardParams p{};
  p.K = model->K; p.du = model->du; p.dv = model->dv; p.N = (int)N; p.mode = mode;
  p.MT = model->MT; p.m = model->m; p.dt = model->dt; p.sd = model->sd; p.lognorm = model->lognorm;
  p.keys = keys; p.vs = vs; p.uss = uss; p.log_wss = log_wss; p.B = B; p.shared = shared_history ? 1 : 0; p.xs = xs_star; p.bs = bs_star;
  const size_t smem = ((size_t)p.du * p.du + 2 * (size_t)p.du + (size_t)N) * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("backward_sample: du=%d N=%lld needs %zu B of shared memory", p.du, (long long)N, smem);
    return FBS_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(backward_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("backward_sample: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, backward_sample_kernel, 256, smem);
  if (occ < 1) occ = 1;
  int64_t grid = B;
  const int64_t cap = (int64_t)sm_count() * occ;
  if (grid > cap) grid = cap;
  backward_sample_kernel<<<(int)grid, 256, smem, as_stream(s)>>>(p);
  return check_launch("backward_sample_kernel");
}

}  // extern "C"
