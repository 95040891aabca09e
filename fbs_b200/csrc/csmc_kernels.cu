// Persistent whole-sweep kernels for the affine-Gaussian model: the K-step CSMC forward pass
// (fbs/samplers/csmc/csmc.py:80-164) and the pMCMC particle filter (fbs/samplers/smc.py:115-158).
//
// One CTA owns G chains for all K steps; the particle set lives in shared memory, transposed
// ([du][rows], rows = G*N) so the drift GEMM reads it with 128-bit loads.  Per step:
//   1. parents : for every current particle p,  mean[p] = u_p + dt (M_uu u_p + c_u)  and the
//                Gaussian log-likelihood LW[p] of v_k given parent p (one drift evaluation feeds
//                both, SURVEY.md finding 6c)
//   2. weights : CSMC  -> A = cond_resample(exp(lw_prev)); lw[n] = LW[A[n]], normalise
//                pMCMC -> lw = LW; log_ell += logsumexp(lw) - log N; inds = resample(exp(lw - c))
//   3. children: u'[n] = mean[A[n]] + sd * eps[n]   (in-kernel threefry normals), pin reference
// Nothing but the optional history touches HBM inside the loop.
#include <stdlib.h>
#include "fbs_common.cuh"
#include "fbs_resample.cuh"
#include "fbs_sweep.cuh"

namespace fbs {

// SweepParams, MODE_* : fbs_sweep.cuh

constexpr int TI = 8;  // outputs per thread tile
constexpr int TN = 4;  // particle rows per thread tile

struct SmemLayout {
  int Rp, D, tiles_i, tiles_r, tiles_v0;
  size_t P, Q, cvec, LW, lw, w, cum, part, idx, tmp, vbuf, keys, scal, total;
};

__host__ __device__ inline SmemLayout make_layout(int G, int N, int du, int dv) {
  SmemLayout L;
  const int R = G * N;
  L.D = du + dv;
  L.Rp = (R + TN - 1) / TN * TN;
  L.tiles_i = (L.D + TI - 1) / TI;
  L.tiles_r = L.Rp / TN;
  L.tiles_v0 = du / TI;  // first i-tile that can contain a v output
  size_t o = 0;
  auto take = [&](size_t nfloats) {
    size_t r = o;
    o += (nfloats + 3) / 4 * 4;
    return r;
  };
  L.P = take((size_t)du * L.Rp);
  L.Q = take((size_t)du * L.Rp);
  L.cvec = take((size_t)G * L.D);
  L.LW = take(L.Rp);
  L.lw = take(L.Rp);
  L.w = take(L.Rp);
  L.cum = take((size_t)L.Rp + G);
  L.part = take((size_t)(L.tiles_i - L.tiles_v0) * L.Rp);
  L.idx = take(L.Rp);
  L.tmp = take((size_t)L.Rp + G);
  L.vbuf = take((size_t)2 * G * dv);
  L.keys = take((size_t)6 * G);
  L.scal = take((size_t)4 * G);
  L.total = o;
  return L;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// lw[0..n) -= logsumexp(lw); returns logsumexp to all lanes.  One warp.
__device__ __forceinline__ float warp_normalise(float* lw, int n, int lane) {
  float m = -INFINITY;
  for (int q = lane; q < n; q += 32) m = fmaxf(m, lw[q]);
  m = warp_max(m);
  if (!(fabsf(m) < INFINITY)) m = 0.f;  // jax logsumexp: non-finite max -> 0
  float s = 0.f;
  for (int q = lane; q < n; q += 32) s += expf(lw[q] - m);
  s = warp_sum(s);
  const float lse = logf(s) + m;
  for (int q = lane; q < n; q += 32) lw[q] -= lse;
  __syncwarp();
  return lse;
}

// Phase 1.  For all parent rows: Q[i][r] = P[i][r] + dt (M_uu P[:, r] + cvec_u)_i   (i < du)
//                                LW[r]   = -0.5 (sum_i (rbase_i - dt (M_vu P[:, r])_i)^2 / sd^2 + lognorm)
// cvec[g][i<du] = m_k[i] + (M_uv v_prev)_i ;  cvec[g][du+i'] = v_i' - v_prev_i' - dt (m_k + M_vv v_prev)_{du+i'}
__device__ void parents_phase(const SweepParams& p, const SmemLayout& L, float* sm, int k, const float* v_of_g,
                              const float* vprev_of_g, int64_t chain0, int nchains) {
  const int du = p.du, dv = p.dv, D = L.D, N = p.N, Rp = L.Rp, R = nchains * N;
  const int tid = threadIdx.x, NT = blockDim.x;
  float* P = sm + L.P;
  float* Q = sm + L.Q;
  float* cvec = sm + L.cvec;
  float* part = sm + L.part;
  float* LW = sm + L.LW;
  float* vbuf = sm + L.vbuf;  // [0, G*dv): v_prev ; [G*dv, 2G*dv): v
  const float* MTk = p.MT + (size_t)k * D * D;
  const float* mk = p.m + (size_t)k * D;
  const float dt = p.dt[k], sd = p.sd[k], lognorm = p.lognorm[k];
  const float inv_s2 = 1.0f / (sd * sd);

  // stage v_prev, v of every chain
  for (int t = tid; t < nchains * dv; t += NT) {
    const int g = t / dv, q = t - g * dv;
    vbuf[t] = vprev_of_g[(size_t)g * (p.K + 1) * dv + q];
    vbuf[p.G * dv + t] = v_of_g[(size_t)g * (p.K + 1) * dv + q];
  }
  __syncthreads();
  // per-chain step vector: cvec[g][i] = m_k[i] + sum_j MT_k[du + j][i] * v_prev[g][j]
  for (int t = tid; t < nchains * D; t += NT) {
    const int g = t / D, i = t - g * D;
    const float* col = MTk + (size_t)du * D + i;
    const float* vp = vbuf + g * dv;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = 0;
    for (; j + 4 <= dv; j += 4) {
      a0 = fmaf(__ldg(col + (size_t)(j + 0) * D), vp[j + 0], a0);
      a1 = fmaf(__ldg(col + (size_t)(j + 1) * D), vp[j + 1], a1);
      a2 = fmaf(__ldg(col + (size_t)(j + 2) * D), vp[j + 2], a2);
      a3 = fmaf(__ldg(col + (size_t)(j + 3) * D), vp[j + 3], a3);
    }
    for (; j < dv; ++j) a0 = fmaf(__ldg(col + (size_t)j * D), vp[j], a0);
    float c = mk[i] + ((a0 + a1) + (a2 + a3));
    if (i >= du) {
      const int q = i - du;
      c = (vbuf[p.G * dv + g * dv + q] - vp[q]) - dt * c;
    }
    cvec[t] = c;
  }
  __syncthreads();

  // tiled GEMM over the u inputs
  const int ntiles = L.tiles_i * L.tiles_r;
  for (int tile = tid; tile < ntiles; tile += NT) {
    const int ti = tile % L.tiles_i, tr = tile / L.tiles_i;
    const int i0 = ti * TI, r0 = tr * TN;
    float acc[TI][TN];
#pragma unroll
    for (int a = 0; a < TI; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
    const bool full_i = (i0 + TI <= D) && ((D & 3) == 0);
    for (int j = 0; j < du; ++j) {
      const float4 pv = *reinterpret_cast<const float4*>(P + (size_t)j * Rp + r0);
      float a[TI];
      const float* row = MTk + (size_t)j * D + i0;
      if (full_i) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(row));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(row + 4));
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else {
#pragma unroll
        for (int q = 0; q < TI; ++q) a[q] = (i0 + q < D) ? __ldg(row + q) : 0.f;
      }
      const float b[TN] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
      for (int q = 0; q < TI; ++q)
#pragma unroll
        for (int s = 0; s < TN; ++s) acc[q][s] = fmaf(a[q], b[s], acc[q][s]);
    }
    float ss[TN] = {0.f, 0.f, 0.f, 0.f};
    bool has_v = false;
#pragma unroll
    for (int q = 0; q < TI; ++q) {
      const int i = i0 + q;
      if (i >= D) break;
#pragma unroll
      for (int s = 0; s < TN; ++s) {
        const int r = r0 + s;
        if (r >= R) continue;
        const int g = r / N;
        const float c = cvec[g * D + i];
        if (i < du) {
          Q[(size_t)i * Rp + r] = P[(size_t)i * Rp + r] + dt * (acc[q][s] + c);
        } else {
          const float resid = c - dt * acc[q][s];
          ss[s] = fmaf(resid, resid, ss[s]);
          has_v = true;
        }
      }
    }
    if (has_v || ti >= L.tiles_v0) {
#pragma unroll
      for (int s = 0; s < TN; ++s) part[(size_t)(ti - L.tiles_v0) * Rp + r0 + s] = ss[s];
    }
  }
  __syncthreads();
  for (int r = tid; r < R; r += NT) {
    float s = 0.f;
    for (int t = 0; t < L.tiles_i - L.tiles_v0; ++t) s += part[(size_t)t * Rp + r];
    LW[r] = -0.5f * (s * inv_s2 + lognorm);
  }
  __syncthreads();
}

// Phase 3.  P[i][gN+n] = Q[i][gN + A[n]] + sd_k * normal(key_tr[g], (N, du))[n][i]
__device__ void children_phase(const SweepParams& p, const SmemLayout& L, float* sm, int k, int nchains) {
  const int du = p.du, N = p.N, Rp = L.Rp;
  const int tid = threadIdx.x, NT = blockDim.x;
  float* P = sm + L.P;
  const float* Q = sm + L.Q;
  const int* idx = reinterpret_cast<const int*>(sm + L.idx);
  const Key* ktr = reinterpret_cast<const Key*>(sm + L.keys) + 2 * p.G;
  const float sd = p.sd[k];
  const uint32_t nel = (uint32_t)N * du;
  const uint32_t h = (nel + 1u) >> 1;
  if (p.mode == MODE_BOOTSTRAP) {
    // smc.py:63,72: us_new = mean(us) + sd eps (row n of the noise belongs to particle n), then us = us_new[inds]
    float* Qw = sm + L.Q;
    for (int t = tid; t < nchains * (int)h; t += NT) {
      const int g = t / (int)h;
      const uint32_t b = (uint32_t)(t - g * (int)h);
      uint32_t y0, y1;
      random_bits_block(ktr[g], nel, b, y0, y1);
      {
        const int n = b / du, i = b - n * du;
        Qw[(size_t)i * Rp + g * N + n] += sd * bits_to_normal(y0);
      }
      const uint32_t e = b + h;
      if (e < nel) {
        const int n = e / du, i = e - n * du;
        Qw[(size_t)i * Rp + g * N + n] += sd * bits_to_normal(y1);
      }
    }
    __syncthreads();
    for (int t = tid; t < nchains * N * du; t += NT) {
      const int i = t / (nchains * N), r = t - i * (nchains * N), g = r / N;
      P[(size_t)i * Rp + r] = Q[(size_t)i * Rp + g * N + idx[r]];
    }
    __syncthreads();
    return;
  }
  if ((N & 1) == 0) {
    // N even: element (n, i) pairs with (n + N/2, i); map threads with n fastest (bank-conflict free).
    const int hn = N >> 1;
    const int per_chain = hn * du;
    for (int t = tid; t < nchains * per_chain; t += NT) {
      const int g = t / per_chain, rem = t - g * per_chain;
      const int i = rem / hn, n = rem - i * hn;
      uint32_t y0, y1;
      random_bits_block(ktr[g], nel, (uint32_t)n * du + i, y0, y1);
      const int r0 = g * N + n, r1 = r0 + hn;
      P[(size_t)i * Rp + r0] = Q[(size_t)i * Rp + g * N + idx[r0]] + sd * bits_to_normal(y0);
      P[(size_t)i * Rp + r1] = Q[(size_t)i * Rp + g * N + idx[r1]] + sd * bits_to_normal(y1);
    }
  } else {
    for (int t = tid; t < nchains * (int)h; t += NT) {
      const int g = t / (int)h;
      const uint32_t b = (uint32_t)(t - g * (int)h);
      uint32_t y0, y1;
      random_bits_block(ktr[g], nel, b, y0, y1);
      {
        const int n = b / du, i = b - n * du, r = g * N + n;
        P[(size_t)i * Rp + r] = Q[(size_t)i * Rp + g * N + idx[r]] + sd * bits_to_normal(y0);
      }
      const uint32_t e = b + h;
      if (e < nel) {
        const int n = e / du, i = e - n * du, r = g * N + n;
        P[(size_t)i * Rp + r] = Q[(size_t)i * Rp + g * N + idx[r]] + sd * bits_to_normal(y1);
      }
    }
  }
  __syncthreads();
}

// Copy the transposed particle buffer of every chain to a [*, N, du] global array slice.
__device__ void store_particles(const float* P, int Rp, int N, int du, int nchains, float* dst, size_t chain_stride) {
  const int per = N * du;
  for (int t = threadIdx.x; t < nchains * per; t += blockDim.x) {
    const int g = t / per, e = t - g * per;
    const int n = e / du, i = e - n * du;
    dst[(size_t)g * chain_stride + e] = P[(size_t)i * Rp + g * N + n];
  }
}

__global__ void __launch_bounds__(1024, 1) sweep_affine_kernel(const SweepParams p) {
  extern __shared__ __align__(16) float sm[];
  const SmemLayout L = make_layout(p.G, p.N, p.du, p.dv);
  const int du = p.du, dv = p.dv, N = p.N, K = p.K, Rp = L.Rp;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  float* P = sm + L.P;
  float* LW = sm + L.LW;
  float* lw = sm + L.lw;
  float* w = sm + L.w;
  float* cum = sm + L.cum;
  int* idx = reinterpret_cast<int*>(sm + L.idx);
  int* tmp = reinterpret_cast<int*>(sm + L.tmp);
  Key* kbase = reinterpret_cast<Key*>(sm + L.keys);  // [0,G): sweep key, [G,2G): resampling, [2G,3G): transition
  float* scal = sm + L.scal;                         // [0,G): log_ell accumulator
  const float logN = logf((float)N);

  for (int64_t chain0 = (int64_t)blockIdx.x * p.G; chain0 < p.B; chain0 += (int64_t)gridDim.x * p.G) {
    const int nchains = (int)min((int64_t)p.G, p.B - chain0);
    const int R = nchains * N;
    const float* vs0 = p.vs + (size_t)chain0 * (K + 1) * dv;

    // ---------------- initialisation ----------------
    for (int t = tid; t < du * Rp; t += NT) P[t] = 0.f;
    if (tid < nchains) {
      Key key{p.keys[2 * (chain0 + tid)], p.keys[2 * (chain0 + tid) + 1]};
      if (p.mode == MODE_CSMC) {
        Key key_init, key_scan;
        split2(key, key_init, key_scan);  // csmc.py:150
        kbase[tid] = key_scan;
        kbase[2 * p.G + tid] = key_init;  // transition slot doubles as the init key
      } else {
        kbase[tid] = key;  // smc.py:154 splits the kernel key itself
      }
      scal[tid] = 0.f;
    }
    __syncthreads();
    if (p.mode != MODE_CSMC) {
      for (int t = tid; t < R * du; t += NT) {
        const int r = t / du, i = t - r * du;
        P[(size_t)i * Rp + r] = p.u0s[(size_t)chain0 * N * du + t];
      }
      __syncthreads();
    } else {
      if (p.init_mode == FBS_INIT_DEGENERATE) {  // gibbs.py:140-144
        for (int t = tid; t < R * du; t += NT) {
          const int r = t / du, i = t - r * du, g = r / N;
          P[(size_t)i * Rp + r] = p.us_star[(size_t)(chain0 + g) * (K + 1) * du + i];
        }
        for (int r = tid; r < R; r += NT) lw[r] = p.init_log_w;
        __syncthreads();
      } else {  // gibbs.py:133-137: N(0, I) draws, reference pinned, weights = likelihood(vs[0] | u0, vs[1], ts[0])
        const uint32_t nel = (uint32_t)N * du, h = (nel + 1u) >> 1;
        for (int t = tid; t < nchains * (int)h; t += NT) {
          const int g = t / (int)h;
          const uint32_t b = (uint32_t)(t - g * (int)h);
          uint32_t y0, y1;
          random_bits_block(kbase[2 * p.G + g], nel, b, y0, y1);
          {
            const int n = b / du, i = b - n * du;
            P[(size_t)i * Rp + g * N + n] = bits_to_normal(y0);
          }
          if (b + h < nel) {
            const int n = (b + h) / du, i = (b + h) - n * du;
            P[(size_t)i * Rp + g * N + n] = bits_to_normal(y1);
          }
        }
        __syncthreads();
        for (int t = tid; t < nchains * du; t += NT) {  // csmc.py:152
          const int g = t / du, i = t - g * du;
          const int b0 = clamp_index(p.bs_star[(size_t)(chain0 + g) * (K + 1)], N);
          P[(size_t)i * Rp + g * N + b0] = p.us_star[(size_t)(chain0 + g) * (K + 1) * du + i];
        }
        __syncthreads();
        // v = vs[0], v_prev = vs[1] (gibbs.py:136-137 as called from csmc.py:154), coefficients of step 0
        parents_phase(p, L, sm, 0, vs0, vs0 + dv, chain0, nchains);
        for (int r = tid; r < R; r += NT) lw[r] = LW[r];
        __syncthreads();
      }
      for (int g = warp; g < nchains; g += nwarps) warp_normalise(lw + g * N, N, lane);  // csmc.py:155
      __syncthreads();
      if (p.uss) store_particles(P, Rp, N, du, nchains, p.uss + (size_t)chain0 * (K + 1) * N * du, (size_t)(K + 1) * N * du);
      if (p.log_wss)
        for (int t = tid; t < R; t += NT) {
          const int g = t / N, n = t - g * N;
          p.log_wss[(size_t)(chain0 + g) * (K + 1) * N + n] = lw[t];
        }
    }

    // ---------------- the K-step sweep ----------------
    for (int k = 0; k < K; ++k) {
      // step keys (one thread per chain)
      if (tid < nchains) {
        const Key key_k = split_key(kbase[tid], (uint32_t)K, (uint32_t)k);  // csmc.py:157 / smc.py:154
        Key a, b;
        split2(key_k, a, b);
        if (p.mode == MODE_CSMC) {  // csmc.py:136: (key_resampling, key_transition)
          kbase[p.G + tid] = a;
          kbase[2 * p.G + tid] = b;
        } else {  // smc.py:142: (key_proposal, key_resampling)
          kbase[2 * p.G + tid] = a;
          kbase[p.G + tid] = b;
        }
      }
      // 1. per-parent mean and log-likelihood:  v = vs[k+1], v_prev = vs[k]
      parents_phase(p, L, sm, k, vs0 + (size_t)(k + 1) * dv, vs0 + (size_t)k * dv, chain0, nchains);

      // 2. weights + ancestors, one warp per chain
      for (int g = warp; g < nchains; g += nwarps) {
        float* lwg = lw + g * N;
        float* wg = w + g * N;
        float* cumg = cum + g * (N + 1);
        int* idxg = idx + g * N;
        int* tmpg = tmp + g * (N + 1);
        const Key kres = kbase[p.G + g];
        if (p.mode == MODE_CSMC) {
          for (int q = lane; q < N; q += 32) wg[q] = expf(lwg[q]);  // csmc.py:139
          __syncwarp();
          const int32_t* bs = p.bs_star + (size_t)(chain0 + g) * (K + 1);
          const int bi = bs[k], bj = bs[k + 1];
          if (p.scheme == FBS_RESAMPLE_KILLING)
            warp_cond_killing(kres, wg, N, bi, bj, true, cumg, tmpg, idxg, lane);
          else
            warp_cond_multinomial(kres, wg, N, bi, bj, true, cumg, idxg, lane);
          for (int q = lane; q < N; q += 32) lwg[q] = LW[g * N + idxg[q]];  // csmc.py:145 on the resampled parents
          __syncwarp();
          warp_normalise(lwg, N, lane);  // csmc.py:146
        } else {
          for (int q = lane; q < N; q += 32) lwg[q] = LW[g * N + q];  // smc.py:144
          __syncwarp();
          if (p.lw_hist)
            for (int q = lane; q < N; q += 32) p.lw_hist[((size_t)(chain0 + g) * K + k) * N + q] = lwg[q];
          const float c = warp_normalise(lwg, N, lane);  // smc.py:145,147
          if (lane == 0) scal[g] = p.mode == MODE_BOOTSTRAP ? scal[g] - (c - logN)   // smc.py:67 (negative log-likelihood)
                                                            : (scal[g] - logN) + c;  // smc.py:146
          for (int q = lane; q < N; q += 32) wg[q] = expf(lwg[q]);
          __syncwarp();
          if (p.scheme == FBS_RESAMPLE_KILLING)
            warp_cond_killing(kres, wg, N, 0, 0, false, cumg, tmpg, idxg, lane);
          else if (p.scheme == FBS_RESAMPLE_MULTINOMIAL)
            warp_sorted_multinomial(kres, wg, N, cumg, reinterpret_cast<float*>(tmpg), idxg, lane);
          else
            warp_systematic_or_stratified(kres, wg, N, p.scheme == FBS_RESAMPLE_SYSTEMATIC, true, cumg, idxg, lane);
        }
      }
      __syncthreads();

      // 3. children
      children_phase(p, L, sm, k, nchains);
      if (p.mode == MODE_CSMC) {
        for (int t = tid; t < nchains * du; t += NT) {  // csmc.py:143
          const int g = t / du, i = t - g * du;
          const int bj = clamp_index(p.bs_star[(size_t)(chain0 + g) * (K + 1) + k + 1], N);
          P[(size_t)i * Rp + g * N + bj] = p.us_star[((size_t)(chain0 + g) * (K + 1) + k + 1) * du + i];
        }
        __syncthreads();
        if (p.As)
          for (int t = tid; t < R; t += NT) {
            const int g = t / N, n = t - g * N;
            p.As[((size_t)(chain0 + g) * K + k) * N + n] = idx[t];
          }
        if (p.log_wss)
          for (int t = tid; t < R; t += NT) {
            const int g = t / N, n = t - g * N;
            p.log_wss[((size_t)(chain0 + g) * (K + 1) + k + 1) * N + n] = lw[t];
          }
        if (p.uss)
          store_particles(P, Rp, N, du, nchains, p.uss + ((size_t)chain0 * (K + 1) + k + 1) * N * du,
                          (size_t)(K + 1) * N * du);
      } else {
        if (p.inds)
          for (int t = tid; t < R; t += NT) {
            const int g = t / N, n = t - g * N;
            p.inds[((size_t)(chain0 + g) * K + k) * N + n] = idx[t];
          }
        if (p.us_hist)
          store_particles(P, Rp, N, du, nchains, p.us_hist + ((size_t)chain0 * K + k) * N * du, (size_t)K * N * du);
      }
    }

    // ---------------- final state ----------------
    if (p.mode == MODE_CSMC) {
      if (p.us_last) store_particles(P, Rp, N, du, nchains, p.us_last + (size_t)chain0 * N * du, (size_t)N * du);
      if (p.log_ws_last)
        for (int t = tid; t < R; t += NT) p.log_ws_last[(size_t)chain0 * N + t] = lw[t];
    } else {
      if (p.uT) store_particles(P, Rp, N, du, nchains, p.uT + (size_t)chain0 * N * du, (size_t)N * du);
      if (p.log_ell && tid < nchains) p.log_ell[chain0 + tid] = scal[tid];
    }
    __syncthreads();
  }
}

static int launch_sweep(fbs_stream_t s, SweepParams& p) {
  const int D = p.du + p.dv;
  {
    const int impl = p.mode == MODE_BOOTSTRAP ? 1 : debug_opt(OPT_SWEEP_IMPL);
    // preference: tcgen05 kernel (v3) -> tiled CUDA-core kernel (v2) -> general kernel (v1); each returns -1 when the
    // shape is not eligible.  FBS_SWEEP_IMPL = v1 | v2 | v3 pins the choice (tests / A-B measurements).
    const bool only1 = impl == 1, only2 = impl == 2;
    const bool verbose = debug_opt(OPT_SWEEP_VERBOSE) != 0;
    if ((impl == 0 || impl == 4) && p.mode != MODE_BOOTSTRAP) {  // narrow states: one warp per chain (sweep_warp.cu)
      const int rc = launch_sweep_warp(s, p);
      if (verbose) fprintf(stderr, "[fbs] sweep v4 (warp per chain) -> %d\n", rc);
      if (rc >= 0) return rc;
      if (impl == 4) {
        set_error("sweep: the warp-per-chain kernel was pinned (sweep_impl = 4) but the shape N=%lld du=%d dv=%d is not eligible",
                  (long long)p.N, p.du, p.dv);
        return FBS_ERR_UNSUPPORTED;
      }
    }
    if (!only1 && !only2) {
      const int rc = launch_sweep_v3(s, p);
      if (verbose) fprintf(stderr, "[fbs] sweep v3 (tcgen05) -> %d\n", rc);
      if (rc >= 0) return rc;
      if (impl == 3) {
        set_error("sweep: the tcgen05 kernel was pinned (sweep_impl = 3) but the shape N=%lld du=%d dv=%d is not eligible",
                  (long long)p.N, p.du, p.dv);
        return FBS_ERR_UNSUPPORTED;
      }
    }
    if (!only1) {
      const int rc = launch_sweep_v2(s, p);
      if (verbose) fprintf(stderr, "[fbs] sweep v2 (tiled) -> %d\n", rc);
      if (rc >= 0) return rc;
    }
    if (verbose) fprintf(stderr, "[fbs] sweep v1 (general)\n");
  }
  // chains per CTA: fill ~128 particle rows when N is small
  int G = 1;
  if (p.N < 64) G = (128 + p.N - 1) / p.N;
  if (G > 32) G = 32;
  if ((int64_t)G > p.B) G = (int)p.B;
  SmemLayout L;
  for (;; --G) {
    L = make_layout(G, p.N, p.du, p.dv);
    if (L.total * sizeof(float) <= 220 * 1024 || G == 1) break;
  }
  const size_t smem = L.total * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("sweep: N=%d du=%d dv=%d needs %zu B of shared memory per chain; use the per-step kernels", p.N, p.du,
              p.dv, smem);
    return FBS_ERR_UNSUPPORTED;
  }
  p.G = G;
  int ntiles = L.tiles_i * L.tiles_r;
  int threads = (ntiles + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 64) threads = 64;
  (void)D;
  cudaError_t e = cudaFuncSetAttribute(sweep_affine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("sweep: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_affine_kernel, threads, smem);
  if (occ < 1) occ = 1;
  int64_t groups = (p.B + G - 1) / G;
  int64_t grid = groups;
  const int64_t cap = (int64_t)sm_count() * occ;
  if (grid > cap) grid = cap;
  sweep_affine_kernel<<<(int)grid, threads, smem, as_stream(s)>>>(p);
  return check_launch("sweep_affine_kernel");
}

static int check_model(const fbs_affine_model_t* m) {
  FBS_REQUIRE(m, "model is null");
  FBS_REQUIRE(m->K >= 1 && m->du >= 1 && m->dv >= 1, "model: bad dims K=%d du=%d dv=%d", m->K, m->du, m->dv);
  FBS_REQUIRE(m->MT && m->m && m->dt && m->sd && m->lognorm, "model: null coefficient array");
  return FBS_OK;
}

}  // namespace fbs

using namespace fbs;

extern "C" {

int fbs_debug_umma_gemm(fbs_stream_t s, const float* A, const float* Bimg, int32_t K8, int32_t nout, float* D) {
  return launch_umma_selftest(s, A, Bimg, K8, nout, D);
}

size_t fbs_sweep_workspace_bytes(const fbs_affine_model_t* model, int64_t B) {
  if (!model || B <= 0) return 0;
  return sweep_v2_workspace_bytes(B, model->K, model->du, model->dv);
}

int fbs_csmc_forward_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, const uint32_t* keys,
                                const float* us_star, const int32_t* bs_star, const float* vs, int init_mode,
                                float init_log_w, int scheme, int64_t B, int64_t N, int32_t* As, float* log_wss,
                                float* uss, float* log_ws_last, float* us_last, void* workspace,
                                size_t workspace_bytes) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  int rc = check_model(model);
  if (rc) return rc;
  FBS_REQUIRE(keys && us_star && bs_star && vs, "csmc_forward: null input");
  FBS_REQUIRE(B >= 0 && N >= 1 && N < (1 << 20), "csmc_forward: bad sizes B=%lld N=%lld", (long long)B, (long long)N);
  FBS_REQUIRE(init_mode == FBS_INIT_DEGENERATE || init_mode == FBS_INIT_NORMAL, "csmc_forward: bad init_mode");
  FBS_REQUIRE(scheme == FBS_RESAMPLE_KILLING || scheme == FBS_RESAMPLE_MULTINOMIAL,
              "csmc_forward: conditional resampling scheme must be killing or multinomial (got %d)", scheme);
  if (B == 0) return FBS_OK;
  SweepParams p{};
  p.K = model->K; p.du = model->du; p.dv = model->dv;
  p.MT = model->MT; p.m = model->m; p.dt = model->dt; p.sd = model->sd; p.lognorm = model->lognorm;
  p.keys = keys; p.us_star = us_star; p.bs_star = bs_star; p.vs = vs;
  p.mode = MODE_CSMC; p.init_mode = init_mode; p.scheme = scheme; p.init_log_w = init_log_w;
  p.B = B; p.N = (int)N;
  p.As = As; p.log_wss = log_wss; p.uss = uss; p.log_ws_last = log_ws_last; p.us_last = us_last;
  if (model->MTp && workspace && workspace_bytes >= sweep_v2_workspace_bytes(B, p.K, p.du, p.dv)) {
    p.MTp = model->MTp;
    p.MTc = model->MTc;
    p.ws = static_cast<float*>(workspace);
  }
  return launch_sweep(s, p);
}

int fbs_bootstrap_filter_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, const uint32_t* step_keys,
                                    const float* vs, const float* u0s, int scheme, int64_t B, int64_t N, float* uT,
                                    float* log_nell, int32_t* inds, float* log_ws_hist, float* us_hist) {
  if (B == 0) return FBS_OK;
  int rc = check_model(model);
  if (rc) return rc;
  FBS_REQUIRE(step_keys && vs && u0s, "bootstrap_filter: null input");
  FBS_REQUIRE(N >= 1 && N < (1 << 20), "bootstrap_filter: bad sizes");
  FBS_REQUIRE(scheme >= FBS_RESAMPLE_MULTINOMIAL && scheme <= FBS_RESAMPLE_STRATIFIED, "bootstrap_filter: bad scheme %d",
              scheme);
  SweepParams p{};
  p.K = model->K; p.du = model->du; p.dv = model->dv;
  p.MT = model->MT; p.m = model->m; p.dt = model->dt; p.sd = model->sd; p.lognorm = model->lognorm;
  p.keys = step_keys; p.vs = vs; p.u0s = u0s;
  p.mode = MODE_BOOTSTRAP; p.scheme = scheme;
  p.B = B; p.N = (int)N;
  p.uT = uT; p.log_ell = log_nell; p.inds = inds; p.lw_hist = log_ws_hist; p.us_hist = us_hist;
  return launch_sweep(s, p);
}

int fbs_pmcmc_filter_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, const uint32_t* keys, const float* vs,
                                const float* u0s, int scheme, int64_t B, int64_t N, float* uT, float* log_ell,
                                int32_t* inds, float* log_ws_hist, float* us_hist, void* workspace,
                                size_t workspace_bytes) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  int rc = check_model(model);
  if (rc) return rc;
  FBS_REQUIRE(keys && vs && u0s, "pmcmc_filter: null input");
  FBS_REQUIRE(B >= 0 && N >= 1 && N < (1 << 20), "pmcmc_filter: bad sizes");
  FBS_REQUIRE(scheme >= FBS_RESAMPLE_MULTINOMIAL && scheme <= FBS_RESAMPLE_STRATIFIED, "pmcmc_filter: bad scheme %d",
              scheme);
  if (B == 0) return FBS_OK;
  SweepParams p{};
  p.K = model->K; p.du = model->du; p.dv = model->dv;
  p.MT = model->MT; p.m = model->m; p.dt = model->dt; p.sd = model->sd; p.lognorm = model->lognorm;
  p.keys = keys; p.vs = vs; p.u0s = u0s;
  p.mode = MODE_PMCMC; p.scheme = scheme;
  p.B = B; p.N = (int)N;
  p.uT = uT; p.log_ell = log_ell; p.inds = inds; p.lw_hist = log_ws_hist; p.us_hist = us_hist;
  if (model->MTp && workspace && workspace_bytes >= sweep_v2_workspace_bytes(B, p.K, p.du, p.dv)) {
    p.MTp = model->MTp;
    p.MTc = model->MTc;
    p.ws = static_cast<float*>(workspace);
  }
  return launch_sweep(s, p);
}

}  // extern "C"
