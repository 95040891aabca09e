// Whole-sweep kernel for NARROW states ("v4"): ONE WARP PER CHAIN, particles in registers.
//
// The Gaussian Schroedinger-bridge configuration (experiments/sb/gibbs.py: d = 10, N <= 64, K = 100) is far too small for the
// CTA-per-chain-group kernels: per step a chain owns 64 x 10 values, the tiled kernel (sweep_v2.cu) spends its time in
// block barriers and instruction-cache misses (profiles/r1_v2_sb_sweep_full.md: issue 34 %, "no instruction" 1.15 stalls per
// issue, 12 warps per SM).  Here a chain never leaves its warp:
//   * particle n lives in lane n % 32, slot n / 32 (P = N / 32 slots, state u[P][DU] in registers);
//   * the step matrices of ALL K steps are staged once per CTA in shared memory ([K][du][DP], the MTp image) and read with
//     broadcast 16-byte loads -- a lane's drift is a du x (du + dv) register GEMM per slot;
//   * weights / conditional resampling run on a per-warp shared-memory scratch with the SAME device functions as the other
//     sweep kernels (fbs_resample.cuh: sequential cumulative sums, identical random streams) -> identical ancestors;
//   * the parents' transition means travel through a per-warp shared-memory tile; the transition noise of particle
//     (n, n + N/2) -- the two outputs of one threefry block -- belongs to the same lane (slots s and s + P/2);
//   * no block-wide barrier inside the sweep: 16 independent warps per SM hide each other's latencies.
// Reference: fbs/samplers/csmc/csmc.py:80-164, fbs/samplers/smc.py:115-158.
#include "fbs_common.cuh"
#include "fbs_resample.cuh"
#include "fbs_sweep.cuh"

namespace fbs {

namespace v4 {

constexpr int WARPS = 16;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// lw[0..n) -= logsumexp(lw); returns logsumexp (same operation order as warp_normalise of csmc_kernels.cu)
__device__ __forceinline__ float wnormalise(float* lw, int n, int lane) {
  float m = -INFINITY;
  for (int q = lane; q < n; q += 32) m = fmaxf(m, lw[q]);
  m = warp_max(m);
  if (!(fabsf(m) < INFINITY)) m = 0.f;
  float s = 0.f;
  for (int q = lane; q < n; q += 32) s += expf(lw[q] - m);
  s = wsum(s);
  const float lse = logf(s) + m;
  for (int q = lane; q < n; q += 32) lw[q] -= lse;
  __syncwarp();
  return lse;
}

// four threefry blocks b .. b + 3 of normal(key, (2 hblk,)): lo[c] = element b + c, hi[c] = element b + c + hblk
struct N8 {
  float lo[4], hi[4];
};
__device__ __noinline__ N8 noise4(uint32_t k0, uint32_t k1, uint32_t b, uint32_t hblk) {
  uint32_t x0[4] = {b, b + 1u, b + 2u, b + 3u};
  uint32_t x1[4] = {b + hblk, b + hblk + 1u, b + hblk + 2u, b + hblk + 3u};
  threefry2x32_x4(k0, k1, x0, x1);
  N8 r;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    r.lo[c] = bits_to_normal(x0[c]);
    r.hi[c] = bits_to_normal(x1[c]);
  }
  return r;
}

// DQ = padded state width (du, dv <= DQ, multiples of 4 in the MTp image), P = particle slots per lane (N = 32 P)
// MSM: the matrices of all K steps fit shared memory (else they are read through L1 / L2 with broadcast loads)
template <int DQ, int P, bool MSM>
__global__ void __launch_bounds__(32 * WARPS, 1) sweep_warp_kernel(const SweepParams p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int DP = 2 * DQ;   // row of the MTp image: [u outputs | v outputs]
  constexpr int N = 32 * P;
  const int du = p.du, dv = p.dv, K = p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // shared memory: Msm [K][du][DP] | per warp: mean [N][DQ] | cvs [DP] | lwraw [N] | lw [N] | w [N] | cum [N + 1] | idx [N] | tmp [N + 1]
  float* Msm = sm;
  const size_t per_warp = (size_t)N * DQ + DP + 3 * N + (N + 4) + N + (N + 4);
  float* wb = sm + (MSM ? (size_t)K * du * DP : 0) + (size_t)warp * per_warp;
  float* mean = wb;
  float* cvs = mean + (size_t)N * DQ;
  float* lwraw = cvs + DP;
  float* lw = lwraw + N;
  float* w = lw + N;
  float* cum = w + N;
  int* idx = reinterpret_cast<int*>(cum + N + 4);
  int* tmp = idx + N;
  if (MSM) {
    for (int t = tid; t < K * du * DP / 4; t += blockDim.x)
      reinterpret_cast<float4*>(Msm)[t] = __ldg(reinterpret_cast<const float4*>(p.MTp) + t);
    __syncthreads();
  }

  const float logN = logf((float)N);
  const uint32_t hblk = (uint32_t)(N / 2) * du;  // random_bits(key, N du): elements e and e + hblk share a block

  for (int64_t chain = (int64_t)blockIdx.x * WARPS + warp; chain < p.B; chain += (int64_t)gridDim.x * WARPS) {
    float u[P][DQ];
    Key kscan, kinit{0u, 0u};
    {
      const Key key{p.keys[2 * chain], p.keys[2 * chain + 1]};
      if (p.mode == MODE_CSMC) split2(key, kinit, kscan);  // csmc.py:150
      else kscan = key;                                    // smc.py:154 splits the kernel key itself
    }
    float log_ell = 0.f;
    const float* wschain = p.ws + (size_t)chain * (K + 1) * DP;

    // drift of this lane's particles for step matrix k and workspace slot `slot`: mean -> shared tile, log-likelihood -> lwraw
    auto drift_phase = [&](int k, int slot) {
      if (lane < DP / 4) reinterpret_cast<float4*>(cvs)[lane] = __ldg(reinterpret_cast<const float4*>(wschain + (size_t)slot * DP) + lane);
      __syncwarp();
      const float dt = __ldg(p.dt + k), sdk = __ldg(p.sd + k), lognorm = __ldg(p.lognorm + k);
      const float inv_var = 1.0f / (sdk * sdk);
      const float4* Mk = reinterpret_cast<const float4*>((MSM ? Msm : p.MTp) + (size_t)k * du * DP);
#pragma unroll 1
      for (int s = 0; s < P; ++s) {
        // the slot's state through compile-time register indices only (a runtime u[s] would force the array into local memory)
        float us[DQ];
#pragma unroll
        for (int q = 0; q < P; ++q)
          if (q == s) {
#pragma unroll
            for (int i = 0; i < DQ; ++i) us[i] = u[q][i];
          }
        float acc[DP];
#pragma unroll
        for (int c = 0; c < DP; ++c) acc[c] = 0.f;
#pragma unroll
        for (int j = 0; j < DQ; ++j) {
          if (j < du) {
            const float uj = us[j];
#pragma unroll
            for (int c4 = 0; c4 < DP / 4; ++c4) {
              const float4 m4 = MSM ? Mk[j * (DP / 4) + c4] : __ldg(Mk + j * (DP / 4) + c4);
              acc[4 * c4 + 0] = fmaf(m4.x, uj, acc[4 * c4 + 0]);
              acc[4 * c4 + 1] = fmaf(m4.y, uj, acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = fmaf(m4.z, uj, acc[4 * c4 + 2]);
              acc[4 * c4 + 3] = fmaf(m4.w, uj, acc[4 * c4 + 3]);
            }
          }
        }
        const int n = lane + 32 * s;
        float ss = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < DQ / 4; ++c4) {
          const float4 cu = reinterpret_cast<const float4*>(cvs)[c4];
          const float4 cv = reinterpret_cast<const float4*>(cvs + DQ)[c4];
          float4 mu;
          mu.x = us[4 * c4 + 0] + dt * (acc[4 * c4 + 0] + cu.x);
          mu.y = us[4 * c4 + 1] + dt * (acc[4 * c4 + 1] + cu.y);
          mu.z = us[4 * c4 + 2] + dt * (acc[4 * c4 + 2] + cu.z);
          mu.w = us[4 * c4 + 3] + dt * (acc[4 * c4 + 3] + cu.w);
          reinterpret_cast<float4*>(mean + (size_t)n * DQ)[c4] = mu;
          // padded v columns: cv = 0 and the matrix columns are 0 -> residual 0
          const float r0 = cv.x - dt * acc[DQ + 4 * c4 + 0], r1 = cv.y - dt * acc[DQ + 4 * c4 + 1];
          const float r2 = cv.z - dt * acc[DQ + 4 * c4 + 2], r3 = cv.w - dt * acc[DQ + 4 * c4 + 3];
          ss = fmaf(r0, r0, ss);
          ss = fmaf(r1, r1, ss);
          ss = fmaf(r2, r2, ss);
          ss = fmaf(r3, r3, ss);
        }
        lwraw[n] = -0.5f * (ss * inv_var + lognorm);
      }
      __syncwarp();
    };
    // u[s][:] = (use_parents ? mean[parent of n] : 0) + scale * normal(key, (N, du))[n][:]   for this lane's particles
    auto noise_phase = [&](Key ktr, float scale, bool use_parents) {
#pragma unroll 1
      for (int s = 0; s < P / 2; ++s) {
        const int n0 = lane + 32 * s, n1 = n0 + N / 2;
        const float* m0 = mean + (size_t)(use_parents ? idx[n0] : 0) * DQ;
        const float* m1 = mean + (size_t)(use_parents ? idx[n1] : 0) * DQ;
        float xa[DQ], xb[DQ];
#pragma unroll
        for (int c4 = 0; c4 < DQ / 4; ++c4) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          if (4 * c4 < du) {
            const N8 z = noise4(ktr.k0, ktr.k1, (uint32_t)n0 * du + 4u * c4, hblk);
            if (use_parents) {
              a = reinterpret_cast<const float4*>(m0)[c4];
              b = reinterpret_cast<const float4*>(m1)[c4];
            }
            a.x += scale * z.lo[0]; a.y += scale * z.lo[1]; a.z += scale * z.lo[2]; a.w += scale * z.lo[3];
            b.x += scale * z.hi[0]; b.y += scale * z.hi[1]; b.z += scale * z.hi[2]; b.w += scale * z.hi[3];
          }
          // columns >= du stay exactly zero (the tail of the last 4-block belongs to the next particle's row)
          xa[4 * c4 + 0] = 4 * c4 + 0 < du ? a.x : 0.f; xa[4 * c4 + 1] = 4 * c4 + 1 < du ? a.y : 0.f;
          xa[4 * c4 + 2] = 4 * c4 + 2 < du ? a.z : 0.f; xa[4 * c4 + 3] = 4 * c4 + 3 < du ? a.w : 0.f;
          xb[4 * c4 + 0] = 4 * c4 + 0 < du ? b.x : 0.f; xb[4 * c4 + 1] = 4 * c4 + 1 < du ? b.y : 0.f;
          xb[4 * c4 + 2] = 4 * c4 + 2 < du ? b.z : 0.f; xb[4 * c4 + 3] = 4 * c4 + 3 < du ? b.w : 0.f;
        }
#pragma unroll
        for (int q = 0; q < P / 2; ++q)
          if (q == s) {
#pragma unroll
            for (int i = 0; i < DQ; ++i) {
              u[q][i] = xa[i];
              u[q + P / 2][i] = xb[i];
            }
          }
      }
    };
    auto pin_particle = [&](int slot_row, const float* src) {  // u[slot_row] = src[0..du)  (csmc.py:143,152)
      if ((slot_row & 31) == lane) {
        const int s = slot_row >> 5;
#pragma unroll
        for (int q = 0; q < P; ++q)
          if (q == s) {
#pragma unroll
            for (int i = 0; i < DQ; ++i) u[q][i] = i < du ? __ldg(src + i) : 0.f;
          }
      }
    };
    auto store_particles = [&](float* dst) {  // [N][du] row-major
#pragma unroll
      for (int s = 0; s < P; ++s) {
        float* row = dst + (size_t)(lane + 32 * s) * du;
#pragma unroll
        for (int i = 0; i < DQ; ++i)
          if (i < du) row[i] = u[s][i];
      }
    };

    // =============================== initialisation ===============================
    if (p.mode == MODE_PMCMC) {
#pragma unroll
      for (int s = 0; s < P; ++s) {
        const float* row = p.u0s + ((size_t)chain * N + lane + 32 * s) * du;
#pragma unroll
        for (int i = 0; i < DQ; ++i) u[s][i] = i < du ? __ldg(row + i) : 0.f;
      }
    } else {
      const float* u0 = p.us_star + (size_t)chain * (K + 1) * du;
      if (p.init_mode == FBS_INIT_DEGENERATE) {  // gibbs.py:140-144
#pragma unroll
        for (int s = 0; s < P; ++s)
#pragma unroll
          for (int i = 0; i < DQ; ++i) u[s][i] = i < du ? __ldg(u0 + i) : 0.f;
        for (int q = lane; q < N; q += 32) lw[q] = p.init_log_w;
        __syncwarp();
      } else {  // gibbs.py:133-137: N(0, I) draws, reference pinned, weights = likelihood(vs[0] | u0, vs[1], ts[0])
        noise_phase(kinit, 1.0f, false);
        pin_particle(clamp_index(p.bs_star[(size_t)chain * (K + 1)], N), u0);  // csmc.py:152
        drift_phase(0, K);  // workspace slot K: (v, v_prev) = (vs[0], vs[1]) with the step-0 coefficients
        for (int q = lane; q < N; q += 32) lw[q] = lwraw[q];
        __syncwarp();
      }
      wnormalise(lw, N, lane);  // csmc.py:155
      if (p.uss) store_particles(p.uss + (size_t)chain * (K + 1) * N * du);
      if (p.log_wss)
        for (int q = lane; q < N; q += 32) p.log_wss[(size_t)chain * (K + 1) * N + q] = lw[q];
    }

    // =============================== the K-step sweep ===============================
    Key ka_l{0u, 0u}, kb_l{0u, 0u};  // the two keys of step (k & ~31) + lane, derived 32 steps at a time (one step per lane)
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      if ((k & 31) == 0) {
        const int kk = k + lane < K ? k + lane : K - 1;
        const Key key_k = split_key(kscan, (uint32_t)K, (uint32_t)kk);  // csmc.py:157 / smc.py:154
        split2(key_k, ka_l, kb_l);
      }
      Key ka, kb;
      ka.k0 = __shfl_sync(0xffffffffu, ka_l.k0, k & 31);
      ka.k1 = __shfl_sync(0xffffffffu, ka_l.k1, k & 31);
      kb.k0 = __shfl_sync(0xffffffffu, kb_l.k0, k & 31);
      kb.k1 = __shfl_sync(0xffffffffu, kb_l.k1, k & 31);
      const Key kres = p.mode == MODE_CSMC ? ka : kb;  // csmc.py:136 (resampling, transition); smc.py:142 (proposal, resampling)
      const Key ktr = p.mode == MODE_CSMC ? kb : ka;
      drift_phase(k, k);
      if (p.mode == MODE_CSMC) {
        for (int q = lane; q < N; q += 32) w[q] = expf(lw[q]);  // csmc.py:139
        __syncwarp();
        const int32_t* bs = p.bs_star + (size_t)chain * (K + 1);
        if (p.scheme == FBS_RESAMPLE_KILLING)
          warp_cond_killing(kres, w, N, bs[k], bs[k + 1], true, cum, tmp, idx, lane);
        else
          warp_cond_multinomial(kres, w, N, bs[k], bs[k + 1], true, cum, idx, lane);
        for (int q = lane; q < N; q += 32) lw[q] = lwraw[idx[q]];  // csmc.py:145 on the resampled parents
        __syncwarp();
        wnormalise(lw, N, lane);  // csmc.py:146
        if (p.As)
          for (int q = lane; q < N; q += 32) p.As[((size_t)chain * K + k) * N + q] = idx[q];
        if (p.log_wss)
          for (int q = lane; q < N; q += 32) p.log_wss[((size_t)chain * (K + 1) + k + 1) * N + q] = lw[q];
      } else {
        for (int q = lane; q < N; q += 32) lw[q] = lwraw[q];  // smc.py:144
        __syncwarp();
        if (p.lw_hist)
          for (int q = lane; q < N; q += 32) p.lw_hist[((size_t)chain * K + k) * N + q] = lw[q];
        const float c = wnormalise(lw, N, lane);  // smc.py:145,147
        log_ell = (log_ell - logN) + c;            // smc.py:146
        for (int q = lane; q < N; q += 32) w[q] = expf(lw[q]);
        __syncwarp();
        if (p.scheme == FBS_RESAMPLE_KILLING)
          warp_cond_killing(kres, w, N, 0, 0, false, cum, tmp, idx, lane);
        else if (p.scheme == FBS_RESAMPLE_MULTINOMIAL)
          warp_sorted_multinomial(kres, w, N, cum, reinterpret_cast<float*>(tmp), idx, lane);
        else
          warp_systematic_or_stratified(kres, w, N, p.scheme == FBS_RESAMPLE_SYSTEMATIC, true, cum, idx, lane);
        if (p.inds)
          for (int q = lane; q < N; q += 32) p.inds[((size_t)chain * K + k) * N + q] = idx[q];
      }
      __syncwarp();
      noise_phase(ktr, __ldg(p.sd + k), true);  // children = parent's mean + noise
      if (p.mode == MODE_CSMC) {
        pin_particle(clamp_index(p.bs_star[(size_t)chain * (K + 1) + k + 1], N),
                     p.us_star + ((size_t)chain * (K + 1) + k + 1) * du);  // csmc.py:143
        if (p.uss) store_particles(p.uss + ((size_t)chain * (K + 1) + k + 1) * N * du);
      } else if (p.us_hist) {
        store_particles(p.us_hist + ((size_t)chain * K + k) * N * du);
      }
      __syncwarp();  // every lane has read its parents' means before the next step overwrites the tile
    }

    // =============================== final state ===============================
    if (p.mode == MODE_CSMC) {
      if (p.us_last) store_particles(p.us_last + (size_t)chain * N * du);
      if (p.log_ws_last)
        for (int q = lane; q < N; q += 32) p.log_ws_last[(size_t)chain * N + q] = lw[q];
    } else {
      if (p.uT) store_particles(p.uT + (size_t)chain * N * du);
      if (p.log_ell && lane == 0) p.log_ell[chain] = log_ell;
    }
    __syncwarp();
  }
}

template <int DQ, int P, bool MSM>
static int launch_t(cudaStream_t st, const SweepParams& p, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(sweep_warp_kernel<DQ, P, MSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("sweep_warp: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  const int64_t groups = (p.B + WARPS - 1) / WARPS;
  const int grid = (int)(groups < sm_count() ? groups : sm_count());
  sweep_warp_kernel<DQ, P, MSM><<<grid, 32 * WARPS, smem, st>>>(p);
  return check_launch("sweep_warp_kernel");
}

}  // namespace v4

// Host: eligibility + launch.  Returns FBS_OK, an error, or -1 when the shape is not eligible.
int launch_sweep_warp(void* stream, SweepParams& p) {
  using namespace v4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.MTp == nullptr || p.ws == nullptr) return -1;
  if (p.N != 64 && p.N != 128) return -1;  // P = 2 or 4 slots per lane (the threefry pairing keeps (n, n + N/2) in one lane)
  const int dup = (p.du + 3) / 4 * 4, dvp = (p.dv + 3) / 4 * 4;
  if (dup != dvp || dup > 16) return -1;
  const int P = p.N / 32;
  if (P == 4 && dup > 12) return -1;  // register budget: u[P][DQ] + the accumulators must stay under 128 registers
  const size_t per_warp = (size_t)p.N * dup + 2 * dup + 3 * p.N + (p.N + 4) + p.N + (p.N + 4);
  const size_t msm_bytes = (size_t)p.K * p.du * 2 * dup * sizeof(float);
  const bool msm = msm_bytes + WARPS * per_warp * sizeof(float) <= 220 * 1024;
  const size_t smem = (msm ? msm_bytes : 0) + WARPS * per_warp * sizeof(float);
  if (smem > 220 * 1024) return -1;
  {
    const int rc = launch_stepvec(stream, p);
    if (rc) return rc;
  }
#define FBS_V4(DQ_, P_) \
  if (dup == DQ_ && P == P_) return msm ? launch_t<DQ_, P_, true>(st, p, smem) : launch_t<DQ_, P_, false>(st, p, smem);
  FBS_V4(4, 2) FBS_V4(8, 2) FBS_V4(12, 2) FBS_V4(16, 2) FBS_V4(4, 4) FBS_V4(8, 4) FBS_V4(12, 4)
#undef FBS_V4
  return -1;
}

}  // namespace fbs
