// Forward-noising paths and the small per-sweep pieces of gibbs_kernel / pmcmc_kernel.
// Reference: fbs/sdes/linear.py:190-225, fbs/sdes/simulators.py:53-106, fbs/samplers/gibbs.py:171-214,
// fbs/samplers/smc.py:161-168,244-258, experiments/toy/gp_gibbs.py:138-141.
#include <math.h>
#include <stdlib.h>
#include "fbs_common.cuh"
#include "fbs_resample.cuh"

namespace fbs {

// Store element d of the path at time index k, optionally reversed in time and split into (u, v).
__device__ __forceinline__ void store_path(float x, int64_t b, int k, int d, int K, int D, int du, int rev, float* out_u,
                                           float* out_v) {
  if (!rev) {
    out_u[(b * (K + 1) + k) * D + d] = x;
  } else {
    const int kr = K - k;
    if (d < du) {
      if (out_u) out_u[(b * (K + 1) + kr) * du + d] = x;
    } else if (out_v) {
      out_v[(b * (K + 1) + kr) * (D - du) + (d - du)] = x;
    }
  }
}

// simulate_cond_forward(keep_path=True): rnds = normal(key, (K, D)); x_{k+1} = F_k x_k + sqrtQ_k rnds[k].
// One thread per (chain, coordinate); the K steps are sequential.
__global__ void ou_forward_path_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ x0, int x0_batched,
                                       const float* __restrict__ F, const float* __restrict__ sqrtQ, int64_t B, int K,
                                       int D, int du, int rev, float* __restrict__ out_u, float* __restrict__ out_v) {
  const int64_t total = B * D;
  const uint32_t nel = (uint32_t)K * D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / D;
    const int d = (int)(t - b * D);
    const Key key{keys[2 * b], keys[2 * b + 1]};
    float x = x0[(x0_batched ? b * D : 0) + d];
    store_path(x, b, 0, d, K, D, du, rev, out_u, out_v);
    for (int k = 0; k < K; ++k) {
      const float eps = bits_to_normal(random_bits_elem(key, nel, (uint32_t)k * D + d));
      x = __fadd_rn(__fmul_rn(F[k], x), __fmul_rn(sqrtQ[k], eps));  // linear.py:216
      store_path(x, b, k + 1, d, K, D, du, rev, out_u, out_v);
    }
  }
}

// euler_maruyama with an affine drift.  One CTA per chain, state in shared memory, thread i owns
// coordinate i.  keys = split(key, K); rnds = normal(keys[k], (m, D)).
__global__ void em_affine_path_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ x0, int x0_batched,
                                      const float* __restrict__ AT, const float* __restrict__ a,
                                      const float* __restrict__ ddt, const float* __restrict__ disp, int64_t B, int K,
                                      int m, int D, int du, int rev, float* __restrict__ out_u,
                                      float* __restrict__ out_v) {
  extern __shared__ float xs[];  // [2][D]
  const int tid = threadIdx.x;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const Key key{keys[2 * b], keys[2 * b + 1]};
    float* cur = xs;
    float* nxt = xs + D;
    for (int d = tid; d < D; d += blockDim.x) {
      const float x = x0[(x0_batched ? b * D : 0) + d];
      cur[d] = x;
      store_path(x, b, 0, d, K, D, du, rev, out_u, out_v);
    }
    __syncthreads();
    const uint32_t nel = (uint32_t)m * D;
    for (int k = 0; k < K; ++k) {
      const Key key_k = split_key(key, (uint32_t)K, (uint32_t)k);  // simulators.py:81
      const float h = ddt[k];
      const float sq = sqrtf(h);
      for (int q = 0; q < m; ++q) {
        const int kq = k * m + q;
        const float* ATk = AT + (size_t)kq * D * D;
        const float g = disp[kq];
        for (int i = tid; i < D; i += blockDim.x) {
          float acc = 0.f;
          for (int j = 0; j < D; ++j) acc = fmaf(__ldg(ATk + (size_t)j * D + i), cur[j], acc);
          const float drift = acc + a[(size_t)kq * D + i];
          const float eps = bits_to_normal(random_bits_elem(key_k, nel, (uint32_t)q * D + i));
          nxt[i] = cur[i] + drift * h + g * sq * eps;  // simulators.py:87
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
      }
      for (int d = tid; d < D; d += blockDim.x) store_path(cur[d], b, k + 1, d, K, D, du, rev, out_u, out_v);
    }
    __syncthreads();
  }
}

// euler_maruyama with an affine drift, ONE THREAD PER CHAIN (small D, many chains -- the Gaussian Schroedinger-bridge
// forward sampler of experiments/sb/gibbs.py:141-143 batched over conditioning targets): the state lives in registers, the
// m sub-step matrices of an interval are staged in shared memory with cp.async one interval ahead (all threads read the
// same element: broadcast), and both outputs of every threefry block are used -- element e of normal(keys[k], (m, D)) is
// drawn together with element e + m D / 2, which a later sub-step reads back from a per-thread shared-memory stash.
template <int DT>
__global__ void __launch_bounds__(64) em_affine_path_tpc_kernel(
    const uint32_t* __restrict__ keys, const float* __restrict__ x0, int x0_batched, const float* __restrict__ AT,
    const float* __restrict__ a, const float* __restrict__ ddt, const float* __restrict__ disp, int64_t B, int K, int m,
    int D, int du, int rev, float* __restrict__ out_u, float* __restrict__ out_v) {
  extern __shared__ __align__(16) float sm[];
  const int nt = blockDim.x, tid = threadIdx.x;
  const uint32_t n = (uint32_t)m * D, h = (n + 1u) >> 1;
  const int per_k = m * D * D, per_ka = m * D, per_buf = (per_k + per_ka + 3) & ~3;
  float* mat = sm;                       // [2][per_buf]: AT of the m sub-steps, then a of the m sub-steps
  float* stash = sm + 2 * per_buf;       // [h][nt]: second outputs of the threefry blocks, read by a later sub-step
  float* epsb = stash + (size_t)h * nt;  // [D][nt]: this sub-step's normals
  const int64_t b = blockIdx.x * (int64_t)nt + tid;
  const bool active = b < B;
  const int64_t bb = active ? b : B - 1;  // idle threads of the last CTA shadow a real chain (no stores)
  const Key key{keys[2 * bb], keys[2 * bb + 1]};
  float x[DT];
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    x[i] = i < D ? x0[(x0_batched ? bb * D : 0) + i] : 0.f;
    if (active && i < D) store_path(x[i], b, 0, i, K, D, du, rev, out_u, out_v);
  }
  auto stage = [&](int k, int buf) {
    float* dst = mat + buf * per_buf;
    const float* srcA = AT + (size_t)k * per_k;
    const float* srca = a + (size_t)k * per_ka;
    for (int t = tid; t < per_k + per_ka; t += nt) {
      const float* src = t < per_k ? srcA + t : srca + (t - per_k);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + t)), "l"(src)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, 0);
  for (int k = 0; k < K; ++k) {
    if (k + 1 < K) {
      stage(k + 1, (k + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* buf = mat + (k & 1) * per_buf;
    const Key key_k = split_key(key, (uint32_t)K, (uint32_t)k);  // simulators.py:81
    const float hh = ddt[k];
    const float sq = sqrtf(hh);
    for (int q = 0; q < m; ++q) {
      const float* Aq = buf + q * D * D;
      const float* aq = buf + per_k + q * D;
      const float g = disp[k * m + q];
      float acc[DT];
#pragma unroll
      for (int i = 0; i < DT; ++i) acc[i] = 0.f;
      if ((D & 3) == 0) {
#pragma unroll
        for (int j = 0; j < DT; ++j) {
          if (j < D) {
            const float xj = x[j];
#pragma unroll
            for (int i = 0; i < DT; i += 4) {
              if (i < D) {
                const float4 w = *reinterpret_cast<const float4*>(Aq + j * D + i);
                acc[i] = fmaf(w.x, xj, acc[i]);
                if (i + 1 < DT) acc[i + 1] = fmaf(w.y, xj, acc[i + 1]);
                if (i + 2 < DT) acc[i + 2] = fmaf(w.z, xj, acc[i + 2]);
                if (i + 3 < DT) acc[i + 3] = fmaf(w.w, xj, acc[i + 3]);
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < DT; ++j) {
          if (j < D) {
            const float xj = x[j];
#pragma unroll
            for (int i = 0; i < DT; ++i)
              if (i < D) acc[i] = fmaf(Aq[j * D + i], xj, acc[i]);
          }
        }
      }
      // the sub-step's normals -> epsb (a ROLLED loop: unrolled over D the kernel outgrows the instruction cache, and with
      // one warp per scheduler every fetch miss is exposed); four threefry blocks in lockstep where the group allows it
#pragma unroll 1
      for (int i0 = 0; i0 < D; i0 += 4) {
        const uint32_t e0 = (uint32_t)q * D + i0;
        const int cnt = min(4, D - i0);
        if (cnt == 4 && e0 + 3u < h && e0 + 3u + h < n) {
          uint32_t y0[4] = {e0, e0 + 1u, e0 + 2u, e0 + 3u};
          uint32_t y1[4] = {e0 + h, e0 + h + 1u, e0 + h + 2u, e0 + h + 3u};
          threefry2x32_x4(key_k.k0, key_k.k1, y0, y1);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            epsb[(size_t)(i0 + r) * nt + tid] = bits_to_normal(y0[r]);
            stash[(size_t)(e0 + r) * nt + tid] = bits_to_normal(y1[r]);
          }
        } else {
          for (int r = 0; r < cnt; ++r) {
            const uint32_t e = e0 + r;
            float eps;
            if (e < h) {
              uint32_t y0, y1;
              random_bits_block(key_k, n, e, y0, y1);
              eps = bits_to_normal(y0);
              if (e + h < n) stash[(size_t)e * nt + tid] = bits_to_normal(y1);
            } else {
              eps = stash[(size_t)(e - h) * nt + tid];
            }
            epsb[(size_t)(i0 + r) * nt + tid] = eps;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < DT; ++i) {
        if (i < D) {
          const float drift = acc[i] + aq[i];
          x[i] = x[i] + drift * hh + g * sq * epsb[(size_t)i * nt + tid];  // simulators.py:87
        }
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < DT; ++i)
        if (i < D) store_path(x[i], b, k + 1, i, K, D, du, rev, out_u, out_v);
    }
    __syncthreads();  // buffer (k & 1) is refilled by the next iteration's stage(k + 2)
  }
}

// The same integrator with a chain split over T = 2 or 4 adjacent lanes (lane t owns coordinates [t DO, (t + 1) DO),
// DO = D / T <= 8): T times as many warps for the same number of chains -- what the serial sub-step chain needs to
// hide its latencies when there are only tens of thousands of chains.  The matvec exchanges the state by warp shuffles.
// DOC: the coordinates per lane D / T as a compile-time constant (0: run time, loops padded to 8) -- with it the 8-wide
// predicated loops of the matvec (64 FMA slots per exchanged lane for 25 products at D = 20, T = 4) shrink to the products.
template <int T, int DOC>
__global__ void __launch_bounds__(128) em_affine_path_split_kernel(
    const uint32_t* __restrict__ keys, const float* __restrict__ x0, int x0_batched, const float* __restrict__ AT,
    const float* __restrict__ a, const float* __restrict__ ddt, const float* __restrict__ disp, int64_t B, int K, int m,
    int D, int du, int rev, float* __restrict__ out_u, float* __restrict__ out_v) {
  constexpr int DOT = DOC ? DOC : 8;
  extern __shared__ __align__(16) float sm[];
  const int nt = blockDim.x, tid = threadIdx.x, lane = tid & 31;
  const int cpc = nt / T, cl = tid / T, t = tid % T, lbase = lane & ~(T - 1);
  const int DO = DOC ? DOC : D / T;
  const uint32_t n = (uint32_t)m * D, h = (n + 1u) >> 1;
  const int per_k = m * D * D, per_ka = m * D, per_buf = (per_k + per_ka + 3) & ~3;
  float* mat = sm;                   // [2][per_buf]
  float* stash = sm + 2 * per_buf;   // [h][cpc + 1]
  const int sst = cpc + 1;
  const int64_t b = blockIdx.x * (int64_t)cpc + cl;
  const bool active = b < B;
  const int64_t bb = active ? b : B - 1;
  const Key key{keys[2 * bb], keys[2 * bb + 1]};
  float x[DOT];
#pragma unroll
  for (int r = 0; r < DOT; ++r) {
    x[r] = r < DO ? x0[(x0_batched ? bb * D : 0) + t * DO + r] : 0.f;
    if (active && r < DO) store_path(x[r], b, 0, t * DO + r, K, D, du, rev, out_u, out_v);
  }
  auto stage = [&](int k, int buf) {
    float* dst = mat + buf * per_buf;
    const float* srcA = AT + (size_t)k * per_k;
    const float* srca = a + (size_t)k * per_ka;
    for (int i = tid; i < per_k + per_ka; i += nt) {
      const float* src = i < per_k ? srcA + i : srca + (i - per_k);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i)), "l"(src)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, 0);
  for (int k = 0; k < K; ++k) {
    if (k + 1 < K) {
      stage(k + 1, (k + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* buf = mat + (k & 1) * per_buf;
    const Key key_k = split_key(key, (uint32_t)K, (uint32_t)k);  // simulators.py:81
    const float hh = ddt[k];
    const float sq = sqrtf(hh);
    for (int q = 0; q < m; ++q) {
      const float* Aq = buf + q * D * D + t * DO;
      const float* aq = buf + per_k + q * D + t * DO;
      const float g = disp[k * m + q];
      float acc[DOT];
#pragma unroll
      for (int r = 0; r < DOT; ++r) acc[r] = 0.f;
#pragma unroll
      for (int tt = 0; tt < T; ++tt) {
#pragma unroll
        for (int r = 0; r < DOT; ++r) {
          const float xj = __shfl_sync(0xffffffffu, x[r], lbase + tt);  // coordinate j = tt DO + r of this chain
          if (r < DO) {
            const float* row = Aq + (tt * DO + r) * D;
#pragma unroll
            for (int ro = 0; ro < DOT; ++ro)
              if (ro < DO) acc[ro] = fmaf(row[ro], xj, acc[ro]);
          }
        }
      }
      // normals: element e = q D + i comes from threefry block e (first output) or block e - h (second output, stashed)
      float eps[DOT];
      const uint32_t e0 = (uint32_t)q * D + t * DO;
#pragma unroll
      for (int r = 0; r < DOT; ++r) {
        eps[r] = 0.f;
        if (r < DO && e0 + r < h) {
          uint32_t y0, y1;
          random_bits_block(key_k, n, e0 + r, y0, y1);
          eps[r] = bits_to_normal(y0);
          if (e0 + r + h < n) stash[(size_t)(e0 + r) * sst + cl] = bits_to_normal(y1);
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < DOT; ++r) {
        if (r < DO) {
          if (e0 + r >= h) eps[r] = stash[(size_t)(e0 + r - h) * sst + cl];
          const float drift = acc[r] + aq[r];
          x[r] = x[r] + drift * hh + g * sq * eps[r];  // simulators.py:87
        }
      }
    }
    if (active) {
#pragma unroll
      for (int r = 0; r < DOT; ++r)
        if (r < DO) store_path(x[r], b, k + 1, t * DO + r, K, D, du, rev, out_u, out_v);
    }
    __syncthreads();
  }
}

// force_move + x0 selection.  One warp per chain.
__global__ void force_move_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ log_ws, int is_log,
                                  const float* __restrict__ us_last, const int32_t* __restrict__ kk, int64_t B, int N,
                                  int du, int32_t* __restrict__ idx_out, float* __restrict__ alpha_out,
                                  float* __restrict__ x0_out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* w = smem + (size_t)warp * 2 * N;
  float* cum = w + N;
  for (int64_t b = blockIdx.x * (int64_t)nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    const Key key{keys[2 * b], keys[2 * b + 1]};
    Key key_1, key_2;
    split2(key, key_1, key_2);  // gibbs.py:197
    for (int q = lane; q < N; q += 32) w[q] = is_log ? expf(log_ws[b * N + q]) : log_ws[b * N + q];  // gibbs.py:152
    __syncwarp();
    const int k = kk[b];
    const float w_k = w[k];
    const float temp = 1.0f - w_k;
    const bool regular = w_k < 1.0f;  // threshold max(1 - exp(-M), 1 - 1e-12) == 1.0f in float32 (gibbs.py:203)
    const float unif = 1.0f / (float)N;
    float acc = 0.f, asum = 0.f;
    for (int q = lane; q < N; q += 32) {
      const float rest = regular ? ((q == k) ? 0.f : w[q]) / temp : unif;
      const float term = temp * rest / (1.0f - w[q]);
      if (term == term) asum += term;  // nansum, gibbs.py:211
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
    for (int q = lane; q < N; q += 32) cum[q] = regular ? __fdiv_rn((q == k) ? 0.f : w[q], temp) : unif;
    __syncwarp();
    acc = warp_seq_cumsum(cum, cum, N, lane);  // the contract's summation order (fbs_resample.cuh)
    uint32_t x0 = 0u, x1 = 0u;
    threefry2x32(key_1.k0, key_1.k1, x0, x1);
    int i = choice_from_cum(cum, N, bits_to_unit(x0));  // gibbs.py:207
    uint32_t z0 = 0u, z1 = 0u;
    threefry2x32(key_2.k0, key_2.k1, z0, z1);
    const float u = bits_to_unit(z0);  // gibbs.py:208
    const bool accept = __fmul_rn(u, __fsub_rn(1.0f, w[i < N ? i : N - 1])) < temp;  // gibbs.py:209
    i = accept ? i : k;
    if (lane == 0) {
      idx_out[b] = i;
      if (alpha_out) alpha_out[b] = fminf(fmaxf(asum, 0.f), 1.f);
    }
    if (x0_out)
      for (int d = lane; d < du; d += 32) x0_out[b * du + d] = us_last[(b * N + i) * du + d];
    __syncwarp();
  }
}

__global__ void pcn_combine_kernel(float beta, float s0, float omb, float s1, const float* __restrict__ x,
                                   const float* __restrict__ mean, const float* __restrict__ r0,
                                   const float* __restrict__ r1, int64_t B, int64_t n, float* __restrict__ out) {
  const int64_t total = B * n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const float mu = mean[t % n];
    const float p = x[t] + s0 * (r0[t] - mu);          // smc.py:167
    out[t] = beta * p + omb * mu + s1 * (r1[t] - mu);  // smc.py:168
  }
}

// One CTA per chain: MH accept + select of the (uT, log_ell, ys) state.
__global__ void mh_accept_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ prop_uTs,
                                 const float* __restrict__ prop_log_ell, const float* __restrict__ prop_ys, int64_t B,
                                 int N, int du, int64_t ny, int which_u, float* __restrict__ uT,
                                 float* __restrict__ log_ell, float* __restrict__ ys, float* __restrict__ acc_prob,
                                 uint8_t* __restrict__ is_acc) {
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const Key key{keys[2 * b], keys[2 * b + 1]};
    uint32_t x0 = 0u, x1 = 0u;
    threefry2x32(key.k0, key.k1, x0, x1);
    const float z = bits_to_unit(x0);                          // smc.py:248
    const float le = log_ell[b], ple = prop_log_ell[b];
    // smc.py:246: jnp.minimum propagates NaN (CUDA's fminf drops it), so a NaN evidence -- all weights -inf, or a
    // non-finite score -- rejects the proposal and the chain keeps its last finite state, as upstream
    const float dl = ple - le;
    const float log_acc = (dl == dl) ? fminf(0.f, dl) : dl;
    const bool acc = logf(z) < log_acc;                        // smc.py:249 (false when log_acc is NaN)
    __syncthreads();  // every thread has read log_ell[b] before thread 0 may overwrite it
    if (threadIdx.x == 0) {
      if (acc_prob) acc_prob[b] = expf(log_acc);
      if (is_acc) is_acc[b] = acc ? 1 : 0;
      if (acc) log_ell[b] = ple;
    }
    if (acc) {
      for (int d = threadIdx.x; d < du; d += blockDim.x) uT[b * du + d] = prop_uTs[((int64_t)b * N + which_u) * du + d];
      for (int64_t q = threadIdx.x; q < ny; q += blockDim.x) ys[b * ny + q] = prop_ys[b * ny + q];
    }
    __syncthreads();
  }
}

// ref_sampler: out[b, n, :] = a + Bm (yT[b] - c) + eps[b, n, :] @ L,   eps = normal(key_b, (N, du)).
__global__ void gaussian_ref_sample_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ yT,
                                           const float* __restrict__ a, const float* __restrict__ Bm,
                                           const float* __restrict__ c, const float* __restrict__ Lm, int64_t B, int N,
                                           int du, int dv, float* __restrict__ out) {
  extern __shared__ float smem[];
  float* mean = smem;              // [du]
  float* eps = smem + du;          // [N * du]
  const int tid = threadIdx.x, NT = blockDim.x;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const Key key{keys[2 * b], keys[2 * b + 1]};
    for (int i = tid; i < du; i += NT) {
      float acc = 0.f;
      for (int j = 0; j < dv; ++j) acc = fmaf(Bm[(size_t)i * dv + j], yT[b * dv + j] - c[j], acc);
      mean[i] = a[i] + acc;
    }
    const uint32_t nel = (uint32_t)N * du, h = (nel + 1u) >> 1;
    for (uint32_t blk = tid; blk < h; blk += NT) {
      uint32_t y0, y1;
      random_bits_block(key, nel, blk, y0, y1);
      eps[blk] = bits_to_normal(y0);
      if (blk + h < nel) eps[blk + h] = bits_to_normal(y1);
    }
    __syncthreads();
    for (int t = tid; t < N * du; t += NT) {
      const int n = t / du, i = t - n * du;
      float acc = 0.f;
      for (int j = 0; j < du; ++j) acc = fmaf(eps[n * du + j], __ldg(Lm + (size_t)j * du + i), acc);
      out[(b * N + n) * du + i] = mean[i] + acc;
    }
    __syncthreads();
  }
}


// The same sampler as a register-tiled GEMM (du % 4 == 0): persistent CTAs keep L in shared memory, a chain's N x du
// normals are drawn into shared memory (both outputs of every threefry block used) and every thread forms 4 x 4 blocks of
// eps @ L from one 16-byte and four broadcast shared loads per 16 FMAs.  Same summation order (j ascending, then + mean).
__global__ void __launch_bounds__(256, 2) gaussian_ref_sample_tiled_kernel(
    const uint32_t* __restrict__ keys, const float* __restrict__ yT, const float* __restrict__ a,
    const float* __restrict__ Bm, const float* __restrict__ c, const float* __restrict__ Lm, int64_t B, int N, int du, int dv,
    float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* Ls = smem;                              // [du][du]
  float* eps = Ls + (size_t)du * du;             // [N4][du], rows >= N zero
  float* mean = eps + (size_t)((N + 3) / 4 * 4) * du;  // [du]
  const int tid = threadIdx.x, NT = blockDim.x;
  const int n4 = (N + 3) / 4, i4 = du / 4;
  for (int t = tid; t < du * du; t += NT) Ls[t] = __ldg(Lm + t);
  for (int t = N * du + tid; t < n4 * 4 * du; t += NT) eps[t] = 0.f;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const Key key{keys[2 * b], keys[2 * b + 1]};
    __syncthreads();
    for (int i = tid; i < du; i += NT) {
      float acc = 0.f;
      for (int j = 0; j < dv; ++j) acc = fmaf(Bm[(size_t)i * dv + j], yT[b * dv + j] - c[j], acc);
      mean[i] = a[i] + acc;
    }
    const uint32_t nel = (uint32_t)N * du, h = (nel + 1u) >> 1;
    for (uint32_t blk = tid; blk < h; blk += NT) {
      uint32_t y0, y1;
      random_bits_block(key, nel, blk, y0, y1);
      eps[blk] = bits_to_normal(y0);
      if (blk + h < nel) eps[blk + h] = bits_to_normal(y1);
    }
    __syncthreads();
    for (int t = tid; t < n4 * i4; t += NT) {
      const int nb = t / i4, ib = t - nb * i4;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
      const float* e0 = eps + (size_t)(4 * nb) * du;
#pragma unroll 4
      for (int j = 0; j < du; ++j) {
        const float4 l4 = *reinterpret_cast<const float4*>(Ls + (size_t)j * du + 4 * ib);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float e = e0[r * du + j];
          acc[r][0] = fmaf(e, l4.x, acc[r][0]);
          acc[r][1] = fmaf(e, l4.y, acc[r][1]);
          acc[r][2] = fmaf(e, l4.z, acc[r][2]);
          acc[r][3] = fmaf(e, l4.w, acc[r][3]);
        }
      }
      const float4 m4 = *reinterpret_cast<const float4*>(mean + 4 * ib);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = 4 * nb + r;
        if (n < N)
          *reinterpret_cast<float4*>(out + (b * N + n) * du + 4 * ib) =
              make_float4(m4.x + acc[r][0], m4.y + acc[r][1], m4.z + acc[r][2], m4.w + acc[r][3]);
      }
    }
  }
}

// backward_scanning_pass (csmc.py:230-270): B_T ~ Cat(normalise(log_w_T)), then B_{t-1} = A_t[B_t].
// One warp per chain.
__global__ void backward_scan_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ As,
                                     const float* __restrict__ uss, const float* __restrict__ log_w_T, int64_t B, int K,
                                     int N, int du, float* __restrict__ xs_star, int32_t* __restrict__ bs_star) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* w = smem + (size_t)warp * N;
  for (int64_t b = blockIdx.x * (int64_t)nwarps + warp; b < B; b += (int64_t)gridDim.x * nwarps) {
    const float* lw = log_w_T + b * N;
    float m = -INFINITY;
    for (int q = lane; q < N; q += 32) m = fmaxf(m, lw[q]);
    m = warp_max(m);
    if (!(fabsf(m) < INFINITY)) m = 0.f;
    float sacc = 0.f;
    for (int q = lane; q < N; q += 32) sacc += expf(lw[q] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
    const float lse = logf(sacc) + m;
    for (int q = lane; q < N; q += 32) w[q] = expf(lw[q] - lse);  // normalise(log_w_T), csmc.py:257
    __syncwarp();
    warp_seq_cumsum(w, w, N, lane);
    const Key key{keys[2 * b], keys[2 * b + 1]};
    uint32_t x0 = 0u, x1 = 0u;
    threefry2x32(key.k0, key.k1, x0, x1);
    int Bt = choice_from_cum(w, N, bits_to_unit(x0));  // barker_move, csmc.py:295-297
    if (Bt >= N) Bt = N - 1;
    for (int t = K; t >= 0; --t) {
      if (lane == 0) bs_star[b * (K + 1) + t] = Bt;
      const float* row = uss + (((size_t)b * (K + 1) + t) * N + Bt) * du;
      for (int d = lane; d < du; d += 32) xs_star[((size_t)b * (K + 1) + t) * du + d] = row[d];
      if (t > 0) Bt = As[((size_t)b * K + (t - 1)) * N + Bt];  // csmc.py:262
    }
    __syncwarp();
  }
}

static int grid1d(int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace fbs

using namespace fbs;

extern "C" {

int fbs_ou_forward_path_f32(fbs_stream_t s, const uint32_t* keys, const float* x0, int x0_batched, const float* F,
                            const float* sqrtQ, int64_t B, int64_t K, int64_t D, int64_t du, int rev, float* out_u,
                            float* out_v) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(keys && x0 && F && sqrtQ, "ou_forward_path: null input");
  FBS_REQUIRE(B >= 0 && K >= 1 && D >= 1 && K * D < 0xFFFFFFFFll, "ou_forward_path: bad sizes");
  FBS_REQUIRE(rev ? (du >= 0 && du <= D && (out_u || out_v)) : (out_u != nullptr && du == D),
              "ou_forward_path: bad output configuration (rev=%d du=%lld D=%lld)", rev, (long long)du, (long long)D);
  if (B == 0) return FBS_OK;
  ou_forward_path_kernel<<<grid1d(B * D, 128), 128, 0, as_stream(s)>>>(keys, x0, x0_batched, F, sqrtQ, B, (int)K, (int)D,
                                                                      (int)du, rev, out_u, out_v);
  return check_launch("ou_forward_path_kernel");
}

int fbs_em_affine_path_f32(fbs_stream_t s, const uint32_t* keys, const float* x0, int x0_batched, const float* AT,
                           const float* a, const float* ddt, const float* disp, int64_t B, int64_t K, int64_t m,
                           int64_t D, int64_t du, int rev, float* out_u, float* out_v) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(keys && x0 && AT && a && ddt && disp, "em_affine_path: null input");
  FBS_REQUIRE(B >= 0 && K >= 1 && m >= 1 && D >= 1 && D <= 8192 && m * D < 0xFFFFFFFFll, "em_affine_path: bad sizes");
  FBS_REQUIRE(rev ? (du >= 0 && du <= D && (out_u || out_v)) : (out_u != nullptr && du == D),
              "em_affine_path: bad output configuration");
  if (B == 0) return FBS_OK;
  {
    // many chains of a small system: one thread per chain
    const int nt = 32;
    const size_t per_buf = ((size_t)(m * D * D + m * D) + 3) & ~(size_t)3;
    const size_t smem_tpc = (2 * per_buf + (size_t)((m * D + 1) / 2 + D) * nt) * sizeof(float);
    const int impl = debug_opt(OPT_EM_IMPL);  // 1 / 2 pin the CTA-per-chain / thread-per-chain kernels (tests)
    const bool pinned_cta = impl == 1;
    const bool pinned_tpc = impl == 2;
    const int T = (D % 4 == 0 && D <= 32) ? 4 : ((D % 2 == 0 && D <= 16) ? 2 : 0);
    if (T != 0 && B >= 1024 && !pinned_cta && !pinned_tpc) {
      // a chain over T lanes: 32 chains (T = 4) or 64 chains (T = 2) per 128-thread CTA
      const int cpc = 128 / T;
      const size_t smem_sp = (2 * per_buf + (size_t)((m * D + 1) / 2) * (cpc + 1)) * sizeof(float);
      if (smem_sp <= 100 * 1024) {
        const int grid_sp = (int)((B + cpc - 1) / cpc);
#define FBS_EM_SPLIT(TT, DD)                                                                                                   \
  do {                                                                                                                        \
    if (smem_sp > 48 * 1024)                                                                                                  \
      cudaFuncSetAttribute(em_affine_path_split_kernel<TT, DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sp);   \
    em_affine_path_split_kernel<TT, DD><<<grid_sp, 128, smem_sp, as_stream(s)>>>(keys, x0, x0_batched, AT, a, ddt, disp, B, (int)K, \
                                                                                 (int)m, (int)D, (int)du, rev, out_u, out_v); \
  } while (0)
        if (T == 4 && D == 20) FBS_EM_SPLIT(4, 5);        // the Gaussian Schroedinger bridge (sb/gibbs.py: d = 10, state 2 d)
        else if (T == 4) FBS_EM_SPLIT(4, 0);
        else FBS_EM_SPLIT(2, 0);
#undef FBS_EM_SPLIT
        return check_launch("em_affine_path_split_kernel");
      }
    }
    if (D <= 32 && B >= 1024 && smem_tpc <= 100 * 1024 && !pinned_cta) {
      const int grid_tpc = (int)((B + nt - 1) / nt);
#define FBS_EM_TPC(DT)                                                                                               \
  do {                                                                                                               \
    if (smem_tpc > 48 * 1024)                                                                                        \
      cudaFuncSetAttribute(em_affine_path_tpc_kernel<DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tpc); \
    em_affine_path_tpc_kernel<DT><<<grid_tpc, nt, smem_tpc, as_stream(s)>>>(keys, x0, x0_batched, AT, a, ddt, disp, B,  \
                                                                           (int)K, (int)m, (int)D, (int)du, rev, out_u, \
                                                                           out_v);                                   \
  } while (0)
      if (D <= 4) FBS_EM_TPC(4);
      else if (D <= 8) FBS_EM_TPC(8);
      else if (D <= 12) FBS_EM_TPC(12);
      else if (D <= 16) FBS_EM_TPC(16);
      else if (D <= 20) FBS_EM_TPC(20);
      else if (D <= 24) FBS_EM_TPC(24);
      else FBS_EM_TPC(32);
#undef FBS_EM_TPC
      return check_launch("em_affine_path_tpc_kernel");
    }
  }
  int threads = (int)((D + 31) / 32 * 32);
  if (threads > 256) threads = 256;
  const size_t smem = 2 * (size_t)D * sizeof(float);
  int64_t grid = B;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  em_affine_path_kernel<<<(int)grid, threads, smem, as_stream(s)>>>(keys, x0, x0_batched, AT, a, ddt, disp, B, (int)K,
                                                                   (int)m, (int)D, (int)du, rev, out_u, out_v);
  return check_launch("em_affine_path_kernel");
}

int fbs_force_move_f32(fbs_stream_t s, const uint32_t* keys, const float* log_ws_last, int weights_are_log,
                       const float* us_last, const int32_t* k, int64_t B, int64_t N, int64_t du, int32_t* idx,
                       float* alpha, float* x0) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(keys && log_ws_last && k && idx, "force_move: null input");
  FBS_REQUIRE(!x0 || us_last, "force_move: x0 requested without us_last");
  FBS_REQUIRE(B >= 0 && N >= 1 && du >= 0, "force_move: bad sizes");
  if (B == 0) return FBS_OK;
  const size_t per_warp = (size_t)2 * N * sizeof(float);
  if (per_warp > 200 * 1024) {
    set_error("force_move: N=%lld too large for the single-warp kernel", (long long)N);
    return FBS_ERR_UNSUPPORTED;
  }
  int warps = (int)(32 * 1024 / per_warp);
  warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
  const size_t smem = per_warp * warps;
  if (smem > 48 * 1024) cudaFuncSetAttribute(force_move_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t blocks = (B + warps - 1) / warps;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  force_move_kernel<<<(int)blocks, warps * 32, smem, as_stream(s)>>>(keys, log_ws_last, weights_are_log, us_last, k, B, (int)N, (int)du,
                                                                    idx, alpha, x0);
  return check_launch("force_move_kernel");
}

int fbs_pcn_combine_f32(fbs_stream_t s, double delta, const float* x, const float* mean, const float* r0,
                        const float* r1, int64_t B, int64_t n, float* out) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(x && mean && r0 && r1 && out, "pcn_combine: null pointer");
  FBS_REQUIRE(B >= 0 && n >= 1 && delta > 0, "pcn_combine: bad arguments");
  if (B == 0) return FBS_OK;
  const double beta = 2.0 / (2.0 + delta);  // smc.py:164 (python floats -> float32 weak-typed constants)
  pcn_combine_kernel<<<grid1d(B * n, 256), 256, 0, as_stream(s)>>>((float)beta, (float)sqrt(delta / 2.0),
                                                                  (float)(1.0 - beta), (float)sqrt(1.0 - beta), x, mean,
                                                                  r0, r1, B, n, out);
  return check_launch("pcn_combine_kernel");
}

int fbs_mh_accept_f32(fbs_stream_t s, const uint32_t* keys_mh, const float* prop_uTs, const float* prop_log_ell,
                      const float* prop_ys, int64_t B, int64_t N, int64_t du, int64_t ny, int32_t which_u, float* uT,
                      float* log_ell, float* ys, float* acceptance_prob, uint8_t* is_accepted) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(keys_mh && prop_uTs && prop_log_ell && prop_ys && uT && log_ell && ys, "mh_accept: null pointer");
  FBS_REQUIRE(B >= 0 && N >= 1 && which_u >= 0 && which_u < N, "mh_accept: bad sizes");
  if (B == 0) return FBS_OK;
  int64_t grid = B;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (grid > cap) grid = cap;
  mh_accept_kernel<<<(int)grid, 128, 0, as_stream(s)>>>(keys_mh, prop_uTs, prop_log_ell, prop_ys, B, (int)N, (int)du, ny,
                                                       which_u, uT, log_ell, ys, acceptance_prob, is_accepted);
  return check_launch("mh_accept_kernel");
}

int fbs_gaussian_ref_sample_f32(fbs_stream_t s, const uint32_t* keys, const float* yT, const float* a, const float* Bm,
                                const float* c, const float* L, int64_t B, int64_t N, int64_t du, int64_t dv,
                                float* out) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(keys && yT && a && Bm && c && L && out, "gaussian_ref_sample: null pointer");
  FBS_REQUIRE(B >= 0 && N >= 1 && du >= 1 && dv >= 1, "gaussian_ref_sample: bad sizes");
  if (B == 0) return FBS_OK;
  {
    const size_t smem_t = ((size_t)du * du + (size_t)((N + 3) / 4 * 4) * du + (size_t)du) * sizeof(float);
    if (du % 4 == 0 && smem_t <= 110 * 1024 && B >= 64) {
      cudaFuncSetAttribute(gaussian_ref_sample_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
      const int64_t slots = 2 * (int64_t)sm_count();
      gaussian_ref_sample_tiled_kernel<<<(int)(B < slots ? B : slots), 256, smem_t, as_stream(s)>>>(
          keys, yT, a, Bm, c, L, B, (int)N, (int)du, (int)dv, out);
      return check_launch("gaussian_ref_sample_tiled_kernel");
    }
  }
  const size_t smem = ((size_t)du + (size_t)N * du) * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("gaussian_ref_sample: N*du=%lld too large for shared memory", (long long)(N * du));
    return FBS_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(gaussian_ref_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t grid = B;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  gaussian_ref_sample_kernel<<<(int)grid, 256, smem, as_stream(s)>>>(keys, yT, a, Bm, c, L, B, (int)N, (int)du, (int)dv,
                                                                    out);
  return check_launch("gaussian_ref_sample_kernel");
}

int fbs_backward_scan_f32(fbs_stream_t s, const uint32_t* keys, const int32_t* As, const float* uss,
                          const float* log_w_T, int64_t B, int64_t K, int64_t N, int64_t du, float* xs_star,
                          int32_t* bs_star) {
  if (B == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(keys && As && uss && log_w_T && xs_star && bs_star, "backward_scan: null pointer");
  FBS_REQUIRE(B >= 0 && K >= 1 && N >= 1 && du >= 1, "backward_scan: bad sizes");
  if (B == 0) return FBS_OK;
  const size_t per_warp = (size_t)N * sizeof(float);
  if (per_warp > 200 * 1024) {
    set_error("backward_scan: N=%lld too large for the single-warp kernel", (long long)N);
    return FBS_ERR_UNSUPPORTED;
  }
  int warps = (int)(32 * 1024 / per_warp);
  warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
  const size_t smem = per_warp * warps;
  if (smem > 48 * 1024) cudaFuncSetAttribute(backward_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t blocks = (B + warps - 1) / warps;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  backward_scan_kernel<<<(int)blocks, warps * 32, smem, as_stream(s)>>>(keys, As, uss, log_w_T, B, (int)K, (int)N, (int)du,
                                                                       xs_star, bs_star);
  return check_launch("backward_scan_kernel");
}

}  // extern "C"
