// Whole-sweep kernel, tiled fast path ("v2") for the affine-Gaussian model.
//
// Same algorithm and random streams as csmc_kernels.cu (csmc.py:80-164 / smc.py:115-158); what changes
// is where the bytes live and who computes what:
//   * the drift matrix of the step, M_k[:, :du] packed as [du][DP], is brought into shared memory by ONE
//     elected thread with TMA bulk copies (cp.async.bulk + mbarrier), issued one step ahead, so the GEMM
//     never waits on L2;
//   * the per-chain step vectors (M_k[:, du:] v_prev + m_k and the v residual base) are precomputed for all
//     K steps by a separate GEMM kernel into caller workspace -- they depend only on the inputs;
//   * every thread owns a fixed 8-output x 8-particle register tile for the whole sweep (4 u + 4 v outputs;
//     particles n..n+3 and n+N/2..n+N/2+3, the pairs that share threefry blocks), so there is no index
//     arithmetic in the loop, the transition noise is generated in registers by the thread that consumes
//     it, and both words of every threefry block are used;
//   * the new particles overwrite the old in place (gather through registers), so one particle buffer.
#include <stdlib.h>
#include "fbs_common.cuh"
#include "fbs_resample.cuh"
#include "fbs_sweep.cuh"

namespace fbs {

// Four threefry blocks -> 8 scaled normals: elements b .. b + 3 and b + hblk .. b + hblk + 3 of normal(key, (2 hblk,)).
struct V2Noise8 {
  float lo[4], hi[4];
};
static __device__ __noinline__ V2Noise8 v2_noise_task(uint32_t k0, uint32_t k1, uint32_t b, uint32_t hblk, float scale) {
  uint32_t x0[4] = {b, b + 1u, b + 2u, b + 3u};
  uint32_t x1[4] = {b + hblk, b + hblk + 1u, b + hblk + 2u, b + hblk + 3u};
  threefry2x32_x4(k0, k1, x0, x1);
  V2Noise8 r;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    r.lo[c] = scale * bits_to_normal(x0[c]);
    r.hi[c] = scale * bits_to_normal(x1[c]);
  }
  return r;
}


// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA bulk copy (global -> shared)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
struct V2Layout {
  int half, hp, RC, dup, dvp, DP, nti, rt_chain, ntr, ntiles;
  size_t P, MTs, part, lw, w, cum, idx, tmp, keys, scal, bar, total;  // offsets in floats
};

__host__ __device__ inline V2Layout make_v2_layout(int G, int N, int du, int dv) {
  V2Layout L;
  L.half = N / 2;
  L.hp = (L.half + 3) / 4 * 4;
  L.RC = G * 2 * L.hp;
  L.dup = (du + 3) / 4 * 4;
  L.dvp = (dv + 3) / 4 * 4;
  L.DP = L.dup + L.dvp;
  L.nti = (L.dup > L.dvp ? L.dup : L.dvp) / 4;
  L.rt_chain = L.hp / 4;
  L.ntr = G * L.rt_chain;
  L.ntiles = L.nti * L.ntr;
  size_t o = 0;
  auto take = [&](size_t nfloats) {
    size_t r = o;
    o += (nfloats + 31) / 32 * 32;  // 128-byte granules
    return r;
  };
  L.MTs = take((size_t)du * L.DP);
  L.P = take((size_t)du * L.RC);
  L.part = take((size_t)L.nti * L.RC);
  L.lw = take((size_t)G * N);
  L.w = take((size_t)G * N);
  L.cum = take((size_t)G * (N + 1));
  L.idx = take((size_t)G * N);
  L.tmp = take((size_t)G * (N + 1));
  L.keys = take((size_t)6 * G);
  L.scal = take((size_t)G);
  L.bar = take(4);
  L.total = o;
  return L;
}

size_t sweep_v2_workspace_bytes(int64_t B, int K, int du, int dv) {
  const int DP = (du + 3) / 4 * 4 + (dv + 3) / 4 * 4;
  return (size_t)B * (K + 1) * DP * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// step-vector precompute: ws[b][k][0:dup]   = m_k[:du] + M_k[:du, du:] v_prev            (drift offset of u)
//                         ws[b][k][dup:DP]  = (v - v_prev) - dt_k (m_k[du:] + M_k[du:, du:] v_prev)   (residual base)
// with (v, v_prev) = (vs[k+1], vs[k]) for k < K and the initial-weight slot k == K: coefficients of step 0,
// (v, v_prev) = (vs[0], vs[1])  (gibbs.py:136-137 as called from csmc.py:154).
// One CTA per (slot, 32 chains); thread i owns output i for the 32 chains.
// ------------------------------------------------------------------------------------------------
constexpr int CV_CH = 32;
__global__ void __launch_bounds__(256) stepvec_kernel(const SweepParams p, int dup, int DP) {
  extern __shared__ float sv[];  // [dv][CV_CH] v_prev ; [dv][CV_CH] v
  const int du = p.du, dv = p.dv, D = du + dv, K = p.K;
  const int slot = blockIdx.y;
  const int k = slot < K ? slot : 0;
  const int kv = slot < K ? slot + 1 : 0, kp = slot < K ? slot : 1;
  const int64_t b0 = (int64_t)blockIdx.x * CV_CH;
  const int nb = (int)min((int64_t)CV_CH, p.B - b0);
  float* vp = sv;
  float* vc = sv + (size_t)dv * CV_CH;
  for (int t = threadIdx.x; t < dv * CV_CH; t += blockDim.x) {
    const int c = t / dv, j = t - c * dv;
    float a = 0.f, b = 0.f;
    if (c < nb) {
      a = p.vs[((size_t)(b0 + c) * (K + 1) + kp) * dv + j];
      b = p.vs[((size_t)(b0 + c) * (K + 1) + kv) * dv + j];
    }
    vp[j * CV_CH + c] = a;
    vc[j * CV_CH + c] = b;
  }
  __syncthreads();
  const float* MTk = p.MT + (size_t)k * D * D + (size_t)du * D;  // rows du.. : inputs v
  const float dt = p.dt[k];
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float acc[CV_CH];
#pragma unroll
    for (int c = 0; c < CV_CH; ++c) acc[c] = 0.f;
    for (int j = 0; j < dv; ++j) {
      const float a = __ldg(MTk + (size_t)j * D + i);
      const float4* v4 = reinterpret_cast<const float4*>(vp + j * CV_CH);
#pragma unroll
      for (int c4 = 0; c4 < CV_CH / 4; ++c4) {
        const float4 x = v4[c4];
        acc[4 * c4 + 0] = fmaf(a, x.x, acc[4 * c4 + 0]);
        acc[4 * c4 + 1] = fmaf(a, x.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(a, x.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(a, x.w, acc[4 * c4 + 3]);
      }
    }
    const float mi = p.m[(size_t)k * D + i];
    const int col = i < du ? i : dup + (i - du);
#pragma unroll
    for (int c = 0; c < CV_CH; ++c) {
      if (c < nb) {
        float r = mi + acc[c];
        if (i >= du) r = (vc[(i - du) * CV_CH + c] - vp[(i - du) * CV_CH + c]) - dt * r;
        p.ws[((size_t)(b0 + c) * (K + 1) + slot) * DP + col] = r;
      }
    }
  }
  // zero the padding columns
  for (int t = threadIdx.x; t < nb * (DP - D); t += blockDim.x) {
    const int c = t / (DP - D), q = t - c * (DP - D);
    const int col = q < dup - du ? du + q : dup + dv + (q - (dup - du));
    p.ws[((size_t)(b0 + c) * (K + 1) + slot) * DP + col] = 0.f;
  }
}

// The same precompute as a register-tiled GEMM: per slot, ws[:, slot, :] = epilogue(vs[:, kp, :] (chains x dv) * Mv_k (dv x DP)).
// A CTA keeps the slot's matrix in shared memory (columns already in the workspace's padded order) and walks over tiles of
// 64 chains; a thread owns 8 chains x 8 outputs (two groups of 4 columns): four 16-byte shared loads per 64 FMAs.  Same summation order
// as stepvec_kernel (j ascending, then + m), hence the same bits.
constexpr int SV_CH = 64;
__global__ void __launch_bounds__(256, 2) stepvec_tiled_kernel(const SweepParams p, int dup, int DP, int DP8) {
  extern __shared__ __align__(16) float sv2[];
  const int du = p.du, dv = p.dv, D = du + dv, K = p.K;
  const int vst = dv | 1;                   // odd row stride of the chain-major v tile
  float* Ws = sv2;                          // [dv][DP8]
  float* Vt = sv2 + (size_t)dv * DP8;       // [SV_CH][vst]  v_prev of the tile's chains
  const int tid = threadIdx.x, ng = DP8 / 8;
  const int og = tid % ng, cg = tid / ng;   // output group (2 x 4 columns), chain group (8 chains)
  const bool worker = cg < SV_CH / 8;
  const int64_t ntiles = (p.B + SV_CH - 1) / SV_CH;
  // work items (slot, chain tile), slot major, split into contiguous runs over a persistent grid: the slot's matrix is
  // reloaded only when the slot changes
  const int64_t items = (int64_t)(K + 1) * ntiles;
  const int64_t per = (items + gridDim.x - 1) / gridDim.x;
  const int64_t it0 = blockIdx.x * per, it1 = it0 + per < items ? it0 + per : items;
  int cur_slot = -1;
  for (int64_t it = it0; it < it1; ++it) {
    const int slot = (int)(it / ntiles);
    const int64_t tile = it - (int64_t)slot * ntiles;
    const int k = slot < K ? slot : 0;
    const int kv = slot < K ? slot + 1 : 0, kp = slot < K ? slot : 1;
    const float dt = p.dt[k];
    const int64_t b0 = tile * SV_CH;
    const int nb = (int)min((int64_t)SV_CH, p.B - b0);
    __syncthreads();
    if (slot != cur_slot) {
      const float* MTk = p.MT + (size_t)k * D * D + (size_t)du * D;  // rows du.. : inputs v
      for (int j = tid / 32; j < dv; j += blockDim.x / 32)
        for (int col = tid % 32; col < DP8; col += 32) {
          float w = 0.f;
          if (col < du) w = __ldg(MTk + (size_t)j * D + col);
          else if (col >= dup && col - dup < dv) w = __ldg(MTk + (size_t)j * D + du + (col - dup));
          Ws[(size_t)j * DP8 + col] = w;
        }
      cur_slot = slot;
    }
    if ((dv & 3) == 0) {
      // 16-byte loads, all of a warp's rows in flight before the first shared store
      const int nw = blockDim.x / 32, q4 = dv / 4;
      for (int c0 = tid / 32; c0 < SV_CH; c0 += 4 * nw) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * nw;
          x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < nb && (tid % 32) < q4)
            x[u] = __ldg(reinterpret_cast<const float4*>(p.vs + ((size_t)(b0 + c) * (K + 1) + kp) * dv) + (tid % 32));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + u * nw;
          if (c < SV_CH) {
            for (int j4 = tid % 32; j4 < q4; j4 += 32) {
              float4 y = x[u];
              if (j4 != (tid % 32)) {
                y = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < nb) y = __ldg(reinterpret_cast<const float4*>(p.vs + ((size_t)(b0 + c) * (K + 1) + kp) * dv) + j4);
              }
              float* d = Vt + c * vst + 4 * j4;
              d[0] = y.x; d[1] = y.y; d[2] = y.z; d[3] = y.w;
            }
          }
        }
      }
    } else {
      for (int c = tid / 32; c < SV_CH; c += blockDim.x / 32) {
        const float* src = p.vs + ((size_t)(b0 + (c < nb ? c : 0)) * (K + 1) + kp) * dv;
        for (int j = tid % 32; j < dv; j += 32) Vt[c * vst + j] = c < nb ? src[j] : 0.f;
      }
    }
    __syncthreads();
    if (!worker) continue;
    float acc[8][8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[c][o] = 0.f;
    const float* vbase = Vt + (size_t)(8 * cg) * vst;
#pragma unroll 2
    for (int j = 0; j < dv; ++j) {
      const float4 wa = *reinterpret_cast<const float4*>(Ws + (size_t)j * DP8 + 4 * og);          // columns [4 og, 4 og + 4)
      const float4 wb = *reinterpret_cast<const float4*>(Ws + (size_t)j * DP8 + 4 * (og + ng));   // and [4 (og + ng), ..): 16-byte lane stride
      const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float vv = vbase[c * vst + j];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[c][o] = fmaf(ww[o], vv, acc[c][o]);
      }
    }
    // epilogue: u columns  m + acc;  v columns  (v - v_prev) - dt (m + acc);  padding columns 0
    if (((du | dv | dup) & 3) == 0) {
      // every 4-column group is entirely u, v or padding: the group's m once per tile, v as ONE 16-byte load per chain,
      // v_prev from the shared-memory tile
      float4 mg[2];
      int kind[2], q0[2];  // 0 padding, 1 u, 2 v; first column inside its part
#pragma unroll
      for (int hq = 0; hq < 2; ++hq) {
        const int col = hq == 0 ? 4 * og : 4 * (og + ng);
        kind[hq] = col < du ? 1 : ((col >= dup && col - dup < dv) ? 2 : 0);
        q0[hq] = kind[hq] == 2 ? col - dup : col;
        mg[hq] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kind[hq]) mg[hq] = __ldg(reinterpret_cast<const float4*>(p.m + (size_t)k * D + (kind[hq] == 2 ? du : 0) + q0[hq]));
      }
      float4 vcur[8];
      if (kind[0] == 2 || kind[1] == 2) {
        const int qv = kind[0] == 2 ? q0[0] : q0[1];  // (a thread has at most one v group when dup >= 4 ng ... both may be v)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int ch = 8 * cg + c;
          vcur[c] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ch < nb && !(kind[0] == 2 && kind[1] == 2))
            vcur[c] = __ldg(reinterpret_cast<const float4*>(p.vs + ((size_t)(b0 + ch) * (K + 1) + kv) * dv + qv));
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int ch = 8 * cg + c;
        if (ch >= nb) continue;
        float* dst = p.ws + ((size_t)(b0 + ch) * (K + 1) + slot) * DP;
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          const int col = hq == 0 ? 4 * og : 4 * (og + ng);
          if (col + 4 > DP) continue;
          const float a0 = acc[c][4 * hq + 0], a1 = acc[c][4 * hq + 1], a2 = acc[c][4 * hq + 2], a3 = acc[c][4 * hq + 3];
          float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kind[hq] == 1) {
            r4 = make_float4(mg[hq].x + a0, mg[hq].y + a1, mg[hq].z + a2, mg[hq].w + a3);
          } else if (kind[hq] == 2) {
            float4 vc4 = vcur[c];
            if (kind[0] == 2 && kind[1] == 2)
              vc4 = __ldg(reinterpret_cast<const float4*>(p.vs + ((size_t)(b0 + ch) * (K + 1) + kv) * dv + q0[hq]));
            const float* vp = Vt + (size_t)ch * vst + q0[hq];
            r4.x = (vc4.x - vp[0]) - dt * (mg[hq].x + a0);
            r4.y = (vc4.y - vp[1]) - dt * (mg[hq].y + a1);
            r4.z = (vc4.z - vp[2]) - dt * (mg[hq].z + a2);
            r4.w = (vc4.w - vp[3]) - dt * (mg[hq].w + a3);
          }
          *reinterpret_cast<float4*>(dst + col) = r4;
        }
      }
      continue;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int ch = 8 * cg + c;
      if (ch >= nb) continue;
      const float* vrow_p = p.vs + ((size_t)(b0 + ch) * (K + 1) + kp) * dv;
      const float* vrow_c = p.vs + ((size_t)(b0 + ch) * (K + 1) + kv) * dv;
      float r[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int col = o < 4 ? 4 * og + o : 4 * (og + ng) + (o - 4);
        float x = 0.f;
        if (col < du) {
          x = p.m[(size_t)k * D + col] + acc[c][o];
        } else if (col >= dup && col - dup < dv) {
          const int q = col - dup;
          x = (vrow_c[q] - vrow_p[q]) - dt * (p.m[(size_t)k * D + du + q] + acc[c][o]);
        }
        r[o] = x;
      }
      float* dst = p.ws + ((size_t)(b0 + ch) * (K + 1) + slot) * DP;
      if (4 * og + 4 <= DP) *reinterpret_cast<float4*>(dst + 4 * og) = make_float4(r[0], r[1], r[2], r[3]);
      if (4 * (og + ng) + 4 <= DP) *reinterpret_cast<float4*>(dst + 4 * (og + ng)) = make_float4(r[4], r[5], r[6], r[7]);
    }
  }
}

__device__ __forceinline__ float warp_sum_v2(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_normalise_v2(float* lw, int n, int lane) {
  float m = -INFINITY;
  for (int q = lane; q < n; q += 32) m = fmaxf(m, lw[q]);
  m = warp_max(m);
  if (!(fabsf(m) < INFINITY)) m = 0.f;
  float s = 0.f;
  for (int q = lane; q < n; q += 32) s += expf(lw[q] - m);
  s = warp_sum_v2(s);
  const float lse = logf(s) + m;
  for (int q = lane; q < n; q += 32) lw[q] -= lse;
  __syncwarp();
  return lse;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) sweep_v2_kernel(const SweepParams p) {
  extern __shared__ __align__(128) float sm[];
  const V2Layout L = make_v2_layout(p.G, p.N, p.du, p.dv);
  const int du = p.du, dv = p.dv, N = p.N, K = p.K, RC = L.RC, DP = L.DP, hp = L.hp, half = L.half, dup = L.dup;
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  float* P = sm + L.P;
  float* MTs = sm + L.MTs;
  float* part = sm + L.part;
  float* lw = sm + L.lw;
  float* w = sm + L.w;
  float* cum = sm + L.cum;
  int* idx = reinterpret_cast<int*>(sm + L.idx);
  int* tmp = reinterpret_cast<int*>(sm + L.tmp);
  Key* kbase = reinterpret_cast<Key*>(sm + L.keys);
  float* scal = sm + L.scal;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
  const float logN = logf((float)N);
  const uint32_t mt_bytes = (uint32_t)((size_t)du * DP * sizeof(float));

  // ---- fixed tile ownership: row tile varies fastest across lanes (conflict-free P accesses, broadcast M) ----
  const bool has_tile = tid < L.ntiles;
  const int trg = has_tile ? tid % L.ntr : 0;
  const int ti = has_tile ? tid / L.ntr : 0;
  const int g = trg / L.rt_chain;
  const int trl = trg - g * L.rt_chain;
  const int col_lo = g * 2 * hp + 4 * trl;  // particles n0 .. n0+3
  const int col_hi = col_lo + hp;           // particles half + n0 .. half + n0 + 3
  const int n0 = 4 * trl;
  const int uoff = min(4 * ti, dup - 4);
  const int voff = dup + min(4 * ti, L.dvp - 4);
  const bool u_tile = has_tile && 4 * ti < dup;
  const bool v_tile = has_tile && 4 * ti < L.dvp;
  const uint32_t nel = (uint32_t)N * du;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t mt_phase = 0;
  auto wait_mt = [&]() {  // every issue_mt() opens one mbarrier phase; waits consume them in order
    mbar_wait(bar, mt_phase);
    mt_phase ^= 1u;
  };

  auto issue_mt = [&](int k) {  // one elected thread: TMA bulk copies of MTp[k] into shared memory
    fence_proxy_async();
    mbar_arrive_expect_tx(bar, mt_bytes);
    const char* src = reinterpret_cast<const char*>(p.MTp + (size_t)k * du * DP);
    char* dst = reinterpret_cast<char*>(MTs);
    for (uint32_t off = 0; off < mt_bytes; off += 16384u) {
      const uint32_t n = min(16384u, mt_bytes - off);
      bulk_g2s(dst + off, src + off, n, bar);
    }
  };

  for (int64_t chain0 = (int64_t)blockIdx.x * p.G; chain0 < p.B; chain0 += (int64_t)gridDim.x * p.G) {
    const int nchains = (int)min((int64_t)p.G, p.B - chain0);
    const bool live = has_tile && g < nchains;
    const int64_t chain = chain0 + g;

    // =============================== initialisation ===============================
    if (tid == 0) issue_mt(0);
    for (int t = tid; t < du * RC; t += NT) P[t] = 0.f;
    if (tid < nchains) {
      Key key{p.keys[2 * (chain0 + tid)], p.keys[2 * (chain0 + tid) + 1]};
      if (p.mode == MODE_CSMC) {
        Key key_init, key_scan;
        split2(key, key_init, key_scan);  // csmc.py:150
        kbase[tid] = key_scan;
        kbase[2 * p.G + tid] = key_init;
      } else {
        kbase[tid] = key;
      }
      scal[tid] = 0.f;
    }
    __syncthreads();

    // own-element I/O helpers -------------------------------------------------------------------
    // particle index of tile row r (0..7): r < 4 -> n0 + r ; else half + n0 + r - 4 ; valid iff n0 + (r & 3) < half
    auto load_own = [&](const float* src /* [N][du] of this chain */) {
      if (!(live && u_tile)) return;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int nl = n0 + (r & 3);
        if (nl >= half) continue;
        const int n = r < 4 ? nl : half + nl;
        const int col = (r < 4 ? col_lo : col_hi) + (r & 3);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (4 * ti + q < du) P[(size_t)(4 * ti + q) * RC + col] = src[(size_t)n * du + 4 * ti + q];
      }
    };
    auto store_own = [&](float* dst /* [N][du] of this chain */) {
      if (!(live && u_tile)) return;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int nl = n0 + (r & 3);
        if (nl >= half) continue;
        const int n = r < 4 ? nl : half + nl;
        const int col = (r < 4 ? col_lo : col_hi) + (r & 3);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (4 * ti + q < du) dst[(size_t)n * du + 4 * ti + q] = P[(size_t)(4 * ti + q) * RC + col];
      }
    };

    // the tile GEMM + epilogue:  P <- P + dt (M_uu P + cu)   (in place, own elements);  part <- partial residual sums
    auto parents = [&](int k, int slot) {
      float cu[4] = {0.f, 0.f, 0.f, 0.f}, cv[4] = {0.f, 0.f, 0.f, 0.f};
      if (live) {  // issued before the GEMM, consumed after it: the L2 latency hides behind the FMAs
        const float* wsrow = p.ws + ((size_t)chain * (K + 1) + slot) * DP;
        const float4 a = __ldg(reinterpret_cast<const float4*>(wsrow + uoff));
        const float4 b = __ldg(reinterpret_cast<const float4*>(wsrow + voff));
        cu[0] = a.x; cu[1] = a.y; cu[2] = a.z; cu[3] = a.w;
        cv[0] = b.x; cv[1] = b.y; cv[2] = b.z; cv[3] = b.w;
      }
      const float dt = p.dt[k];
      float au[4][8], av[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 8; ++r) au[q][r] = av[q][r] = 0.f;
      if (has_tile) {
        const float* pl = P + col_lo;
        const float* ph = P + col_hi;
        const float* mu = MTs + uoff;
        const float* mv = MTs + voff;
#pragma unroll 2
        for (int j = 0; j < du; ++j) {
          const float4 xl = *reinterpret_cast<const float4*>(pl + (size_t)j * RC);
          const float4 xh = *reinterpret_cast<const float4*>(ph + (size_t)j * RC);
          const float4 a4 = *reinterpret_cast<const float4*>(mu + (size_t)j * DP);
          const float4 b4 = *reinterpret_cast<const float4*>(mv + (size_t)j * DP);
          const float x[8] = {xl.x, xl.y, xl.z, xl.w, xh.x, xh.y, xh.z, xh.w};
          const float a[4] = {a4.x, a4.y, a4.z, a4.w};
          const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              au[q][r] = fmaf(a[q], x[r], au[q][r]);
              av[q][r] = fmaf(b[q], x[r], av[q][r]);
            }
        }
      }
      __syncthreads();  // every read of P and MTs by the GEMM is done
      if (tid == 0 && k + 1 < K && slot < K) issue_mt(k + 1);  // next step's matrix lands during phases 2-3
      if (has_tile) {
        if (u_tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (4 * ti + q >= du) continue;
            float* row = P + (size_t)(4 * ti + q) * RC;
            float4 ol = *reinterpret_cast<float4*>(row + col_lo);
            float4 oh = *reinterpret_cast<float4*>(row + col_hi);
            ol.x += dt * (au[q][0] + cu[q]); ol.y += dt * (au[q][1] + cu[q]);
            ol.z += dt * (au[q][2] + cu[q]); ol.w += dt * (au[q][3] + cu[q]);
            oh.x += dt * (au[q][4] + cu[q]); oh.y += dt * (au[q][5] + cu[q]);
            oh.z += dt * (au[q][6] + cu[q]); oh.w += dt * (au[q][7] + cu[q]);
            *reinterpret_cast<float4*>(row + col_lo) = ol;
            *reinterpret_cast<float4*>(row + col_hi) = oh;
          }
        }
        float ss[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (v_tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (4 * ti + q >= dv) continue;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const float resid = cv[q] - dt * av[q][r];
              ss[r] = fmaf(resid, resid, ss[r]);
            }
          }
        }
        float* prow = part + (size_t)ti * RC;
        *reinterpret_cast<float4*>(prow + col_lo) = make_float4(ss[0], ss[1], ss[2], ss[3]);
        *reinterpret_cast<float4*>(prow + col_hi) = make_float4(ss[4], ss[5], ss[6], ss[7]);
      }
      __syncthreads();
    };

    // LW[n] of chain gg into dst[n] (one warp)
    auto reduce_lw = [&](int gg, int k, float* dst) {
      const float sd = p.sd[k];
      const float inv_s2 = 1.0f / (sd * sd), lognorm = p.lognorm[k];
      for (int n = lane; n < N; n += 32) {
        const int col = gg * 2 * hp + (n < half ? n : hp + (n - half));
        float s = 0.f;
        for (int t = 0; t < L.nti; ++t) s += part[(size_t)t * RC + col];
        dst[n] = -0.5f * (s * inv_s2 + lognorm);
      }
      __syncwarp();
    };

    // transition noise of the tile in registers: element (n, i), n < half shares its threefry block with (n + half, i)
    float nz[4][8];
    // (four threefry blocks in lockstep per call of a NON-inlined task: inlined 16 times the noise alone is ~40 KB of
    //  straight-line code and the step loop stalls on instruction fetch; out-of-range elements of the tile draw from
    //  counters nobody reads)
    auto make_noise = [&](Key ktr, float scale) {
      const uint32_t hblk = (uint32_t)half * (uint32_t)du;  // nel / 2: element e shares its block with e + hblk
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const V2Noise8 r8 = v2_noise_task(ktr.k0, ktr.k1, (uint32_t)(n0 + s) * du + 4u * ti, hblk, scale);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          nz[q][s] = r8.lo[q];
          nz[q][4 + s] = r8.hi[q];
        }
      }
    };

    if (p.mode == MODE_PMCMC) {
      load_own(p.u0s + (size_t)chain * N * du);
    } else if (p.init_mode == FBS_INIT_DEGENERATE) {  // gibbs.py:140-144
      if (live && u_tile) {
        const float* u0 = p.us_star + (size_t)chain * (K + 1) * du;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (4 * ti + q >= du) continue;
          const float x = u0[4 * ti + q];
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (n0 + (r & 3) < half) P[(size_t)(4 * ti + q) * RC + (r < 4 ? col_lo : col_hi) + (r & 3)] = x;
        }
      }
      for (int t = tid; t < nchains * N; t += NT) lw[t] = p.init_log_w;
    } else {  // gibbs.py:133-137
      if (live && u_tile) {
        make_noise(kbase[2 * p.G + g], 1.0f);
        const int b0 = clamp_index(p.bs_star[(size_t)chain * (K + 1)], N);
        const float* u0 = p.us_star + (size_t)chain * (K + 1) * du;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (4 * ti + q >= du) continue;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int nl = n0 + (r & 3);
            if (nl >= half) continue;
            const int n = r < 4 ? nl : half + nl;
            P[(size_t)(4 * ti + q) * RC + (r < 4 ? col_lo : col_hi) + (r & 3)] = (n == b0) ? u0[4 * ti + q] : nz[q][r];  // csmc.py:152
          }
        }
      }
    }
    __syncthreads();
    if (p.mode == MODE_CSMC) {
      if (p.uss) store_own(p.uss + (size_t)chain * (K + 1) * N * du);
      if (p.init_mode == FBS_INIT_NORMAL) {
        // initial weights: likelihood with the step-0 coefficients and (v, v_prev) = (vs[0], vs[1])  -> workspace slot K.
        // The in-place mean update of `parents` must not touch the particles here, so save and restore them.
        float keep[4][8];
        if (live && u_tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 8; ++r)
              keep[q][r] = P[(size_t)min(4 * ti + q, du - 1) * RC + (r < 4 ? col_lo : col_hi) + (r & 3)];
        }
        wait_mt();  // MT_0 landed (kept for step 0: slot K does not re-issue)
        parents(0, K);
        if (live && u_tile) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * ti + q < du)
#pragma unroll
              for (int r = 0; r < 8; ++r) P[(size_t)(4 * ti + q) * RC + (r < 4 ? col_lo : col_hi) + (r & 3)] = keep[q][r];
        }
        for (int gg = warp; gg < nchains; gg += nwarps) reduce_lw(gg, 0, lw + gg * N);
        __syncthreads();
      }
      for (int gg = warp; gg < nchains; gg += nwarps) warp_normalise_v2(lw + gg * N, N, lane);  // csmc.py:155
      __syncthreads();
      if (p.log_wss)
        for (int t = tid; t < nchains * N; t += NT) {
          const int gg = t / N, n = t - gg * N;
          p.log_wss[(size_t)(chain0 + gg) * (K + 1) * N + n] = lw[t];
        }
    }
    const bool mt0_consumed = (p.mode == MODE_CSMC && p.init_mode == FBS_INIT_NORMAL);

    // =============================== the K-step sweep ===============================
    for (int k = 0; k < K; ++k) {
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
      // 1. parents: wait for M_k (TMA), GEMM, in-place means, residual partial sums
      if (!(k == 0 && mt0_consumed)) wait_mt();
      parents(k, k);

      // 2. weights + ancestors (one warp per chain) overlapped with the noise generation of the other warps
      for (int gg = warp; gg < nchains; gg += nwarps) {
        float* lwg = lw + gg * N;
        float* wg = w + gg * N;
        float* cumg = cum + gg * (N + 1);
        int* idxg = idx + gg * N;
        int* tmpg = tmp + gg * (N + 1);
        const Key kres = kbase[p.G + gg];
        if (p.mode == MODE_CSMC) {
          for (int q = lane; q < N; q += 32) wg[q] = expf(lwg[q]);  // csmc.py:139
          __syncwarp();
          const int32_t* bs = p.bs_star + (size_t)(chain0 + gg) * (K + 1);
          const int bi = bs[k], bj = bs[k + 1];
          if (p.scheme == FBS_RESAMPLE_KILLING)
            warp_cond_killing(kres, wg, N, bi, bj, true, cumg, tmpg, idxg, lane);
          else
            warp_cond_multinomial(kres, wg, N, bi, bj, true, cumg, idxg, lane);
          reduce_lw(gg, k, wg);                                         // LW of every parent ...
          for (int q = lane; q < N; q += 32) lwg[q] = wg[idxg[q]];      // ... gathered: csmc.py:145
          __syncwarp();
          warp_normalise_v2(lwg, N, lane);                              // csmc.py:146
        } else {
          reduce_lw(gg, k, lwg);                                        // smc.py:144
          if (p.lw_hist)
            for (int q = lane; q < N; q += 32) p.lw_hist[((size_t)(chain0 + gg) * K + k) * N + q] = lwg[q];
          const float c = warp_normalise_v2(lwg, N, lane);              // smc.py:145,147
          if (lane == 0) scal[gg] = (scal[gg] - logN) + c;              // smc.py:146
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
      if (live && u_tile) make_noise(kbase[2 * p.G + g], p.sd[k]);
      __syncthreads();

      // 3. children: gather the parent means through registers, add the noise, write in place, pin the reference
      float val[4][8];
      if (live && u_tile) {
        int pc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int nl = n0 + (r & 3);
          const int n = r < 4 ? nl : half + nl;
          const int a = nl < half ? idx[g * N + n] : 0;
          pc[r] = g * 2 * hp + (a < half ? a : hp + (a - half));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* row = P + (size_t)min(4 * ti + q, du - 1) * RC;
#pragma unroll
          for (int r = 0; r < 8; ++r) val[q][r] = row[pc[r]] + nz[q][r];
        }
      }
      __syncthreads();
      if (live && u_tile) {
        int bj = -1;
        const float* ustar = nullptr;
        if (p.mode == MODE_CSMC) {
          bj = clamp_index(p.bs_star[(size_t)chain * (K + 1) + k + 1], N);
          ustar = p.us_star + ((size_t)chain * (K + 1) + k + 1) * du;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (4 * ti + q >= du) continue;
          float* row = P + (size_t)(4 * ti + q) * RC;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int nl = n0 + (r & 3);
            if (nl >= half) continue;
            const int n = r < 4 ? nl : half + nl;
            row[(r < 4 ? col_lo : col_hi) + (r & 3)] = (n == bj) ? ustar[4 * ti + q] : val[q][r];  // csmc.py:143
          }
        }
      }
      __syncthreads();

      // optional history
      if (p.mode == MODE_CSMC) {
        if (p.As)
          for (int t = tid; t < nchains * N; t += NT) {
            const int gg = t / N, n = t - gg * N;
            p.As[((size_t)(chain0 + gg) * K + k) * N + n] = idx[t];
          }
        if (p.log_wss)
          for (int t = tid; t < nchains * N; t += NT) {
            const int gg = t / N, n = t - gg * N;
            p.log_wss[((size_t)(chain0 + gg) * (K + 1) + k + 1) * N + n] = lw[t];
          }
        if (p.uss) store_own(p.uss + ((size_t)chain * (K + 1) + k + 1) * N * du);
      } else {
        if (p.inds)
          for (int t = tid; t < nchains * N; t += NT) {
            const int gg = t / N, n = t - gg * N;
            p.inds[((size_t)(chain0 + gg) * K + k) * N + n] = idx[t];
          }
        if (p.us_hist) store_own(p.us_hist + ((size_t)chain * K + k) * N * du);
      }
    }

    // =============================== final state ===============================
    if (p.mode == MODE_CSMC) {
      if (p.us_last) store_own(p.us_last + (size_t)chain * N * du);
      if (p.log_ws_last)
        for (int t = tid; t < nchains * N; t += NT) p.log_ws_last[(size_t)chain0 * N + t] = lw[t];
    } else {
      if (p.uT) store_own(p.uT + (size_t)chain * N * du);
      if (p.log_ell && tid < nchains) p.log_ell[chain0 + tid] = scal[tid];
    }
    __syncthreads();
  }
}

template <int MAXT>
static int launch_variant(cudaStream_t st, SweepParams& p, int grid, int threads, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(sweep_v2_kernel<MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("sweep_v2: cudaFuncSetAttribute(%zu B) failed: %s", smem, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  sweep_v2_kernel<MAXT><<<grid, threads, smem, st>>>(p);
  return check_launch("sweep_v2_kernel");
}

int launch_stepvec(void* stream, SweepParams& p) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int dup = (p.du + 3) / 4 * 4, DP = dup + (p.dv + 3) / 4 * 4;
  {
    const int DP8 = (DP + 7) / 8 * 8, ng = DP8 / 8;
    const size_t sm2 = ((size_t)p.dv * DP8 + (size_t)SV_CH * (p.dv | 1)) * sizeof(float);
    const bool old_impl = debug_opt(OPT_STEPVEC_IMPL) == 1;  // pins the thread-per-output kernel
    // (narrow systems keep the thread-per-output kernel: with fewer than 8 output groups most of a tiled CTA idles)
    if (ng >= 8 && ng * (SV_CH / 8) <= 256 && sm2 <= 110 * 1024 && p.B >= SV_CH && !old_impl) {
      cudaFuncSetAttribute(stepvec_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
      const int64_t items = (int64_t)(p.K + 1) * ((p.B + SV_CH - 1) / SV_CH);
      const int64_t slots = 2 * (int64_t)sm_count();  // two CTAs per SM, persistent
      int threads = ng * (SV_CH / 8);
      threads = (threads + 31) / 32 * 32;
      stepvec_tiled_kernel<<<(unsigned)(items < slots ? items : slots), threads, sm2, st>>>(p, dup, DP, DP8);
      return check_launch("stepvec_tiled_kernel");
    }
  }
  dim3 grid((unsigned)((p.B + CV_CH - 1) / CV_CH), (unsigned)(p.K + 1));
  const size_t sm = (size_t)2 * p.dv * CV_CH * sizeof(float);
  if (sm > 48 * 1024) cudaFuncSetAttribute(stepvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  // thread i owns output i: a CTA wider than D only idles (narrow systems: 32 or 64 threads, many CTAs per SM)
  int sv_threads = (p.du + p.dv + 31) / 32 * 32;
  if (sv_threads > 256) sv_threads = 256;
  stepvec_kernel<<<grid, sv_threads, sm, st>>>(p, dup, DP);
  return check_launch("stepvec_kernel");
}

int launch_sweep_v2(void* stream, SweepParams& p) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.MTp == nullptr || p.ws == nullptr) return -1;
  if (p.N < 2 || (p.N & 1)) return -1;  // the threefry pairing needs an even particle count
  const int gforce = debug_opt(OPT_SWEEP_G);
  // chains per CTA: as many as fit 384 threads (one 8x8 register tile each, <= 168 registers) and shared memory;
  // FBS_SWEEP_G overrides (up to 704 threads) for experiments
  const int max_tiles = gforce ? 704 : 384;
  int bestG = 0;
  V2Layout L{};
  for (int G = 1; G <= 64; ++G) {
    V2Layout c = make_v2_layout(G, p.N, p.du, p.dv);
    if (c.ntiles > max_tiles || c.total * sizeof(float) > 225 * 1024) break;
    if ((int64_t)G > p.B && bestG > 0) break;
    bestG = G;
    L = c;
    if (gforce && G == gforce) break;
  }
  if (bestG == 0) return -1;
  p.G = bestG;
  const size_t smem = L.total * sizeof(float);
  int threads = (L.ntiles + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  {
    const int rc = launch_stepvec(stream, p);
    if (rc) return rc;
  }
  p.dup = L.dup;
  p.dvp = L.dvp;
  int64_t groups = (p.B + p.G - 1) / p.G;
  int64_t grid = groups < sm_count() ? groups : sm_count();
  if (threads <= 384) return launch_variant<384>(st, p, (int)grid, threads, smem);
  return launch_variant<704>(st, p, (int)grid, threads, smem);
}

}  // namespace fbs
