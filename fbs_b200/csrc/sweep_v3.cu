// Whole-sweep kernel on the 5th-generation tensor cores ("v3"): tcgen05.mma + TMEM + TMA.
//
// Per chain-step the joint drift of all particles is ONE GEMM  D[128 x Nout] = U[128 x du] * Mu_k^T  with
// Nout = du8 + dv8 outputs (u-drift | v-drift).  float32 parity is kept by a split-TF32 product
//   U M = Uhi Mhi + Ulo Mhi + Uhi Mlo        (hi = round-to-nearest tf32, lo = exact float32 remainder)
// accumulated in float32 in TENSOR MEMORY.  While the tensor core runs, the CUDA cores generate the step's
// transition noise (threefry2x32) in registers -- the two pipes overlap.
//
//   warp 0        : control.  One elected thread streams the packed (hi, lo) K-blocks of M_k through a ring of
//                   shared-memory stages with TMA bulk copies (cp.async.bulk + mbarrier), issues the
//                   tcgen05.mma's, frees stages with tcgen05.commit, and signals the accumulator barrier.
//   warps 1..16   : workers.  During the GEMM: noise.  After it: warps 1..8 read the accumulators from TMEM
//                   (tcgen05.ld, one particle row per thread: u-half / v-half), write the transition means and
//                   the per-row Gaussian log-likelihood; one warp resamples; all workers gather + add noise and
//                   write the new particles (hi / lo split) in the UMMA K-major core-matrix layout.
//
// Shared-memory operand layout (no swizzle, K-major "interleave"): element (row r, k) of an operand lives at
//   (k / 4) * LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 4) * 4  bytes,   LBO = (#rows / 8) * 128,
// i.e. 8-row x 16-byte core matrices; one MMA consumes K = 8 (two k-chunks, LBO apart), SBO = 128.
//
// Reference: fbs/samplers/csmc/csmc.py:80-164, fbs/samplers/smc.py:115-158 (same algorithm and random streams
// as csmc_kernels.cu / sweep_v2.cu).
#include <stdlib.h>
#include "fbs_common.cuh"
#include "fbs_resample.cuh"
#include "fbs_sweep.cuh"

namespace fbs {

namespace v3 {

constexpr int ROWS = 128;            // MMA M: particle rows per chain (N <= 128)
constexpr int NWORK = 16;            // worker warps
constexpr int NTHREADS = 32 * (1 + NWORK);
constexpr int STAGES = 6;            // ring of K-blocks of the step matrix
constexpr int GPC_MAX = 4;           // k-groups (of 4 columns) per worker in the children phase
constexpr uint32_t A_LBO = (ROWS / 8) * 128;  // 2048 bytes between k-chunks of the particle operand

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- tcgen05 --------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives TMEM lane (32 * (warp % 4) + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// byte offset of particle element (row r, column k) in the A operands
__device__ __forceinline__ uint32_t a_off(int r, int k) {
  return (uint32_t)(k >> 2) * A_LBO + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

struct Layout {
  int du8, dv8, nout, nkb, dup4;
  uint32_t b_lbo, blk_bytes, stage_bytes;  // one K-block image of M (hi or lo), one ring stage (hi + lo)
  uint32_t a_bytes;
  size_t Ahi, Alo, ring, cvs, lwraw, lw, w, cum, idx, tmp, keys, scal, bars, tmem, total;  // byte offsets
};

__host__ __device__ inline Layout make_layout(int N, int du, int dv) {
  Layout L;
  L.du8 = (du + 7) / 8 * 8;
  L.dv8 = (dv + 7) / 8 * 8;
  L.nout = L.du8 + L.dv8;
  if (L.nout % 16) L.nout += 8;
  L.nkb = L.du8 / 8;
  L.dup4 = (du + 3) / 4 * 4;
  L.b_lbo = (uint32_t)(L.nout / 8) * 128u;
  L.blk_bytes = 2u * L.b_lbo;
  L.stage_bytes = 2u * L.blk_bytes;
  L.a_bytes = (uint32_t)(L.du8 / 4) * A_LBO;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += (bytes + 127) / 128 * 128;
    return r;
  };
  L.Ahi = take(L.a_bytes);
  L.Alo = take(L.a_bytes);
  L.ring = take((size_t)STAGES * L.stage_bytes);
  L.cvs = take((size_t)(L.du8 + L.dv8) * 4);
  L.lwraw = take(ROWS * 4);
  L.lw = take(ROWS * 4);
  L.w = take(ROWS * 4);
  L.cum = take((ROWS + 1) * 4);
  L.idx = take(ROWS * 4);
  L.tmp = take((ROWS + 1) * 4);
  L.keys = take(64);
  L.scal = take(16);
  L.bars = take((2 * STAGES + 1) * 8);
  L.tmem = take(16);
  L.total = o;
  return L;
}

}  // namespace v3

using namespace v3;

__device__ __forceinline__ float warp_sum_v3(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_normalise_v3(float* lw, int n, int lane) {
  float m = -INFINITY;
  for (int q = lane; q < n; q += 32) m = fmaxf(m, lw[q]);
  m = warp_max(m);
  if (!(fabsf(m) < INFINITY)) m = 0.f;
  float s = 0.f;
  for (int q = lane; q < n; q += 32) s += expf(lw[q] - m);
  s = warp_sum_v3(s);
  const float lse = logf(s) + m;
  for (int q = lane; q < n; q += 32) lw[q] -= lse;
  __syncwarp();
  return lse;
}

__global__ void __launch_bounds__(v3::NTHREADS, 1) sweep_v3_kernel(const SweepParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const Layout L = make_layout(p.N, p.du, p.dv);
  const int du = p.du, dv = p.dv, N = p.N, K = p.K, half = N / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* Ahi = smem + L.Ahi;
  unsigned char* Alo = smem + L.Alo;  // doubles as the transition-mean buffer between the GEMM and the gather
  unsigned char* ring = smem + L.ring;
  float* cvs = reinterpret_cast<float*>(smem + L.cvs);  // [0, du8): cu, [du8, du8 + dv8): cv of this step
  float* lwraw = reinterpret_cast<float*>(smem + L.lwraw);
  float* lw = reinterpret_cast<float*>(smem + L.lw);
  float* w = reinterpret_cast<float*>(smem + L.w);
  float* cum = reinterpret_cast<float*>(smem + L.cum);
  int* idx = reinterpret_cast<int*>(smem + L.idx);
  int* tmp = reinterpret_cast<int*>(smem + L.tmp);
  Key* kbase = reinterpret_cast<Key*>(smem + L.keys);  // [0]: sweep key, [1]: resampling, [2]: transition
  float* scal = reinterpret_cast<float*>(smem + L.scal);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + STAGES;
  uint64_t* accum = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);
  const float logN = logf((float)N);
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(L.nout >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);

  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, 1);
      }
      mbar_init(accum, 1);
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- worker ownership for noise + children: row pair (n, n + half) x a chunk of k-groups ------------
  const int wt = tid - 32;  // worker thread index, < 0 for the control warp
  const int ngroups = L.dup4 / 4;
  const int per_pair = max(1, (NWORK * 32) / max(half, 1));
  const int gpc = (ngroups + per_pair - 1) / per_pair;  // <= GPC_MAX (checked on the host)
  const int nchunk = (ngroups + gpc - 1) / gpc;
  const int own_pair = wt >= 0 ? wt / nchunk : -1;
  const int own_chunk = wt >= 0 ? wt - own_pair * nchunk : 0;
  const bool owner = wt >= 0 && own_pair < half;
  const int g0 = own_chunk * gpc;  // first k-group
  const uint32_t nel = (uint32_t)N * du;

  // ring bookkeeping (control thread): global K-block counter over the whole launch
  uint32_t prod_blk = 0, cons_blk = 0;  // blocks issued to TMA / consumed by MMA
  uint32_t accum_phase = 0;

  for (int64_t chain = blockIdx.x; chain < p.B; chain += gridDim.x) {
    const float* vs0 = p.vs + (size_t)chain * (K + 1) * dv;
    (void)vs0;
    // total K-blocks this chain consumes: one GEMM for the initial weights (explicit_final) + K steps
    const bool init_gemm = (p.mode == MODE_CSMC && p.init_mode == FBS_INIT_NORMAL);
    const uint32_t chain_blocks = (uint32_t)(K + (init_gemm ? 1 : 0)) * L.nkb;
    uint32_t chain_prod = 0;  // blocks of this chain already requested
    auto block_src = [&](uint32_t b) {  // b-th block of this chain -> step matrix index, K-block
      uint32_t step = b / L.nkb;
      const uint32_t kb = b - step * L.nkb;
      if (init_gemm) step = step == 0 ? 0 : step - 1;  // initial weights use the step-0 matrix
      return reinterpret_cast<const unsigned char*>(p.MTc) + ((size_t)step * L.nkb + kb) * L.stage_bytes;
    };
    auto produce = [&]() {  // control thread: fill the next ring stage if this chain still has blocks to fetch
      if (chain_prod >= chain_blocks) return;
      const uint32_t s = prod_blk % STAGES;
      if (prod_blk >= STAGES) mbar_wait(empty + s, ((prod_blk / STAGES) - 1) & 1);
      mbar_expect_tx(full + s, L.stage_bytes);
      bulk_g2s(ring + (size_t)s * L.stage_bytes, block_src(chain_prod), L.stage_bytes, full + s);
      ++prod_blk;
      ++chain_prod;
    };

    // =============================== initialisation ===============================
    if (tid == 0) {
      fence_proxy_async();
      for (int s = 0; s < STAGES - 1; ++s) produce();
      Key key{p.keys[2 * chain], p.keys[2 * chain + 1]};
      if (p.mode == MODE_CSMC) {
        Key key_init, key_scan;
        split2(key, key_init, key_scan);  // csmc.py:150
        kbase[0] = key_scan;
        kbase[2] = key_init;
      } else {
        kbase[0] = key;
      }
      scal[0] = 0.f;
    }
    for (int t = tid; t < (int)(L.a_bytes / 4); t += NTHREADS) {
      reinterpret_cast<float*>(Ahi)[t] = 0.f;
      reinterpret_cast<float*>(Alo)[t] = 0.f;
    }
    __syncthreads();

    // write particle value x of (row, column) as the hi / lo pair
    auto put = [&](int r, int k, float x) {
      const float hi = tf32_rn(x);
      const uint32_t off = a_off(r, k);
      *reinterpret_cast<float*>(Ahi + off) = hi;
      *reinterpret_cast<float*>(Alo + off) = x - hi;
    };
    auto get = [&](int r, int k) {
      const uint32_t off = a_off(r, k);
      return *reinterpret_cast<const float*>(Ahi + off) + *reinterpret_cast<const float*>(Alo + off);
    };

    float nz[2][4 * GPC_MAX];  // noise of (n, n + half) x the owned columns
    auto make_noise = [&](Key ktr, float scale) {
#pragma unroll
      for (int q = 0; q < 4 * GPC_MAX; ++q) {
        const int k = 4 * g0 + q;
        uint32_t y0 = 0u, y1 = 0u;
        if (owner && q < 4 * gpc && k < du) random_bits_block(ktr, nel, (uint32_t)own_pair * du + k, y0, y1);
        nz[0][q] = scale * bits_to_normal(y0);
        nz[1][q] = scale * bits_to_normal(y1);
      }
    };

    if (p.mode == MODE_PMCMC) {
      const float* src = p.u0s + (size_t)chain * N * du;
      for (int t = tid; t < N * du; t += NTHREADS) put(t / du, t % du, src[t]);
    } else if (p.init_mode == FBS_INIT_DEGENERATE) {  // gibbs.py:140-144
      const float* u0 = p.us_star + (size_t)chain * (K + 1) * du;
      for (int t = tid; t < N * du; t += NTHREADS) put(t / du, t % du, u0[t % du]);
      for (int t = tid; t < N; t += NTHREADS) lw[t] = p.init_log_w;
    } else {  // gibbs.py:133-137
      make_noise(kbase[2], 1.0f);
      if (owner) {
        const int b0 = p.bs_star[(size_t)chain * (K + 1)];
        const float* u0 = p.us_star + (size_t)chain * (K + 1) * du;
#pragma unroll
        for (int q = 0; q < 4 * GPC_MAX; ++q) {
          const int k = 4 * g0 + q;
          if (q < 4 * gpc && k < du) {
            put(own_pair, k, own_pair == b0 ? u0[k] : nz[0][q]);                  // csmc.py:152
            put(own_pair + half, k, own_pair + half == b0 ? u0[k] : nz[1][q]);
          }
        }
      }
    }
    fence_proxy_async();
    __syncthreads();

    // history helper: particles of this chain -> dst [N][du]
    auto store_particles = [&](float* dst) {
      for (int t = tid; t < N * du; t += NTHREADS) dst[t] = get(t / du, t % du);
    };

    // ---- one GEMM + epilogue: Alo <- transition means, lwraw <- per-row Gaussian log-likelihood -------
    //      k: coefficient step, slot: workspace slot of the per-chain step vectors, restore: keep the particles
    auto gemm_and_epilogue = [&](int k, int slot, bool with_noise, bool restore) {
      if (warp == 0) {
        if (lane == 0) {
          tc_fence_after();
          for (int kb = 0; kb < L.nkb; ++kb) {
            produce();  // refill the stage freed by the PREVIOUS K-block's MMAs (keeps two batches in flight)
            const uint32_t s = cons_blk % STAGES;
            mbar_wait(full + s, (cons_blk / STAGES) & 1);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(Ahi) + (uint32_t)kb * 2u * A_LBO;
            const uint32_t a_lo = smem_u32(Alo) + (uint32_t)kb * 2u * A_LBO;
            const uint32_t b_hi = smem_u32(ring + (size_t)s * L.stage_bytes);
            const uint32_t b_lo = b_hi + L.blk_bytes;
            const uint64_t dAh = make_desc(a_hi, A_LBO, 128), dAl = make_desc(a_lo, A_LBO, 128);
            const uint64_t dBh = make_desc(b_hi, L.b_lbo, 128), dBl = make_desc(b_lo, L.b_lbo, 128);
            umma_tf32(tmem_base, dAh, dBh, idesc, kb > 0 ? 1u : 0u);
            umma_tf32(tmem_base, dAl, dBh, idesc, 1u);
            umma_tf32(tmem_base, dAh, dBl, idesc, 1u);
            umma_commit(empty + s);  // stage free once these MMAs have read it
            ++cons_blk;
          }
          umma_commit(accum);  // accumulator complete
        }
        __syncwarp();
      } else {
        // stage the per-chain step vectors, then noise while the tensor core works
        const float* wsrow = p.ws + ((size_t)chain * (K + 1) + slot) * (size_t)(L.dup4 + (dv + 3) / 4 * 4);
        for (int t = wt; t < L.du8 + L.dv8; t += NWORK * 32) {
          float x = 0.f;
          if (t < L.du8) {
            if (t < du) x = wsrow[t];
          } else if (t - L.du8 < dv) {
            x = wsrow[L.dup4 + (t - L.du8)];
          }
          cvs[t] = x;
        }
        if (with_noise) make_noise(kbase[2], p.sd[k]);
      }
      mbar_wait(accum, accum_phase);
      accum_phase ^= 1u;
      tc_fence_after();
      __syncthreads();  // cvs visible; every thread past the accumulator barrier
      if (warp >= 1 && warp <= 8) {
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int r = 32 * q + lane;
        const bool vhalf = warp > 4;
        const float dt = p.dt[k];
        const uint32_t trow = tmem_base + ((uint32_t)(32 * q) << 16);
        float acc[32];
        if (!vhalf) {
          for (int c0 = 0; c0 < L.du8; c0 += 32) {
            const int nc = min(32, L.du8 - c0);
            if (nc == 32) {
              tmem_ld32(trow + c0, acc);
            } else {
              for (int c = 0; c < nc; c += 8) tmem_ld8(trow + c0 + c, acc + c);
            }
            if (r < N) {
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                if (c < nc && c0 + c < du) {
                  const uint32_t off = a_off(r, c0 + c);
                  const float4 hi = *reinterpret_cast<const float4*>(Ahi + off);
                  float4 lo = *reinterpret_cast<const float4*>(Alo + off);
                  const float4 cu = *reinterpret_cast<const float4*>(cvs + c0 + c);
                  // mean = x + dt (drift + offset); when restoring, the particles stay untouched
                  if (!restore) {
                    lo.x = (hi.x + lo.x) + dt * (acc[c + 0] + cu.x);
                    lo.y = (hi.y + lo.y) + dt * (acc[c + 1] + cu.y);
                    lo.z = (hi.z + lo.z) + dt * (acc[c + 2] + cu.z);
                    lo.w = (hi.w + lo.w) + dt * (acc[c + 3] + cu.w);
                    *reinterpret_cast<float4*>(Alo + off) = lo;
                  }
                }
              }
            }
          }
        } else {
          float ss = 0.f;
          for (int c0 = 0; c0 < L.dv8; c0 += 32) {
            const int nc = min(32, L.dv8 - c0);
            if (nc == 32) {
              tmem_ld32(trow + L.du8 + c0, acc);
            } else {
              for (int c = 0; c < nc; c += 8) tmem_ld8(trow + L.du8 + c0 + c, acc + c);
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              if (c < nc && c0 + c < dv) {
                const float resid = cvs[L.du8 + c0 + c] - dt * acc[c];
                ss = fmaf(resid, resid, ss);
              }
            }
          }
          const float sd = p.sd[k];
          if (r < N) lwraw[r] = -0.5f * (ss / (sd * sd) + p.lognorm[k]);
        }
        tc_fence_before();
      }
      __syncthreads();
    };

    if (p.mode == MODE_CSMC) {
      if (p.uss) store_particles(p.uss + (size_t)chain * (K + 1) * N * du);
      if (p.init_mode == FBS_INIT_NORMAL) {
        gemm_and_epilogue(0, K, false, true);  // gibbs.py:136-137: (v, v_prev) = (vs[0], vs[1]) -> workspace slot K
        for (int t = tid; t < N; t += NTHREADS) lw[t] = lwraw[t];
        __syncthreads();
      }
      if (warp == 1) warp_normalise_v3(lw, N, lane);  // csmc.py:155
      __syncthreads();
      if (p.log_wss)
        for (int t = tid; t < N; t += NTHREADS) p.log_wss[(size_t)chain * (K + 1) * N + t] = lw[t];
    }

    // =============================== the K-step sweep ===============================
    for (int k = 0; k < K; ++k) {
      if (tid == 32) {
        const Key key_k = split_key(kbase[0], (uint32_t)K, (uint32_t)k);  // csmc.py:157 / smc.py:154
        Key a, b;
        split2(key_k, a, b);
        if (p.mode == MODE_CSMC) {  // csmc.py:136: (key_resampling, key_transition)
          kbase[1] = a;
          kbase[2] = b;
        } else {  // smc.py:142: (key_proposal, key_resampling)
          kbase[2] = a;
          kbase[1] = b;
        }
      }
      __syncthreads();
      // 1. GEMM on the tensor core || noise on the CUDA cores; epilogue: means -> Alo, log-likelihood -> lwraw
      gemm_and_epilogue(k, k, true, false);

      // 2. weights + ancestors
      if (warp == 1) {
        const Key kres = kbase[1];
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
          warp_normalise_v3(lw, N, lane);  // csmc.py:146
        } else {
          for (int q = lane; q < N; q += 32) lw[q] = lwraw[q];  // smc.py:144
          __syncwarp();
          if (p.lw_hist)
            for (int q = lane; q < N; q += 32) p.lw_hist[((size_t)chain * K + k) * N + q] = lw[q];
          const float c = warp_normalise_v3(lw, N, lane);  // smc.py:145,147
          if (lane == 0) scal[0] = (scal[0] - logN) + c;   // smc.py:146
          for (int q = lane; q < N; q += 32) w[q] = expf(lw[q]);
          __syncwarp();
          if (p.scheme == FBS_RESAMPLE_KILLING)
            warp_cond_killing(kres, w, N, 0, 0, false, cum, tmp, idx, lane);
          else if (p.scheme == FBS_RESAMPLE_MULTINOMIAL)
            warp_sorted_multinomial(kres, w, N, cum, reinterpret_cast<float*>(tmp), idx, lane);
          else
            warp_systematic_or_stratified(kres, w, N, p.scheme == FBS_RESAMPLE_SYSTEMATIC, true, cum, idx, lane);
        }
      }
      __syncthreads();

      // 3. children: gather the parents' means (through registers), add the noise, write hi / lo, pin the reference
      float val[2][4 * GPC_MAX];
      if (owner) {
        const int a0 = idx[own_pair], a1 = idx[own_pair + half];
#pragma unroll
        for (int gq = 0; gq < GPC_MAX; ++gq) {
          const int kk = 4 * (g0 + gq);
          if (gq < gpc && kk < du) {
            const float4 m0 = *reinterpret_cast<const float4*>(Alo + a_off(a0, kk));
            const float4 m1 = *reinterpret_cast<const float4*>(Alo + a_off(a1, kk));
            val[0][4 * gq + 0] = m0.x + nz[0][4 * gq + 0]; val[0][4 * gq + 1] = m0.y + nz[0][4 * gq + 1];
            val[0][4 * gq + 2] = m0.z + nz[0][4 * gq + 2]; val[0][4 * gq + 3] = m0.w + nz[0][4 * gq + 3];
            val[1][4 * gq + 0] = m1.x + nz[1][4 * gq + 0]; val[1][4 * gq + 1] = m1.y + nz[1][4 * gq + 1];
            val[1][4 * gq + 2] = m1.z + nz[1][4 * gq + 2]; val[1][4 * gq + 3] = m1.w + nz[1][4 * gq + 3];
          }
        }
      }
      __syncthreads();
      if (owner) {
        int bj = -1;
        const float* ustar = nullptr;
        if (p.mode == MODE_CSMC) {
          bj = p.bs_star[(size_t)chain * (K + 1) + k + 1];
          ustar = p.us_star + ((size_t)chain * (K + 1) + k + 1) * du;
        }
#pragma unroll
        for (int q = 0; q < 4 * GPC_MAX; ++q) {
          const int kk = 4 * g0 + q;
          if (q < 4 * gpc && kk < du) {
            put(own_pair, kk, own_pair == bj ? ustar[kk] : val[0][q]);  // csmc.py:143
            put(own_pair + half, kk, own_pair + half == bj ? ustar[kk] : val[1][q]);
          }
        }
      }
      fence_proxy_async();  // the new particles are read by the tensor core (async proxy) next step
      __syncthreads();

      // optional history
      if (p.mode == MODE_CSMC) {
        if (p.As)
          for (int t = tid; t < N; t += NTHREADS) p.As[((size_t)chain * K + k) * N + t] = idx[t];
        if (p.log_wss)
          for (int t = tid; t < N; t += NTHREADS) p.log_wss[((size_t)chain * (K + 1) + k + 1) * N + t] = lw[t];
        if (p.uss) store_particles(p.uss + ((size_t)chain * (K + 1) + k + 1) * N * du);
      } else {
        if (p.inds)
          for (int t = tid; t < N; t += NTHREADS) p.inds[((size_t)chain * K + k) * N + t] = idx[t];
        if (p.us_hist) store_particles(p.us_hist + ((size_t)chain * K + k) * N * du);
      }
    }

    // =============================== final state ===============================
    if (p.mode == MODE_CSMC) {
      if (p.us_last) store_particles(p.us_last + (size_t)chain * N * du);
      if (p.log_ws_last)
        for (int t = tid; t < N; t += NTHREADS) p.log_ws_last[(size_t)chain * N + t] = lw[t];
    } else {
      if (p.uT) store_particles(p.uT + (size_t)chain * N * du);
      if (p.log_ell && tid == 0) p.log_ell[chain] = scal[0];
    }
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------
// Self-test of the UMMA plumbing (descriptors, operand layout, TMEM read-back): D = A * B^T with the same
// split-TF32 scheme, operands given as float32 row-major A [128, K8] and the packed B image of one step.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bimg,
                                                               int K8, int nout, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nkb = K8 / 8;
  const uint32_t a_bytes = (uint32_t)(K8 / 4) * A_LBO;
  const uint32_t b_lbo = (uint32_t)(nout / 8) * 128u, blk = 2u * b_lbo, stage = 2u * blk;
  unsigned char* Ahi = smem;
  unsigned char* Alo = smem + a_bytes;
  unsigned char* Bst = smem + 2 * a_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bst + stage);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    tmem_alloc(slot, 256);
    if (lane == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
    }
  }
  for (int t = tid; t < ROWS * K8; t += 128) {
    const int r = t / K8, k = t - r * K8;
    const float x = A[t], hi = tf32_rn(x);
    *reinterpret_cast<float*>(Ahi + a_off(r, k)) = hi;
    *reinterpret_cast<float*>(Alo + a_off(r, k)) = x - hi;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nout >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
  uint32_t phase = 0;
  for (int kb = 0; kb < nkb; ++kb) {
    for (uint32_t t = tid; t < stage / 4; t += 128)
      reinterpret_cast<float*>(Bst)[t] = Bimg[(size_t)kb * (stage / 4) + t];
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi = smem_u32(Ahi) + (uint32_t)kb * 2u * A_LBO, a_lo = smem_u32(Alo) + (uint32_t)kb * 2u * A_LBO;
      const uint32_t b_hi = smem_u32(Bst), b_lo = b_hi + blk;
      umma_tf32(tbase, make_desc(a_hi, A_LBO, 128), make_desc(b_hi, b_lbo, 128), idesc, kb > 0 ? 1u : 0u);
      umma_tf32(tbase, make_desc(a_lo, A_LBO, 128), make_desc(b_hi, b_lbo, 128), idesc, 1u);
      umma_tf32(tbase, make_desc(a_hi, A_LBO, 128), make_desc(b_lo, b_lbo, 128), idesc, 1u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    __syncthreads();
  }
  const int r = 32 * warp + lane;
  for (int c0 = 0; c0 < nout; c0 += 8) {
    float v[8];
    tmem_ld8(tbase + ((uint32_t)(32 * warp) << 16) + c0, v);
    for (int c = 0; c < 8; ++c) D[(size_t)r * nout + c0 + c] = v[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

int launch_umma_selftest(void* stream, const float* A, const float* Bimg, int K8, int nout, float* D) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (K8 % 8 || nout % 16 || nout > 256 || K8 < 8) {
    set_error("umma_selftest: need K8 %% 8 == 0, nout %% 16 == 0, nout <= 256");
    return FBS_ERR_INVALID_ARGUMENT;
  }
  const size_t smem = (size_t)2 * (K8 / 4) * A_LBO + (size_t)4 * (nout / 8) * 128 + 64;
  if (smem > 227 * 1024) {
    set_error("umma_selftest: K8=%d too large", K8);
    return FBS_ERR_UNSUPPORTED;
  }
  cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_selftest_kernel<<<1, 128, smem, st>>>(A, Bimg, K8, nout, D);
  return check_launch("umma_selftest_kernel");
}

// Host: eligibility + launch.  p.MTc is the tensor-core image of the step matrices; p.ws the step-vector workspace.
int launch_sweep_v3(void* stream, SweepParams& p) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.MTc == nullptr || p.ws == nullptr) return -1;
  if (p.N < 2 || (p.N & 1) || p.N > ROWS) return -1;
  if (p.du % 4 != 0) return -1;
  const Layout L = make_layout(p.N, p.du, p.dv);
  if (L.nout > 256 || L.nkb < 1) return -1;
  const int half = p.N / 2, ngroups = L.dup4 / 4;
  const int per_pair = (NWORK * 32) / half;
  if (per_pair < 1) return -1;
  const int gpc = (ngroups + per_pair - 1) / per_pair;
  if (gpc > GPC_MAX) return -1;
  if (L.total > 227 * 1024) return -1;
  cudaError_t e = cudaFuncSetAttribute(sweep_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) {
    set_error("sweep_v3: cudaFuncSetAttribute(%zu B) failed: %s", L.total, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  {
    const int rc = launch_stepvec(stream, p);
    if (rc) return rc;
  }
  const int64_t grid = p.B < sm_count() ? p.B : sm_count();
  sweep_v3_kernel<<<(int)grid, NTHREADS, L.total, st>>>(p);
  return check_launch("sweep_v3_kernel");
}

}  // namespace fbs
