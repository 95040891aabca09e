// Per-timestep fused transition + weight kernel on the 5th-generation tensor cores, for particle sets in global memory
// with a LARGE state dimension (du >= 32): the body of fbs/samplers/csmc/csmc.py:140-146 for one step k,
//
//   parent = us_prev[A[n]]                         (gather through the ancestor indices)
//   D[128 x Nout] = parent[128 x du] * Mu_k^T      (u-drift | v-drift, split-TF32 on tcgen05, float32 in TMEM)
//   us_out[n] = parent + dt (D_u + c_u) + sd eps   (eps = jax.random.normal(key_tr, (N, du)), in-kernel threefry)
//   lw_out[n] = -0.5 (sum_v (c_v - dt D_v)^2 / sd^2 + lognorm)
//
// At du = dv = 100 the drift is 40 kFLOP per particle against 816 algorithmic bytes: on the CUDA cores the kernel is
// compute bound at a few percent of the HBM roofline, on the tensor cores the GEMM disappears behind the in-kernel RNG.
//
// A CTA is persistent over tiles of 128 particle rows = 64 PAIRS (n, n + N/2) of one chain -- the two particles whose
// noise comes from the same threefry blocks (jax's counter layout pairs elements half an array apart).
//   warp 0      MMA issuer: per 8-input K-block  Ahi Bhi + Alo Bhi + Ahi Blo  (tcgen05.mma kind::tf32)
//   warp 1      TMA producer: streams the packed (hi, lo) K-blocks of the step matrix through a shared-memory ring
//   warps 2-9   gather + tf32 split of the parents into the UMMA K-major operand, then the tile's normals in the
//               shadow of the GEMM (written to a row-major staging tile), then the epilogue: tcgen05.ld of the
//               accumulator (thread = particle row, two threads per row split the columns), children formed in place
//               in the staging tile, coalesced 16-byte stores.
// Operand layout, descriptors and the split-TF32 scheme are those of sweep_v3.cu (element (r, k) at
// (k / 4) * LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 4) * 4, LBO = 2048 for the 128-row particle operand).
#include <stdlib.h>
#include "fbs_common.cuh"
#include "fbs_rng.cuh"

namespace fbs {
namespace steptc {

constexpr int ROWS = 128, PAIRS = 64;
constexpr int MAX_WW = 16;  // worker warps: 8 (parents of the next tile prefetched into registers) or 16 (no prefetch)
constexpr int MAX_STAGES = 4;
constexpr uint32_t A_LBO = ROWS * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {  // off the critical path: back off
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done) __nanosleep(64);
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns, issued without the wait (two loads are kept in flight)
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t a_off(int r, int cg) {
  return (uint32_t)cg * A_LBO + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
}

// barrier of the MMA warp and the workers (the TMA producer warp runs free)
__device__ __forceinline__ void cta_sync(int nthreads) { asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); }

struct Params {
  int du, dv, N, k;
  int64_t B;
  const float *MT, *m, *dt, *sd, *lognorm, *MTc;
  const uint32_t* step_keys;
  const float* us_prev;
  const int32_t* A;
  const float *v, *v_prev, *u_star;
  const int32_t* b_cur;
  float *us_out, *lw_out;
  int stages, tiles_per_chain;
  long long* dbg;  // optional phase timers (cycles, CTA 0, first worker thread): see fbs_debug_step_tc_timers
};

struct Layout {
  int du8, nout, nkb, ncg, ncgA, nzs;
  uint32_t b_lbo, blk_bytes, stage_bytes, a_bytes;
  uint32_t Ahi, Alo, nz, ring, cs, ss, bars, misc, total;
};

__host__ __device__ inline Layout make_layout(int du, int dv, int stages) {
  Layout L;
  L.du8 = (du + 7) / 8 * 8;
  const int dv8 = (dv + 7) / 8 * 8;
  L.nout = L.du8 + dv8;
  if (L.nout % 16) L.nout += 8;
  L.nkb = L.du8 / 8;
  L.ncg = du / 4;
  L.ncgA = L.du8 / 4;
  L.nzs = (L.ncg & 1) ? du : du + 4;  // staging row stride (floats): an odd number of 16-byte groups, conflict free
  L.b_lbo = (uint32_t)(L.nout / 8) * 128u;
  L.blk_bytes = 2u * L.b_lbo;
  L.stage_bytes = 2u * L.blk_bytes;
  L.a_bytes = (uint32_t)L.ncgA * A_LBO;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t r = o;
    o += (bytes + 127u) / 128u * 128u;
    return r;
  };
  L.Ahi = take(L.a_bytes);
  L.Alo = take(L.a_bytes);
  L.nz = take((uint32_t)ROWS * L.nzs * 4u);
  L.ring = take((uint32_t)stages * L.stage_bytes);
  L.cs = take((uint32_t)L.nout * 4u);
  L.ss = take((MAX_WW / 4) * ROWS * 4);
  L.bars = take((2 * MAX_STAGES + 1) * 8);
  L.misc = take(64);
  L.total = o;
  return L;
}

__device__ __forceinline__ bool elect_one_lane() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

template <int WW, bool PRE>
__global__ void __launch_bounds__(32 * WW + 64, 1) step_transition_tc_kernel(const Params p) {
  constexpr int WORKERS = 32 * WW, NTHREADS = WORKERS + 64, HS = WW / 4, QSLOTS = 4 * WW;
  extern __shared__ __align__(1024) unsigned char smem[];
  const Layout L = make_layout(p.du, p.dv, p.stages);
  const int du = p.du, dv = p.dv, N = p.N, half = N / 2, D = du + dv;
  // (warp index through a shuffle: provably warp-uniform, so that the MMA / TMA warps' descriptors stay in uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  unsigned char* Ahi = smem + L.Ahi;
  unsigned char* Alo = smem + L.Alo;
  float* nz = reinterpret_cast<float*>(smem + L.nz);
  unsigned char* ring = smem + L.ring;
  float* cs = reinterpret_cast<float*>(smem + L.cs);
  float* ssp = reinterpret_cast<float*>(smem + L.ss);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint32_t* misc = reinterpret_cast<uint32_t*>(smem + L.misc);  // [0] tmem base, [2..3] key_tr

  if (warp == 0) {
    tmem_alloc(misc, 256);
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, 1);
      }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, misc[0], 0);
  const bool leader = elect_one_lane();  // the lane of the MMA / TMA warps that issues the asynchronous instructions
  const float dt = p.dt[p.k], sd = p.sd[p.k], lognorm = p.lognorm[p.k];
  const float inv_s2 = 1.0f / (sd * sd);
  const float* MTk = p.MT + (size_t)p.k * D * D;
  const float* mk = p.m + (size_t)p.k * D;
  const unsigned char* img = reinterpret_cast<const unsigned char*>(p.MTc) + (size_t)p.k * L.nkb * L.stage_bytes;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(L.nout >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
  const int64_t tiles = p.B * p.tiles_per_chain;
  // a CTA owns a contiguous run of tiles: consecutive tiles mostly belong to the same chain (its constants are reused)
  const int64_t per_cta = (tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_begin = blockIdx.x * per_cta;
  const int64_t tile_end = tile_begin + per_cta < tiles ? tile_begin + per_cta : tiles;
  const int wt = tid - 64;            // worker thread index (warps 2..9)
  const int step_r = WORKERS / L.ncg, step_c = WORKERS - step_r * L.ncg;  // (row, column group) stride of a worker
  long long tprev = 0;
  const bool timing = p.dbg != nullptr && blockIdx.x == 0 && tid == 64;
#define FBS_TICK(slot)                                 \
  if (timing) {                                        \
    const long long tnow = clock64();                  \
    p.dbg[slot] += tnow - tprev;                       \
    tprev = tnow;                                      \
  }
  uint32_t g = 0;                     // K-blocks streamed / consumed so far (ring position)
  uint32_t it = 0;                    // tiles done by this CTA (accumulator barrier parity)
  int64_t prev_b = -1;

  if (warp == 1) {
    // ---- TMA producer (free running, synchronised with the MMA warp through the ring's mbarriers only): the K-blocks
    //      of the step matrix once per tile, in consumption order
    // (all lanes run the loop, one elected lane issues)
    for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
      for (int kb = 0; kb < L.nkb; ++kb, ++g) {
        const uint32_t slot = g % (uint32_t)p.stages, use = g / (uint32_t)p.stages;
        if (use > 0) mbar_wait_sleep(empty + slot, (use - 1) & 1u);
        if (leader) {
          mbar_expect_tx(full + slot, L.stage_bytes);
          bulk_g2s(ring + slot * L.stage_bytes, img + (size_t)kb * L.stage_bytes, L.stage_bytes, full + slot);
        }
      }
    }
  } else
  {
  // ---- gather plumbing (workers).  A quarter-warp covers 8 distinct rows (conflict-free 16-byte shared stores); a thread
  //      always works on the same tile row and loads full 32-byte sectors (two column groups) of its parent row.  The
  //      parents of the NEXT tile are fetched into registers while the current tile's noise / epilogue run.
  constexpr int NJ = 16 / (QSLOTS / 16);  // du8 <= 128: at most 16 column-group pairs per row, QSLOTS / 16 quarter-slots per row group
  const int qslot = warp >= 2 ? (warp - 2) * 4 + (lane >> 3) : 0;  // < QSLOTS
  const int grow = (qslot & 15) * 8 + (lane & 7);  // this thread's tile row in the gather
  const int ncp = (L.ncgA + 1) / 2;
  float4 pre[NJ][2];
  int pre_idx = 0;
  auto tile_coords = [&](int64_t tile, int64_t& b, int& p0, int& npairs) {
    b = tile / p.tiles_per_chain;
    p0 = (int)(tile - b * p.tiles_per_chain) * PAIRS;
    npairs = min(PAIRS, half - p0);
  };
  auto fetch_index = [&](int64_t tile) {  // ancestor of this thread's gather row in `tile`
    pre_idx = -1;
    if (tile < tile_end) {
      int64_t b; int p0, npairs;
      tile_coords(tile, b, p0, npairs);
      if ((grow & (PAIRS - 1)) < npairs) {
        const int n = grow < PAIRS ? p0 + grow : half + p0 + (grow - PAIRS);
        pre_idx = __ldg(p.A + (size_t)b * N + n);
      }
    }
  };
  auto fetch_parent = [&](int64_t tile) {
    const int64_t b = tile < tile_end ? tile / p.tiles_per_chain : 0;
    const float4* parent = reinterpret_cast<const float4*>(p.us_prev + ((size_t)b * N + (pre_idx < 0 ? 0 : pre_idx)) * du);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int cp = (qslot >> 4) + (QSLOTS / 16) * j, cg0 = 2 * cp, cg1 = cg0 + 1;
      pre[j][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      pre[j][1] = pre[j][0];
      if (pre_idx >= 0 && cp < ncp) {
        if (cg0 < L.ncg) pre[j][0] = __ldg(parent + cg0);
        if (cg1 < L.ncg) pre[j][1] = __ldg(parent + cg1);
      }
    }
  };
  auto store_operands = [&]() {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int cp = (qslot >> 4) + (QSLOTS / 16) * j, cg0 = 2 * cp, cg1 = cg0 + 1;
      if (cp < ncp) {
        const float4 x0 = pre[j][0], x1 = pre[j][1];
        {
          const float4 h = make_float4(tf32_rn(x0.x), tf32_rn(x0.y), tf32_rn(x0.z), tf32_rn(x0.w));
          *reinterpret_cast<float4*>(Ahi + a_off(grow, cg0)) = h;
          *reinterpret_cast<float4*>(Alo + a_off(grow, cg0)) = make_float4(x0.x - h.x, x0.y - h.y, x0.z - h.z, x0.w - h.w);
        }
        if (cg1 < L.ncgA) {
          const float4 h = make_float4(tf32_rn(x1.x), tf32_rn(x1.y), tf32_rn(x1.z), tf32_rn(x1.w));
          *reinterpret_cast<float4*>(Ahi + a_off(grow, cg1)) = h;
          *reinterpret_cast<float4*>(Alo + a_off(grow, cg1)) = make_float4(x1.x - h.x, x1.y - h.y, x1.z - h.z, x1.w - h.w);
        }
      }
    }
  };
  if (PRE && warp >= 2) {
    fetch_index(tile_begin);
    fetch_parent(tile_begin);
  }

  for (int64_t tile = tile_begin; tile < tile_end; ++tile, ++it) {
    int64_t b; int p0, npairs;
    tile_coords(tile, b, p0, npairs);
    // row r of the tile <-> particle n(r); rows r in [npairs, 64) and [64 + npairs, 128) are empty
    auto row_particle = [&](int r) { return r < PAIRS ? p0 + r : half + p0 + (r - PAIRS); };
    auto row_valid = [&](int r) { return (r & (PAIRS - 1)) < npairs; };

    if (timing) tprev = clock64();
    if (warp >= 2) {
      if (!PRE) {
        fetch_index(tile);
        fetch_parent(tile);
      }
      store_operands();  // tf32 split of the (prefetched) parents into the UMMA operands
      // ---- the chain's constant vectors: u rows  c_u = m + M[:, du:] v_prev;  v rows  (v - v_prev) - dt (m + M[:, du:] v_prev)
      if (b != prev_b) {
        for (int o = wt; o < L.nout; o += WORKERS) {
          const bool isu = o < L.du8;
          const int i = isu ? o : du + (o - L.du8);
          const bool ok = isu ? o < du : (o - L.du8) < dv;
          float acc = 0.f;
          if (ok) {
            acc = mk[i];
            const float* vp = p.v_prev + (size_t)b * dv;
#pragma unroll 4
            for (int j = 0; j < dv; ++j) acc = fmaf(__ldg(MTk + (size_t)(du + j) * D + i), vp[j], acc);
            if (!isu) acc = (p.v[(size_t)b * dv + (o - L.du8)] - vp[o - L.du8]) - dt * acc;
          }
          cs[o] = acc;
        }
        if (wt == 0) {
          Key key_res, key_tr;
          split2(Key{p.step_keys[2 * b], p.step_keys[2 * b + 1]}, key_res, key_tr);  // csmc.py:136
          misc[2] = key_tr.k0;
          misc[3] = key_tr.k1;
        }
      }
      fence_proxy_async();  // the operand stores must be visible to the tensor core (async proxy)
      FBS_TICK(0)  // gather + split + constants
    }
    prev_b = b;
    tc_fence_before();
    cta_sync(NTHREADS - 32);  // #1: operands, constants and the transition key are in place; the previous tile's stores are done
    tc_fence_after();
    FBS_TICK(1)  // barrier #1

    if (warp == 0) {
      // ---- MMA issuer
      {
        uint32_t gm = g;
        for (int kb = 0; kb < L.nkb; ++kb, ++gm) {
          const uint32_t slot = gm % (uint32_t)p.stages, use = gm / (uint32_t)p.stages;
          mbar_wait(full + slot, use & 1u);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(Ahi) + (uint32_t)kb * 2u * A_LBO, a_lo = smem_u32(Alo) + (uint32_t)kb * 2u * A_LBO;
          const uint32_t b_hi = smem_u32(ring + slot * L.stage_bytes), b_lo = b_hi + L.blk_bytes;
          const uint64_t dAh = make_desc(a_hi, A_LBO, 128), dAl = make_desc(a_lo, A_LBO, 128);
          const uint64_t dBh = make_desc(b_hi, L.b_lbo, 128), dBl = make_desc(b_lo, L.b_lbo, 128);
          if (leader) {
            umma_tf32(tbase, dAh, dBh, idesc, kb > 0 ? 1u : 0u);
            umma_tf32(tbase, dAl, dBh, idesc, 1u);
            umma_tf32(tbase, dAh, dBl, idesc, 1u);
            umma_commit(empty + slot);
          }
        }
        if (leader) umma_commit(acc_full);
      }
      g += (uint32_t)L.nkb;
    } else if (warp >= 2) {
      // ---- noise in the shadow of the GEMM: task (pair, column group) = 4 threefry blocks = 4 normals for particle
      //      p0 + pair and 4 for its partner, element e = n * du + i of normal(key_tr, (N, du))
      if (PRE) fetch_index(tile + 1);
      const uint32_t k0 = misc[2], k1 = misc[3];
      const uint32_t hblk = (uint32_t)half * (uint32_t)du;
      const int ntasks = npairs * L.ncg;
      int pr = wt / L.ncg, cg = wt - pr * L.ncg;
      for (int t = wt; t < ntasks; t += WORKERS, pr += step_r, cg += step_c) {
        if (cg >= L.ncg) { cg -= L.ncg; ++pr; }
        const uint32_t e = (uint32_t)(p0 + pr) * (uint32_t)du + 4u * (uint32_t)cg;
        uint32_t x0[4] = {e, e + 1u, e + 2u, e + 3u};
        uint32_t x1[4] = {e + hblk, e + hblk + 1u, e + hblk + 2u, e + hblk + 3u};
        threefry2x32_x4(k0, k1, x0, x1);
        float4 lo4, hi4;
        lo4.x = sd * bits_to_normal(x0[0]); lo4.y = sd * bits_to_normal(x0[1]);
        lo4.z = sd * bits_to_normal(x0[2]); lo4.w = sd * bits_to_normal(x0[3]);
        hi4.x = sd * bits_to_normal(x1[0]); hi4.y = sd * bits_to_normal(x1[1]);
        hi4.z = sd * bits_to_normal(x1[2]); hi4.w = sd * bits_to_normal(x1[3]);
        *reinterpret_cast<float4*>(nz + (size_t)pr * L.nzs + 4 * cg) = lo4;
        *reinterpret_cast<float4*>(nz + (size_t)(pr + PAIRS) * L.nzs + 4 * cg) = hi4;
      }
      FBS_TICK(2)  // noise
      if (PRE) fetch_parent(tile + 1);  // in flight during the epilogue
      // the noise rows are read by other threads (row owners) below: a worker-only barrier
      asm volatile("bar.sync 1, %0;" ::"r"(WORKERS) : "memory");
      FBS_TICK(3)  // worker barrier
      mbar_wait(acc_full, it & 1u);
      tc_fence_after();
      FBS_TICK(4)  // accumulator wait
      // ---- epilogue: thread = accumulator row (TMEM lane quadrant = warp % 4), the two threads of a row split the
      //      8-column chunks
      const int r = 32 * (warp & 3) + lane;
      const int hs = (warp - 2) >> 2;
      const bool valid = row_valid(r);
      const int n = row_particle(r);
      const bool pinned = valid && n == p.b_cur[b];
      const uint32_t trow = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
      const int nuc = L.du8 / 8, nvc = (L.nout - L.du8) / 8;
      const int uc0 = nuc * hs / HS, uc1 = nuc * (hs + 1) / HS;
      const int vc0 = nvc * hs / HS, vc1 = nvc * (hs + 1) / HS;
      float ss = 0.f;
      for (int c = vc0; c < vc1; c += 4) {  // four 8-column loads in flight per wait
        uint32_t acc[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c + u < vc1) tmem_ld8_issue(trow + (uint32_t)(L.du8 + 8 * (c + u)), acc[u]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c + u < vc1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float resid = cs[L.du8 + 8 * (c + u) + q] - dt * __uint_as_float(acc[u][q]);
              ss = fmaf(resid, resid, ss);
            }
          }
        }
      }
      ssp[hs * ROWS + r] = ss;
      for (int c2 = uc0; c2 < uc1; c2 += 2) {  // two 8-column loads in flight per wait
        uint32_t acc2[2][8];
        tmem_ld8_issue(trow + (uint32_t)(8 * c2), acc2[0]);
        if (c2 + 1 < uc1) tmem_ld8_issue(trow + (uint32_t)(8 * (c2 + 1)), acc2[1]);
        tmem_ld_wait();
#pragma unroll
        for (int hq2 = 0; hq2 < 4; ++hq2) {
          const int c = c2 + (hq2 >> 1), hq = hq2 & 1;
          const uint32_t* acc = acc2[hq2 >> 1];
          const int cg = c < uc1 ? 2 * c + hq : L.ncg;
          if (cg < L.ncg) {
            const float4 ph = *reinterpret_cast<const float4*>(Ahi + a_off(r, cg));
            const float4 pl = *reinterpret_cast<const float4*>(Alo + a_off(r, cg));
            const float4 cu = *reinterpret_cast<const float4*>(cs + 4 * cg);
            float4* dst = reinterpret_cast<float4*>(nz + (size_t)r * L.nzs + 4 * cg);
            float4 x = *dst;
            x.x += (ph.x + pl.x) + dt * (__uint_as_float(acc[4 * hq + 0]) + cu.x);
            x.y += (ph.y + pl.y) + dt * (__uint_as_float(acc[4 * hq + 1]) + cu.y);
            x.z += (ph.z + pl.z) + dt * (__uint_as_float(acc[4 * hq + 2]) + cu.z);
            x.w += (ph.w + pl.w) + dt * (__uint_as_float(acc[4 * hq + 3]) + cu.w);
            if (pinned) x = __ldg(reinterpret_cast<const float4*>(p.u_star + (size_t)b * du) + cg);  // csmc.py:143
            *dst = x;
          }
        }
      }
    }
    FBS_TICK(5)  // epilogue
    tc_fence_before();
    cta_sync(NTHREADS - 32);  // #2: accumulator and operands are free again, the children are staged
    tc_fence_after();
    FBS_TICK(6)  // barrier #2
    if (warp >= 2) {
      // ---- coalesced stores: the two runs of npairs consecutive particle rows, 16 bytes per thread
      const int nvec = 2 * npairs * L.ncg;
      int rr = wt / L.ncg, cg = wt - rr * L.ncg;
      for (int t = wt; t < nvec; t += WORKERS, rr += step_r, cg += step_c) {
        if (cg >= L.ncg) { cg -= L.ncg; ++rr; }
        const int r = rr < npairs ? rr : PAIRS + (rr - npairs);
        const int n = row_particle(r);
        *(reinterpret_cast<float4*>(p.us_out + ((size_t)b * N + n) * du) + cg) =
            *reinterpret_cast<const float4*>(nz + (size_t)r * L.nzs + 4 * cg);
      }
      if (wt < ROWS && row_valid(wt))
      {
        float ss = ssp[wt];
#pragma unroll
        for (int h2 = 1; h2 < HS; ++h2) ss += ssp[h2 * ROWS + wt];
        p.lw_out[(size_t)b * N + row_particle(wt)] = -0.5f * (ss * inv_s2 + lognorm);
      }
      FBS_TICK(7)  // stores
    }
  }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

#undef FBS_TICK
static long long* g_step_tc_dbg = nullptr;

}  // namespace steptc

// Test / profiling hook: device buffer of 8 int64 that CTA 0 accumulates its per-phase cycle counts into (NULL = off).
extern "C" int fbs_debug_step_tc_timers(long long* dev_buf) {
  steptc::g_step_tc_dbg = dev_buf;
  return FBS_OK;
}

// Returns FBS_OK, an error, or -1 when the shape is not eligible (the caller falls back to the CUDA-core kernels).
int launch_step_transition_tc(cudaStream_t st, const fbs_affine_model_t* model, int k, const uint32_t* step_keys,
                              const float* us_prev, const int32_t* A, const float* v, const float* v_prev,
                              const float* u_star, const int32_t* b_cur, int64_t B, int64_t N, float* us_out,
                              float* lw_out) {
  using namespace steptc;
  const int du = model->du, dv = model->dv;
  if (model->MTc == nullptr || du % 4 != 0 || du < 32 || du > 128 || (N & 1) || N < 2 || N >= (1 << 24)) return -1;
  if ((uint64_t)N * (uint64_t)du >= (1ull << 32)) return -1;
  int stages = MAX_STAGES;
  Layout L = make_layout(du, dv, stages);
  while (L.total > 227 * 1024 && stages > 2) L = make_layout(du, dv, --stages);
  if (L.total > 227 * 1024 || L.nout > 256) return -1;
  Params p;
  p.du = du; p.dv = dv; p.N = (int)N; p.k = k; p.B = B;
  p.MT = model->MT; p.m = model->m; p.dt = model->dt; p.sd = model->sd; p.lognorm = model->lognorm; p.MTc = model->MTc;
  p.step_keys = step_keys; p.us_prev = us_prev; p.A = A; p.v = v; p.v_prev = v_prev; p.u_star = u_star; p.b_cur = b_cur;
  p.us_out = us_out; p.lw_out = lw_out;
  p.stages = stages;
  p.dbg = g_step_tc_dbg;
  p.tiles_per_chain = (int)((N / 2 + PAIRS - 1) / PAIRS);
  const int64_t tiles = B * p.tiles_per_chain;
  const int64_t per_cta = (tiles + sm_count() - 1) / sm_count();
  const int grid = (int)((tiles + per_cta - 1) / per_cta);
  const bool eight = debug_opt(OPT_STEP_TC_WARPS) == 8;  // eight worker warps with register prefetch; default sixteen
  auto kern = eight ? step_transition_tc_kernel<8, true> : step_transition_tc_kernel<16, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) {
    set_error("step_tc: cudaFuncSetAttribute(%u B) failed: %s", L.total, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  kern<<<grid, eight ? 320 : 576, L.total, st>>>(p);
  return check_launch("step_transition_tc_kernel");
}

}  // namespace fbs
