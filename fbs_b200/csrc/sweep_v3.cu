// Whole-sweep kernel on the 5th-generation tensor cores ("v3"): tcgen05.mma + TMEM + TMA, two chains per CTA.
//
// Per chain-step the joint drift of all particles is ONE GEMM  D[128 x Nout] = U[128 x du] * Mu_k^T  with
// Nout = du8 + dv8 outputs (u-drift | v-drift).  float32 parity is kept by a split-TF32 product
//   U M = Uhi Mhi + Ulo Mhi + Uhi Mlo        (hi = round-to-nearest tf32, lo = exact float32 remainder)
// accumulated in float32 in TENSOR MEMORY.
//
// A CSMC step is a strictly serial chain of phases (GEMM -> weights -> ancestors -> gather), so one CTA runs TWO
// independent chains, one per group of 7 warps, each with its own particle operands, accumulator (256 TMEM columns)
// and named barriers; the groups drift half a step apart and fill each other's bubbles.  16 warps = 512 threads is
// the largest CTA that still gets 128 registers per thread.
//
//   warp 0   : MMA.  One elected thread issues the tcgen05.mma's of both groups in a fixed alternating order (GEMM t of
//              group 0, GEMM t of group 1, ...), each when that group has signalled "particles ready" (mbarrier).
//   warp 15  : TMA producer.  One elected thread streams the packed (hi, lo) K-blocks of M_k through a ring of
//              shared-memory slots with bulk copies (cp.async.bulk + mbarrier), in the order the MMA warp consumes them.
//   group g (warps 1 + 7g .. 7 + 7g), by role:
//     E x4   : epilogue + noise.  Noise tasks (threefry2x32 -> normals, kept in registers) in the shadow of the GEMM;
//              then tcgen05.ld of the accumulator, one particle row per thread: v-half -> Gaussian log-likelihood
//              (releases R), u-half -> transition means; then the rest of their noise.
//     R x1   : resampling only (normalise, sequential cumsum, searches), register resident; the step keys of the next
//              step and this step's uniforms are computed before the weights arrive, off the critical path.
//     X x2   : noise only (two tasks more than an E warp: what the epilogue costs).
//   After the group barrier all noise warps gather the parents' means + their noise and write the new particles
//   (hi / lo split) in the UMMA K-major core-matrix layout.
//
// Shared-memory operand layout (no swizzle, K-major "interleave"): element (row r, k) of an operand lives at
//   (k / 4) * LBO + (r / 8) * 128 + (r % 8) * 16 + (k % 4) * 4  bytes,
// i.e. 8-row x 16-byte core matrices; one MMA consumes K = 8 (two k-chunks, LBO apart), SBO = 128.  For the particle
// operand LBO = ceil(N / 8) * 128: the MMA always reads M = 128 rows, rows >= N alias the next k-chunk (finite
// garbage) and only produce accumulator rows >= N, which are never read.
//
// Reference: fbs/samplers/csmc/csmc.py:80-164, fbs/samplers/smc.py:115-158 (same algorithm and random streams
// as csmc_kernels.cu / sweep_v2.cu).
#include <stdlib.h>
#include <type_traits>
#include "fbs_common.cuh"
#include "fbs_resample.cuh"
#include "fbs_sweep.cuh"

namespace fbs {

namespace v3 {

constexpr int ROWS = 128;            // MMA M: particle rows per chain (N <= 128)
constexpr int GROUPS = 2;            // chains in flight per CTA
constexpr int GWARPS = 7;            // warps per group: 4 epilogue + noise, 1 resampling, 2 noise only
constexpr int GTHREADS = 32 * GWARPS;
constexpr int NOISE_THREADS = 32 * (GWARPS - 1);
constexpr int NWARPS = 2 + GROUPS * GWARPS;  // MMA warp, two groups, TMA producer warp
constexpr int NTHREADS = 32 * NWARPS;
constexpr int MAX_STAGES = 8;        // ring of K-blocks of the step matrix (2..4, chosen by the host to fit)
constexpr int TMEM_COLS_PER_GROUP = 256;
constexpr uint32_t SELFTEST_LBO = (ROWS / 8) * 128;

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
// the same wait with a suspend-time hint: the thread sleeps inside try_wait (woken by the phase completion) instead of
// re-issuing the probe -- the MMA / TMA warps' polls were ~9 % of the executed instructions (profiles/r1_v3_sweep_hot_lines.txt)
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITH_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAITH_DONE;\n\t"
      "bra WAITH_LOOP;\n\t"
      "WAITH_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
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
__device__ __forceinline__ bool elect_one() {
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

// polling wait that backs off: the waiting warps must not steal issue slots from the other group's noise
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done) __nanosleep(200);
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of the 4 consecutive columns 4 cg .. 4 cg + 3 of particle row r in an A operand
__device__ __forceinline__ uint32_t a_off(uint32_t lbo, int r, int cg) {
  return (uint32_t)cg * lbo + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
}

// Four threefry blocks -> 8 scaled normals: elements b .. b + 3 (lo) and b + hblk .. b + hblk + 3 (hi) of
// normal(key, (2 * hblk,)).  Deliberately NOT inlined: the sweep's hot loop has to stay inside the 32 KB instruction
// cache while two warp groups in different phases execute it; inlined per task it is > 60 KB of straight-line code.
struct Noise8 {
  float lo[4], hi[4];
};
__device__ __noinline__ Noise8 noise_task(uint32_t k0, uint32_t k1, uint32_t b, uint32_t hblk, float scale) {
  uint32_t x0[4] = {b, b + 1u, b + 2u, b + 3u};
  uint32_t x1[4] = {b + hblk, b + hblk + 1u, b + hblk + 2u, b + hblk + 3u};
  threefry2x32_x4(k0, k1, x0, x1);
  Noise8 r;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    r.lo[c] = scale * bits_to_normal(x0[c]);
    r.hi[c] = scale * bits_to_normal(x1[c]);
  }
  return r;
}

struct Layout {
  int du8, dv8, nout, nkb, ncg, stages;
  int nu_pass, nv_pass;  // output rows of the second (u) and first (v + the last u rows) GEMM pass, multiples of 16
  uint32_t a_lbo, b_lbo, blk_bytes, img_bytes, stage_bytes, a_bytes;
  // byte offsets.  Group g: A operands at A + g * 2 * a_bytes (hi, then lo), small arrays at grp + g * grp_bytes + <field>
  uint32_t A, ring, bars, tmem, grp, grp_bytes;
  uint32_t cvs, lwraw, lw, w, cum, idx, tmp, keys, scal, skeys, pin, ub;
  uint32_t total;
};

__host__ __device__ inline Layout make_layout(int N, int du, int dv, int stages_flags) {
  const int stages = stages_flags & 0xFF;
  const bool twopass = (stages_flags >> 8) & 1;
  Layout L;
  L.du8 = (du + 7) / 8 * 8;
  L.dv8 = (dv + 7) / 8 * 8;
  L.nout = L.du8 + L.dv8;
  if (L.nout % 16) L.nout += 8;
  L.nkb = L.du8 / 8;
  L.ncg = du / 4;
  L.stages = stages;
  L.a_lbo = (uint32_t)((N + 7) / 8) * 128u;
  L.b_lbo = (uint32_t)(L.nout / 8) * 128u;
  L.blk_bytes = 2u * L.b_lbo;
  L.img_bytes = 2u * L.blk_bytes;           // one K-block of the image in global memory: (hi | lo) x 2 k-chunks x nout rows
  L.nu_pass = twopass ? (L.du8 / 16) * 16 : 0;
  L.nv_pass = L.nout - L.nu_pass;
  L.stage_bytes = 4u * (uint32_t)(L.nv_pass / 8) * 128u;  // ring slot: the rows of ONE pass, (hi | lo) x 2 k-chunks, compact
  L.a_bytes = (uint32_t)(L.du8 / 4) * L.a_lbo;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t r = o;
    o += (bytes + 127u) / 128u * 128u;
    return r;
  };
  L.A = take(2u * GROUPS * L.a_bytes);
  L.ring = take((uint32_t)stages * L.stage_bytes);  // also absorbs the M = 128 over-read of the last k-chunk
  L.bars = take((2 * MAX_STAGES + 3 * GROUPS) * 8);
  L.tmem = take(16);
  L.grp = o;
  o = 0;
  L.cvs = take((uint32_t)(L.du8 + L.dv8) * 4u);
  L.lwraw = take(ROWS * 4);
  L.lw = take(ROWS * 4);
  L.w = take(ROWS * 4);
  L.cum = take((ROWS + 1) * 4);
  L.idx = take(ROWS * 4);
  L.tmp = take((ROWS + 1) * 4);
  L.keys = take(64);
  L.scal = take(16);
  L.skeys = take(64);                       // 2 slots (step parity) x (resampling key, transition key)
  L.pin = take((uint32_t)(du + 4) * 4u);    // pinned reference particle of the step + its slot
  L.ub = take((2 * ROWS + 4) * 4);          // resampling uniforms of the step, drawn before the weights arrive
  L.grp_bytes = o;
  L.total = L.grp + GROUPS * L.grp_bytes;
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

// NE / NX: noise tasks per thread of the E (epilogue + noise) and X (noise only) warps (the two epilogues cost an E warp
// about two tasks: <6, 7, 2> at d = 100, N = 100).
template <int NE, int NX, int NR, bool DBG>
__global__ void __launch_bounds__(v3::NTHREADS, 1) sweep_v3_kernel(const SweepParams p, const int stages_flags) {
  const int stages = stages_flags & 0xFF;
  extern __shared__ __align__(1024) unsigned char smem[];
  // an E warp generates NE1 of its tasks in the shadow of the first GEMM pass, up to NE2 in the shadow of the second
  // pass (while the resampling warp works), the rest after its u-epilogue
  constexpr int NT0 = NE > NX ? NE : NX, NT = NT0 > NR ? NT0 : NR, NE1 = NE / 2, NE2 = NE > 0 ? NE - 1 : 0;
  // NR > 0: the resampling warp generates NR tasks per lane too, in the window in which it otherwise waits for the weights,
  // and takes part in the gather / operand stores with them
  constexpr int TASK_THREADS = NR > 0 ? GTHREADS : NOISE_THREADS;
  const Layout L = make_layout(p.N, p.du, p.dv, stages_flags);
  const int du = p.du, dv = p.dv, N = p.N, K = p.K, half = N / 2;
  // (warp index through a shuffle: provably warp-uniform, so that the role branches are uniform control flow and the MMA / TMA
  // warps' descriptors and addresses stay in uniform registers)
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  unsigned char* ring = smem + L.ring;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* accum = empty + MAX_STAGES;  // [g]: first pass (v columns) of group g's accumulator complete
  uint64_t* accum_u = accum + GROUPS;    // [g]: second pass (u columns) complete
  uint64_t* ready = accum_u + GROUPS;    // [g]: particles of group g written, accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem);
  if (warp == 0) {
    tmem_alloc(tmem_slot, GROUPS * TMEM_COLS_PER_GROUP);
    if (lane == 0) {
      for (int s = 0; s < MAX_STAGES; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, 1);
      }
      for (int g = 0; g < GROUPS; ++g) {
        mbar_init(accum + g, 1);
        mbar_init(accum_u + g, 1);
        mbar_init(ready + g, 1);
      }
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const bool leader = elect_one();  // the one lane of the MMA / TMA warps that issues the asynchronous instructions
  // phase time stamps (profiling hook, scripts/v3_timeline.py): CTA 0, first chain pair, steps [DBG_K0, DBG_K0 + DBG_NS),
  // lane 0 of one warp per role; layout dbg[role][step][stamp]
  constexpr int DBG_K0 = 64, DBG_NS = 4, DBG_MAXS = 16;
  const bool dbg_cta = DBG && p.dbg != nullptr && blockIdx.x == 0 && lane == 0;

  // chains of this CTA: pair P = blockIdx.x + i * gridDim.x, group g runs chain 2 P + g
  const bool init_gemm = (p.mode == MODE_CSMC && p.init_mode == FBS_INIT_NORMAL);
  const uint32_t GP = (uint32_t)K + (init_gemm ? 1u : 0u);  // GEMMs per chain
  uint32_t nch[GROUPS];
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    const int64_t lim = (p.B - g + 1) / 2;  // pairs P with 2 P + g < B
    nch[g] = lim > (int64_t)blockIdx.x ? (uint32_t)((lim - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  }

  const uint32_t G0 = nch[0] * GP, G1 = nch[1] * GP;  // G0 >= G1; GEMM order: (0,0) (1,0) (0,1) (1,1) ...
  // Each GEMM runs as two passes over the K-blocks: first the output rows [nu_pass, nout) -- every v row, what the
  // weights need -- then the rows [0, nu_pass), so that the resampling overlaps the second pass.
  const int npass = L.nu_pass > 0 ? 2 : 1;

  if (warp == NWARPS - 1) {
    // =============================== TMA producer warp ===============================
    // streams the K-blocks of the step matrices, in the order the MMA warp consumes them, through the ring
    // all lanes run the loop (uniform registers for the addresses); the copies are issued by one elected lane
    {
      uint32_t ps = 0, pph = 1;  // slot, parity of its "empty" barrier (1: passes on a fresh barrier)
      fence_proxy_async();
      for (uint32_t t = 0; t < G0; ++t) {
        uint32_t step = t % GP;
        if (init_gemm) step = step == 0 ? 0 : step - 1;  // the initial weights use the step-0 matrix
        const unsigned char* img = reinterpret_cast<const unsigned char*>(p.MTc) + (size_t)step * L.nkb * L.img_bytes;
        for (int g = 0; g < GROUPS; ++g) {
          if (g == 1 && t >= G1) break;
          for (int pass = 0; pass < npass; ++pass) {
            const uint32_t r0 = pass == 0 ? (uint32_t)L.nu_pass : 0u;
            const uint32_t np = pass == 0 ? (uint32_t)L.nv_pass : (uint32_t)L.nu_pass;
            const uint32_t run = (np / 8u) * 128u;  // bytes of this pass's rows inside one (part, k-chunk) of the image
            const unsigned char* src = img + (size_t)(r0 / 8u) * 128u;
            for (int kb = 0; kb < L.nkb; ++kb, src += L.img_bytes) {
              if (stages_flags & 0x200) mbar_wait_hint(empty + ps, pph); else mbar_wait(empty + ps, pph);
              unsigned char* dst = ring + (size_t)ps * L.stage_bytes;
              if (leader) {
                mbar_expect_tx(full + ps, 4u * run);
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c)  // (hi, chunk 0) (hi, chunk 1) (lo, chunk 0) (lo, chunk 1)
                  bulk_g2s(dst + c * run, src + (size_t)c * L.b_lbo, run, full + ps);
              }
              if (++ps == (uint32_t)stages) {
                ps = 0;
                pph ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // =============================== MMA warp ===============================
    // all lanes run the loop, one elected lane issues: inside an `if (lane == 0)` region every descriptor was moved from a
    // vector to a uniform register in front of each tcgen05.mma (measured on conv_gemm_kernel: ~75 cycles per MMA issued)
    {
      uint32_t cs = 0, cph = 0;  // slot, parity of its "full" barrier
      const uint32_t ring16 = smem_u32(ring) >> 4, stage16 = L.stage_bytes >> 4;
      const uint64_t desc_hi_A = ((uint64_t)((L.a_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
      for (uint32_t t = 0; t < G0; ++t) {
        for (int g = 0; g < GROUPS; ++g) {
          if (g == 1 && t >= G1) break;
          if (stages_flags & 0x200) mbar_wait_hint(ready + g, t & 1u); else mbar_wait(ready + g, t & 1u);
          tc_fence_after();
          const int dbg_k = (int)(t % GP) - (init_gemm ? 1 : 0) - DBG_K0;  // the step this GEMM belongs to
          const bool dbg_mma = dbg_cta && t < GP && dbg_k >= 0 && dbg_k < DBG_NS;
          if (dbg_mma) p.dbg[(6 * DBG_NS + dbg_k) * DBG_MAXS + 2 * g] = clock64();
          const uint32_t a_hi0 = smem_u32(smem + L.A) + (uint32_t)g * 2u * L.a_bytes;
          const uint32_t d_tmem = tmem_base + (uint32_t)g * TMEM_COLS_PER_GROUP;
          for (int pass = 0; pass < npass; ++pass) {
            const uint32_t r0 = pass == 0 ? (uint32_t)L.nu_pass : 0u;
            const uint32_t np = pass == 0 ? (uint32_t)L.nv_pass : (uint32_t)L.nu_pass;
            const uint32_t run = (np / 8u) * 128u;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((np >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
            const uint64_t desc_hi_B = ((uint64_t)((run >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
            const uint32_t dt = d_tmem + r0;
            // descriptors advance by plain adds on the (address >> 4) field; all operands sit below 256 KB, no carry
            uint64_t dAh = desc_hi_A | (uint64_t)((a_hi0 >> 4) & 0x3FFFu);
            uint64_t dAl = desc_hi_A | (uint64_t)(((a_hi0 + L.a_bytes) >> 4) & 0x3FFFu);
            const uint64_t a_inc = (uint64_t)((2u * L.a_lbo) >> 4), b_lo_off = (uint64_t)((2u * run) >> 4);
            for (int kb = 0; kb < L.nkb; ++kb) {
              if (stages_flags & 0x200) mbar_wait_hint(full + cs, cph); else mbar_wait(full + cs, cph);
              tc_fence_after();
              const uint64_t dBh = desc_hi_B | (uint64_t)((ring16 + cs * stage16) & 0x3FFFu);
              const uint64_t dBl = dBh + b_lo_off;
              if (leader) {
                umma_tf32(dt, dAh, dBh, idesc, kb > 0 ? 1u : 0u);
                umma_tf32(dt, dAl, dBh, idesc, 1u);
                umma_tf32(dt, dAh, dBl, idesc, 1u);
                umma_commit(empty + cs);  // slot free once these MMAs have read it
              }
              dAh += a_inc;
              dAl += a_inc;
              if (++cs == (uint32_t)stages) {
                cs = 0;
                cph ^= 1u;
              }
            }
            if (leader) umma_commit((pass == 0 ? accum : accum_u) + g);  // accumulator columns of this pass complete
          }
          if (dbg_mma) p.dbg[(6 * DBG_NS + dbg_k) * DBG_MAXS + 2 * g + 1] = clock64();
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== worker group g ===============================
    // roles inside a group (gw = warp of the group): 0..3 "E" epilogue + noise, 4 "R" resampling only, 5.. "X" noise only
    // warp -> (group, role).  A warp lives on scheduler (warp % 4); the four light warps (MMA 0, R 5 and 10, TMA 15) sit on
    // four different schedulers and every scheduler gets exactly three of the twelve noise warps -- with the plain
    // "group g = warps 1 + 7 g .." numbering one scheduler had four noise warps and another two, and the step waited for
    // the crowded one.  E warps of a group keep four distinct TMEM lane quadrants (warp % 4).
    //   warp:   1  2  3  4  5  6  7 |  8  9 10 11 12 13 14
    //   group:  0  0  0  0  0  0  0 |  1  1  1  1  1  1  1
    //   gw:     0  1  2  3  4  5  6 |  0  1  4  3  5  6  2        (0..3 E, 4 R, 5..6 X)
    const int g = warp >= 8 ? 1 : 0;
    const int gw = g == 0 ? warp - 1 : ((0x2653410 >> (4 * (warp - 8))) & 0xF);
    const int gt = 32 * gw + lane;
    const bool is_E = gw < 4, is_R = gw == 4, is_noise = !is_R;
    const int nt = gw < 4 ? gt : gt - 32;  // index among the NOISE_THREADS noise threads
    const int xt = gt - 160;               // index among the 64 X threads
    unsigned char* Ahi = smem + L.A + (size_t)g * 2u * L.a_bytes;
    unsigned char* Alo = Ahi + L.a_bytes;  // doubles as the transition-mean buffer between the GEMM and the gather
    unsigned char* gb = smem + L.grp + (size_t)g * L.grp_bytes;
    float* cvs = reinterpret_cast<float*>(gb + L.cvs);  // [0, du8): cu, [du8, du8 + dv8): cv of this step
    float* lwraw = reinterpret_cast<float*>(gb + L.lwraw);
    float* lw = reinterpret_cast<float*>(gb + L.lw);
    float* w = reinterpret_cast<float*>(gb + L.w);
    float* cum = reinterpret_cast<float*>(gb + L.cum);
    int* idx = reinterpret_cast<int*>(gb + L.idx);
    int* tmp = reinterpret_cast<int*>(gb + L.tmp);
    Key* kbase = reinterpret_cast<Key*>(gb + L.keys);  // [0]: sweep key, [2]: init key
    float* scal = reinterpret_cast<float*>(gb + L.scal);
    Key* skeys = reinterpret_cast<Key*>(gb + L.skeys);  // slot k & 1: [0] resampling key, [1] transition key of step k
    float* pin = reinterpret_cast<float*>(gb + L.pin);  // [0, du): u*_{k+1}, [du]: its slot b*_{k+1} (as int bits)
    float* ub = reinterpret_cast<float*>(gb + L.ub);    // killing: U1 at [0, N), U2 at [ROWS, ROWS + N), U_J at [2 ROWS]
    const float logN = logf((float)N);
    const uint32_t tmem_g = tmem_base + (uint32_t)g * TMEM_COLS_PER_GROUP;
    const uint32_t lbo = L.a_lbo;
    const int ncg = L.ncg;
    const uint32_t hblk = (uint32_t)half * du;  // random_bits(key, N * du): element e and e + hblk share a block

    // named barriers of this group
    auto bar_all = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(GTHREADS) : "memory"); };
    auto bar_noise = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(3 + g), "n"(TASK_THREADS) : "memory"); };
    auto bar_E = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(5 + g), "n"(128) : "memory"); };
    auto bar_ER_arrive = [&]() { asm volatile("bar.arrive %0, %1;" ::"r"(7 + g), "n"(160) : "memory"); };
    auto bar_ER_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(7 + g), "n"(160) : "memory"); };

    // noise tasks of this thread: (row pair (n, n + half), column group cg), pairs fastest so that a warp touches
    // consecutive rows of one core-matrix column (conflict-free 16-byte accesses)
    const int ntasks = half * ncg;
    uint32_t task[NT];
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      int t = ntasks;
      if (is_E && i < NE) t = gt + 128 * i;
      if (gw > 4 && i < NX) t = 128 * NE + xt + 64 * i;
      if (is_R && i < NR) t = 128 * NE + 64 * NX + lane + 32 * i;
      task[i] = t < ntasks ? ((uint32_t)(t % half) | ((uint32_t)(t / half) << 16)) : 0xFFFFFFFFu;
    }
    float nz[2][4 * NT];  // noise, then the children, of the owned (rows, columns)

    uint32_t gcount = 0;  // GEMMs of this group so far (phase of accum[g] / ready[g])

    for (uint32_t ci = 0; ci < nch[g]; ++ci) {
      const int64_t chain = 2 * ((int64_t)blockIdx.x + (int64_t)ci * gridDim.x) + g;

      // write 4 consecutive particle values as the hi / lo pair
      auto put4 = [&](uint32_t off, float4 x) {
        float4 hi;
        hi.x = tf32_rn(x.x); hi.y = tf32_rn(x.y); hi.z = tf32_rn(x.z); hi.w = tf32_rn(x.w);
        *reinterpret_cast<float4*>(Ahi + off) = hi;
        *reinterpret_cast<float4*>(Alo + off) = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
      };
      auto get4 = [&](uint32_t off) {
        const float4 hi = *reinterpret_cast<const float4*>(Ahi + off);
        const float4 lo = *reinterpret_cast<const float4*>(Alo + off);
        return make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
      };
      // global [N][du] <-> operands, row-major linear over `nthr` threads (coalesced on the global side)
      auto load_particles = [&](const float* src, int row_stride, int t0, int nthr) {  // row_stride = 0: every row is src[0..du)
        for (int t = t0; t < N * ncg; t += nthr) {
          const int r = t / ncg, cg = t - r * ncg;
          put4(a_off(lbo, r, cg), *reinterpret_cast<const float4*>(src + (size_t)r * row_stride + 4 * cg));
        }
      };
      auto store_particles = [&](float* dst, int t0, int nthr) {
        for (int t = t0; t < N * ncg; t += nthr) {
          const int r = t / ncg, cg = t - r * ncg;
          *reinterpret_cast<float4*>(dst + (size_t)r * du + 4 * cg) = get4(a_off(lbo, r, cg));
        }
      };
      // transition noise of tasks [I0, I1): element (row, column) of normal(key, (N, du)) scaled
      auto make_noise = [&](Key ktr, float scale, auto I0, auto I1) {
#pragma unroll
        for (int i = decltype(I0)::value; i < decltype(I1)::value; ++i) {
          if (task[i] != 0xFFFFFFFFu) {
            const uint32_t b = (task[i] & 0xFFFFu) * du + 4u * (task[i] >> 16);
            const Noise8 r = noise_task(ktr.k0, ktr.k1, b, hblk, scale);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              nz[0][4 * i + c] = r.lo[c];
              nz[1][4 * i + c] = r.hi[c];
            }
          }
        }
      };
      using I_0 = std::integral_constant<int, 0>;
      using I_1 = std::integral_constant<int, NE1>;
      using I_2 = std::integral_constant<int, NE2>;
      using I_3 = std::integral_constant<int, NT>;

      // E warps: stage the per-chain step vectors of workspace slot `slot` (global loads first, shared stores later so
      // that the load latency hides behind the noise)
      float cv_reg[2];
      auto load_cvs = [&](int slot) {
        const float* wsrow = p.ws + ((size_t)chain * (K + 1) + slot) * (size_t)(du + (dv + 3) / 4 * 4);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int t = gt + 128 * j;
          float x = 0.f;
          if (t < L.du8) {
            if (t < du) x = __ldg(wsrow + t);
          } else if (t - L.du8 < dv) {
            x = __ldg(wsrow + du + (t - L.du8));
          }
          cv_reg[j] = x;
        }
      };
      auto store_cvs = [&]() {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (gt + 128 * j < L.du8 + L.dv8) cvs[gt + 128 * j] = cv_reg[j];
      };
      // E warps, one particle row per thread.  v-half of the accumulator -> lwraw (per-row Gaussian log-likelihood)
      const int erow = 32 * (warp & 3) + lane;  // TMEM lane quadrant this warp may access
      const uint32_t trow = tmem_g + ((uint32_t)(32 * (warp & 3)) << 16);
      auto epilogue_v = [&](int k) {
        const float dt = __ldg(p.dt + k), sdk = __ldg(p.sd + k), lognorm = __ldg(p.lognorm + k);
        const float inv_var = 1.0f / (sdk * sdk);
        float ss = 0.f;
        const float* cv = cvs + L.du8;
        auto v_chunk = [&](const float* acc, int c0, int nc) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            if (c < nc) {
              const float4 cc = *reinterpret_cast<const float4*>(cv + c0 + c);
              // columns >= dv: cv = 0 and the accumulator column is exactly 0 (zero rows of M)
              const float r0 = cc.x - dt * acc[c + 0], r1 = cc.y - dt * acc[c + 1];
              const float r2 = cc.z - dt * acc[c + 2], r3 = cc.w - dt * acc[c + 3];
              ss = fmaf(r0, r0, ss);
              ss = fmaf(r1, r1, ss);
              ss = fmaf(r2, r2, ss);
              ss = fmaf(r3, r3, ss);
            }
          }
        };
        int c0 = 0;
        for (; c0 + 32 <= L.dv8; c0 += 32) {
          float acc[32];
          tmem_ld32(trow + L.du8 + c0, acc);
          v_chunk(acc, c0, 32);
        }
        for (; c0 < L.dv8; c0 += 8) {
          float acc[32];
          tmem_ld8(trow + L.du8 + c0, acc);
          v_chunk(acc, c0, 8);
        }
        if (erow < N) lwraw[erow] = -0.5f * (ss * inv_var + lognorm);
      };
      // u-half: mean = x + dt (drift + offset) written over the lo operand (in place, own row)
      auto epilogue_u = [&](int k) {
        const float dt = __ldg(p.dt + k);
        auto u_chunk = [&](const float* acc, int c0, int nc) {
          if (erow >= N) return;
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            if (c < nc && c0 + c < du) {
              const uint32_t off = a_off(lbo, erow, (c0 + c) >> 2);
              const float4 hi = *reinterpret_cast<const float4*>(Ahi + off);
              float4 lo = *reinterpret_cast<const float4*>(Alo + off);
              const float4 cu = *reinterpret_cast<const float4*>(cvs + c0 + c);
              lo.x = (hi.x + lo.x) + dt * (acc[c + 0] + cu.x);
              lo.y = (hi.y + lo.y) + dt * (acc[c + 1] + cu.y);
              lo.z = (hi.z + lo.z) + dt * (acc[c + 2] + cu.z);
              lo.w = (hi.w + lo.w) + dt * (acc[c + 3] + cu.w);
              *reinterpret_cast<float4*>(Alo + off) = lo;
            }
          }
        };
        int c0 = 0;
        for (; c0 + 32 <= L.du8; c0 += 32) {
          float acc[32];
          tmem_ld32(trow + c0, acc);
          u_chunk(acc, c0, 32);
        }
        for (; c0 < L.du8; c0 += 8) {
          float acc[32];
          tmem_ld8(trow + c0, acc);
          u_chunk(acc, c0, 8);
        }
      };

      // =============================== initialisation ===============================
      if (gt == 0) {
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
      for (uint32_t t = gt; t < L.a_bytes / 4; t += GTHREADS) {
        reinterpret_cast<float*>(Ahi)[t] = 0.f;
        reinterpret_cast<float*>(Alo)[t] = 0.f;
      }
      bar_all();
      // keys of step k -> slot k & 1.  Step 0 here; step k + 1 by the resampling warp while it waits in step k.
      auto step_keys = [&](int k) {
        const Key key_k = split_key(kbase[0], (uint32_t)K, (uint32_t)k);  // csmc.py:157 / smc.py:154
        Key a, b;
        split2(key_k, a, b);
        Key* dst = skeys + 2 * (k & 1);
        if (p.mode == MODE_CSMC) {  // csmc.py:136: (key_resampling, key_transition)
          dst[0] = a;
          dst[1] = b;
        } else {  // smc.py:142: (key_proposal, key_resampling)
          dst[1] = a;
          dst[0] = b;
        }
      };
      if (gt == GTHREADS - 1) step_keys(0);
      if (p.mode == MODE_PMCMC) {
        load_particles(p.u0s + (size_t)chain * N * du, du, gt, GTHREADS);
      } else if (p.init_mode == FBS_INIT_DEGENERATE) {  // gibbs.py:140-144
        load_particles(p.us_star + (size_t)chain * (K + 1) * du, 0, gt, GTHREADS);
        for (int t = gt; t < N; t += GTHREADS) lw[t] = p.init_log_w;
      } else {  // gibbs.py:133-137
        make_noise(kbase[2], 1.0f, I_0{}, I_3{});
        const int b0 = clamp_index(p.bs_star[(size_t)chain * (K + 1)], N);
        const float* u0 = p.us_star + (size_t)chain * (K + 1) * du;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          if (task[i] != 0xFFFFFFFFu) {
            const int pr = task[i] & 0xFFFFu, cg = task[i] >> 16;
            const float4 ref = *reinterpret_cast<const float4*>(u0 + 4 * cg);
            put4(a_off(lbo, pr, cg), pr == b0 ? ref : make_float4(nz[0][4 * i], nz[0][4 * i + 1], nz[0][4 * i + 2], nz[0][4 * i + 3]));  // csmc.py:152
            put4(a_off(lbo, pr + half, cg),
                 pr + half == b0 ? ref : make_float4(nz[1][4 * i], nz[1][4 * i + 1], nz[1][4 * i + 2], nz[1][4 * i + 3]));
          }
        }
      }
      fence_proxy_async();  // the particles are read by the tensor core (async proxy)
      bar_all();
      if (gt == 0) mbar_arrive(ready + g);

      if (p.mode == MODE_CSMC) {
        if (p.uss) store_particles(p.uss + (size_t)chain * (K + 1) * N * du, gt, GTHREADS);
        if (init_gemm) {
          // gibbs.py:136-137: weights of the initial particles, (v, v_prev) = (vs[0], vs[1]) -> workspace slot K
          if (is_E) {
            load_cvs(K);
            store_cvs();
            mbar_wait_sleep(accum + g, gcount & 1u);
            tc_fence_after();
            bar_E();
            epilogue_v(0);
            tc_fence_before();
          }
          ++gcount;
          bar_all();
          for (int t = gt; t < N; t += GTHREADS) lw[t] = lwraw[t];
          bar_all();
          if (gt == 0) mbar_arrive(ready + g);  // particles unchanged, accumulator drained
        }
        if (is_R) warp_normalise_v3(lw, N, lane);  // csmc.py:155
        bar_all();
        if (p.log_wss)
          for (int t = gt; t < N; t += GTHREADS) p.log_wss[(size_t)chain * (K + 1) * N + t] = lw[t];
      }

      // =============================== the K-step sweep ===============================
      for (int k = 0; k < K; ++k) {
        const int dbg_role = 3 * g + (gw == 0 ? 0 : (gw == 5 ? 1 : (gw == 4 ? 2 : -1)));
        const bool dbg_on = dbg_cta && ci == 0 && dbg_role >= 3 * g && k >= DBG_K0 && k < DBG_K0 + DBG_NS;
        long long* dbg_row = p.dbg + ((size_t)(dbg_on ? dbg_role : 0) * DBG_NS + (dbg_on ? k - DBG_K0 : 0)) * DBG_MAXS;
#define FBS_STAMP(i) do { if (DBG && dbg_on) dbg_row[i] = clock64(); } while (0)
        FBS_STAMP(0);
        if (is_noise) {
          // ---- noise on the CUDA cores || GEMM on the tensor core (first part) and || resampling (second part) ----
          if (is_E) load_cvs(k);
          if (p.mode == MODE_CSMC) {  // the pinned reference particle of this step
            const float* ustar = p.us_star + ((size_t)chain * (K + 1) + k + 1) * du;
            for (int t = nt; t < du; t += NOISE_THREADS) pin[t] = ustar[t];
            if (nt == 0) reinterpret_cast<int*>(pin)[du] = clamp_index(p.bs_star[(size_t)chain * (K + 1) + k + 1], N);
          }
          const Key ktr = skeys[2 * (k & 1) + 1];
          const float sd = __ldg(p.sd + k);
          make_noise(ktr, sd, I_0{}, I_1{});
          FBS_STAMP(1);
          if (is_E) {
            store_cvs();
            mbar_wait_sleep(accum + g, gcount & 1u);  // first pass: every v column
            tc_fence_after();
            FBS_STAMP(2);
            bar_E();         // cvs visible to the four E warps
            FBS_STAMP(3);
            epilogue_v(k);   // log-likelihood -> lwraw
            bar_ER_arrive(); // ... releases the resampling warp
          }
          FBS_STAMP(4);
          make_noise(ktr, sd, I_1{}, I_2{});
          FBS_STAMP(5);
          if (is_E) {
            if (L.nu_pass > 0) mbar_wait_sleep(accum_u + g, gcount & 1u);  // second pass: the u columns
            tc_fence_after();
            epilogue_u(k);   // means -> Alo
            tc_fence_before();
          }
          FBS_STAMP(6);
          make_noise(ktr, sd, I_2{}, I_3{});
          FBS_STAMP(7);
        } else {
          // ---- weights + ancestors (one warp) ----
          const bool fast = p.mode == MODE_PMCMC && (p.scheme == FBS_RESAMPLE_STRATIFIED || p.scheme == FBS_RESAMPLE_SYSTEMATIC);
          float* ubuf = reinterpret_cast<float*>(tmp);
          // everything that does not depend on the weights happens BEFORE the barrier, off the step's critical path:
          // the keys of the next step and (fast path) this step's resampling uniforms
          if (lane == 0 && k + 1 < K) step_keys(k + 1);
          if (fast) {
            const Key kr = skeys[2 * (k & 1)];
            if (p.scheme == FBS_RESAMPLE_SYSTEMATIC) {  // uniform(key, ()) = random_bits(key, 1) word 0
              uint32_t x0 = 0u, x1 = 0u;
              threefry2x32(kr.k0, kr.k1, x0, x1);
              if (lane == 0) ubuf[0] = bits_to_unit(x0);
            } else {
              const uint32_t h = ((uint32_t)N + 1u) >> 1;
              for (uint32_t b = lane; b < h; b += 32) {
                uint32_t c0, c1;
                random_bits_block(kr, N, b, c0, c1);
                ubuf[b] = bits_to_unit(c0);
                if (b + h < (uint32_t)N) ubuf[b + h] = bits_to_unit(c1);
              }
            }
            __syncwarp();
          }
          const bool fastk = p.mode == MODE_CSMC && p.scheme == FBS_RESAMPLE_KILLING;
          if (fastk) {  // the three uniform streams of conditional killing (resamplings.py:66,71,74,84)
            Key k1, k2, k3;
            split3(skeys[2 * (k & 1)], k1, k2, k3);
            const uint32_t h = ((uint32_t)N + 1u) >> 1;
            for (uint32_t b = lane; b < h; b += 32) {
              uint32_t a0, a1, c0, c1;
              random_bits_block(k1, N, b, a0, a1);
              random_bits_block(k2, N, b, c0, c1);
              ub[b] = bits_to_unit(a0);
              ub[ROWS + b] = bits_to_unit(c0);
              if (b + h < (uint32_t)N) {
                ub[b + h] = bits_to_unit(a1);
                ub[ROWS + b + h] = bits_to_unit(c1);
              }
            }
            uint32_t x0 = 0u, x1 = 0u;  // choice(key_3, N, (), p): random_bits(key_3, 1) = block (0, 0), word 0
            threefry2x32(k3.k0, k3.k1, x0, x1);
            if (lane == 0) ub[2 * ROWS] = bits_to_unit(x0);
            __syncwarp();
          }
          if (NR > 0) make_noise(skeys[2 * (k & 1) + 1], __ldg(p.sd + k), I_0{}, std::integral_constant<int, NR>{});
          FBS_STAMP(1);
          bar_ER_sync();
          FBS_STAMP(2);
          const Key kres = skeys[2 * (k & 1)];
          if (fastk) {
            // conditional killing (resamplings.py:40-88) + weights of the resampled parents (csmc.py:139-146), the warp's 4
            // elements per lane (q = lane + 32 j) in registers; the same float operations in the same order as
            // warp_cond_killing / warp_normalise_v3, hence the same bits
            const int32_t* bsp = p.bs_star + (size_t)chain * (K + 1);
            const int ci = clamp_index(bsp[k], N), cj = clamp_index(bsp[k + 1], N);
            const float fn = (float)N;
            float wv[4];
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              wv[j] = 0.f;
              if (q < N) {
                wv[j] = expf(lw[q]);  // csmc.py:139
                w[q] = wv[j];
                m = fmaxf(m, wv[j]);
              }
            }
            const float w_max = warp_max(m);
            // J_prob (:79) goes to the int scratch first so that its SUM (:81) rides in lane 1 of the serial pass that builds
            // cumsum(w) in lane 0 -- one dependent chain of N additions less on the step's critical path (this warp is what
            // the twelve noise warps wait for, scripts/v3_timeline.py); the order of every sum is unchanged
            float* jpf = reinterpret_cast<float*>(tmp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              if (q < N) jpf[q] = (q == ci) ? 0.f : __fdiv_rn(__fsub_rn(1.0f, __fdiv_rn(wv[j], w_max)), fn);
            }
            __syncwarp();
            float sacc2 = 0.f;
            if (lane < 2) {
              const bool l0 = lane == 0;
              const float* src = l0 ? w : jpf;
              int i = 0;
              for (; i + 8 <= N; i += 8) {
                float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
                a.x = sacc2 = __fadd_rn(sacc2, a.x);
                a.y = sacc2 = __fadd_rn(sacc2, a.y);
                a.z = sacc2 = __fadd_rn(sacc2, a.z);
                a.w = sacc2 = __fadd_rn(sacc2, a.w);
                b.x = sacc2 = __fadd_rn(sacc2, b.x);
                b.y = sacc2 = __fadd_rn(sacc2, b.y);
                b.z = sacc2 = __fadd_rn(sacc2, b.z);
                b.w = sacc2 = __fadd_rn(sacc2, b.w);
                if (l0) {
                  *reinterpret_cast<float4*>(cum + i) = a;
                  *reinterpret_cast<float4*>(cum + i + 4) = b;
                }
              }
              for (; i < N; ++i) {
                sacc2 = __fadd_rn(sacc2, src[i]);
                if (l0) cum[i] = sacc2;
              }
            }
            __syncwarp();
            const float total = __shfl_sync(0xffffffffu, sacc2, 0);
            const float jsum = __shfl_sync(0xffffffffu, sacc2, 1);
            float r[4];
            int lo[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              const bool killed = q < N && __fmul_rn(ub[q < N ? q : 0], w_max) >= wv[j];
              r[j] = __fmul_rn(total, __fsub_rn(1.0f, ub[ROWS + (q < N ? q : 0)]));
              lo[j] = 0;
              hi[j] = killed ? N : 0;
              if (!killed) lo[j] = q;
            }
#pragma unroll 1
            for (int it = 0; it < 8; ++it) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (lo[j] < hi[j]) {
                  const int mid = (lo[j] + hi[j]) >> 1;
                  if (cum[mid] < r[j]) lo[j] = mid + 1; else hi[j] = mid;
                }
              }
            }
            __syncwarp();  // every search done before cum is reused
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              if (q < N) {
                cum[q] = (q == ci) ? fmaxf(__fsub_rn(1.0f, jsum), 0.f) : jpf[q];  // :80-82
                tmp[q] = lo[j];  // (the same thread read jpf[q] = this word just above)
              }
            }
            __syncwarp();
            warp_seq_cumsum(cum, cum, N, lane);
            const int J = choice_from_cum(cum, N, ub[2 * ROWS]);  // :84
            int shift = (cj - J) % N;
            if (shift < 0) shift += N;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              v[j] = -INFINITY;
              if (q < N) {
                int src = q - shift;
                if (src < 0) src += N;
                const int a = (q == cj) ? ci : tmp[src];  // roll + pin (:85-86)
                idx[q] = a;
                v[j] = lwraw[a];  // csmc.py:145 on the resampled parents
                if (p.As) p.As[((size_t)chain * K + k) * N + q] = a;
              }
            }
            float mm = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
            mm = warp_max(mm);
            if (!(fabsf(mm) < INFINITY)) mm = 0.f;
            float sacc = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (lane + 32 * j < N) sacc += expf(v[j] - mm);
            sacc = warp_sum_v3(sacc);
            const float lse = logf(sacc) + mm;  // csmc.py:146
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              if (q < N) {
                lw[q] = v[j] - lse;
                if (p.log_wss) p.log_wss[((size_t)chain * (K + 1) + k + 1) * N + q] = v[j] - lse;
              }
            }
          } else if (fast) {
            // pmcmc_filter_step (smc.py:144-148) with the warp's 4 elements per lane (q = lane + 32 j) in registers;
            // same operation order as warp_normalise_v3 / warp_systematic_or_stratified, hence the same bits
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              v[j] = q < N ? lwraw[q] : -INFINITY;  // smc.py:144
            }
            if (p.lw_hist) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (lane + 32 * j < N) p.lw_hist[((size_t)chain * K + k) * N + lane + 32 * j] = v[j];
            }
            float m = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
            m = warp_max(m);
            if (!(fabsf(m) < INFINITY)) m = 0.f;
            float sacc = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (lane + 32 * j < N) sacc += expf(v[j] - m);
            sacc = warp_sum_v3(sacc);
            const float lse = logf(sacc) + m;  // smc.py:145,147
            if (lane == 0) scal[0] = (scal[0] - logN) + lse;  // smc.py:146
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              if (q < N) {
                v[j] -= lse;
                lw[q] = v[j];
                w[q] = expf(v[j]);
              }
            }
            __syncwarp();
            warp_seq_cumsum(w, cum, N, lane);
            // ancestors: clip(searchsorted(cumsum(w), (arange(n) + u) / n), 0, n - 1)   (resampling.py:43-59), the four
            // binary searches of a lane interleaved
            const float fn = (float)N;
            const bool sys = p.scheme == FBS_RESAMPLE_SYSTEMATIC;
            float r[4];
            int lo[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              const float u = ubuf[sys ? 0 : (q < N ? q : 0)];
              r[j] = __fdiv_rn(__fadd_rn((float)q, u), fn);
              lo[j] = 0;
              hi[j] = q < N ? N : 0;
            }
#pragma unroll 1
            for (int it = 0; it < 8; ++it) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (lo[j] < hi[j]) {
                  const int mid = (lo[j] + hi[j]) >> 1;
                  if (cum[mid] < r[j]) lo[j] = mid + 1; else hi[j] = mid;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lane + 32 * j;
              if (q < N) {
                const int id = min(max(lo[j], 0), N - 1);
                idx[q] = id;
                if (p.inds) p.inds[((size_t)chain * K + k) * N + q] = id;
              }
            }
          } else
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
            if (p.As)
              for (int q = lane; q < N; q += 32) p.As[((size_t)chain * K + k) * N + q] = idx[q];
            if (p.log_wss)
              for (int q = lane; q < N; q += 32) p.log_wss[((size_t)chain * (K + 1) + k + 1) * N + q] = lw[q];
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
            if (p.inds)
              for (int q = lane; q < N; q += 32) p.inds[((size_t)chain * K + k) * N + q] = idx[q];
          }
        }
        ++gcount;
        if (!is_noise) FBS_STAMP(7);
        bar_all();  // ancestors, means and noise complete
        FBS_STAMP(8);

        if (is_noise || NR > 0) {
          // ---- children: gather the parents' means, add the noise (in the noise registers) ----
#pragma unroll
          for (int i = 0; i < NT; ++i) {
            if (task[i] != 0xFFFFFFFFu) {
              const int pr = task[i] & 0xFFFFu, cg = task[i] >> 16;
              const float4 m0 = *reinterpret_cast<const float4*>(Alo + a_off(lbo, idx[pr], cg));
              const float4 m1 = *reinterpret_cast<const float4*>(Alo + a_off(lbo, idx[pr + half], cg));
              nz[0][4 * i + 0] += m0.x; nz[0][4 * i + 1] += m0.y; nz[0][4 * i + 2] += m0.z; nz[0][4 * i + 3] += m0.w;
              nz[1][4 * i + 0] += m1.x; nz[1][4 * i + 1] += m1.y; nz[1][4 * i + 2] += m1.z; nz[1][4 * i + 3] += m1.w;
            }
          }
          FBS_STAMP(9);
          bar_noise();  // every mean read before the operands are overwritten
          FBS_STAMP(10);
          {
            const int bj = p.mode == MODE_CSMC ? reinterpret_cast<const int*>(pin)[du] : -1;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
              if (task[i] != 0xFFFFFFFFu) {
                const int pr = task[i] & 0xFFFFu, cg = task[i] >> 16;
                float4 x0 = make_float4(nz[0][4 * i], nz[0][4 * i + 1], nz[0][4 * i + 2], nz[0][4 * i + 3]);
                float4 x1 = make_float4(nz[1][4 * i], nz[1][4 * i + 1], nz[1][4 * i + 2], nz[1][4 * i + 3]);
                if (pr == bj) x0 = *reinterpret_cast<const float4*>(pin + 4 * cg);  // csmc.py:143
                if (pr + half == bj) x1 = *reinterpret_cast<const float4*>(pin + 4 * cg);
                put4(a_off(lbo, pr, cg), x0);
                put4(a_off(lbo, pr + half, cg), x1);
              }
            }
          }
          fence_proxy_async();  // the new particles are read by the tensor core (async proxy) next step
          FBS_STAMP(11);
          bar_noise();
          FBS_STAMP(12);
          if (is_noise && nt == 0 && (k + 1 < K)) mbar_arrive(ready + g);
          // optional history
          if (is_noise) {
            if (p.mode == MODE_CSMC) {
              if (p.uss) store_particles(p.uss + ((size_t)chain * (K + 1) + k + 1) * N * du, nt, NOISE_THREADS);
            } else {
              if (p.us_hist) store_particles(p.us_hist + ((size_t)chain * K + k) * N * du, nt, NOISE_THREADS);
            }
          }
        }
      }

      // =============================== final state ===============================
      if (is_noise) {
        float* dst = p.mode == MODE_CSMC ? p.us_last : p.uT;
        if (dst) store_particles(dst + (size_t)chain * N * du, nt, NOISE_THREADS);
      } else {
        if (p.mode == MODE_CSMC) {
          if (p.log_ws_last)
            for (int t = lane; t < N; t += 32) p.log_ws_last[(size_t)chain * N + t] = lw[t];
        } else {
          if (p.log_ell && lane == 0) p.log_ell[chain] = scal[0];
        }
      }
      bar_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, GROUPS * TMEM_COLS_PER_GROUP);
}

// selftest operand layout: element (row r, column k), LBO = SELFTEST_LBO
__device__ __forceinline__ uint32_t st_off(int r, int k) {
  return (uint32_t)(k >> 2) * SELFTEST_LBO + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

// ---------------------------------------------------------------------------------------------------
// Self-test of the UMMA plumbing (descriptors, operand layout, TMEM read-back): D = A * B^T with the same
// split-TF32 scheme, operands given as float32 row-major A [128, K8] and the packed B image of one step.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bimg,
                                                               int K8, int nout, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nkb = K8 / 8;
  const uint32_t a_bytes = (uint32_t)(K8 / 4) * SELFTEST_LBO;
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
    *reinterpret_cast<float*>(Ahi + st_off(r, k)) = hi;
    *reinterpret_cast<float*>(Alo + st_off(r, k)) = x - hi;
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
      const uint32_t a_hi = smem_u32(Ahi) + (uint32_t)kb * 2u * SELFTEST_LBO, a_lo = smem_u32(Alo) + (uint32_t)kb * 2u * SELFTEST_LBO;
      const uint32_t b_hi = smem_u32(Bst), b_lo = b_hi + blk;
      umma_tf32(tbase, make_desc(a_hi, SELFTEST_LBO, 128), make_desc(b_hi, b_lbo, 128), idesc, kb > 0 ? 1u : 0u);
      umma_tf32(tbase, make_desc(a_lo, SELFTEST_LBO, 128), make_desc(b_hi, b_lbo, 128), idesc, 1u);
      umma_tf32(tbase, make_desc(a_hi, SELFTEST_LBO, 128), make_desc(b_lo, b_lbo, 128), idesc, 1u);
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
  const size_t smem = (size_t)2 * (K8 / 4) * SELFTEST_LBO + (size_t)4 * (nout / 8) * 128 + 64;
  if (smem > 227 * 1024) {
    set_error("umma_selftest: K8=%d too large", K8);
    return FBS_ERR_UNSUPPORTED;
  }
  cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_selftest_kernel<<<1, 128, smem, st>>>(A, Bimg, K8, nout, D);
  return check_launch("umma_selftest_kernel");
}

static long long* g_v3_dbg = nullptr;

// Host: eligibility + launch.  p.MTc is the tensor-core image of the step matrices; p.ws the step-vector workspace.
template <int NE, int NX, int NR, bool DBG>
static cudaError_t launch_v3_nt(cudaStream_t st, int grid, size_t smem, const SweepParams& p, int stages) {
  cudaError_t e = cudaFuncSetAttribute(sweep_v3_kernel<NE, NX, NR, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  sweep_v3_kernel<NE, NX, NR, DBG><<<grid, NTHREADS, smem, st>>>(p, stages);
  return cudaSuccess;
}

int launch_sweep_v3(void* stream, SweepParams& p) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.MTc == nullptr || p.ws == nullptr) return -1;
  if (p.N < 2 || (p.N & 1) || p.N > ROWS) return -1;
  if (p.du % 4 != 0 || p.du < 4) return -1;
  int stages = MAX_STAGES;
  const int flags = (debug_opt(OPT_V3_TWOPASS) == 1 ? 0x100 : 0) | ((debug_opt(OPT_V3_VARIANT) & 1) ? 0 : 0x200);  // v3_variant bit 0: the un-hinted mbarrier polls (A/B)  // two-pass GEMM (v columns first): slower, the small-N MMAs are bound by operand fetch
  Layout L = make_layout(p.N, p.du, p.dv, stages | flags);
  while (L.total > 227 * 1024 && stages > 2) L = make_layout(p.N, p.du, p.dv, --stages | flags);
  if (L.total > 227 * 1024) return -1;
  if (L.nout > TMEM_COLS_PER_GROUP || L.nkb < 1) return -1;
  const int ntasks = (p.N / 2) * L.ncg;
  if (((p.du + 7) / 8 * 8 + (p.dv + 7) / 8 * 8) > 2 * 128) return -1;  // cv_reg: 2 values per E thread
  // task capacity of a group: 128 NE (four E warps) + 64 NX (two X warps)
  if (ntasks > 128 * 6 + 64 * 8) return -1;
  {
    const int rc = launch_stepvec(stream, p);
    if (rc) return rc;
  }
  const int64_t pairs = (p.B + 1) / 2;
  const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
  p.dbg = g_v3_dbg;
  cudaError_t e;
  // noise tasks per lane of the E / X / resampling warps: <6, 7, 2> (capacity 128 * 6 + 64 * 7 + 32 * 2 = 1280 tasks).  The
  // resampling warp generates its two tasks while it would otherwise wait for the weights, which takes one task off every E
  // thread -- the E warps (noise + both epilogues) are the step's critical path.  Measured at the benchmarked shape after the
  // MMA-issue fix (ms per Gibbs sweep): <6, 7, 2> 50.8, <7, 6, 0> 52.8, <7, 7, 0> 53.1, <6, 8, 0> 56.3 (eight tasks per thread
  // push the noise registers into local memory), <5, 10, 0> 62.4, <8, 4, 0> 63.8.  v3_variant bit 1 selects <7, 6, 0> (A/B).
  // (the pMCMC sweep -- stratified resampling, a shorter resampling phase -- measured 25.0 ms with <7, 6, 0>, 25.6 ms with <6, 7, 2>)
  const bool no_r_noise = (debug_opt(OPT_V3_VARIANT) & 2) != 0 || p.mode != MODE_CSMC;
  if (ntasks <= 128 * 2 + 64 * 4) e = launch_v3_nt<2, 4, 0, false>(st, grid, L.total, p, stages | flags);
  else if (p.dbg != nullptr) e = launch_v3_nt<6, 7, 2, true>(st, grid, L.total, p, stages | flags);  // time-stamped build (profiling hook)
  else if (no_r_noise) e = launch_v3_nt<7, 6, 0, false>(st, grid, L.total, p, stages | flags);
  else e = launch_v3_nt<6, 7, 2, false>(st, grid, L.total, p, stages | flags);
  if (e != cudaSuccess) {
    set_error("sweep_v3: cudaFuncSetAttribute(%u B) failed: %s", L.total, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  return check_launch("sweep_v3_kernel");
}

}  // namespace fbs

// Profiling hook: a device buffer of 7 x 4 x 16 int64 into which CTA 0 writes clock64() stamps of its phases for steps 64..67
// of its first chain pair (roles: E / X / R warp of group 0, of group 1, MMA warp); NULL switches it off (default).
// Not thread safe; for scripts/v3_timeline.py only.
extern "C" int fbs_debug_v3_timeline(long long* dev_buf) {
  fbs::g_v3_dbg = dev_buf;
  return FBS_OK;
}
