// Implicit-GEMM convolution of the score network on the 5th-generation tensor cores.
//
//   out[n, h, w, :] = bias + sum_{ty, tx, c} in[n, h + off_h + ty, w + off_w + tx, c] * W[ty, tx, c, :]     (NHWC, stride 1)
//
// is the GEMM  D[M = pixels, N = Cout] = A[M, K] * B[N, K]^T  with K = taps x channels.  Nothing is im2col'ed:
// for every (tap, 64-channel chunk) ONE 4-D TMA box load (64 channels x tile width x tile rows x tile samples,
// SWIZZLE_128B) of the bf16 activation tensor, shifted by the tap offset, lands directly in the canonical K-major
// UMMA layout; out-of-bounds pixels (the zero padding) are filled with zeros by the TMA unit.  The channel axis may
// be split over TWO source tensors (the U-Net's skip concatenations are never materialised).
//
//   warp 0 : TMA producer (activation box + weight tile per K-block, ring of mbarrier-guarded stages)
//   warp 1 : MMA issuer   (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 accumulator in tensor memory)
//   warps 2..5 : epilogue (tcgen05.ld, transpose through shared memory so that global stores are row-contiguous, + bias,
//                + optional fp32 residual, fp32 and / or bf16 store; optional pixel-shuffle addressing
//                'b h w (h2 w2 c) -> b (h h2) (w w2) c')
//
// Reference: flax.linen.Conv call sites of fbs/nn/unet.py (3x3 / 1x1 convolutions :50,68,70,97-124,165,183,205,219,242,
// 317,351,363); the 4x4 stride-2 Downsample (:50) is run as a 2x2 convolution on a space-to-depth copy (nn_ops.cu).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include "fbs_common.cuh"

namespace fbs {
namespace nnconv {

constexpr int TILE_M = 128;
constexpr int KBLK = 64;                       // bf16 channels per K-block = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = TILE_M * 128;    // 16 KB
constexpr int N_EPI_WARPS = 8;
constexpr int MAX_ACC = 4;                     // accumulator buffers in tensor memory (even: buffer b belongs to epilogue group b & 1)
constexpr int NTHREADS = 32 * (2 + N_EPI_WARPS);
constexpr int MAX_STAGES = 6;
constexpr int STG_WARP_BYTES = 32 * 128;       // one epilogue warp's transpose buffer: 32 tile rows x 32 fp32 columns
constexpr int RING_BUDGET = 192 * 1024;        // operand ring (+ resident weights); 227 KB - staging - barriers - alignment

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// explicit shared-window accesses for the epilogue's transpose buffer (through the carved-up dynamic buffer the compiler only
// sees generic pointers: LD.E / ST.E on the global path, ~200 cycles per dependent round trip)
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
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
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of the accumulator -> 32 registers per thread; asynchronous until tmem_ld_wait, which takes the
// registers as in/out operands so that the compiler cannot read (or copy) them before the wait
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B (8 rows x 128 bytes atoms, SBO = 1024), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct Params {
  int B, H, W;        // output pixels per sample
  int C0, C1, Cout;   // source channels (multiples of 64; C1 = 0: one source), output channels
  int ntile;          // N tile: multiple of 16, <= 256, divides Cout
  int kh, kw, off_h, off_w;
  int BW, BH, BNb;    // pixel box of one M tile: BW * BH * BNb <= 128; BW == W (per-tap boxes) or W + 2 (haloed box)
  int h_tiles, m_tiles, stages;
  int pixel_shuffle;
  int halo;           // 1: haloed activation box + resident weights (3x3, stride 1, "same"), see the kernel comment
  uint32_t a_stage_bytes, a_tx_bytes;  // ring pitch / bytes one activation box load delivers
  uint32_t wres_bytes;                  // resident weight image in front of the ring (halo mode)
  const float* bias;
  const float* residual;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  float* gn_partials;  // GroupNorm partial sums of the output (host: only single-sample tiles, no residual / pixel shuffle), or NULL
  long long* dbg;     // profiling hook (fbs_debug_conv_timeline): CTA 0 stamps clock64() at its phase boundaries; NULL = off
};

#define CONV_STAMP(i)                                                                                        \
  do {                                                                                                       \
    if (DBG && p.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (i) < 256) p.dbg[(i)] = clock64(); \
  } while (0)

// Copy-out of one 32-column chunk: this lane's 16-byte pieces of rows 4 i + sub (+ bias, + optional fp32 residual), as fp32
// and / or bf16.  Specialised on the outputs: with run-time pointers every store dragged its own null tests, predicated
// address arithmetic and parameter loads along, and the epilogue -- ~540 instructions per chunk and warp -- was plainly
// instruction bound.
// STATS: also the sums and sums of squares of this lane's four columns over the warp's valid rows -- the statistics of the
// GroupNorm that follows the convolution (flax's GroupNorm forms var = E[x^2] - E[x]^2 from exactly these two means) -- reduced
// over the warp's rows and written to part[colq] (float2 per channel quad); fixed order: deterministic.
template <bool F32, bool BF16, bool RES, bool STATS>
__device__ __forceinline__ void store_rows(const float4 (&v)[8], const uint32_t (&rowbase)[8], uint32_t col4, bool col_ok, float4 b4,
                                           const float* __restrict__ res_p, float* __restrict__ out32_p,
                                           __nv_bfloat16* __restrict__ out16_p, float2* __restrict__ part, int lane) {
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (rowbase[i] != 0xFFFFFFFFu && col_ok) {
      const uint32_t e4 = rowbase[i] + col4;  // in units of 4 elements
      float4 y = make_float4(v[i].x + b4.x, v[i].y + b4.y, v[i].z + b4.z, v[i].w + b4.w);
      if (STATS) {
        s += (y.x + y.y) + (y.z + y.w);
        ss += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
      }
      if (RES) {
        const float4 rv = reinterpret_cast<const float4*>(res_p)[e4];
        y.x += rv.x; y.y += rv.y; y.z += rv.z; y.w += rv.w;
      }
      if (F32) reinterpret_cast<float4*>(out32_p)[e4] = y;
      if (BF16) {
        __align__(8) __nv_bfloat162 o[2] = {__floats2bfloat162_rn(y.x, y.y), __floats2bfloat162_rn(y.z, y.w)};
        reinterpret_cast<uint2*>(out16_p)[e4] = *reinterpret_cast<const uint2*>(o);
      }
    }
  }
  if (STATS) {
    // lanes sub = 0..3 hold rows 4 i + sub of the same four columns
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    ss += __shfl_xor_sync(0xffffffffu, ss, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    ss += __shfl_xor_sync(0xffffffffu, ss, 16);
    if (lane < 8 && col_ok) part[lane] = make_float2(s, ss);
  }
}

// Two operand schemes share the kernel.
//
// per-tap (any kernel size): K-block = (tap, 64-channel chunk); ring stage = activation box shifted by the tap + the weight tile.
//
// haloed (3x3 "same", resident weights of an N tile of 64 or 32): a 3x3 layer at 28x28 re-read its input nine times through L2 and its
// weights once per M tile -- 198 KB per 128 x 64 tile, which bound the layer by the SM's L2 port, not by the tensor pipe.
// Here the box carries a one-pixel halo (BH + 2 rows of W + 2 pixels: the row pitch in shared memory is W + 2 pixels) and
// is loaded ONCE per 64-channel chunk; M row m = hh * (W + 2) + w, so that tap (ty, tx) is the SAME shared-memory image
// read from (ty * (W + 2) + tx) rows further on -- an operand-descriptor start address, nothing moves (SWIZZLE_128B is a
// function of the absolute address bits, which a whole-row shift preserves).  The w >= W rows of the tile are junk and never
// stored.  The weights of the CTA's N tile (9 taps x chunks x 8 or 4 KB) are loaded once per CTA and stay resident; a CTA keeps
// its N tile for all its M tiles.
template <bool DBG>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const Params p) {
  // Persistent: a CTA keeps N tile blockIdx.x % n_ntiles and runs M tiles blockIdx.x / n_ntiles, + gridDim.x / n_ntiles, ...;
  // the accumulator is double buffered in tensor memory so that the epilogue of tile i overlaps the TMA / MMA main loop of
  // tile i + 1.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // warp index through a shuffle: provably warp-uniform for the compiler, so that the role branches below are uniform
  // control flow and the TMA / MMA operands can be kept in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) CONV_STAMP(0);
  const uint32_t b_stage_bytes = p.halo ? 0u : (uint32_t)p.ntile * 128u;
  const uint32_t stage_bytes = p.a_stage_bytes + b_stage_bytes;
  unsigned char* wres = smem;                   // halo mode: [tap][chunk][64 x 128 bytes]
  unsigned char* ring = smem + p.wres_bytes;
  unsigned char* stg_base = ring + (size_t)p.stages * stage_bytes;  // epilogue staging: 8 warps x (32 rows x 128 bytes)
  uint64_t* full = reinterpret_cast<uint64_t*>(stg_base + N_EPI_WARPS * STG_WARP_BYTES);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;    // [MAX_ACC]: accumulator buffer written by the MMA warp
  uint64_t* acc_empty = acc_full + MAX_ACC;   // [MAX_ACC]: ... drained by the four epilogue warps of the buffer's group
  uint64_t* wbar = acc_empty + MAX_ACC;       // resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int ctot = p.C0 + p.C1;
  const int cblocks = ctot / KBLK;
  const int kblocks = p.kh * p.kw * cblocks;
  const int n_ntiles = p.Cout / p.ntile;
  const int nt = blockIdx.x % n_ntiles;
  const int mt_first = blockIdx.x / n_ntiles, mt_step = gridDim.x / n_ntiles;  // host: gridDim.x % n_ntiles == 0
  uint32_t acc_cols = 32;  // columns of one accumulator buffer (power of two >= ntile)
  while (acc_cols < (uint32_t)p.ntile) acc_cols <<= 1;
  // accumulator buffers in tensor memory (512 columns): the MMA warp runs up to nacc tiles ahead of the epilogue
  const uint32_t nacc = 512u / acc_cols < (uint32_t)MAX_ACC ? 512u / acc_cols : (uint32_t)MAX_ACC;

  if (warp == 0) {
    tmem_alloc(tmem_slot, nacc * acc_cols);
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, 1);
      }
      for (int i = 0; i < MAX_ACC; ++i) {
        mbar_init(acc_full + i, 1);
        mbar_init(acc_empty + i, N_EPI_WARPS / 2);  // one arrival per epilogue warp of the group that owns the buffer
      }
      mbar_init(wbar, 1);
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (threadIdx.x == 0) CONV_STAMP(1);

  // The producer and the MMA issuer run their loops with ALL lanes of their warp and predicate only the asynchronous
  // instructions on one lane: addresses and descriptors then live in uniform registers.  (Inside an `if (lane == 0)` region
  // every TMA / MMA operand was moved vector -> uniform register with its own scoreboard wait: ~70 cycles per MMA issued,
  // twice the 32 cycles the tensor pipe needs for a 128 x 64 x 16 MMA -- the 3x3 layers were bound by that.)
  const bool leader = elect_one();  // (evaluated by every warp in converged code; only warps 0 and 1 use it)
  if (warp == 0) {
    uint32_t s = 0, ph = 1;
    int dbg_it = 0;
    if (p.halo) {
      if (leader) mbar_expect_tx(wbar, p.wres_bytes);
      const uint32_t wtile = (uint32_t)p.ntile * 128u;
      // the weight map is 3-D here, (channel, output channel, tap): ONE copy per 64-channel chunk delivers all nine taps as
      // [tap][N tile][64 channels] (nine 2-D copies cost ~180 cycles of issue each at the head of every launch)
      for (int cb = 0; cb < cblocks; ++cb)
        if (leader) tma_load_3d(wres + (size_t)cb * 9u * wtile, &tmB, wbar, cb * KBLK, nt * p.ntile, 0);
      CONV_STAMP(2);
      for (int mt = mt_first; mt < p.m_tiles; mt += mt_step) {
        const int n0 = mt / p.h_tiles, h0 = (mt % p.h_tiles) * p.BH;
        for (int c = 0; c < ctot; c += KBLK) {
          mbar_wait(empty + s, ph);
          if (c == 0) {
            CONV_STAMP(8 + 8 * dbg_it);
            ++dbg_it;
          }
          unsigned char* dst = ring + (size_t)s * stage_bytes;
          if (leader) {
            mbar_expect_tx(full + s, p.a_tx_bytes);
            if (c < p.C0)
              tma_load_4d(dst, &tmA0, full + s, c, -1, h0 - 1, n0);
            else
              tma_load_4d(dst, &tmA1, full + s, c - p.C0, -1, h0 - 1, n0);
          }
          if (++s == (uint32_t)p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    } else {
      for (int mt = mt_first; mt < p.m_tiles; mt += mt_step) {
        const int n0 = (mt / p.h_tiles) * p.BNb, h0 = (mt % p.h_tiles) * p.BH;
        int kcol = 0;  // column of this K-block in the weight matrix
        for (int ty = 0; ty < p.kh; ++ty) {
          for (int tx = 0; tx < p.kw; ++tx) {
            for (int c = 0; c < ctot; c += KBLK, kcol += KBLK) {
              mbar_wait(empty + s, ph);
              if (kcol == 0) {
                CONV_STAMP(8 + 8 * dbg_it);
                ++dbg_it;
              }
              unsigned char* dst = ring + (size_t)s * stage_bytes;
              if (leader) {
                mbar_expect_tx(full + s, p.a_tx_bytes + b_stage_bytes);
                if (c < p.C0)
                  tma_load_4d(dst, &tmA0, full + s, c, p.off_w + tx, h0 + p.off_h + ty, n0);
                else
                  tma_load_4d(dst, &tmA1, full + s, c - p.C0, p.off_w + tx, h0 + p.off_h + ty, n0);
                tma_load_2d(dst + p.a_stage_bytes, &tmB, full + s, kcol, nt * p.ntile);
              }
              if (++s == (uint32_t)p.stages) {
                s = 0;
                ph ^= 1u;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // instruction descriptor: D fp32, A / B bf16, both K-major, N, M
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.ntile >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
    uint32_t s = 0, ph = 0;
    uint32_t it = 0;
    if (p.halo) {
      mbar_wait(wbar, 0);
      tc_fence_after();
    }
    CONV_STAMP(3);
    const uint32_t ring_u = smem_u32(ring), wres_u = smem_u32(wres);
    const uint32_t wtile16 = (uint32_t)p.ntile * 8u;  // one resident weight tile, in descriptor units of 16 bytes
    for (int mt = mt_first; mt < p.m_tiles; mt += mt_step, ++it) {
      const uint32_t buf = it % nacc;
      mbar_wait(acc_empty + buf, ((it / nacc) & 1u) ^ 1u);  // passes on a fresh barrier; then waits for the epilogue of tile it - nacc
      tc_fence_after();
      CONV_STAMP(8 + 8 * it + 1);
      const uint32_t d_tmem = tmem_base + buf * acc_cols;
      if (p.halo) {
        for (int cb = 0; cb < cblocks; ++cb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          if (cb == 0) CONV_STAMP(8 + 8 * it + 2);
          // descriptors differ from a base one only in the 14-bit start-address field (no carry: shared memory < 256 KB)
          const uint64_t da0 = make_desc_sw128(ring_u + s * stage_bytes);
          const uint64_t db0 = make_desc_sw128(wres_u) + (uint64_t)((uint32_t)cb * 9u * wtile16);  // resident weights: [chunk][tap]
          uint32_t acc_flag = cb ? 1u : 0u;
#pragma unroll
          for (int ty = 0; ty < 3; ++ty) {
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
              // tap (ty, tx) = the haloed image read (ty * pitch + tx) rows further on: + 8 descriptor units per row
              const uint64_t da = da0 + (uint64_t)((uint32_t)(ty * p.BW + tx) * 8u);
              const uint64_t db = db0 + (uint64_t)((uint32_t)(ty * 3 + tx) * wtile16);
#pragma unroll
              for (int k = 0; k < KBLK / 16; ++k) {
                if (leader) umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc_flag);
                acc_flag = 1u;
              }
            }
          }
          if (leader) umma_commit(empty + s);
          if (++s == (uint32_t)p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      } else {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          if (kb == 0) CONV_STAMP(8 + 8 * it + 2);
          const uint32_t a_addr = ring_u + s * stage_bytes;
          const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + p.a_stage_bytes);
#pragma unroll
          for (int k = 0; k < KBLK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row: +2 in the address field
            if (leader) umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          if (leader) umma_commit(empty + s);
          if (++s == (uint32_t)p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
      if (leader) umma_commit(acc_full + buf);
      CONV_STAMP(8 + 8 * it + 3);
    }
    __syncwarp();
  } else {
    // epilogue: TMEM lane = tile row = pixel.  A thread owns one ROW of the accumulator, so storing it directly would be 32
    // scattered 16-byte requests per warp instruction (the store-request rate, not the bytes, bound the 1x1 layers); instead
    // each warp transposes 32 rows x 32 columns through its own shared-memory buffer (16-byte pieces XOR-swizzled by the row:
    // conflict-free both ways) and writes rows out with 8 lanes per 128-byte row segment.
    // Two groups of four warps (a warp may only read TMEM lanes 32 * (warp % 4) ...) take alternate tiles, so that one group's
    // tensor-memory reads (64 bytes per cycle per SM: 512 cycles for a 128 x 64 tile) overlap the other's stores; inside a
    // warp the tcgen05.ld of the next 32-column chunk is in flight while the current one is written out.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    // the output pointers in registers: left in the parameter bank they cost two or three dependent constant loads in front
    // of EVERY store (the loop below is predicated, the compiler does not hoist them): ~230 cycles per 16-byte row piece
    const float* const res_p = p.residual;
    float* const out32_p = p.out_f32;
    __nv_bfloat16* const out16_p = p.out_bf16;
    const float* const bias_p = p.bias;
    float* const gn_p = p.gn_partials;
    const int ntile = p.ntile;
    const int r = 32 * q + lane;
    const int w = r % p.BW, hh = (r / p.BW) % p.BH, nn = r / (p.BW * p.BH);
    const uint32_t stg = smem_u32(stg_base + (warp - 2) * STG_WARP_BYTES);
    const int piece = lane & 7, sub = lane >> 3;
    const int Cq = p.Cout >> 2;
    // byte offsets inside the transpose buffer: own row (written), rows 4 i + sub (read); pieces XOR-swizzled by row & 7
    uint32_t st_addr[8];  // own row, pieces XOR-swizzled
#pragma unroll
    for (int j = 0; j < 8; ++j) st_addr[j] = stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
    const int out_mode = (out32_p ? 1 : 0) | (out16_p ? 2 : 0) | (res_p ? 4 : 0);
    const uint32_t ld_even = stg + (uint32_t)sub * 128u + (uint32_t)((piece ^ sub) << 4);        // rows 8 j + sub
    const uint32_t ld_odd = stg + (uint32_t)(4 + sub) * 128u + (uint32_t)((piece ^ (4 + sub)) << 4);  // rows 8 j + 4 + sub
    for (uint32_t it = (uint32_t)grp; mt_first + (int)it * mt_step < p.m_tiles; it += 2) {
      const int mt = mt_first + (int)it * mt_step;
      const int n0 = (mt / p.h_tiles) * p.BNb, h0 = (mt % p.h_tiles) * p.BH;
      const int h = h0 + hh, n = n0 + nn;
      const bool valid = r < p.BW * p.BH * p.BNb && w < p.W && h < p.H && n < p.B;
      // row base in units of 4 elements, without the chunk's channel offset (Cout % 16 == 0)
      uint32_t mybase = 0xFFFFFFFFu;
      if (valid)
        mybase = p.pixel_shuffle ? (uint32_t)(((((size_t)n * (2 * p.H) + 2 * h) * (2 * p.W) + 2 * w) * Cq) >> 2)
                                 : (uint32_t)(((((size_t)n * p.H + h) * p.W + w) * p.Cout) >> 2);
      uint32_t rowbase[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rowbase[i] = __shfl_sync(0xffffffffu, mybase, 4 * i + sub);
      // GroupNorm partials of this (sample, row tile, lane quadrant): [sample][slot = 4 row tile + q][Cout / 4] float2
      float2* const part_tile =
          gn_p ? reinterpret_cast<float2*>(gn_p) + ((size_t)(mt / p.h_tiles) * (4 * p.h_tiles) + 4 * (mt % p.h_tiles) + q) * (size_t)Cq : nullptr;
      const uint32_t buf = it % nacc;
      mbar_wait(acc_full + buf, (it / nacc) & 1u);
      tc_fence_after();
      if (q == 2 && lane == 0) CONV_STAMP(8 + 8 * it + 4);
      const uint32_t trow = tmem_base + buf * acc_cols + ((uint32_t)(32 * q) << 16);
      uint32_t acc[32];
      tmem_ld32_issue(trow, acc);
      for (int c0 = 0; c0 < ntile; c0 += 32) {
        tmem_ld_wait(acc);
        if (q == 2 && c0 == 0) CONV_STAMP(8 + 8 * it + 6);
#pragma unroll
        for (int j = 0; j < 8; ++j) sts_v4(st_addr[j], acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        if (c0 + 32 < ntile) tmem_ld32_issue(trow + c0 + 32, acc);  // (the stores above have read their registers)
        const int cg = nt * ntile + c0;  // first global output channel of this chunk
        const bool col_ok = 4 * piece < ntile - c0;
        uint32_t coff = (uint32_t)cg;  // element offset of the chunk inside a row
        if (p.pixel_shuffle) {
          // channel = (h2 * 2 + w2) * Cq + cq  ->  pixel (2h + h2, 2w + w2), channel cq   (fbs/nn/utils.py:53-57)
          const int blk = cg / Cq, cq = cg - blk * Cq;  // a 32-channel chunk never straddles a block (Cq % 32 == 0)
          coff = (uint32_t)((((blk >> 1) * 2 * p.W) + (blk & 1)) * Cq + cq);
        }
        const uint32_t col4 = (coff >> 2) + (uint32_t)piece;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias_p && col_ok) b4 = __ldg(reinterpret_cast<const float4*>(bias_p + cg + 4 * piece));
        __syncwarp();
        float4 v[8];  // all eight row pieces first: independent shared loads, then the stores
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lds_v4(((i & 1) ? ld_odd : ld_even) + (uint32_t)(i >> 1) * 1024u);
        float2* const part = part_tile ? part_tile + (cg >> 2) : nullptr;  // this chunk's channel quads
        switch (part_tile ? 8 + out_mode : out_mode) {
          case 1: store_rows<true, false, false, false>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          case 2: store_rows<false, true, false, false>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          case 3: store_rows<true, true, false, false>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          case 5: store_rows<true, false, true, false>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          case 6: store_rows<false, true, true, false>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          case 7: store_rows<true, true, true, false>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          case 9: store_rows<true, false, false, true>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          case 10: store_rows<false, true, false, true>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
          default: store_rows<true, true, false, true>(v, rowbase, col4, col_ok, b4, res_p, out32_p, out16_p, part, lane); break;
        }
        if (q == 2 && c0 == 0) CONV_STAMP(8 + 8 * it + 7);
        __syncwarp();
      }
      // this warp's TMEM reads of the buffer are complete (the last tmem_ld_wait): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + buf);
      if (q == 2 && lane == 0) CONV_STAMP(8 + 8 * it + 5);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) CONV_STAMP(4);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, nacc * acc_cols);
  }
}

// ---- host: tensor maps -------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

static int make_act_map(CUtensorMap* tm, const void* base, int B, int H, int W, int C, int BW, int BH, int BNb) {
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)KBLK, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BNb};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
// halo mode: the weight matrix [Cout][9 taps x ctot] seen as (channel, output channel, tap); box = 64 channels x N tile x 9 taps
static int make_w_map3(CUtensorMap* tm, const void* base, int ctot, int Cout, int ntile) {
  const cuuint64_t dims[3] = {(cuuint64_t)ctot, (cuuint64_t)Cout, 9};
  const cuuint64_t strides[2] = {(cuuint64_t)9 * ctot * 2, (cuuint64_t)ctot * 2};
  const cuuint32_t box[3] = {(cuuint32_t)KBLK, (cuuint32_t)ntile, 9};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
static int make_w_map(CUtensorMap* tm, const void* base, int K, int Cout, int ntile) {
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)KBLK, (cuuint32_t)ntile};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace nnconv
}  // namespace fbs

using namespace fbs;

static long long* g_conv_dbg = nullptr;
// Profiling hook: a device buffer of 256 int64 into which CTA 0 of every following convolution launch writes clock64()
// stamps ([0] entry, [1] prologue done, [2] weights issued, [3] weights landed, [4] exit; per tile i at 8 + 8 i: producer
// first load, MMA got the accumulator, first operands landed, all MMAs issued, epilogue start, epilogue end); NULL = off.
// Not thread safe; for scripts/conv_timeline.py only.
extern "C" int fbs_debug_conv_timeline(long long* dev_buf) {
  g_conv_dbg = dev_buf;
  return FBS_OK;
}

// Validation and tiling of one convolution call (shared by the launch and by fbs_nn_conv_gn_layout).
static int plan_conv(const fbs_nn_conv_t* a, fbs::nnconv::Params& p, size_t& smem_out) {
  using namespace fbs::nnconv;
  FBS_REQUIRE(a != nullptr && a->in0 != nullptr && a->weight != nullptr, "nn_conv: null argument");
  FBS_REQUIRE(a->C0 > 0 && a->C0 % KBLK == 0 && a->C1 >= 0 && a->C1 % KBLK == 0, "nn_conv: source channels must be multiples of 64");
  FBS_REQUIRE(a->C1 == 0 || a->in1 != nullptr, "nn_conv: in1 missing");
  FBS_REQUIRE(a->Cout % 16 == 0 && a->Cout >= 16, "nn_conv: Cout must be a multiple of 16");
  FBS_REQUIRE(a->W >= 1 && a->W <= TILE_M && a->H >= 1 && a->B >= 1, "nn_conv: need 1 <= W <= 128");
  FBS_REQUIRE(a->out_f32 != nullptr || a->out_bf16 != nullptr, "nn_conv: no output");
  p.B = a->B; p.H = a->H; p.W = a->W;
  p.C0 = a->C0; p.C1 = a->C1; p.Cout = a->Cout;
  p.kh = a->kh; p.kw = a->kw; p.off_h = a->off_h; p.off_w = a->off_w;
  p.pixel_shuffle = a->pixel_shuffle;
  p.bias = a->bias; p.residual = a->residual; p.out_f32 = a->out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.dbg = g_conv_dbg;
  p.gn_partials = a->gn_partials;
  const int Hin = a->Hin > 0 ? a->Hin : a->H, Win = a->Win > 0 ? a->Win : a->W;
  const int ctot = a->C0 + a->C1;
  // N tile: the largest of 256 / 192 / 128 / 64 / ... that divides Cout
  int ntile = a->Cout;
  if (ntile > 256) {
    ntile = 256;
    while (a->Cout % ntile) ntile -= 16;
  }
  if (p.pixel_shuffle) FBS_REQUIRE((a->Cout / 4) % 32 == 0, "nn_conv: pixel shuffle needs Cout / 4 to be a multiple of 32");
  // M tile: full rows; as many rows (then samples) as fit 128 pixels
  p.BW = a->W;
  p.BH = TILE_M / a->W;
  if (p.BH > a->H) p.BH = a->H;
  p.BNb = 1;
  if (p.BH == a->H) {
    p.BNb = TILE_M / (a->W * a->H);
    if (p.BNb > a->B) p.BNb = a->B;
    if (p.BNb < 1) p.BNb = 1;
  }
  // haloed scheme (kernel comment): 3x3 "same" layers with up to 128 output channels whose N-tile weights (N tile 64) fit
  // beside two activation boxes, and whose tiles keep at least 3/4 of the pixels the per-tap tiling would have.  Measured
  // (101 samples, us, haloed / per-tap): 28x28 64->64 12.0 / 14.7, 128->64 16.5 / 21.2; 14x14 128->128 12.1 / 14.4; but
  // 14x14 128->512 18.4 / 16.3 (per-tap runs N tiles of 128: a 128 x 64 x 16 MMA is bound by its shared-memory operand reads,
  // 48 cycles for 32 cycles of tensor work), and 7x7 256->256 with N tile 32 and one sample per tile 25.8 / 14.7
  p.halo = 0;
  p.wres_bytes = 0;
  const int PW = a->W + 2;
  if (debug_opt(OPT_CONV_IMPL) != 1 && a->kh == 3 && a->kw == 3 && a->off_h == -1 && a->off_w == -1 && Hin == a->H && Win == a->W &&
      a->Cout % 32 == 0 && PW <= TILE_M) {
    int hBH = TILE_M / PW;
    if (hBH > a->H) hBH = a->H;
    const size_t a_box = (size_t)(hBH + 2) * PW * 128;
    const size_t a_stage = (a_box + 1023) & ~(size_t)1023;
    // (compared at the un-capped samples-per-tile, so that the choice -- hence the summation order -- does not depend on B)
    const int per_tap_pixels = p.BH == a->H ? a->W * a->H * (TILE_M / (a->W * a->H)) : a->W * p.BH;
    const int hnt = (a->Cout % 64 == 0 && a->Cout <= 128 && (size_t)9 * (ctot / KBLK) * 64 * 128 + 2 * a_stage <= (size_t)RING_BUDGET) ? 64 : 0;
    // rows an MMA may touch past the box (junk rows of the last taps) stay inside the staging area that follows the ring
    if (hBH >= 1 && 4 * hBH * a->W >= 3 * per_tap_pixels && hnt) {
      p.halo = 1;
      ntile = hnt;
      p.BW = PW; p.BH = hBH; p.BNb = 1;
      p.a_stage_bytes = (uint32_t)a_stage;
      p.a_tx_bytes = (uint32_t)a_box;
      p.wres_bytes = (uint32_t)((size_t)9 * (ctot / KBLK) * hnt * 128);
    }
  }
  p.h_tiles = (a->H + p.BH - 1) / p.BH;
  const int n_tiles = (a->B + p.BNb - 1) / p.BNb;
  p.m_tiles = n_tiles * p.h_tiles;
  int stages;
  if (p.halo) {
    stages = (int)((RING_BUDGET - p.wres_bytes) / p.a_stage_bytes);
  } else {
    // N tile among ntile, ntile / 2, ... (>= 64): the one with the least (waves of tiles over the SMs) x (cycles of one
    // 128 x N x 16 MMA = max(tensor pipe N / 2, shared-memory operand reads 32 + N / 4)).  7x7 256 -> 256: 51 M tiles, N = 128
    // (one wave of 102 tiles) 11.1 us against 14.7 us for N = 64
    {
      int best = ntile;
      int64_t best_cost = -1;
      for (int cand = ntile; cand >= 64; cand /= 2) {
        const int64_t tiles = (int64_t)p.m_tiles * (a->Cout / cand);
        const int64_t waves = (tiles + sm_count() - 1) / sm_count();
        const int64_t clk = cand / 2 > 32 + cand / 4 ? cand / 2 : 32 + cand / 4;
        if (best_cost < 0 || waves * clk < best_cost) {
          best_cost = waves * clk;
          best = cand;
        }
        if (cand % 32 != 0) break;
      }
      ntile = best;
    }
    p.a_stage_bytes = A_STAGE_BYTES;
    p.a_tx_bytes = (uint32_t)(KBLK * p.BW * p.BH * p.BNb) * 2u;
    stages = (int)(RING_BUDGET / ((size_t)A_STAGE_BYTES + (size_t)ntile * 128));
  }
  p.ntile = ntile;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) {
    set_error("nn_conv: tile does not fit shared memory");
    return FBS_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const size_t stage = (size_t)p.a_stage_bytes + (p.halo ? 0 : (size_t)ntile * 128);
  smem_out = 1024 + p.wres_bytes + stages * stage + N_EPI_WARPS * STG_WARP_BYTES + (2 * MAX_STAGES + 2 * MAX_ACC + 1) * 8 + 16;
  if (p.gn_partials != nullptr)
    FBS_REQUIRE(p.BNb == 1 && !p.pixel_shuffle && p.residual == nullptr,
                "nn_conv: GroupNorm partials need single-sample tiles, no residual and no pixel shuffle (fbs_nn_conv_gn_layout)");
  return FBS_OK;
}

extern "C" int fbs_nn_conv_gn_layout(const fbs_nn_conv_t* a, int32_t* slots_per_sample) {
  using namespace fbs::nnconv;
  FBS_REQUIRE(a != nullptr && slots_per_sample != nullptr, "nn_conv_gn_layout: null argument");
  fbs_nn_conv_t b = *a;
  b.gn_partials = nullptr;
  Params p;
  size_t smem;
  const int rc = plan_conv(&b, p, smem);
  if (rc != FBS_OK) return rc;
  // (the answer must not depend on the batch size: a per-tap tiling that WOULD pack two samples per tile says no even for B = 1)
  const bool packs_samples = !p.halo && 2 * a->H * a->W <= TILE_M;
  *slots_per_sample = (p.BNb == 1 && !packs_samples && !p.pixel_shuffle && p.residual == nullptr) ? 4 * p.h_tiles : 0;
  return FBS_OK;
}

extern "C" int fbs_nn_conv_bf16(fbs_stream_t s, const fbs_nn_conv_t* a) {
  using namespace fbs::nnconv;
  if (encode_fn() == nullptr) {
    set_error("nn_conv: cuTensorMapEncodeTiled is not available from this driver");
    return FBS_ERR_CUDA;
  }
  Params p;
  size_t smem;
  {
    const int rc = plan_conv(a, p, smem);
    if (rc != FBS_OK) return rc;
  }
  const int Hin = a->Hin > 0 ? a->Hin : a->H, Win = a->Win > 0 ? a->Win : a->W;
  const int ctot = a->C0 + a->C1;
  const int ntile = p.ntile;
  CUtensorMap tmA0, tmA1, tmB;
  const int boxH = p.halo ? p.BH + 2 : p.BH;
  int rc = make_act_map(&tmA0, a->in0, a->B, Hin, Win, a->C0, p.BW, boxH, p.BNb);
  if (!rc) rc = make_act_map(&tmA1, a->C1 ? a->in1 : a->in0, a->B, Hin, Win, a->C1 ? a->C1 : a->C0, p.BW, boxH, p.BNb);
  if (!rc) rc = p.halo ? make_w_map3(&tmB, a->weight, ctot, a->Cout, ntile) : make_w_map(&tmB, a->weight, a->kh * a->kw * ctot, a->Cout, ntile);
  if (rc) {
    set_error("nn_conv: cuTensorMapEncodeTiled failed with CUresult %d", rc);
    return FBS_ERR_CUDA;
  }
  cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && p.dbg) e = cudaFuncSetAttribute(conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("nn_conv: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  const int n_ntiles = a->Cout / ntile;
  const int64_t tiles = (int64_t)p.m_tiles * n_ntiles;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  grid -= grid % n_ntiles;  // a CTA keeps one N tile (n_ntiles <= 16 <= the SM count)
  if (grid < n_ntiles) grid = n_ntiles;
  if (p.dbg)
    conv_gemm_kernel<true><<<grid, NTHREADS, smem, as_stream(s)>>>(tmA0, tmA1, tmB, p);
  else
    conv_gemm_kernel<false><<<grid, NTHREADS, smem, as_stream(s)>>>(tmA0, tmA1, tmB, p);
  return check_launch("conv_gemm_kernel");
}
