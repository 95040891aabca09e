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
//   warps 2..5 : epilogue (tcgen05.ld, + bias, + optional fp32 residual, fp32 and / or bf16 store; optional
//                pixel-shuffle addressing 'b h w (h2 w2 c) -> b (h h2) (w w2) c')
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
constexpr int NTHREADS = 192;
constexpr int MAX_STAGES = 6;

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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// shared-memory matrix descriptor: K-major, SWIZZLE_128B (8 rows x 128 bytes atoms, SBO = 1024), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct Params {
  int B, H, W;        // output pixels per sample
  int C0, C1, Cout;   // source channels (multiples of 64; C1 = 0: one source), output channels
  int ntile;          // N tile: multiple of 16, <= 256, divides Cout
  int kh, kw, off_h, off_w;
  int BW, BH, BNb;    // pixel box of one M tile: BW * BH * BNb <= 128, BW == W
  int h_tiles, m_tiles, stages;
  int pixel_shuffle;
  const float* bias;
  const float* residual;
  float* out_f32;
  __nv_bfloat16* out_bf16;
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const Params p) {
  // Persistent: CTA c runs tiles c, c + gridDim.x, ...; the accumulator is double buffered in tensor memory so that the
  // epilogue of tile i overlaps the TMA / MMA main loop of tile i + 1.
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_stage_bytes = (uint32_t)p.ntile * 128u;
  const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  unsigned char* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;   // [2]: accumulator buffer written by the MMA warp
  uint64_t* acc_empty = acc_full + 2;        // [2]: ... drained by the four epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int ctot = p.C0 + p.C1;
  const int kblocks = p.kh * p.kw * (ctot / KBLK);
  const int n_ntiles = p.Cout / p.ntile;
  const int ntiles = p.m_tiles * n_ntiles;
  uint32_t acc_cols = 32;  // columns of one accumulator buffer (power of two >= ntile)
  while (acc_cols < (uint32_t)p.ntile) acc_cols <<= 1;

  if (warp == 0) {
    tmem_alloc(tmem_slot, 2 * acc_cols);
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full + s, 1);
        mbar_init(empty + s, 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(acc_full + i, 1);
        mbar_init(acc_empty + i, 4);  // one arrival per epilogue warp
      }
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t a_bytes = (uint32_t)(KBLK * p.BW * p.BH * p.BNb) * 2u;
      uint32_t s = 0, ph = 1;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int mt = t / n_ntiles, nt = t - mt * n_ntiles;  // N tiles of one M tile are neighbours: the A boxes hit in L2
        const int n0 = (mt / p.h_tiles) * p.BNb, h0 = (mt % p.h_tiles) * p.BH;
        int kcol = 0;  // column of this K-block in the weight matrix
        for (int ty = 0; ty < p.kh; ++ty) {
          for (int tx = 0; tx < p.kw; ++tx) {
            for (int c = 0; c < ctot; c += KBLK, kcol += KBLK) {
              mbar_wait(empty + s, ph);
              unsigned char* dst = ring + (size_t)s * stage_bytes;
              mbar_expect_tx(full + s, a_bytes + b_stage_bytes);
              if (c < p.C0)
                tma_load_4d(dst, &tmA0, full + s, c, p.off_w + tx, h0 + p.off_h + ty, n0);
              else
                tma_load_4d(dst, &tmA1, full + s, c - p.C0, p.off_w + tx, h0 + p.off_h + ty, n0);
              tma_load_2d(dst + A_STAGE_BYTES, &tmB, full + s, kcol, nt * p.ntile);
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
    if (lane == 0) {
      // instruction descriptor: D fp32, A / B bf16, both K-major, N, M
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.ntile >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      uint32_t s = 0, ph = 0;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const uint32_t buf = it & 1u;
        mbar_wait(acc_empty + buf, ((it >> 1) & 1u) ^ 1u);  // passes on a fresh barrier; then waits for the epilogue of tile it - 2
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * acc_cols;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)s * stage_bytes);
          const uint64_t da = make_desc_sw128(a_addr), db = make_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < KBLK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row: +2 in the address field
            umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(empty + s);
          if (++s == (uint32_t)p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        umma_commit(acc_full + buf);
      }
    }
    __syncwarp();
  } else {
    // epilogue: TMEM lane = tile row = pixel
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int w = r % p.BW, hh = (r / p.BW) % p.BH, nn = r / (p.BW * p.BH);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int mt = t / n_ntiles, nt = t - mt * n_ntiles;
      const int n0 = (mt / p.h_tiles) * p.BNb, h0 = (mt % p.h_tiles) * p.BH;
      const int h = h0 + hh, n = n0 + nn;
      const bool valid = r < p.BW * p.BH * p.BNb && h < p.H && n < p.B;
      const uint32_t buf = it & 1u;
      mbar_wait(acc_full + buf, (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + buf * acc_cols + ((uint32_t)(32 * q) << 16);
      const size_t pix = ((size_t)n * p.H + h) * p.W + w;
      for (int c0 = 0; c0 < p.ntile; c0 += 32) {
        float acc[32];
        tmem_ld32(trow + c0, acc);
        if (!valid) continue;
        const int cg = nt * p.ntile + c0;  // first global output channel of this chunk
        const int nc = min(32, p.ntile - c0);
        if (p.bias) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < nc) acc[c] += __ldg(p.bias + cg + c);
        }
        size_t obase;
        if (p.pixel_shuffle) {
          // channel = (h2 * 2 + w2) * Cq + cq  ->  pixel (2h + h2, 2w + w2), channel cq   (fbs/nn/utils.py:53-57)
          const int Cq = p.Cout >> 2;
          const int blk = cg / Cq, cq = cg - blk * Cq;  // a 32-channel chunk never straddles a block (Cq % 32 == 0)
          const int h2 = blk >> 1, w2 = blk & 1;
          obase = (((size_t)n * (2 * p.H) + (2 * h + h2)) * (2 * p.W) + (2 * w + w2)) * Cq + cq;
        } else {
          obase = pix * p.Cout + cg;
        }
        if (p.residual) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            if (c < nc) {
              const float4 rv = *reinterpret_cast<const float4*>(p.residual + obase + c);
              acc[c] += rv.x; acc[c + 1] += rv.y; acc[c + 2] += rv.z; acc[c + 3] += rv.w;
            }
          }
        }
        if (p.out_f32) {
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            if (c < nc) *reinterpret_cast<float4*>(p.out_f32 + obase + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        }
        if (p.out_bf16) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            if (c < nc) {
              __align__(16) __nv_bfloat162 v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = __floats2bfloat162_rn(acc[c + 2 * j], acc[c + 2 * j + 1]);
              *reinterpret_cast<uint4*>(p.out_bf16 + obase + c) = *reinterpret_cast<const uint4*>(v);
            }
          }
        }
      }
      // this warp's TMEM reads of the buffer are complete (tcgen05.wait::ld inside tmem_ld32): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + buf);
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * acc_cols);
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

extern "C" int fbs_nn_conv_bf16(fbs_stream_t s, const fbs_nn_conv_t* a) {
  using namespace fbs::nnconv;
  FBS_REQUIRE(a != nullptr && a->in0 != nullptr && a->weight != nullptr, "nn_conv: null argument");
  FBS_REQUIRE(a->C0 > 0 && a->C0 % KBLK == 0 && a->C1 >= 0 && a->C1 % KBLK == 0, "nn_conv: source channels must be multiples of 64");
  FBS_REQUIRE(a->C1 == 0 || a->in1 != nullptr, "nn_conv: in1 missing");
  FBS_REQUIRE(a->Cout % 16 == 0 && a->Cout >= 16, "nn_conv: Cout must be a multiple of 16");
  FBS_REQUIRE(a->W >= 1 && a->W <= TILE_M && a->H >= 1 && a->B >= 1, "nn_conv: need 1 <= W <= 128");
  FBS_REQUIRE(a->out_f32 != nullptr || a->out_bf16 != nullptr, "nn_conv: no output");
  if (encode_fn() == nullptr) {
    set_error("nn_conv: cuTensorMapEncodeTiled is not available from this driver");
    return FBS_ERR_CUDA;
  }
  Params p;
  p.B = a->B; p.H = a->H; p.W = a->W;
  p.C0 = a->C0; p.C1 = a->C1; p.Cout = a->Cout;
  p.kh = a->kh; p.kw = a->kw; p.off_h = a->off_h; p.off_w = a->off_w;
  p.pixel_shuffle = a->pixel_shuffle;
  p.bias = a->bias; p.residual = a->residual; p.out_f32 = a->out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  // N tile: the largest of 256 / 192 / 128 / 64 / ... that divides Cout
  int ntile = a->Cout;
  if (ntile > 256) {
    ntile = 256;
    while (a->Cout % ntile) ntile -= 16;
  }
  p.ntile = ntile;
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
  p.h_tiles = (a->H + p.BH - 1) / p.BH;
  const int n_tiles = (a->B + p.BNb - 1) / p.BNb;
  p.m_tiles = n_tiles * p.h_tiles;
  // few M tiles (the 7x7 / 14x14 levels): a CTA streams its K loop through ONE SM's L2 port, so prefer narrower N tiles
  // until there are about two tiles per SM
  while (ntile > 64 && ntile % 32 == 0 && (int64_t)p.m_tiles * (a->Cout / ntile) < 2 * sm_count()) ntile /= 2;
  p.ntile = ntile;
  const size_t stage = (size_t)A_STAGE_BYTES + (size_t)ntile * 128;
  int stages = (int)((200 * 1024) / stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) {
    set_error("nn_conv: tile does not fit shared memory");
    return FBS_ERR_UNSUPPORTED;
  }
  p.stages = stages;
  const size_t smem = 1024 + stages * stage + (2 * MAX_STAGES + 4) * 8 + 16;
  CUtensorMap tmA0, tmA1, tmB;
  const int Hin = a->Hin > 0 ? a->Hin : a->H, Win = a->Win > 0 ? a->Win : a->W;
  int rc = make_act_map(&tmA0, a->in0, a->B, Hin, Win, a->C0, p.BW, p.BH, p.BNb);
  if (!rc) rc = make_act_map(&tmA1, a->C1 ? a->in1 : a->in0, a->B, Hin, Win, a->C1 ? a->C1 : a->C0, p.BW, p.BH, p.BNb);
  if (!rc) rc = make_w_map(&tmB, a->weight, a->kh * a->kw * (a->C0 + a->C1), a->Cout, ntile);
  if (rc) {
    set_error("nn_conv: cuTensorMapEncodeTiled failed with CUresult %d", rc);
    return FBS_ERR_CUDA;
  }
  cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("nn_conv: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  const int64_t tiles = (int64_t)p.m_tiles * (a->Cout / ntile);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  conv_gemm_kernel<<<grid, NTHREADS, smem, as_stream(s)>>>(tmA0, tmA1, tmB, p);
  return check_launch("conv_gemm_kernel");
}
