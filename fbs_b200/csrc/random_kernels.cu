// Standalone jax.random kernels behind the C ABI (include/fbs_b200.h) + library bookkeeping.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "fbs_common.cuh"
#include "fbs_resample.cuh"

namespace fbs {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;
static thread_local int g_sms = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
int sm_count() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

static std::atomic<int> g_opts[OPT_COUNT];
static const char* const g_opt_names[OPT_COUNT] = {"sweep_impl", "sweep_verbose", "step_impl", "step_tc_warps", "stepvec_impl",
                                                    "sweep_g", "v3_twopass", "em_impl", "v3_variant", "conv_impl"};
int debug_opt(DebugOpt which) { return g_opts[which].load(std::memory_order_relaxed); }

enum { OUT_BITS = 0, OUT_UNIFORM = 1, OUT_NORMAL = 2 };

// One thread per threefry block: writes elements b and b + h of each key's stream.
template <int MODE>
__global__ void random_fill_kernel(const uint32_t* __restrict__ keys, int64_t B, uint32_t n, float lo, float hi,
                                   void* __restrict__ out_) {
  const uint32_t h = (n + 1u) >> 1;
  const int64_t total = B * (int64_t)h;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bch = t / h;
    const uint32_t b = (uint32_t)(t - bch * h);
    Key key{keys[2 * bch], keys[2 * bch + 1]};
    uint32_t y0, y1;
    random_bits_block(key, n, b, y0, y1);
    const int64_t base = bch * (int64_t)n;
    if (MODE == OUT_BITS) {
      uint32_t* out = (uint32_t*)out_;
      out[base + b] = y0;
      if (b + h < n) out[base + b + h] = y1;
    } else if (MODE == OUT_UNIFORM) {
      float* out = (float*)out_;
      out[base + b] = bits_to_uniform(y0, lo, hi);
      if (b + h < n) out[base + b + h] = bits_to_uniform(y1, lo, hi);
    } else {
      float* out = (float*)out_;
      out[base + b] = bits_to_normal(y0);
      if (b + h < n) out[base + b + h] = bits_to_normal(y1);
    }
  }
}

// jax.random.randint: two independent 32-bit streams from split(key), combined modulo span.
__global__ void randint_kernel(const uint32_t* __restrict__ keys, int64_t B, uint32_t n, int32_t minval, uint32_t span,
                               uint32_t mult, int32_t* __restrict__ out) {
  const int64_t total = B * (int64_t)n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bch = t / n;
    const uint32_t e = (uint32_t)(t - bch * n);
    Key key{keys[2 * bch], keys[2 * bch + 1]};
    Key k1, k2;
    split2(key, k1, k2);
    const uint32_t hi_bits = random_bits_elem(k1, n, e);
    const uint32_t lo_bits = random_bits_elem(k2, n, e);
    uint32_t off = (hi_bits % span) * mult + (lo_bits % span);  // uint32 wrap-around, as lax.mul/add on uint32
    off %= span;
    out[t] = minval + (int32_t)off;
  }
}

// jax.random.choice(key, N, (n,), p=p): one warp per batch entry; cumsum sequential (lane 0) in smem.
__global__ void choice_kernel(const uint32_t* __restrict__ keys, const float* __restrict__ p, int64_t B, int N, uint32_t n,
                              int32_t* __restrict__ out) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* cum = smem + (size_t)warp * N;
  for (int64_t bch = blockIdx.x * (int64_t)nwarps + warp; bch < B; bch += (int64_t)gridDim.x * nwarps) {
    const float* pb = p + bch * N;
    warp_seq_cumsum(pb, cum, N, lane);  // the contract's summation order (fbs_resample.cuh)
    Key key{keys[2 * bch], keys[2 * bch + 1]};
    const uint32_t h = (n + 1u) >> 1;
    for (uint32_t b = lane; b < h; b += 32) {
      uint32_t y0, y1;
      random_bits_block(key, n, b, y0, y1);
      float r = __fmul_rn(cum[N - 1], __fsub_rn(1.0f, bits_to_unit(y0)));
      out[bch * n + b] = searchsorted_left(cum, N, r);
      if (b + h < n) {
        r = __fmul_rn(cum[N - 1], __fsub_rn(1.0f, bits_to_unit(y1)));
        out[bch * n + b + h] = searchsorted_left(cum, N, r);
      }
    }
    __syncwarp();
  }
}

static int grid_for(int64_t total, int threads) {
  int64_t blocks = (total + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int MODE>
static int launch_fill(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, float lo, float hi, void* out) {
  FBS_REQUIRE(B >= 0 && n >= 0 && n < 0xFFFFFFFFll, "random fill: bad sizes B=%lld n=%lld", (long long)B, (long long)n);
  if (B == 0 || n == 0) return FBS_OK;  // an empty batch is a no-op (its buffers may be NULL)
  FBS_REQUIRE(keys && out, "random fill: null pointer");
  const int64_t total = B * ((n + 1) / 2);
  random_fill_kernel<MODE><<<grid_for(total, 256), 256, 0, as_stream(s)>>>(keys, B, (uint32_t)n, lo, hi, out);
  return check_launch("random_fill_kernel");
}

}  // namespace fbs

using namespace fbs;

extern "C" {

int fbs_version(void) { return 100; }
const char* fbs_last_error(void) { return g_err; }
int64_t fbs_launch_count(void) { return g_launches; }
void fbs_reset_launch_count(void) { g_launches = 0; }

int fbs_debug_set_option(const char* name, int value) {
  FBS_REQUIRE(name != nullptr, "fbs_debug_set_option: null name");
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(name, g_opt_names[i]) == 0) {
      g_opts[i].store(value, std::memory_order_relaxed);
      return FBS_OK;
    }
  set_error("fbs_debug_set_option: unknown option '%s'", name);
  return FBS_ERR_INVALID_ARGUMENT;
}

int fbs_random_bits_u32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, uint32_t* out) {
  return launch_fill<OUT_BITS>(s, keys, B, n, 0.f, 1.f, out);
}

int fbs_random_split(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t num, uint32_t* out) {
  return launch_fill<OUT_BITS>(s, keys, B, 2 * num, 0.f, 1.f, out);
}

int fbs_random_uniform_f32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, float minval, float maxval,
                           float* out) {
  return launch_fill<OUT_UNIFORM>(s, keys, B, n, minval, maxval, out);
}

int fbs_random_normal_f32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, float* out) {
  return launch_fill<OUT_NORMAL>(s, keys, B, n, 0.f, 1.f, out);
}

int fbs_random_randint_i32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, int32_t minval, int32_t maxval,
                           int32_t* out) {
  FBS_REQUIRE(B >= 0 && n >= 0 && n < 0xFFFFFFFFll, "randint: bad sizes");
  if (B == 0 || n == 0) return FBS_OK;
  FBS_REQUIRE(keys && out, "randint: null pointer");
  uint32_t span = maxval > minval ? (uint32_t)((int64_t)maxval - (int64_t)minval) : 1u;
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  randint_kernel<<<grid_for(B * n, 256), 256, 0, as_stream(s)>>>(keys, B, (uint32_t)n, minval, span, mult, out);
  return check_launch("randint_kernel");
}

int fbs_random_choice_f32(fbs_stream_t s, const uint32_t* keys, const float* p, int64_t B, int64_t N, int64_t n,
                          int32_t* out) {
  FBS_REQUIRE(B >= 0 && N >= 1 && n >= 0, "choice: bad sizes");
  if (B == 0 || n == 0) return FBS_OK;
  FBS_REQUIRE(keys && p && out, "choice: null pointer");
  const size_t per_warp = (size_t)N * sizeof(float);
  if (per_warp > 200 * 1024) {
    set_error("choice: N=%lld exceeds the single-warp shared-memory scan limit", (long long)N);
    return FBS_ERR_UNSUPPORTED;
  }
  int warps = (int)(48 * 1024 / per_warp);
  warps = warps < 1 ? 1 : (warps > 4 ? 4 : warps);
  const size_t smem = per_warp * warps;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(choice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t blocks = (B + warps - 1) / warps;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  choice_kernel<<<(int)blocks, warps * 32, smem, as_stream(s)>>>(keys, p, B, (int)N, (uint32_t)n, out);
  return check_launch("choice_kernel");
}

}  // extern "C"
