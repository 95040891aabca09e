// Shared host-side helpers for the fbs_b200 C-ABI library (error string, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fbs_b200.h"

namespace fbs {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return FBS_ERR_CUDA;
  }
  count_launch();
  return FBS_OK;
}

#define FBS_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      ::fbs::set_error(__VA_ARGS__);        \
      return FBS_ERR_INVALID_ARGUMENT;      \
    }                                       \
  } while (0)

inline cudaStream_t as_stream(fbs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Tiled resampling (resample_kernels.cu): chains in shared-memory tiles, one thread per chain for the sequential sums.
// expw: weights are log-weights (exp applied on load); split_first: key_resampling = split(key)[0] (csmc.py:136).
// Returns FBS_ERR_UNSUPPORTED, without touching the error string, when a chain's rows do not fit in shared memory.
int launch_resample_tile(cudaStream_t st, int scheme, const uint32_t* keys, const float* weights, const int32_t* iv,
                         const int32_t* jv, int conditional, int clip, int expw, int split_first, int64_t B, int64_t N,
                         int32_t* out);

// Tensor-core per-timestep transition + weight kernel (step_tc.cu).  Returns FBS_OK, an error, or -1 (not eligible).
int launch_step_transition_tc(cudaStream_t st, const fbs_affine_model_t* model, int k, const uint32_t* step_keys,
                              const float* us_prev, const int32_t* A, const float* v, const float* v_prev,
                              const float* u_star, const int32_t* b_cur, int64_t B, int64_t N, float* us_out,
                              float* lw_out);

// Number of SMs of the current device (cached per thread; B200: 148).
int sm_count();

// Implementation-selection knobs for tests and A/B measurements (fbs_debug_set_option): relaxed atomics, default 0 = the
// library's own choice.  The library never reads the environment.
enum DebugOpt {
  OPT_SWEEP_IMPL = 0,   // 1 / 2 / 3 / 4: pin the general / tiled / tcgen05 / warp-per-chain sweep kernel
  OPT_SWEEP_VERBOSE,    // 1: print the sweep kernel chosen to stderr
  OPT_STEP_IMPL,        // 1: CUDA-core per-timestep transition kernel
  OPT_STEP_TC_WARPS,    // 8: eight-warp variant of the tcgen05 per-timestep kernel
  OPT_STEPVEC_IMPL,     // 1: thread-per-output step-vector kernel
  OPT_SWEEP_G,          // > 0: chains per CTA of the tiled sweep kernel
  OPT_V3_TWOPASS,       // 1: two-pass GEMM in the tcgen05 sweep kernel
  OPT_EM_IMPL,          // 1: CTA-per-chain, 2: thread-per-chain Euler--Maruyama path kernel
  OPT_V3_VARIANT,       // experimental variants of the tcgen05 sweep kernel (A/B measurements)
  OPT_CONV_IMPL,        // 1: per-tap activation boxes for every convolution (no haloed 3x3 path)
  OPT_COUNT
};
int debug_opt(DebugOpt which);

}  // namespace fbs
