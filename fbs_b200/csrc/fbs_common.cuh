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

// Number of SMs of the current device (cached per thread; B200: 148).
int sm_count();

}  // namespace fbs
