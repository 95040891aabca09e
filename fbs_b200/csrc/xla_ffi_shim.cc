// XLA-FFI (jax.ffi) handlers over the C ABI of include/fbs_b200.h.
//
// NOT COMPILED IN THIS IMAGE: JAX / jaxlib are not installed and cannot be (no network), so
// xla/ffi/api/ffi.h does not exist here and this file is excluded from fbs_b200/build.py unless
// XLA_FFI_INCLUDE points at jax.ffi.include_dir().  It is therefore UNTESTED; the tested boundary is the
// plain C ABI underneath.  INTEGRATION.md shows the Python side (jax.ffi.register_ffi_target / ffi_call).
//
// Build (on a machine with jaxlib):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
//        -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//        fbs_b200/csrc/*.cu fbs_b200/csrc/xla_ffi_shim.cc -o libfbs_b200_xla.so
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define FBS_HAVE_XLA_FFI 1
#endif
#endif

#ifdef FBS_HAVE_XLA_FFI
#include <cuda_runtime.h>
#include "xla/ffi/api/ffi.h"
#include "../../include/fbs_b200.h"

namespace ffi = xla::ffi;

static ffi::Error as_error(int rc) {
  if (rc == FBS_OK) return ffi::Error::Success();
  return ffi::Error(rc == FBS_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument
                                                   : rc == FBS_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                                               : ffi::ErrorCode::kInternal,
                    fbs_last_error());
}

static fbs_affine_model_t model_of(ffi::Buffer<ffi::F32> MT, ffi::Buffer<ffi::F32> m, ffi::Buffer<ffi::F32> dt,
                                   ffi::Buffer<ffi::F32> sd, ffi::Buffer<ffi::F32> lognorm, ffi::Buffer<ffi::F32> MTp,
                                   int32_t du) {
  fbs_affine_model_t mod{};
  auto dims = MT.dimensions();  // [K, D, D]
  mod.K = (int32_t)dims[0];
  mod.du = du;
  mod.dv = (int32_t)dims[1] - du;
  mod.MT = MT.typed_data();
  mod.m = m.typed_data();
  mod.dt = dt.typed_data();
  mod.sd = sd.typed_data();
  mod.lognorm = lognorm.typed_data();
  mod.MTp = MTp.typed_data();
  return mod;
}

// forward_pass(key, us_star, bs_star, vs, ...) -> (As, log_wss, uss)      csmc.py:80-164
static ffi::Error CsmcForwardImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::Buffer<ffi::U32> keys,
                                  ffi::Buffer<ffi::F32> us_star, ffi::Buffer<ffi::S32> bs_star, ffi::Buffer<ffi::F32> vs,
                                  ffi::Buffer<ffi::F32> MT, ffi::Buffer<ffi::F32> m, ffi::Buffer<ffi::F32> dt,
                                  ffi::Buffer<ffi::F32> sd, ffi::Buffer<ffi::F32> lognorm, ffi::Buffer<ffi::F32> MTp,
                                  int32_t du, int32_t init_mode, float init_log_w, int32_t scheme,
                                  ffi::ResultBuffer<ffi::S32> As, ffi::ResultBuffer<ffi::F32> log_wss,
                                  ffi::ResultBuffer<ffi::F32> uss) {
  fbs_affine_model_t mod = model_of(MT, m, dt, sd, lognorm, MTp, du);
  const int64_t B = keys.dimensions()[0];
  const int64_t N = As->dimensions()[2];
  const size_t ws_bytes = fbs_sweep_workspace_bytes(&mod, B);
  void* ws = scratch.Allocate(ws_bytes).value_or(nullptr);  // nullptr -> the general kernel runs
  return as_error(fbs_csmc_forward_affine_f32(stream, &mod, keys.typed_data(), us_star.typed_data(),
                                              bs_star.typed_data(), vs.typed_data(), init_mode, init_log_w, scheme, B, N,
                                              As->typed_data(), log_wss->typed_data(), uss->typed_data(), nullptr,
                                              nullptr, ws, ws ? ws_bytes : 0));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_csmc_forward, CsmcForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("du")
                                  .Attr<int32_t>("init_mode")
                                  .Attr<float>("init_log_w")
                                  .Attr<int32_t>("scheme")
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// cond_resampling(key, weights, i, j, True) -> idx                          resamplings.py:10-88
static ffi::Error CondResampleImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> keys, ffi::Buffer<ffi::F32> weights,
                                   ffi::Buffer<ffi::S32> i, ffi::Buffer<ffi::S32> j, int32_t scheme, int32_t conditional,
                                   ffi::ResultBuffer<ffi::S32> idx) {
  const int64_t B = keys.dimensions()[0], N = weights.dimensions()[1];
  return as_error(fbs_cond_resample_f32(stream, scheme, keys.typed_data(), weights.typed_data(), i.typed_data(),
                                        j.typed_data(), conditional, B, N, idx->typed_data()));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_cond_resample, CondResampleImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Attr<int32_t>("scheme")
                                  .Attr<int32_t>("conditional")
                                  .Ret<ffi::Buffer<ffi::S32>>());

// simulate_cond_forward(key, x0, ts) -> path                                linear.py:190-221
static ffi::Error OuForwardImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> keys, ffi::Buffer<ffi::F32> x0,
                                ffi::Buffer<ffi::F32> F, ffi::Buffer<ffi::F32> sqrtQ, ffi::ResultBuffer<ffi::F32> path) {
  const int64_t B = keys.dimensions()[0], K = F.dimensions()[0], D = x0.dimensions().back();
  return as_error(fbs_ou_forward_path_f32(stream, keys.typed_data(), x0.typed_data(), x0.dimensions().size() == 2,
                                          F.typed_data(), sqrtQ.typed_data(), B, K, D, D, 0, path->typed_data(), nullptr));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_ou_forward_path, OuForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
#endif  // FBS_HAVE_XLA_FFI
