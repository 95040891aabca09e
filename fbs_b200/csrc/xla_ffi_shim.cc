// XLA-FFI (jax.ffi) handlers over the C ABI of include/fbs_b200.h: one `XLA_FFI_Error* (*)(XLA_FFI_CallFrame*)` symbol
// per entry point of the hot path, so that the kernels compose under jax.jit / jax.vmap as custom calls.
//
// STATUS.  JAX / jaxlib are not installed in this image and cannot be (no network), so the real xla/ffi/api/ffi.h does
// not exist here: this file is NOT part of libfbs_b200.so and has never RUN.  It is, however, compiled every round by
// tests/test_xla_ffi_shim.py against tests/xla_ffi_stub/ (a minimal stand-in for the binding types that statically
// checks every implementation against the argument list its Bind() chain decodes).  The tested boundary is the plain
// C ABI underneath.  INTEGRATION.md shows the Python side (jax.ffi.register_ffi_target / ffi_call).
//
// Build (on a machine with jaxlib):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
//        -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude \
//        fbs_b200/csrc/*.cu fbs_b200/csrc/xla_ffi_shim.cc -o libfbs_b200_xla.so
//
// Conventions.  Every array carries the leading chain axis B (vmap_method="broadcast_all" maps jax.vmap onto it).  An
// OPTIONAL operand / result is passed as a zero-element buffer and reaches the C ABI as NULL.  The affine model travels as
// seven operands (MT, m, dt, sd, lognorm, MTp, MTc -- the arrays of fbs_affine_model_t; MTp / MTc may be empty, which
// selects the general kernel) plus the attribute `du`.  Workspace comes from ffi::ScratchAllocator; nothing is
// allocated, synchronised or retained, and only the stream XLA hands over is used.
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

using BufF = ffi::Buffer<ffi::F32>;
using BufI = ffi::Buffer<ffi::S32>;
using BufU = ffi::Buffer<ffi::U32>;
using BufH = ffi::Buffer<ffi::BF16>;
using ResF = ffi::ResultBuffer<ffi::F32>;
using ResI = ffi::ResultBuffer<ffi::S32>;
using ResU = ffi::ResultBuffer<ffi::U32>;
using ResB = ffi::ResultBuffer<ffi::U8>;
using ResH = ffi::ResultBuffer<ffi::BF16>;

static ffi::Error as_error(int rc) {
  if (rc == FBS_OK) return ffi::Error::Success();
  return ffi::Error(rc == FBS_ERR_INVALID_ARGUMENT ? ffi::ErrorCode::kInvalidArgument
                                                   : rc == FBS_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                                               : ffi::ErrorCode::kInternal,
                    fbs_last_error());
}

// zero-element buffer -> NULL (optional operands / results)
template <typename B>
static auto opt(B& b) -> decltype(b.typed_data()) {
  return b.element_count() == 0 ? nullptr : b.typed_data();
}
template <typename B>
static void* opt_raw(B& b) {
  return b.element_count() == 0 ? nullptr : b.untyped_data();
}

static fbs_affine_model_t model_of(BufF& MT, BufF& m, BufF& dt, BufF& sd, BufF& lognorm, BufF& MTp, BufF& MTc, int32_t du) {
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
  mod.MTp = opt(MTp);
  mod.MTc = opt(MTc);  // the tcgen05 sweep kernel needs it (sweep_v3.cu); empty -> the tiled / general kernel runs
  return mod;
}

// the seven model operands + `du`, in the order model_of() takes them
#define FBS_MODEL_PARAMS BufF MT, BufF m, BufF dt, BufF sd, BufF lognorm, BufF MTp, BufF MTc
#define FBS_MODEL_ARGS MT, m, dt, sd, lognorm, MTp, MTc
#define FBS_BIND_MODEL() Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>()
#define FBS_BIND_STREAM() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

static void* sweep_workspace(ffi::ScratchAllocator& scratch, const fbs_affine_model_t& mod, int64_t B, size_t* bytes) {
  *bytes = fbs_sweep_workspace_bytes(&mod, B);
  void* ws = *bytes ? scratch.Allocate(*bytes).value_or(nullptr) : nullptr;  // nullptr -> the general kernel runs
  if (!ws) *bytes = 0;
  return ws;
}

// ---------------------------------------------------------------------------------------------------------------
// Whole sweeps
// ---------------------------------------------------------------------------------------------------------------
// forward_pass(key, us_star, bs_star, vs, ...) -> (As, log_wss, uss, log_ws_last, us_last)      csmc.py:80-164
// (history results may be empty: gibbs_kernel's explicit-backward branch only consumes the last step, gibbs.py:148-154)
static ffi::Error CsmcForwardImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, BufU keys, BufF us_star, BufI bs_star,
                                  BufF vs, FBS_MODEL_PARAMS, int32_t du, int32_t init_mode, float init_log_w, int32_t scheme,
                                  int32_t nparticles, ResI As, ResF log_wss, ResF uss, ResF log_ws_last, ResF us_last) {
  fbs_affine_model_t mod = model_of(FBS_MODEL_ARGS, du);
  const int64_t B = keys.dimensions()[0];
  size_t ws_bytes = 0;
  void* ws = sweep_workspace(scratch, mod, B, &ws_bytes);
  return as_error(fbs_csmc_forward_affine_f32(stream, &mod, keys.typed_data(), us_star.typed_data(), bs_star.typed_data(),
                                              vs.typed_data(), init_mode, init_log_w, scheme, B, nparticles, opt(*As),
                                              opt(*log_wss), opt(*uss), opt(*log_ws_last), opt(*us_last), ws, ws_bytes));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_csmc_forward, CsmcForwardImpl,
                              FBS_BIND_STREAM().Ctx<ffi::ScratchAllocator>().Arg<BufU>().Arg<BufF>().Arg<BufI>().Arg<BufF>()
                                  .FBS_BIND_MODEL().Attr<int32_t>("du").Attr<int32_t>("init_mode").Attr<float>("init_log_w")
                                  .Attr<int32_t>("scheme").Attr<int32_t>("nparticles")
                                  .Ret<BufI>().Ret<BufF>().Ret<BufF>().Ret<BufF>().Ret<BufF>());

// pmcmc_filter_step(key, vs_bridge, u0s, ...) -> (uT, log_ell)                                  smc.py:115-158
static ffi::Error PmcmcFilterImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, BufU keys, BufF vs, BufF u0s,
                                  FBS_MODEL_PARAMS, int32_t du, int32_t scheme, ResF uT, ResF log_ell) {
  fbs_affine_model_t mod = model_of(FBS_MODEL_ARGS, du);
  const int64_t B = keys.dimensions()[0], N = u0s.dimensions()[1];
  size_t ws_bytes = 0;
  void* ws = sweep_workspace(scratch, mod, B, &ws_bytes);
  return as_error(fbs_pmcmc_filter_affine_f32(stream, &mod, keys.typed_data(), vs.typed_data(), u0s.typed_data(), scheme, B, N,
                                              uT->typed_data(), log_ell->typed_data(), nullptr, nullptr, nullptr, ws, ws_bytes));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_pmcmc_filter, PmcmcFilterImpl,
                              FBS_BIND_STREAM().Ctx<ffi::ScratchAllocator>().Arg<BufU>().Arg<BufF>().Arg<BufF>()
                                  .FBS_BIND_MODEL().Attr<int32_t>("du").Attr<int32_t>("scheme").Ret<BufF>().Ret<BufF>());

// bootstrap_filter(...) -> (uT, log_nell, us_hist)   (us_hist may be empty: return_last=True)      smc.py:9-88
static ffi::Error BootstrapFilterImpl(cudaStream_t stream, BufU step_keys, BufF vs, BufF u0s, FBS_MODEL_PARAMS, int32_t du,
                                      int32_t scheme, ResF uT, ResF log_nell, ResF us_hist) {
  fbs_affine_model_t mod = model_of(FBS_MODEL_ARGS, du);
  const int64_t B = step_keys.dimensions()[0], N = u0s.dimensions()[1];
  return as_error(fbs_bootstrap_filter_affine_f32(stream, &mod, step_keys.typed_data(), vs.typed_data(), u0s.typed_data(),
                                                  scheme, B, N, uT->typed_data(), log_nell->typed_data(), nullptr, nullptr,
                                                  opt(*us_hist)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_bootstrap_filter, BootstrapFilterImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().FBS_BIND_MODEL().Attr<int32_t>("du")
                                  .Attr<int32_t>("scheme").Ret<BufF>().Ret<BufF>().Ret<BufF>());

// backward_scanning_pass(key, As, xss, log_w_T) -> (xs_star, bs_star)                            csmc.py:230-270
static ffi::Error BackwardScanImpl(cudaStream_t stream, BufU keys, BufI As, BufF uss, BufF log_w_T, ResF xs_star, ResI bs_star) {
  auto d = uss.dimensions();  // [B, K+1, N, du]
  return as_error(fbs_backward_scan_f32(stream, keys.typed_data(), As.typed_data(), uss.typed_data(), log_w_T.typed_data(), d[0],
                                        d[1] - 1, d[2], d[3], xs_star->typed_data(), bs_star->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_backward_scan, BackwardScanImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufI>().Arg<BufF>().Arg<BufF>().Ret<BufF>().Ret<BufI>());

// backward_sampling_pass (mode 0, csmc.py:167-227) / bootstrap_backward_smoother (mode 1, smc.py:91-112)
static ffi::Error BackwardSampleImpl(cudaStream_t stream, BufU keys, BufF vs, BufF uss, BufF log_wss, FBS_MODEL_PARAMS,
                                     int32_t du, int32_t mode, int32_t shared_history, ResF xs_star, ResI bs_star) {
  fbs_affine_model_t mod = model_of(FBS_MODEL_ARGS, du);
  auto d = uss.dimensions();  // [B, K+1, N, du] or [K+1, N, du] (shared)
  const int64_t B = keys.dimensions()[0], N = d[d.size() - 2];
  return as_error(fbs_backward_sample_affine_f32(stream, &mod, mode, keys.typed_data(), vs.typed_data(), uss.typed_data(),
                                                 opt(log_wss), shared_history, B, N, xs_star->typed_data(), opt(*bs_star)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_backward_sample, BackwardSampleImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().FBS_BIND_MODEL()
                                  .Attr<int32_t>("du").Attr<int32_t>("mode").Attr<int32_t>("shared_history").Ret<BufF>().Ret<BufI>());

// ---------------------------------------------------------------------------------------------------------------
// Per-timestep kernels and the pieces of gibbs_kernel / pmcmc_kernel
// ---------------------------------------------------------------------------------------------------------------
// scan body of forward_pass (csmc.py:132-148) for particle sets in global memory -> (A, us, log_ws)
static ffi::Error CsmcStepImpl(cudaStream_t stream, BufU step_keys, BufF us_prev, BufF log_ws, BufF v, BufF v_prev, BufF u_star,
                               BufI b_star_prev, BufI b_star, FBS_MODEL_PARAMS, int32_t du, int32_t k, int32_t scheme, ResI A,
                               ResF us, ResF log_ws_out) {
  fbs_affine_model_t mod = model_of(FBS_MODEL_ARGS, du);
  const int64_t B = step_keys.dimensions()[0], N = us_prev.dimensions()[1];
  return as_error(fbs_csmc_step_affine_f32(stream, &mod, k, scheme, step_keys.typed_data(), us_prev.typed_data(),
                                           log_ws.typed_data(), v.typed_data(), v_prev.typed_data(), u_star.typed_data(),
                                           b_star_prev.typed_data(), b_star.typed_data(), B, N, A->typed_data(), us->typed_data(),
                                           log_ws_out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_csmc_step, CsmcStepImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufI>()
                                  .Arg<BufI>().FBS_BIND_MODEL().Attr<int32_t>("du").Attr<int32_t>("k").Attr<int32_t>("scheme")
                                  .Ret<BufI>().Ret<BufF>().Ret<BufF>());

// the three closures on their own (gp_gibbs.py:120-135); absent operands / results are empty buffers
static ffi::Error AffineEvalImpl(cudaStream_t stream, BufU tr_keys, BufF us_prev, BufF v, BufF v_prev, BufF u_eval,
                                 FBS_MODEL_PARAMS, int32_t du, int32_t k, ResF us_out, ResF lw_out, ResF tlp_out) {
  fbs_affine_model_t mod = model_of(FBS_MODEL_ARGS, du);
  const int64_t B = us_prev.dimensions()[0], N = us_prev.dimensions()[1];
  return as_error(fbs_affine_eval_f32(stream, &mod, k, opt(tr_keys), us_prev.typed_data(), opt(v), v_prev.typed_data(),
                                      opt(u_eval), B, N, opt(*us_out), opt(*lw_out), opt(*tlp_out)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_affine_eval, AffineEvalImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().FBS_BIND_MODEL()
                                  .Attr<int32_t>("du").Attr<int32_t>("k").Ret<BufF>().Ret<BufF>().Ret<BufF>());

// twisted_smc(...) -> (samples, log_weights)                                                   smc.py:261-309
static ffi::Error TwistedSmcImpl(cudaStream_t stream, BufU keys, BufF y, BufF x0, BufF MT, BufF Mrow, BufF m, BufF sd, BufF g2,
                                 float dt, float obs_var, int32_t scheme, ResF samples, ResF log_ws) {
  auto dm = x0.dimensions();  // [B, N, d]
  const int64_t K = sd.dimensions()[0] - 1;
  return as_error(fbs_twisted_smc_affine_f32(stream, MT.typed_data(), Mrow.typed_data(), m.typed_data(), sd.typed_data(),
                                             g2.typed_data(), dt, obs_var, K, dm[2], keys.typed_data(), y.typed_data(),
                                             y.dimensions().size() == 2 && y.dimensions()[0] == dm[0] && dm[0] > 1, x0.typed_data(),
                                             scheme, dm[0], dm[1], samples->typed_data(), log_ws->typed_data(), nullptr, nullptr,
                                             nullptr));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_twisted_smc, TwistedSmcImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>()
                                  .Arg<BufF>().Attr<float>("dt").Attr<float>("obs_var").Attr<int32_t>("scheme").Ret<BufF>().Ret<BufF>());

// force_move(key, weights, k) fused with x0 = uss[-1, idx]  -> (idx, alpha, x0)                 gibbs.py:152-154,171-214
static ffi::Error ForceMoveImpl(cudaStream_t stream, BufU keys, BufF log_ws_last, BufF us_last, BufI k, int32_t weights_are_log,
                                ResI idx, ResF alpha, ResF x0) {
  const int64_t B = keys.dimensions()[0], N = log_ws_last.dimensions()[1];
  const int64_t du = us_last.element_count() ? us_last.dimensions()[2] : 0;
  return as_error(fbs_force_move_f32(stream, keys.typed_data(), log_ws_last.typed_data(), weights_are_log, opt(us_last),
                                     k.typed_data(), B, N, du, idx->typed_data(), opt(*alpha), opt(*x0)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_force_move, ForceMoveImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufI>().Attr<int32_t>("weights_are_log")
                                  .Ret<BufI>().Ret<BufF>().Ret<BufF>());

// pcn_proposal combination step                                                                 smc.py:161-168
static ffi::Error PcnCombineImpl(cudaStream_t stream, BufF x, BufF mean, BufF r0, BufF r1, float delta, ResF out) {
  const int64_t B = x.dimensions()[0], n = (int64_t)mean.element_count();
  return as_error(fbs_pcn_combine_f32(stream, (double)delta, x.typed_data(), mean.typed_data(), r0.typed_data(), r1.typed_data(),
                                      B, n, out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_pcn_combine, PcnCombineImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Attr<float>("delta").Ret<BufF>());

// Metropolis--Hastings accept / select (smc.py:244-258).  The C entry point updates the chain state in place; XLA hands
// out separate result buffers (unless the caller aliases them with input_output_aliases), so the state is copied on the
// stream first.
static ffi::Error MhAcceptImpl(cudaStream_t stream, BufU keys_mh, BufF prop_uTs, BufF prop_log_ell, BufF prop_ys, BufF uT,
                               BufF log_ell, BufF ys, int32_t which_u, ResF uT_out, ResF log_ell_out, ResF ys_out,
                               ResF acceptance_prob, ResB is_accepted) {
  const int64_t B = keys_mh.dimensions()[0], N = prop_uTs.dimensions()[1], du = prop_uTs.dimensions()[2];
  const int64_t ny = (int64_t)ys.element_count() / (B > 0 ? B : 1);
  auto copy = [&](float* dst, const float* src, size_t n) {
    return dst == src || n == 0 ? cudaSuccess : cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, stream);
  };
  if (copy(uT_out->typed_data(), uT.typed_data(), uT.element_count()) != cudaSuccess ||
      copy(log_ell_out->typed_data(), log_ell.typed_data(), log_ell.element_count()) != cudaSuccess ||
      copy(ys_out->typed_data(), ys.typed_data(), ys.element_count()) != cudaSuccess)
    return ffi::Error(ffi::ErrorCode::kInternal, "fbs_xla_mh_accept: state copy failed");
  return as_error(fbs_mh_accept_f32(stream, keys_mh.typed_data(), prop_uTs.typed_data(), prop_log_ell.typed_data(),
                                    prop_ys.typed_data(), B, N, du, ny, which_u, uT_out->typed_data(), log_ell_out->typed_data(),
                                    ys_out->typed_data(), acceptance_prob->typed_data(), is_accepted->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_mh_accept, MhAcceptImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>()
                                  .Attr<int32_t>("which_u").Ret<BufF>().Ret<BufF>().Ret<BufF>().Ret<BufF>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

// ref_sampler(key, yT, n)                                                                       gp_gibbs.py:138-141
static ffi::Error GaussianRefSampleImpl(cudaStream_t stream, BufU keys, BufF yT, BufF a, BufF Bm, BufF c, BufF L, ResF out) {
  auto d = out->dimensions();  // [B, N, du]
  return as_error(fbs_gaussian_ref_sample_f32(stream, keys.typed_data(), yT.typed_data(), a.typed_data(), Bm.typed_data(),
                                              c.typed_data(), L.typed_data(), d[0], d[1], d[2], (int64_t)c.element_count(),
                                              out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_gaussian_ref_sample, GaussianRefSampleImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Ret<BufF>());

// ---------------------------------------------------------------------------------------------------------------
// Resampling and forward noising
// ---------------------------------------------------------------------------------------------------------------
// cond_resampling(key, weights, i, j, conditional) -> idx                                       resamplings.py:10-125
static ffi::Error CondResampleImpl(cudaStream_t stream, BufU keys, BufF weights, BufI i, BufI j, int32_t scheme,
                                   int32_t conditional, ResI idx) {
  const int64_t B = keys.dimensions()[0], N = weights.dimensions()[1];
  return as_error(fbs_cond_resample_f32(stream, scheme, keys.typed_data(), weights.typed_data(), opt(i), opt(j), conditional, B,
                                        N, idx->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_cond_resample, CondResampleImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufI>().Arg<BufI>().Attr<int32_t>("scheme")
                                  .Attr<int32_t>("conditional").Ret<BufI>());

// resampling(weights, key) -> idx                                                               resampling.py:43-101
static ffi::Error ResampleImpl(cudaStream_t stream, BufU keys, BufF weights, int32_t scheme, ResI idx) {
  const int64_t B = keys.dimensions()[0], N = weights.dimensions()[1];
  return as_error(fbs_resample_f32(stream, scheme, keys.typed_data(), weights.typed_data(), B, N, idx->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_resample, ResampleImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Attr<int32_t>("scheme").Ret<BufI>());

// simulate_cond_forward(key, x0, ts) -> path; rev != 0: (us, vs) = (path_x[::-1], path_y[::-1])   linear.py:190-221
static ffi::Error OuForwardImpl(cudaStream_t stream, BufU keys, BufF x0, BufF F, BufF sqrtQ, int32_t du, int32_t rev,
                                ResF out_u, ResF out_v) {
  const int64_t B = keys.dimensions()[0], K = F.dimensions()[0], D = x0.dimensions().back();
  return as_error(fbs_ou_forward_path_f32(stream, keys.typed_data(), x0.typed_data(), x0.dimensions().size() == 2, F.typed_data(),
                                          sqrtQ.typed_data(), B, K, D, rev ? du : D, rev, out_u->typed_data(), opt(*out_v)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_ou_forward_path, OuForwardImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Attr<int32_t>("du")
                                  .Attr<int32_t>("rev").Ret<BufF>().Ret<BufF>());

// euler_maruyama(key, x0, ts, drift, dispersion, integration_nsteps=m, return_path=True), affine drift   simulators.py:53-106
static ffi::Error EmAffinePathImpl(cudaStream_t stream, BufU keys, BufF x0, BufF AT, BufF a, BufF ddt, BufF disp, int32_t du,
                                   int32_t rev, ResF out_u, ResF out_v) {
  const int64_t B = keys.dimensions()[0], K = ddt.dimensions()[0], D = x0.dimensions().back();
  const int64_t m = AT.dimensions()[0] / (K > 0 ? K : 1);
  return as_error(fbs_em_affine_path_f32(stream, keys.typed_data(), x0.typed_data(), x0.dimensions().size() == 2, AT.typed_data(),
                                         a.typed_data(), ddt.typed_data(), disp.typed_data(), B, K, m, D, rev ? du : D, rev,
                                         out_u->typed_data(), opt(*out_v)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_em_affine_path, EmAffinePathImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>()
                                  .Attr<int32_t>("du").Attr<int32_t>("rev").Ret<BufF>().Ret<BufF>());

// ---------------------------------------------------------------------------------------------------------------
// jax.random work-alikes (so a jitted driver can draw inside the same custom-call graph)
// ---------------------------------------------------------------------------------------------------------------
static ffi::Error RandomBitsImpl(cudaStream_t stream, BufU keys, ResU out) {
  const int64_t B = keys.dimensions()[0];
  return as_error(fbs_random_bits_u32(stream, keys.typed_data(), B, (int64_t)out->element_count() / (B > 0 ? B : 1),
                                      out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_random_bits, RandomBitsImpl, FBS_BIND_STREAM().Arg<BufU>().Ret<BufU>());

static ffi::Error RandomSplitImpl(cudaStream_t stream, BufU keys, ResU out) {  // out [B, num, 2]
  return as_error(fbs_random_split(stream, keys.typed_data(), keys.dimensions()[0], out->dimensions()[1], out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_random_split, RandomSplitImpl, FBS_BIND_STREAM().Arg<BufU>().Ret<BufU>());

static ffi::Error RandomUniformImpl(cudaStream_t stream, BufU keys, float minval, float maxval, ResF out) {
  const int64_t B = keys.dimensions()[0];
  return as_error(fbs_random_uniform_f32(stream, keys.typed_data(), B, (int64_t)out->element_count() / (B > 0 ? B : 1), minval,
                                         maxval, out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_random_uniform, RandomUniformImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Attr<float>("minval").Attr<float>("maxval").Ret<BufF>());

static ffi::Error RandomNormalImpl(cudaStream_t stream, BufU keys, ResF out) {
  const int64_t B = keys.dimensions()[0];
  return as_error(fbs_random_normal_f32(stream, keys.typed_data(), B, (int64_t)out->element_count() / (B > 0 ? B : 1),
                                        out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_random_normal, RandomNormalImpl, FBS_BIND_STREAM().Arg<BufU>().Ret<BufF>());

static ffi::Error RandomRandintImpl(cudaStream_t stream, BufU keys, int32_t minval, int32_t maxval, ResI out) {
  const int64_t B = keys.dimensions()[0];
  return as_error(fbs_random_randint_i32(stream, keys.typed_data(), B, (int64_t)out->element_count() / (B > 0 ? B : 1), minval,
                                         maxval, out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_random_randint, RandomRandintImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Attr<int32_t>("minval").Attr<int32_t>("maxval").Ret<BufI>());

static ffi::Error RandomChoiceImpl(cudaStream_t stream, BufU keys, BufF p, ResI out) {
  const int64_t B = keys.dimensions()[0], N = p.dimensions()[1];
  return as_error(fbs_random_choice_f32(stream, keys.typed_data(), p.typed_data(), B, N,
                                        (int64_t)out->element_count() / (B > 0 ? B : 1), out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_random_choice, RandomChoiceImpl, FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Ret<BufI>());

// ---------------------------------------------------------------------------------------------------------------
// NN-score closures (experiments/imgs/inpainting.py:94-147) and the ops of the score network (fbs/nn/unet.py)
// ---------------------------------------------------------------------------------------------------------------
// dataset.concat                                                                                images.py:352-363
static ffi::Error NnAssembleImageImpl(cudaStream_t stream, BufF us, BufF v, BufI unobs_idx, BufI obs_idx, ResF img) {
  auto d = us.dimensions();  // [B, p, c]
  return as_error(fbs_nn_assemble_image_f32(stream, us.typed_data(), v.typed_data(), unobs_idx.typed_data(), obs_idx.typed_data(),
                                            d[0], (int32_t)d[1], (int32_t)obs_idx.element_count(), (int32_t)d[2], img->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_assemble_image, NnAssembleImageImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufI>().Arg<BufI>().Ret<BufF>());

// transition_sampler + likelihood_logpdf from ONE score evaluation -> (us_new, mean, lw)         inpainting.py:122-147
static ffi::Error NnEmStepImpl(cudaStream_t stream, BufF img, BufF score, BufI unobs_idx, BufI obs_idx, BufF v_next, BufU key,
                               BufI pin_row, BufF pin_value, float a, float g2, float dt, float sd, int32_t row_offset,
                               int32_t rows_total, ResF us_new, ResF mean_out, ResF lw) {
  auto d = img.dimensions();  // [B, H, W, c]
  const int32_t p = (int32_t)unobs_idx.element_count(), q = (int32_t)obs_idx.element_count();
  return as_error(fbs_nn_em_step_f32(stream, img.typed_data(), score.typed_data(), unobs_idx.typed_data(), obs_idx.typed_data(),
                                     opt(v_next), opt(key), d[0], p, q, (int32_t)d[3], a, g2, dt, sd, row_offset,
                                     rows_total > 0 ? rows_total : d[0], opt(pin_row), opt(pin_value), opt(*us_new),
                                     opt(*mean_out), opt(*lw)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_em_step, NnEmStepImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufI>().Arg<BufI>().Arg<BufF>().Arg<BufU>()
                                  .Arg<BufI>().Arg<BufF>().Attr<float>("a").Attr<float>("g2").Attr<float>("dt").Attr<float>("sd")
                                  .Attr<int32_t>("row_offset").Attr<int32_t>("rows_total").Ret<BufF>().Ret<BufF>().Ret<BufF>());

// one Euler--Maruyama sub-step with a network drift (forward sampler of the SB image runs)       sb_imgs/supr.py:132-137
static ffi::Error EmDriftStepImpl(cudaStream_t stream, BufU keys, BufF x, BufF drift, float ddt, float gs, ResF out) {
  const int64_t B = keys.dimensions()[0];
  return as_error(fbs_em_drift_step_f32(stream, keys.typed_data(), x.typed_data(), drift.typed_data(), B,
                                        (int64_t)x.element_count() / (B > 0 ? B : 1), ddt, gs, out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_em_drift_step, EmDriftStepImpl,
                              FBS_BIND_STREAM().Arg<BufU>().Arg<BufF>().Arg<BufF>().Attr<float>("ddt").Attr<float>("gs").Ret<BufF>());

// normalise + exp of the step's log-weights                                                     csmc.py:146,139
static ffi::Error NormaliseLogwImpl(cudaStream_t stream, BufF lw, ResF log_w, ResF w) {
  const int64_t N = lw.dimensions().back();
  return as_error(fbs_normalise_logw_f32(stream, lw.typed_data(), (int64_t)lw.element_count() / (N > 0 ? N : 1), N, opt(*log_w),
                                         opt(*w)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_normalise_logw, NormaliseLogwImpl, FBS_BIND_STREAM().Arg<BufF>().Ret<BufF>().Ret<BufF>());

// the ancestor gather                                                                           csmc.py:140
static ffi::Error GatherRowsImpl(cudaStream_t stream, BufF src, BufI idx, ResF dst) {
  const int64_t rows = src.dimensions()[0], row = (int64_t)src.element_count() / (rows > 0 ? rows : 1);
  return as_error(fbs_gather_rows_f32(stream, src.typed_data(), idx.typed_data(), (int64_t)idx.element_count(), row, rows,
                                      dst->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_gather_rows, GatherRowsImpl, FBS_BIND_STREAM().Arg<BufF>().Arg<BufI>().Ret<BufF>());

// flax.linen.Conv call sites of unet.py as implicit GEMMs on tcgen05 -> (out_f32, out_bf16); either result may be empty
static ffi::Error NnConvImpl(cudaStream_t stream, BufH in0, BufH in1, BufH weight, BufF bias, BufF residual, int32_t kh, int32_t kw,
                             int32_t off_h, int32_t off_w, int32_t pixel_shuffle, int32_t H, int32_t W, int32_t Cout, ResF out_f32,
                             ResH out_bf16, ResF gn_partials) {
  auto d = in0.dimensions();  // [B, Hin, Win, C0]
  fbs_nn_conv_t a{};
  a.B = (int32_t)d[0]; a.H = H; a.W = W; a.Hin = (int32_t)d[1]; a.Win = (int32_t)d[2];
  a.C0 = (int32_t)d[3]; a.C1 = in1.element_count() ? (int32_t)in1.dimensions()[3] : 0; a.Cout = Cout;
  a.kh = kh; a.kw = kw; a.off_h = off_h; a.off_w = off_w; a.pixel_shuffle = pixel_shuffle;
  a.in0 = in0.untyped_data(); a.in1 = opt_raw(in1); a.weight = weight.untyped_data();
  a.bias = opt(bias); a.residual = opt(residual); a.out_f32 = opt(*out_f32); a.out_bf16 = opt_raw(*out_bf16);
  a.gn_partials = opt(*gn_partials);  // [B, slots, Cout / 4, 2] (slots: fbs_nn_conv_gn_layout, evaluated at trace time) or empty
  return as_error(fbs_nn_conv_bf16(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_conv, NnConvImpl,
                              FBS_BIND_STREAM().Arg<BufH>().Arg<BufH>().Arg<BufH>().Arg<BufF>().Arg<BufF>().Attr<int32_t>("kh")
                                  .Attr<int32_t>("kw").Attr<int32_t>("off_h").Attr<int32_t>("off_w").Attr<int32_t>("pixel_shuffle")
                                  .Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("Cout").Ret<BufF>().Ret<BufH>().Ret<BufF>());

// GroupNorm + swish with the statistics of the producing convolution (x: fp32 or bf16, the other one empty)
static ffi::Error NnGroupNormStatsImpl(cudaStream_t stream, BufF x_f32, BufH x_bf16, BufF partials, BufF gamma, BufF beta,
                                       BufF time_scale_shift, BufF residual, BufF ln_gamma, int32_t groups, float eps, float ln_eps,
                                       ResF out_f32, ResH out_bf16, ResH ln_out_bf16) {
  auto d = x_f32.element_count() ? x_f32.dimensions() : x_bf16.dimensions();  // [B, P, C]
  return as_error(fbs_nn_groupnorm_swish_stats(stream, opt(x_f32), opt_raw(x_bf16), partials.typed_data(),
                                               (int32_t)partials.dimensions()[1], d[0], (int32_t)d[1], (int32_t)d[2], groups,
                                               gamma.typed_data(), beta.typed_data(), opt(time_scale_shift), opt(residual), eps,
                                               opt(*out_f32), opt_raw(*out_bf16), opt(ln_gamma), ln_eps, opt_raw(*ln_out_bf16)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_groupnorm_swish_stats, NnGroupNormStatsImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufH>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>()
                                  .Arg<BufF>().Attr<int32_t>("groups").Attr<float>("eps").Attr<float>("ln_eps").Ret<BufF>()
                                  .Ret<BufH>().Ret<BufH>());

static ffi::Error NnGroupNormImpl(cudaStream_t stream, BufF x, BufF gamma, BufF beta, BufF time_scale_shift, BufF residual,
                                  int32_t groups, float eps, ResF out_f32, ResH out_bf16) {
  auto d = x.dimensions();  // [B, P, C]
  return as_error(fbs_nn_groupnorm_swish_f32(stream, x.typed_data(), d[0], (int32_t)d[1], (int32_t)d[2], groups, gamma.typed_data(),
                                             beta.typed_data(), opt(time_scale_shift), opt(residual), eps, opt(*out_f32),
                                             opt_raw(*out_bf16)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_groupnorm_swish, NnGroupNormImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Attr<int32_t>("groups")
                                  .Attr<float>("eps").Ret<BufF>().Ret<BufH>());

static ffi::Error NnLayerNormImpl(cudaStream_t stream, BufF x, BufF gamma, BufF residual, float eps, ResF out_f32, ResH out_bf16) {
  const int64_t C = x.dimensions().back(), rows = (int64_t)x.element_count() / (C > 0 ? C : 1);
  return as_error(fbs_nn_layernorm_f32(stream, x.typed_data(), rows, (int32_t)C, gamma.typed_data(), opt(residual), eps,
                                       opt(*out_f32), opt_raw(*out_bf16)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_layernorm, NnLayerNormImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufF>().Attr<float>("eps").Ret<BufF>().Ret<BufH>());

static ffi::Error NnLinearAttentionImpl(cudaStream_t stream, BufH qkv, int32_t heads, int32_t dim_head, ResH out) {
  auto d = qkv.dimensions();  // [B, P, 3 heads dim_head]
  return as_error(fbs_nn_linear_attention_bf16(stream, qkv.untyped_data(), d[0], (int32_t)d[1], heads, dim_head, out->untyped_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_linear_attention, NnLinearAttentionImpl,
                              FBS_BIND_STREAM().Arg<BufH>().Attr<int32_t>("heads").Attr<int32_t>("dim_head").Ret<BufH>());

static ffi::Error NnAttentionImpl(cudaStream_t stream, BufH qkv, int32_t heads, int32_t dim_head, float scale, ResH out) {
  auto d = qkv.dimensions();
  return as_error(fbs_nn_attention_bf16(stream, qkv.untyped_data(), d[0], (int32_t)d[1], heads, dim_head, scale, out->untyped_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_attention, NnAttentionImpl,
                              FBS_BIND_STREAM().Arg<BufH>().Attr<int32_t>("heads").Attr<int32_t>("dim_head").Attr<float>("scale")
                                  .Ret<BufH>());

static ffi::Error NnTimeMlpImpl(cudaStream_t stream, BufF tval, BufF W0, BufF b0, BufF W1, BufF b1, BufF Wcat, BufF bcat, float dt,
                                int32_t dim, ResF table) {
  return as_error(fbs_nn_time_mlp_f32(stream, tval.typed_data(), dt, dim, W0.typed_data(), b0.typed_data(), W1.typed_data(),
                                      b1.typed_data(), Wcat.typed_data(), bcat.typed_data(), (int32_t)table->element_count(),
                                      table->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_time_mlp, NnTimeMlpImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>().Arg<BufF>()
                                  .Attr<float>("dt").Attr<int32_t>("dim").Ret<BufF>());

static ffi::Error NnStemConvImpl(cudaStream_t stream, BufF x, BufF weight, BufF bias, int32_t Cout, ResF out_f32, ResH out_bf16) {
  auto d = x.dimensions();  // [B, H, W, Cin]
  return as_error(fbs_nn_stem_conv_f32(stream, x.typed_data(), d[0], (int32_t)d[1], (int32_t)d[2], (int32_t)d[3], Cout,
                                       weight.typed_data(), bias.typed_data(), opt(*out_f32), opt_raw(*out_bf16)));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_stem_conv, NnStemConvImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufF>().Attr<int32_t>("Cout").Ret<BufF>().Ret<BufH>());

static ffi::Error NnHeadConvImpl(cudaStream_t stream, BufF x, BufF weight, BufF bias, ResF out) {
  const int64_t C = x.dimensions().back(), rows = (int64_t)x.element_count() / (C > 0 ? C : 1);
  return as_error(fbs_nn_head_conv_f32(stream, x.typed_data(), rows, (int32_t)C, (int32_t)out->dimensions().back(),
                                       weight.typed_data(), bias.typed_data(), out->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_head_conv, NnHeadConvImpl,
                              FBS_BIND_STREAM().Arg<BufF>().Arg<BufF>().Arg<BufF>().Ret<BufF>());

static ffi::Error NnSpaceToDepthImpl(cudaStream_t stream, BufH in, ResH out) {
  auto d = in.dimensions();  // [B, H, W, C]
  return as_error(fbs_nn_space_to_depth_bf16(stream, in.untyped_data(), d[0], (int32_t)d[1], (int32_t)d[2], (int32_t)d[3],
                                             out->untyped_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(fbs_xla_nn_space_to_depth, NnSpaceToDepthImpl, FBS_BIND_STREAM().Arg<BufH>().Ret<BufH>());
#endif  // FBS_HAVE_XLA_FFI
