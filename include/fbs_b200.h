/*
 * fbs_b200 -- C ABI of the B200-native CSMC / particle-Gibbs / pMCMC hot path.
 *
 * Every entry point is `extern "C"`, takes a CUDA stream, DEVICE pointers and sizes, and
 * returns an int status (FBS_OK == 0).  No entry point allocates, frees, synchronises the
 * device or retains a pointer; all buffers (inputs, outputs, scratch) are owned by the
 * caller.  The library keeps no mutable global state besides a thread-local error string
 * (and the test-only implementation pins of fbs_debug_set_option), so it is re-entrant
 * from concurrent host threads (the XLA-FFI threading model).
 *
 * The reference (zgbkdlm/fbs, pure JAX) has no FFI of its own; each function below names
 * the reference Python function (file:line under /root/reference) whose arithmetic it
 * replaces.  INTEGRATION.md shows the jax.ffi binding a maintainer would add.
 *
 * Conventions
 *   - leading batch dimension B = independent chains (what the reference vmaps over,
 *     experiments/toy/gp_gibbs.py:172-173); all arrays are dense row-major;
 *   - keys are jax.random threefry keys, uint32[2] per chain;
 *   - floats are float32, indices int32 (jax_enable_x64 = False in the experiments);
 *   - "D" = du + dv, the joint dimension; X (unobserved, u) is the first du coordinates.
 */
#ifndef FBS_B200_H_
#define FBS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fbs_stream_t; /* cudaStream_t */

enum {
  FBS_OK = 0,
  FBS_ERR_INVALID_ARGUMENT = 1,
  FBS_ERR_CUDA = 2,
  FBS_ERR_UNSUPPORTED = 3
};

/* Conditional resamplers: fbs/samplers/csmc/resamplings.py.  Unconditional: fbs/samplers/resampling.py. */
enum {
  FBS_RESAMPLE_MULTINOMIAL = 0, /* resamplings.py:10-37 / resampling.py:62-68 (sorted uniforms) */
  FBS_RESAMPLE_KILLING = 1,     /* resamplings.py:40-88 / resampling.py:71-101 */
  FBS_RESAMPLE_SYSTEMATIC = 2,  /* resamplings.py:120-125 (no clip) / resampling.py:54-55 (clip) */
  FBS_RESAMPLE_STRATIFIED = 3   /* resampling.py:58-59 */
};

/* How the CSMC sweep initialises its particle set (fbs/samplers/gibbs.py:132-144). */
enum {
  FBS_INIT_DEGENERATE = 0, /* explicit_final=False: N copies of us_star[0], uniform weights (:140-144) */
  FBS_INIT_NORMAL = 1      /* explicit_final=True: N(0, I) draws, weights = likelihood at ts[0] (:133-137) */
};

int fbs_version(void);
/* Thread-local, human-readable description of the last non-OK return on this thread. */
const char* fbs_last_error(void);
/* Number of kernel launches issued through this library on this thread since the last reset. */
int64_t fbs_launch_count(void);
void fbs_reset_launch_count(void);
/* Test / measurement knob: pins one of several equivalent kernel implementations ("sweep_impl" = 1 | 2 | 3, "step_impl",
 * "step_tc_warps", "stepvec_impl", "sweep_g", "v3_twopass", "em_impl", "sweep_verbose", "v3_variant"; 0 = the library's own choice).
 * The options are process-wide relaxed atomics -- the only mutable global state of the library -- and never change
 * results beyond the documented float32 summation-order differences between the implementations.  The library does
 * not read the environment. */
int fbs_debug_set_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------
 * jax.random (threefry2x32, jax 0.4.26 non-partitionable) -- third-party semantics the
 * reference depends on at csmc.py:136,150,157; resamplings.py:66,71,74,84; gibbs.py:126,147,156;
 * smc.py:142,154,231,248; linear.py:220; simulators.py:81,91.
 * All are batched over B keys; outputs are [B, n] unless noted.
 * ---------------------------------------------------------------------------------------- */
int fbs_random_bits_u32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, uint32_t* out);
/* jax.random.split(key, num): out [B, num, 2] */
int fbs_random_split(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t num, uint32_t* out);
int fbs_random_uniform_f32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, float minval, float maxval,
                           float* out);
int fbs_random_normal_f32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, float* out);
int fbs_random_randint_i32(fbs_stream_t s, const uint32_t* keys, int64_t B, int64_t n, int32_t minval,
                           int32_t maxval, int32_t* out);
/* jax.random.choice(key, N, (n,), p=p[b]) : p [B, N] (need not be normalised), out [B, n] */
int fbs_random_choice_f32(fbs_stream_t s, const uint32_t* keys, const float* p, int64_t B, int64_t N, int64_t n,
                          int32_t* out);

/* ------------------------------------------------------------------------------------------
 * Resampling.  weights [B, N] normalised, idx_out [B, N].
 * cond: (key, weights, i, j, conditional) of resamplings.py; i, j are [B] (may be NULL when
 * conditional == 0).  Indices are bit-exact against the oracle on identical weights and keys;
 * cumulative sums are sequential float32 (DESIGN.md "summation order").
 * ---------------------------------------------------------------------------------------- */
int fbs_cond_resample_f32(fbs_stream_t s, int scheme, const uint32_t* keys, const float* weights, const int32_t* i,
                          const int32_t* j, int conditional, int64_t B, int64_t N, int32_t* idx_out);
int fbs_resample_f32(fbs_stream_t s, int scheme, const uint32_t* keys, const float* weights, int64_t B, int64_t N,
                     int32_t* idx_out);

/* ------------------------------------------------------------------------------------------
 * Forward noising.
 * ---------------------------------------------------------------------------------------- */
/* make_linear_sde(...)[2] = simulate_cond_forward(key, x0, ts, keep_path=True), fbs/sdes/linear.py:190-221.
 * F, sqrtQ: [K] exact-discretisation coefficients of the K intervals (linear.py:169-184).
 * x0 [B, D] (or [1, D] broadcast when x0_batched == 0).
 * Output layout: if rev == 0, path [B, K+1, D] in forward time.  If rev != 0 the path is
 * written time-REVERSED and split the way gibbs_kernel consumes it (gibbs.py:128-130):
 * out_u [B, K+1, du] = path[::-1][..., :du], out_v [B, K+1, D-du] = path[::-1][..., du:]
 * (either may be NULL).  With rev == 0, out_u receives the full path and du must equal D. */
int fbs_ou_forward_path_f32(fbs_stream_t s, const uint32_t* keys, const float* x0, int x0_batched, const float* F,
                            const float* sqrtQ, int64_t B, int64_t K, int64_t D, int64_t du, int rev, float* out_u,
                            float* out_v);

/* euler_maruyama(key, x0, ts, drift, dispersion, integration_nsteps=m, return_path=True),
 * fbs/sdes/simulators.py:53-106, for an affine drift  drift(x, t_{k,q}) = A_{k,q} x + a_{k,q}:
 * AT [K*m, D, D] with AT[kq][j][i] = A_{k,q}[i][j];  a [K*m, D];  ddt [K] sub-step size;
 * disp [K*m] = dispersion(t_{k,q}).  Output as in fbs_ou_forward_path_f32. */
int fbs_em_affine_path_f32(fbs_stream_t s, const uint32_t* keys, const float* x0, int x0_batched, const float* AT,
                           const float* a, const float* ddt, const float* disp, int64_t B, int64_t K, int64_t m,
                           int64_t D, int64_t du, int rev, float* out_u, float* out_v);

/* ------------------------------------------------------------------------------------------
 * Affine-Gaussian reverse-diffusion model: the closures transition_sampler / likelihood_logpdf
 * of experiments/toy/gp_gibbs.py:94-135 and experiments/sb/gibbs.py:93-132 with the joint
 * reverse drift written as  drift(uv, t_k) = M_k uv + m_k  for the K step times ts[:-1].
 *   u' = u + dt_k * drift_u + sd_k * eps                       (gp_gibbs.py:120-122)
 *   log w = sum_dv logN(v; v_prev + dt_k * drift_v, sd_k)      (gp_gibbs.py:132-135)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t K;        /* number of steps */
  int32_t du;       /* dim of the unobserved part u */
  int32_t dv;       /* dim of the observed part v   */
  int32_t reserved;
  const float* MT;  /* [K, D, D]  MT[k][j][i] = M_k[i][j]  (input-major, for coalesced reads) */
  const float* m;   /* [K, D] */
  const float* dt;  /* [K]   the step the closures multiply the drift by (reference: constant T/nsteps) */
  const float* sd;  /* [K]   sqrt(dt) * dispersion(T - t_k) */
  const float* lognorm; /* [K]  dv * log(2 pi sd_k^2), the Gaussian normaliser of the log-weight */
  const float* MTp; /* optional (may be NULL): [K, du, DP] the u-input rows of MT re-packed for the tiled sweep kernel:
                       row j = [M_k[0:du, j] zero-padded to dup | M_k[du:D, j] zero-padded to dvp], dup/dvp = du/dv
                       rounded up to 4, DP = dup + dvp.  Without it the sweeps use the general kernel. */
  const float* MTc; /* optional (may be NULL): tensor-core image of the same rows for the tcgen05 sweep kernel,
                       [K, nkb, 2 (hi, lo), 2 (k-chunks), nout/8, 8, 4] float32: for step k and K-block kb (8 inputs j),
                       the UMMA K-major core-matrix layout of B[o][j] = M_k[row(o)][j], o < du8: u output o,
                       o >= du8: v output o - du8 (du8/dv8 = du/dv rounded up to 8, nout = du8 + dv8 rounded up to 16),
                       split into hi = tf32-rounded value and lo = float32 remainder. */
} fbs_affine_model_t;

/* Test hook for the tcgen05 plumbing: D [128, nout] = A [128, K8] * B^T with the split-TF32 scheme of the sweep
 * kernel; Bimg is ONE step of the MTc image (nkb = K8 / 8 blocks). */
int fbs_debug_umma_gemm(fbs_stream_t s, const float* A, const float* Bimg, int32_t K8, int32_t nout, float* D);

/* Profiling hook of the tcgen05 per-timestep kernel: a device buffer of 8 int64 into which CTA 0 accumulates the
 * cycles it spends per phase (gather, barrier, noise, worker barrier, accumulator wait, epilogue, barrier, stores);
 * NULL switches it off (default).  Not thread safe; for scripts/step_tc_phases.py only. */
int fbs_debug_step_tc_timers(long long* dev_buf);
/* Profiling hook of the tcgen05 sweep kernel: a device buffer of 7 x 4 x 16 int64 into which CTA 0 writes clock64() stamps
 * of its phases for steps 64..67 of its first chain pair (E / X / R warp of each group, MMA warp); NULL switches it off
 * (default).  Not thread safe; for scripts/v3_timeline.py only. */
int fbs_debug_v3_timeline(long long* dev_buf);

/* Profiling hook of conv_gemm_kernel: a device buffer of 256 int64 receiving CTA 0's clock64() stamps ([0] entry, [1] prologue
 * done, [2] weights issued, [3] weights landed, [4] exit; per tile i at 8 + 8 i: producer's first load, MMA warp got the
 * accumulator buffer, first operands landed, all MMAs issued, epilogue start, epilogue end); NULL switches it off (default).
 * Not thread safe; for scripts/conv_timeline.py only. */
int fbs_debug_conv_timeline(long long* dev_buf);

/* Scratch the tiled sweep kernel needs for B chains (per-chain step vectors of all K steps); pass a device
 * buffer of at least this many bytes as `workspace` to fbs_csmc_forward_affine_f32 / fbs_pmcmc_filter_affine_f32.
 * With workspace == NULL (or too small) the general kernel runs instead. */
size_t fbs_sweep_workspace_bytes(const fbs_affine_model_t* model, int64_t B);

/* One CSMC step (csmc.py:132-148 body) -- the per-timestep fused kernel, for large particle
 * sets held in global memory.  us_prev [B, N, du], log_ws [B, N] (normalised) in;
 * A_out [B, N], us_out [B, N, du], log_ws_out [B, N] out.  step_keys [B, 2] = keys[k] of csmc.py:157. */
int fbs_csmc_step_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, int32_t k, int scheme,
                             const uint32_t* step_keys, const float* us_prev, const float* log_ws,
                             const float* v, const float* v_prev, const float* u_star, const int32_t* b_star_prev,
                             const int32_t* b_star, int64_t B, int64_t N, int32_t* A_out, float* us_out,
                             float* log_ws_out);

/* The three closures of the reference drivers evaluated on their own, for callers that compose the
 * samplers step by step (bootstrap filter / smoother, backward sampling):
 *   us_out  [B,N,du] = transition_sampler(us_prev, v_prev, t_k, key)   gp_gibbs.py:120-122  (needs tr_keys [B,2])
 *   lw_out  [B,N]    = likelihood_logpdf(v, us_prev, v_prev, t_k)      gp_gibbs.py:132-135  (needs v [B,dv])
 *   tlp_out [B,N]    = transition_logpdf(u_eval, us_prev, v_prev, t_k) gp_gibbs.py:125-129  (needs u_eval [B,du])
 * Any output may be NULL.  us_prev [B,N,du], v_prev [B,dv]. */
int fbs_affine_eval_f32(fbs_stream_t s, const fbs_affine_model_t* model, int32_t k, const uint32_t* tr_keys,
                        const float* us_prev, const float* v, const float* v_prev, const float* u_eval, int64_t B,
                        int64_t N, float* us_out, float* lw_out, float* tlp_out);

/* forward_pass(key, us_star, bs_star, vs, ts, init_sampler, init_likelihood_logpdf, transition_sampler,
 *              likelihood_logpdf, cond_resampling, nsamples), csmc.py:80-164 -- the whole K-step sweep
 * in one persistent kernel, particles resident on chip.
 *   keys [B,2]; us_star [B,K+1,du]; bs_star [B,K+1]; vs [B,K+1,dv].
 *   N = number of particles actually carried (nparticles for FBS_INIT_DEGENERATE, nparticles+1 for
 *   FBS_INIT_NORMAL, finding 6d of SURVEY.md); init_log_w: the per-particle initial log-weight BEFORE
 *   normalisation for FBS_INIT_DEGENERATE (reference: -log(nparticles), gibbs.py:144).
 * History outputs (any may be NULL): As [B,K,N], log_wss [B,K+1,N], uss [B,K+1,N,du].
 * Final-state outputs (any may be NULL): log_ws_last [B,N], us_last [B,N,du]. */
int fbs_csmc_forward_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, const uint32_t* keys,
                                const float* us_star, const int32_t* bs_star, const float* vs, int init_mode,
                                float init_log_w, int scheme, int64_t B, int64_t N, int32_t* As, float* log_wss,
                                float* uss, float* log_ws_last, float* us_last, void* workspace,
                                size_t workspace_bytes);

/* backward_scanning_pass(key, As, xss, log_w_T), csmc.py:230-270: B_T ~ Cat(normalise(log_w_T)) (barker_move,
 * :295-297), then ancestor tracing B_{t-1} = A_t[B_t].  keys [B,2] (key_bwd of csmc.py:65); As [B,K,N];
 * uss [B,K+1,N,du]; log_w_T [B,N].  Out: xs_star [B,K+1,du], bs_star [B,K+1]. */
int fbs_backward_scan_f32(fbs_stream_t s, const uint32_t* keys, const int32_t* As, const float* uss,
                          const float* log_w_T, int64_t B, int64_t K, int64_t N, int64_t du, float* xs_star,
                          int32_t* bs_star);

/* pmcmc_filter_step(key, vs_bridge, u0s, ts, transition_sampler, likelihood_logpdf, resampling, nparticles),
 * fbs/samplers/smc.py:115-158.  keys [B,2]; vs [B,K+1,dv]; u0s [B,N,du].
 * Out: uT [B,N,du], log_ell [B].  Optional history (may be NULL): inds [B,K,N], log_ws_hist [B,K,N]
 * (unnormalised log-weights of smc.py:144), us_hist [B,K,N,du]. */
int fbs_pmcmc_filter_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, const uint32_t* keys,
                                const float* vs, const float* u0s, int scheme, int64_t B, int64_t N, float* uT,
                                float* log_ell, int32_t* inds, float* log_ws_hist, float* us_hist, void* workspace,
                                size_t workspace_bytes);

/* bootstrap_filter(transition_sampler, measurement_cond_pdf, vs, ts, init_sampler, key, nparticles, resampling),
 * fbs/samplers/smc.py:9-88 (log-domain branch :58-74), the whole K-step scan in one launch.  step_keys [B,2] =
 * key_steps of smc.py:77 (the caller draws u0s with key_init, :78); vs [B,K+1,dv]; u0s [B,N,du].
 * Every particle is propagated BEFORE the resampling and its noise row travels with it (us = us_new[inds], :63,72).
 * Out: uT [B,N,du] (the last resampled set), log_nell [B] (NEGATIVE log-likelihood, :67).  Optional history (may be
 * NULL): inds [B,K,N], log_ws_hist [B,K,N] (unnormalised, :65), us_hist [B,K,N,du] (the resampled set after each step:
 * with u0s prepended, the `return_last=False` output of :85-88). */
int fbs_bootstrap_filter_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, const uint32_t* step_keys,
                                    const float* vs, const float* u0s, int scheme, int64_t B, int64_t N, float* uT,
                                    float* log_nell, int32_t* inds, float* log_ws_hist, float* us_hist);

/* Backward sampling over a stored particle history, one launch for the whole K-step recursion.
 * mode 0: backward_sampling_pass(key, transition_logpdf, vs, ts, uss, log_ws), fbs/samplers/csmc/csmc.py:167-227 --
 *         B_T ~ Cat(normalise(log_ws[-1])) with keys[-1], then for t = K-1..0: G = transition_logpdf(x, uss[t], vs[t], ts[t]),
 *         w = normalise(G - max G + log_ws[t]), B_t ~ Cat(w) with keys[q] (keys = split(key, K+1), consumed in order).
 * mode 1: bootstrap_backward_smoother(key, filter_us, vs, ts, transition_logpdf), fbs/samplers/smc.py:91-112 --
 *         u_T = filter_us[-1][randint(key)] (the UNSPLIT key, :109, as written upstream), then for k = K-1..0:
 *         w = normalise(transition_logpdf(u, filter_us[k], vs[k], ts[k])), draw with split(split(key)[1], K)[q].
 * keys [B,2]; vs [B,K+1,dv]; uss [B,K+1,N,du]; log_wss [B,K+1,N] (mode 0; NULL in mode 1).  shared_history != 0: vs, uss,
 * log_wss carry no chain axis and all B keys walk the SAME stored history (the reference vmaps the smoother over keys
 * only, tests/test_filters.py:138-141).  Out: xs_star [B,K+1,du]; bs_star [B,K+1] (may be NULL). */
int fbs_backward_sample_affine_f32(fbs_stream_t s, const fbs_affine_model_t* model, int mode, const uint32_t* keys,
                                   const float* vs, const float* uss, const float* log_wss, int shared_history, int64_t B,
                                   int64_t N, float* xs_star, int32_t* bs_star);

/* twisted_smc(key, y, ts, init_sampler, transition_logpdf, twisting_logpdf, twisting_prop_sampler, twisting_prop_logpdf,
 * resampling, nparticles), fbs/samplers/smc.py:261-309, with the affine closures of experiments/toy/gp_twisted.py:66-129
 * (reverse_drift(u, t) = M_t u + m_t; twisting N(y; u + reverse_drift dt, obs_var); proposal drift + g^2 grad log twisting) --
 * the whole scan in one launch.  Coefficient tables over the time indices 0..K (the scan walks ts[1:], the initial twisting
 * uses ts[0]): MT [K+1,d,d] with MT[t][j][i] = M_t[i][j], Mrow [K+1,d,d] = M_t row-major, m [K+1,d], sd [K+1] =
 * sqrt(dt) g_t, g2 [K+1] = g_t^2.  keys [B,2] = key_filter of smc.py:296; y [B,d] (y_batched != 0) or [d]; x0 [B,N,d] =
 * init_sampler(key_init, nparticles) (:299).  Out: samples [B,N,d], log_ws [B,N] (normalised).  Optional history (may be
 * NULL): inds [B,K,N], xs_hist [B,K,N,d], lw_hist [B,K,N]. */
int fbs_twisted_smc_affine_f32(fbs_stream_t s, const float* MT, const float* Mrow, const float* m, const float* sd,
                               const float* g2, float dt, float obs_var, int64_t K, int64_t d, const uint32_t* keys,
                               const float* y, int y_batched, const float* x0, int scheme, int64_t B, int64_t N,
                               float* samples, float* log_ws, int32_t* inds, float* xs_hist, float* lw_hist);

/* force_move(key, weights, k), fbs/samplers/gibbs.py:171-214, fused with the selection
 * x0 = uss[-1, idx] (gibbs.py:152-154).  log_ws_last [B,N]: normalised log-weights when weights_are_log != 0
 * (the kernel exponentiates, as gibbs.py:152 does), else the weights themselves.
 * us_last [B,N,du] (may be NULL when x0 is NULL); k [B].  Out: idx [B], alpha [B] (may be NULL), x0 [B,du] (may be NULL). */
int fbs_force_move_f32(fbs_stream_t s, const uint32_t* keys, const float* log_ws_last, int weights_are_log,
                       const float* us_last, const int32_t* k, int64_t B, int64_t N, int64_t du, int32_t* idx,
                       float* alpha, float* x0);

/* pcn_proposal(key, delta, x, mean, sampler) combination step, smc.py:161-168:
 * out = beta * (x + sqrt(delta/2) (r0 - mean)) + (1 - beta) mean + sqrt(1 - beta) (r1 - mean),
 * all [B, n] except mean [n] (shared).  */
int fbs_pcn_combine_f32(fbs_stream_t s, double delta, const float* x, const float* mean, const float* r0,
                        const float* r1, int64_t B, int64_t n, float* out);

/* Metropolis--Hastings accept/select of pmcmc_kernel, smc.py:244-258.
 * keys_mh [B,2]; in-place update of the chain state (uT [B,du], log_ell [B], ys [B,ny]) from the
 * proposal (prop_uTs [B,N,du] -> particle which_u, prop_log_ell [B], prop_ys [B,ny]).
 * Out: acceptance_prob [B], is_accepted [B] (uint8). */
int fbs_mh_accept_f32(fbs_stream_t s, const uint32_t* keys_mh, const float* prop_uTs, const float* prop_log_ell,
                      const float* prop_ys, int64_t B, int64_t N, int64_t du, int64_t ny, int32_t which_u, float* uT,
                      float* log_ell, float* ys, float* acceptance_prob, uint8_t* is_accepted);

/* ref_sampler(key, yT, n), experiments/toy/gp_gibbs.py:138-141: samples of the Gaussian conditional
 * N(a + Bm (yT - c), L^T L):  out[b, n, :] = a + Bm (yT[b] - c) + eps[b, n, :] @ L,  eps = normal(key, (N, du)),
 * with L = cholesky(cov) (lower, [du, du] row-major) exactly as gp_gibbs.py:141 multiplies it.
 * a [du], Bm [du, dv] row-major, c [dv], yT [B, dv]. */
int fbs_gaussian_ref_sample_f32(fbs_stream_t s, const uint32_t* keys, const float* yT, const float* a,
                                const float* Bm, const float* c, const float* L, int64_t B, int64_t N, int64_t du,
                                int64_t dv, float* out);

/* ------------------------------------------------------------------------------------------
 * Score network (fbs/nn/unet.py) on the tensor cores, and the NN-score closures
 * (experiments/imgs/inpainting.py:94-147, supr.py, experiments/sb_imgs/supr.py:84-129).
 * Activations are NHWC; "bf16" pointers are __nv_bfloat16 (passed as void*); P = H * W.
 * ------------------------------------------------------------------------------------------ */

/* Implicit-GEMM convolution, stride 1 (flax.linen.Conv call sites of unet.py: 3x3 pad 1 -> kh = kw = 3, off = -1;
 * 1x1 -> kh = kw = 1, off = 0; the 4x4 stride-2 Downsample (unet.py:50) as kh = kw = 2, off = 0 on the
 * fbs_nn_space_to_depth_bf16 copy).  The channel axis of the input may be split over two tensors (the U-Net's
 * skip concatenations, unet.py:335,340,355).  weight: bf16 [Cout][kh * kw * (C0 + C1)], K ordered (ty, tx, channel) --
 * the flax HWIO kernel reshaped to [K, Cout] and transposed.  out = conv + bias (+ residual), written as fp32
 * and / or bf16; pixel_shuffle != 0 writes 'b h w (h2 w2 c) -> b (h h2) (w w2) c' (fbs/nn/utils.py:53-57). */
typedef struct {
  int32_t B, H, W;      /* output pixels per sample */
  int32_t Hin, Win;     /* input pixels per sample (0: same as the output) */
  int32_t C0, C1, Cout; /* source channels (multiples of 64; C1 = 0: one source), output channels (multiple of 16) */
  int32_t kh, kw, off_h, off_w;
  int32_t pixel_shuffle, reserved;
  const void* in0;      /* bf16 [B, Hin, Win, C0] */
  const void* in1;      /* bf16 [B, Hin, Win, C1] or NULL */
  const void* weight;
  const float* bias;     /* [Cout] or NULL */
  const float* residual; /* fp32, same layout as the output, or NULL */
  float* out_f32;        /* either output may be NULL */
  void* out_bf16;
  float* gn_partials;    /* NULL, or [B][slots][Cout / 4][2]: sums and sums of squares of the output per (sample, slot, channel
                          * quad) for the GroupNorm that follows (fbs_nn_groupnorm_swish_stats); slots: fbs_nn_conv_gn_layout */
} fbs_nn_conv_t;
int fbs_nn_conv_bf16(fbs_stream_t s, const fbs_nn_conv_t* args);
/* Slots per sample of `gn_partials` for this call's tiling, or 0 when the call cannot produce them (tiles that hold several
 * samples, a residual, pixel shuffle).  Host-side only, no launch. */
int fbs_nn_conv_gn_layout(const fbs_nn_conv_t* args, int32_t* slots_per_sample);

/* swish(GroupNorm(x) * gamma + beta [* (1 + scale) + shift]) [+ residual]   (unet.py:144-155,159-160,172).
 * x fp32 [B, P, C]; time_scale_shift: [2 C] = (scale | shift) shared by the batch, or NULL. */
int fbs_nn_groupnorm_swish_f32(fbs_stream_t s, const float* x, int64_t B, int32_t P, int32_t C, int32_t groups, const float* gamma,
                               const float* beta, const float* time_scale_shift, const float* residual, float eps,
                               float* out_f32, void* out_bf16);
/* The same with the statistics taken from the producing convolution's `gn_partials` (flax forms var = E[x^2] - E[x]^2 from
 * the two means, flax.linen.normalization._compute_stats with use_fast_variance): one streaming pass, no reduction over the
 * activations.  x: fp32 [B, P, C] (x_f32) or bf16 (x_bf16); exactly one of the two is non-NULL.
 * ln_gamma / ln_out_bf16 (both or neither; C <= 128): also LayerNorm(result) * ln_gamma as bf16 -- the pre-normalisation of the
 * attention block that follows a ResnetBlock (unet.py:258), from the values already in registers. */
int fbs_nn_groupnorm_swish_stats(fbs_stream_t s, const float* x_f32, const void* x_bf16, const float* partials, int32_t slots,
                                 int64_t B, int32_t P, int32_t C, int32_t groups, const float* gamma, const float* beta,
                                 const float* time_scale_shift, const float* residual, float eps, float* out_f32, void* out_bf16,
                                 const float* ln_gamma, float ln_eps, void* ln_out_bf16);
/* LayerNorm over channels, scale only (unet.py:243,258) [+ residual (unet.py:264)].  x fp32 [rows, C]. */
int fbs_nn_layernorm_f32(fbs_stream_t s, const float* x, int64_t rows, int32_t C, const float* gamma, const float* residual,
                         float eps, float* out_f32, void* out_bf16);
/* LinearAttention core (unet.py:227-239) and Attention core (unet.py:192-199): qkv bf16 [B, P, 3 heads dim_head]
 * -> bf16 [B, P, heads dim_head] (bf16 so that the projection stays L2 resident between the two kernels). */
int fbs_nn_linear_attention_bf16(fbs_stream_t s, const void* qkv, int64_t B, int32_t P, int32_t heads, int32_t dim_head,
                                 void* out_bf16);
int fbs_nn_attention_bf16(fbs_stream_t s, const void* qkv, int64_t B, int32_t P, int32_t heads, int32_t dim_head, float scale,
                          void* out_bf16);
/* Time embedding (unet.py:293-300, base.py:44-77) + every ResnetBlock's Dense(2 dim)(swish(time_emb)) (:148-149):
 * table[nout] = swish(temb) @ Wcat + bcat; tval: device scalar with the network time. */
int fbs_nn_time_mlp_f32(fbs_stream_t s, const float* tval, float dt, int32_t dim, const float* W0, const float* b0, const float* W1,
                        const float* b1, const float* Wcat, const float* bcat, int32_t nout, float* table);
/* First (7x7, unet.py:286-291) and last (1x1, unet.py:363) convolutions, fp32 on the CUDA cores. */
int fbs_nn_stem_conv_f32(fbs_stream_t s, const float* x, int64_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                         const float* weight, const float* bias, float* out_f32, void* out_bf16);
int fbs_nn_head_conv_f32(fbs_stream_t s, const float* x, int64_t rows, int32_t C, int32_t Cimg, const float* weight,
                         const float* bias, float* out);
/* out[b, i, j, (r, s, c)] = in[b, 2 i - 1 + r, 2 j - 1 + s, c] (zero outside): bf16 [B, H, W, C] -> [B, H/2 + 1, W/2 + 1, 4 C]. */
int fbs_nn_space_to_depth_bf16(fbs_stream_t s, const void* in, int64_t B, int32_t H, int32_t W, int32_t C, void* out);
/* dataset.concat (fbs/data/images.py:352-363): img[b] = scatter(us[b] -> unobs_idx, v -> obs_idx). */
int fbs_nn_assemble_image_f32(fbs_stream_t s, const float* us, const float* v, const int32_t* unobs_idx, const int32_t* obs_idx,
                              int64_t B, int32_t p, int32_t q, int32_t c, float* img);
/* transition_sampler + likelihood_logpdf (inpainting.py:122-147) from ONE score evaluation: rd = -a x + g2 score;
 * us_new = u + rd_u dt + sd normal(key, (B, p, c)) (NULL: skip), mean_out = u + rd_u dt (NULL: skip),
 * lw[b] = sum logN(v_next; v_prev + rd_v dt, sd) (NULL: skip).  The B particles of this call are rows
 * [row_offset, row_offset + B) of a particle set of rows_total rows (a particle-sharded sweep draws its slice of the
 * SAME normal(key, (rows_total, p, c)) array; unsharded: row_offset = 0, rows_total = B).  pin_row / pin_value (both NULL or
 * both given): a DEVICE int32 holding the global row b*_{k+1} of the reference particle and its value u*_{k+1} [p c]; that
 * row of us_new receives pin_value instead of a propagated particle (csmc.py:143) -- no separate scatter launch. */
int fbs_nn_em_step_f32(fbs_stream_t s, const float* img, const float* score, const int32_t* unobs_idx, const int32_t* obs_idx,
                       const float* v_next, const uint32_t* key, int64_t B, int32_t p, int32_t q, int32_t c, float a, float g2,
                       float dt, float sd, int64_t row_offset, int64_t rows_total, const int32_t* pin_row, const float* pin_value,
                       float* us_new, float* mean_out, float* lw);
/* normalise(log_ws, log_space=True) and its exp (csmc.py:146,139) for B weight vectors [B, N] in one launch:
 * log_w = lw - logsumexp(lw), w = exp(log_w) (either output may be NULL).  The per-step glue of the score-network sweeps. */
int fbs_normalise_logw_f32(fbs_stream_t s, const float* lw, int64_t B, int64_t N, float* log_w, float* w);
/* One Euler--Maruyama sub-step of fbs/sdes/simulators.py:87 whose drift has ALREADY been evaluated by a network -- the
 * forward sampler of the Schroedinger-bridge image runs (experiments/sb_imgs/supr.py:132-137: drift = nn_drift(x, t, param_fwd),
 * integration_nsteps = 1):  out[b, e] = x[b, e] + drift[b, e] * ddt + gs * normal(keys[b], (n,))[e]  with
 * gs = dispersion(t) * sqrt(ddt).  keys [B, 2] = the interval's key (split(key, K)[k], simulators.py:81); x, drift, out [B, n]. */
int fbs_em_drift_step_f32(fbs_stream_t s, const uint32_t* keys, const float* x, const float* drift, int64_t B, int64_t n,
                          float ddt, float gs, float* out);
/* dst[b, :] = src[clamp(idx[b], 0, src_rows - 1), :]  (the ancestor gather, csmc.py:140; an out-of-range index -- the
 * unclipped systematic scheme of resamplings.py:120-125 can return N -- is clamped, as a JAX gather does). */
int fbs_gather_rows_f32(fbs_stream_t s, const float* src, const int32_t* idx, int64_t B, int64_t row, int64_t src_rows,
                        float* dst);

/* Particle-sharded sweep (one chain over the GPUs of an NVSwitch box; no upstream counterpart): the ancestor gather over
 * PEER MEMORY.  srcs: device array of one pointer per rank to that rank's particle rows [rows_per_rank, row] (the local
 * buffer for the calling rank, CUDA-IPC imports for the others); idx: GLOBAL parent row of each of the B local children.
 * dst[b, :] = srcs[g / rows_per_rank][g % rows_per_rank, :], g = clamp(idx[b], 0, n_ranks rows_per_rank - 1) -- NVLink loads
 * inside the kernel. */
int fbs_gather_rows_peer_f32(fbs_stream_t s, const float* const* srcs, const int32_t* idx, int64_t B, int64_t row,
                             int64_t rows_per_rank, int64_t n_ranks, float* dst);
/* CUDA IPC plumbing for the above (host side; the 64-byte handles travel through the caller's own channel).  export:
 * handle of the allocation containing dev_ptr and dev_ptr's offset inside it; import: the peer's view of that address
 * (peer access enabled lazily); release: closes an import. */
int fbs_ipc_export(const void* dev_ptr, unsigned char* handle64, int64_t* offset);
int fbs_ipc_import(const unsigned char* handle64, int64_t offset, void** out_ptr);
int fbs_ipc_release(void* imported_ptr, int64_t offset);
int fbs_nn_f32_to_bf16(fbs_stream_t s, const float* x, int64_t n, void* y);

#ifdef __cplusplus
}
#endif
#endif /* FBS_B200_H_ */
