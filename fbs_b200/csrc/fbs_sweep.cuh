// Parameters shared by the whole-sweep kernels (csmc_kernels.cu: general v1; sweep_v2.cu: tiled fast path).
#pragma once
#include <stdint.h>

namespace fbs {

// MODE_BOOTSTRAP = the bootstrap filter of smc.py:9-88: the pMCMC filter's weights / resampling, but every particle is
// propagated BEFORE the resampling (its noise row travels with it: us = us_new[inds], smc.py:63,72) and the evidence is
// accumulated as a NEGATIVE log-likelihood (smc.py:67).  General kernel only.
enum { MODE_CSMC = 0, MODE_PMCMC = 1, MODE_BOOTSTRAP = 2 };

struct SweepParams {
  int K, du, dv;
  const float *MT, *m, *dt, *sd, *lognorm;
  const uint32_t* keys;
  const float* us_star;
  const int32_t* bs_star;
  const float* vs;
  const float* u0s;
  int mode, init_mode, scheme;
  float init_log_w;
  int64_t B;
  int N, G;
  int32_t* As;
  float *log_wss, *uss, *log_ws_last, *us_last;
  float *uT, *log_ell;
  int32_t* inds;
  float *lw_hist, *us_hist;
  // v2 only: packed drift matrices [K][du][DP] (u-input rows; per row [u outputs dup | v outputs dvp], zero padded)
  // and the per-chain step vectors precomputed into caller workspace [B][K+1][DP]
  const float* MTp;
  float* ws;
  int dup, dvp;
  // v3 only: tensor-core image of the step matrices (fbs_b200.h: MTc)
  const float* MTc;
  // v3 only: optional phase time stamps of CTA 0 (fbs_debug_v3_timeline; NULL = off)
  long long* dbg;
};


// sweep_v2.cu.  Returns FBS_OK, an error, or -1 when the shape is not eligible for the fast path.
int launch_sweep_v2(void* stream, SweepParams& p);
// sweep_v3.cu (tcgen05).  Same return convention.
int launch_sweep_v3(void* stream, SweepParams& p);
// sweep_warp.cu (one warp per chain, narrow states).  Same return convention.
int launch_sweep_warp(void* stream, SweepParams& p);
int launch_umma_selftest(void* stream, const float* A, const float* Bimg, int K8, int nout, float* D);
// fills p.ws with the per-chain step vectors of all K + 1 slots (sweep_v2.cu)
int launch_stepvec(void* stream, SweepParams& p);
size_t sweep_v2_workspace_bytes(int64_t B, int K, int du, int dv);

}  // namespace fbs
