#!/usr/bin/env python
"""Headline benchmark: CSMC particle-steps/s of the particle-GIBBS SWEEP (BASELINE.json metric).

Workload (the metric's configuration; experiments/toy/gp_gibbs.py scaled out as BASELINE.json configs[1] scales the toy
sampler out): d = du = dv = 100, K = 200 time steps, N = 100 particles, conditional `killing` resampling
(gibbs.py:149), explicit backward / forced move (eb=True, ef=False as in tests/test_gibbs.py), thousands of independent
chains per GPU, chains partitioned over GPUs with no communication (weak scaling: chains per GPU fixed).

One "step" = one gibbs_kernel application (gibbs.py:68-168) to every chain: forward noising, the K-step conditional SMC
sweep, forced move, second forward noising, next reference indices.  particle-steps per step = chains * N * K.

  value     device-resident: chain state lives in HBM, CUDA-event timed, max over ranks
  e2e       the same step through the public numpy-in / numpy-out API with pinned HOST buffers
            (H2D of keys / x0 / bs_star and D2H of x0, us_star, bs_star, changed inside the timed region)
  roofline  dominant kernel (sweep_v3_kernel behind fbs_csmc_forward_affine_f32) timed live with CUDA events
  cpu_baseline / --impl reference: the oracle port of the same sweep on the host cores, bounded sample, warm pool
  secondary the pMCMC step (round-1 headline), Gaussian-SB Gibbs, stand-alone HBM kernels, score network, sharded sweep
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_TOY, K_STEPS, N_PART, DELTA = 100, 200, 100, 0.005
ALG_BYTES_PER_PARTICLE_STEP = 8 * D_TOY + 16      # read parent + write child + weight + ancestor


def gp_setup(d):
    """experiments/toy/gp_pmcmc.py:28-54 (host, float64) + a synthetic observation y0."""
    zs = np.linspace(0., 5., d)
    cov = np.exp(-np.abs(zs[None, :] - zs[:, None]))
    jm = np.zeros(2 * d)
    jc = np.block([[cov, cov], [cov, cov + np.eye(d)]])
    rng = np.random.default_rng(666)
    y0 = (np.linalg.cholesky(cov) @ rng.standard_normal(d) + rng.standard_normal(d)).astype(np.float32)
    return jm, jc, y0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """Wall-clock mark (start / end of the timed region)."""
        return time.time()

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken inside [t0, t1] (the timed region).  The sampler is started before the warm-up steps
        (nvidia-smi needs a few hundred ms to come up -- longer than a short timed region); if no sample landed inside the
        region itself, the samples of the warm-up steps (same kernels, same load) are used and the result says so."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        inside = [ln for ts_, ln in self.lines if t0 is None or (t0 <= ts_ <= t1 + 0.1)]
        window = 'timed region'
        if not inside:
            inside = [ln for _, ln in self.lines]
            window = 'warm-up + timed region (no sample landed inside the timed region itself)'
        for ln in inside:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def rng_bound(normals_per_s):
    """The kernel's normal draws per second against the stand-alone rate of the jax.random-compatible generator (threefry2x32 +
    XLA's erf_inv) measured by scripts/rng_microbench_normals.cu -- the resource that actually binds."""
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'profiles', 'rng_peaks.json')))['normals_per_s'])
    except Exception:
        return None
    return {'achieved': normals_per_s / 1e9, 'peak': peak / 1e9, 'unit': 'G normals/s per GPU', 'frac': normals_per_s / peak,
            'note': 'every particle-step draws du normals in the kernel; peak = stand-alone generator micro-benchmark'}


def recorded_traffic():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return None


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the same Gibbs sweep, one process per host core
# ----------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(d, K, N):
    """Pool initializer: imports, model set-up (host float64 Cholesky etc.) -- NOT part of any timed region."""
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['OPENBLAS_NUM_THREADS'] = '1'
    os.environ['MKL_NUM_THREADS'] = '1'
    from oracle import jax_random as jr
    from oracle import models as om, sdes as osd, gibbs as ogibbs
    jm, jc, y0 = gp_setup(d)
    ts = np.linspace(0., 1., K + 1)
    sde = osd.StationaryConstLinearSDE(a=-0.5, b=1.)
    _CPU.update(jr=jr, ogibbs=ogibbs, sde=sde, y0=y0, d=d, K=K, N=N,
                model=om.JointGaussianDiffusionModel(sde, jm, jc, d, ts, 1., dtype=np.float32))


def _cpu_worker(args):
    """``nchains`` independent chains x one Gibbs sweep each (gibbs.py:68-168 as restated in oracle/gibbs.py, with the
    Cholesky-per-call closures of gp_gibbs.py:78-135), float32.  Returns the worker's own steady-state loop time."""
    seed, nchains = args
    jr, ogibbs, model, sde, d, K, N = (_CPU[k] for k in ('jr', 'ogibbs', 'model', 'sde', 'd', 'K', 'N'))
    key = jr.PRNGKey(seed)
    x0 = np.zeros(d, np.float32)
    us_star = np.zeros((K + 1, d), np.float32)
    bs = np.zeros(K + 1, np.int32)
    t0 = time.perf_counter()
    for c in range(nchains):
        key, sub = jr.split(key)
        ogibbs.gibbs_kernel(sub, x0, _CPU['y0'], us_star, bs, model.ts, model.fwd_sampler, sde, model.unpack, N,
                            model.transition_sampler, model.transition_logpdf, model.likelihood_logpdf)
    return time.perf_counter() - t0


class CpuPool:
    """A warm pool of one oracle process per host core: process spawn, imports and model set-up happen ONCE, outside every
    timed region (their cost is reported separately as ``spawn_s``)."""

    def __init__(self, d=D_TOY, K=K_STEPS, N=N_PART, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        self.d, self.K, self.N = d, K, N
        t0 = time.perf_counter()
        self.pool = mp.get_context('spawn').Pool(self.cores, initializer=_cpu_init, initargs=(d, K, N))
        self.pool.map(_cpu_worker, [(7 + i, 0) for i in range(self.cores)])      # every worker is up and initialised
        self.spawn_s = time.perf_counter() - t0
        self.calls = 0

    def sweep(self, chains_per_core):
        """One bounded sample: ``cores * chains_per_core`` chains, one sweep each, all cores busy.  Returns
        (particle-steps/s, seconds): the time is the slowest worker's own loop time (steady state, no start-up)."""
        self.calls += 1
        times = self.pool.map(_cpu_worker, [(1000 * self.calls + i, chains_per_core) for i in range(self.cores)])
        t = max(times)
        return self.cores * chains_per_core * self.N * self.K / t, t

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(chains_per_core=8, warm_chains=1, pool=None):
    own = pool is None
    pool = pool or CpuPool()
    pool.sweep(warm_chains)                                     # warm-up: page in numpy / scipy code paths
    v, t = pool.sweep(chains_per_core)
    res = {'value': v, 'unit': 'particle-steps/s', 'cores': pool.cores, 'kind': 'port',
           'sample': f'{pool.cores * chains_per_core} chains x 1 Gibbs sweep (d={pool.d}, K={pool.K}, N={pool.N}, conditional killing) '
                     f'with the NumPy float32 oracle (oracle/gibbs.py, Cholesky-per-call closures as in gp_gibbs.py:78-135), one '
                     f'process per core, {chains_per_core} chains per core, steady-state loop time of the slowest worker '
                     f'{t:.1f} s; warm pool (spawn + imports + model set-up {pool.spawn_s:.1f} s, not timed)',
           'spawn_s': pool.spawn_s,
           'note': 'JAX is not installable in this image, so this is the CPU PORT of the reference (NumPy), not upstream JAX'}
    if own:
        pool.close()
    return res


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the SAME workload (Gibbs sweep, d=100, K=200, N=100, killing).
    JAX cannot be installed here, so this times the oracle port on all host cores; each step is a bounded sample
    (two chains per core)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t_all = time.perf_counter()
    per_core = 2
    pool = CpuPool()
    for _ in range(max(1, min(args.warmup, 3))):
        pool.sweep(1)
    times = []
    for _ in range(max(1, args.steps)):
        times.append(pool.sweep(per_core)[1])
    pool.close()
    psteps = pool.cores * per_core * N_PART * K_STEPS
    ms = 1e3 * float(np.mean(times))
    v = psteps / (ms * 1e-3)
    base = {'value': v, 'unit': 'particle-steps/s', 'cores': pool.cores, 'kind': 'port',
            'sample': f'each step = {pool.cores * per_core} chains x 1 Gibbs sweep with the NumPy float32 oracle, one process per core '
                      f'(warm pool; spawn {pool.spawn_s:.1f} s not timed); steady-state loop time of the slowest worker',
            'spawn_s': pool.spawn_s}
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'particle-steps/s', 'n_gpus': args.gpus,
            'steps': len(times), 'warmup': max(1, min(args.warmup, 3)), 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_NAME, 'd': D_TOY, 'K': K_STEPS, 'N': N_PART, 'resampling': 'conditional killing',
                       'chains_per_step': pool.cores * per_core,
                       'note': 'JAX is not installable in this image: the reference arm is the NumPy oracle PORT of the reference '
                               'on the host cores (cpu_baseline.kind = "port"), same workload, each step a bounded sample of it'},
            'cpu_baseline': base,
            'e2e': {'value': v, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'wall_s': time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# secondary workloads (short; reported under "secondary", never the headline value)
# ----------------------------------------------------------------------------------------------
UNET_FLOPS = {(28, 28, 1): 2.788e9, (64, 64, 3): 14.635e9}   # per evaluation, SURVEY.md App. D


def _score_model(shape, K, seed=0):
    import fbs_b200
    from fbs_b200 import sdes
    from fbs_b200.nn import ScoreUNet, ScoreNetModel, random_unet_params
    H, W, C = shape
    T = 2.0
    ts = np.linspace(0., T, K + 1)
    sde = sdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=5., t0=0., T=T)
    side = 15 if H == 28 else H // 2                                   # inpaint-15 (MNIST) / inpaint-32 (CelebA-64)
    off = (H - side) // 2
    rect = np.array([(i + off) * W + (j + off) for i in range(side) for j in range(side)], dtype=np.int32)
    obs = np.setdiff1d(np.arange(H * W, dtype=np.int32), rect)
    net = ScoreUNet(random_unet_params(seed, C), shape, dt=T / 200)
    return ScoreNetModel(net, sde, ts, T, rect, obs), ts, rect, obs


def secondary_mnist(nparticles=101, steps=12):
    """configs[3]: MNIST 28x28 inpainting-15, nparticles + 1 = 101 particles ('gibbs-eb-ef', imgs_gibbs.sh:37-38), random-init
    score U-Net: CSMC steps (resample + gather + ONE score evaluation + EM step + weights) through forward_pass_nn."""
    import torch
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    from fbs_b200 import random as fr
    shape = (28, 28, 1)
    model, ts, rect, obs = _score_model(shape, steps)
    rng = np.random.default_rng(0)
    us_star = rng.standard_normal((steps + 1, rect.size, 1)).astype(np.float32)
    vs = rng.uniform(size=(steps + 1, obs.size, 1)).astype(np.float32)
    bs = np.zeros((steps + 1,), np.int32)
    init = csmc.DegenerateInit(nparticles)
    args = (fr.PRNGKey(3), us_star, bs, vs, model, init, R.killing.scheme, nparticles)
    csmc.forward_pass_nn(*args, history=False)          # warm-up: CUDA graph capture
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    csmc.forward_pass_nn(*args, history=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # several conditioning targets at once (inpainting.py:205-210 loops over them; here C chains share every score evaluation)
    chains = 8
    us_c = rng.standard_normal((chains, steps + 1, rect.size, 1)).astype(np.float32)
    vs_c = rng.uniform(size=(chains, steps + 1, obs.size, 1)).astype(np.float32)
    bs_c = np.zeros((chains, steps + 1), np.int32)
    keys_c = fr.split(fr.PRNGKey(4), chains)
    args_c = (keys_c, us_c, bs_c, vs_c, model, init, R.killing.scheme, nparticles)
    csmc.forward_pass_nn_chains(*args_c)
    torch.cuda.synchronize()
    e0.record()
    csmc.forward_pass_nn_chains(*args_c)
    e1.record()
    torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1)
    # the network alone (graph replays back to back)
    x = torch.randn(nparticles, *shape, device='cuda')
    model.unet(x, 0.5)
    torch.cuda.synchronize()
    e0.record()
    for i in range(10):
        model.unet(x, 0.5)
    e1.record()
    torch.cuda.synchronize()
    net_ms = e0.elapsed_time(e1) / 10
    flops = UNET_FLOPS[shape] * nparticles
    peak = 1391.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['bf16_tflops_sustained'])
    except Exception:
        pass
    return {'workload': 'configs[3]: MNIST 28x28 inpaint-15, 101 particles, random-init fbs/nn score U-Net (12.99 M parameters), '
                        'score GEMMs on tcgen05 (bf16 operands, fp32 accumulate)',
            'value': nparticles * steps / (ms * 1e-3), 'unit': 'particle-steps/s', 'steps': steps, 'ms_per_csmc_step': ms / steps,
            'dtype': 'bf16', 'score_net_ms_per_eval': net_ms,
            'batched_targets': {'chains': chains, 'value': chains * nparticles * steps / (ms_c * 1e-3), 'unit': 'particle-steps/s',
                                'ms_per_csmc_step': ms_c / steps,
                                'how': 'forward_pass_nn_chains: 8 conditioning targets x 101 particles through one score '
                                       'evaluation per step (bit-identical per chain to the single-target sweep)'},
            'roofline': {'bound': 'tensor', 'kernel': 'conv_gemm_kernel + fused norm / attention kernels (one score evaluation)',
                         'achieved': flops / (net_ms * 1e-3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': flops / (net_ms * 1e-3) / 1e12 / peak, 'traffic': None,
                         'flops_per_particle_step': UNET_FLOPS[shape], 'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained'}}



def secondary_pmcmc(chains=2072, steps=3):   # 7 full waves of 148 SMs x 2 chains per CTA
    """configs[1]: the pseudo-marginal MCMC step of experiments/toy/gp_pmcmc.py (d = 100, K = 200, N = 100, stratified, pCN
    delta = 0.005; round 1's headline), device resident, pmcmc_kernel through the public API."""
    import torch
    import fbs_b200
    from fbs_b200 import sdes, parallel, random as fr
    from fbs_b200.samplers import pmcmc_kernel, stratified
    d, K, N = D_TOY, K_STEPS, N_PART
    jm, jc, y0 = gp_setup(d)
    ts = np.linspace(0., 1., K + 1)
    sde = sdes.StationaryConstLinearSDE(a=-0.5, b=1.)
    model = fbs_b200.AffineGaussianModel.from_linear_sde(sde, jm, jc, d, ts, T=1.)
    dev = torch.device('cuda', torch.cuda.current_device())
    y0_d = torch.from_numpy(y0).to(dev)
    kw = dict(ts=ts, fwd_ys_sampler=model.fwd_ys_sampler, sde=sde, ref_sampler=model.ref_sampler,
              transition_sampler=model.transition_sampler, likelihood_logpdf=model.likelihood_logpdf,
              resampling=stratified, nparticles=N, delta=DELTA)
    ys = model.fwd_ys_sampler(torch.from_numpy(parallel.chain_keys(fr.PRNGKey(800), chains, 0, 1)).to(dev), y0_d)
    state = (torch.zeros((chains, d), device=dev), torch.zeros((chains,), device=dev), ys)

    def step(i, state):
        keys = torch.from_numpy(parallel.chain_keys(fr.PRNGKey(900 + i), chains, 0, 1)).to(dev)
        uT, le, ys_, st = pmcmc_kernel(keys, state[0], state[1], state[2], y0_d, **kw)
        return (uT, le, ys_), st

    for i in range(3):
        state, st = step(i, state)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        state, st = step(3 + i, state)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {'workload': 'configs[1]: pseudo-marginal MCMC step (gp_pmcmc.py; d=100, K=200, N=100, stratified, pCN delta=0.005), '
                        f'{chains} chains, device resident', 'value': chains * N * K / (ms * 1e-3), 'unit': 'particle-steps/s',
            'ms_per_step': ms, 'steps': steps, 'dtype': 'f32',
            'mh_acceptance_rate_last_step': float(st.is_accepted.float().mean().item())}


def secondary_sb(chains=16384, steps=3, d=10, K=100, N=64):
    """configs[2]: particle Gibbs through a Gaussian Schroedinger bridge (experiments/sb/gibbs.py: d = 10, K = 100, T = 1,
    forward sampler = Euler--Maruyama with 10 sub-steps, obs_var = 0.1, reference N(1, a a^T)), batched over conditioning
    targets, device resident, gibbs_kernel through the public API."""
    import torch
    import fbs_b200
    from fbs_b200 import parallel, random as fr
    from fbs_b200.samplers import gibbs_kernel
    rng = np.random.default_rng(3)
    zs = np.linspace(0., 5., d)
    cov = np.exp(-np.abs(zs[None] - zs[:, None]))
    obs_var = 0.1
    jm = np.zeros(2 * d)
    jc = np.block([[cov, cov], [cov, cov + obs_var * np.eye(d)]])
    a_ = rng.normal(size=(2 * d, 2 * d))
    ts = np.linspace(0., 1., K + 1)
    model = fbs_b200.AffineGaussianModel.from_gaussian_sb(jm, jc, np.ones(2 * d), a_ @ a_.T, d, ts, sig=1., em_nsteps=10)
    dev = torch.device('cuda', torch.cuda.current_device())
    # one conditioning target per chain (sb/gibbs.py loops over 100 ids; here they are batched)
    y0 = (rng.normal(size=(chains, d)) @ np.linalg.cholesky(cov + obs_var * np.eye(d)).T).astype(np.float32)
    y0_d = torch.from_numpy(y0).to(dev)
    x0 = torch.zeros((chains, d), device=dev)
    bs = torch.zeros((chains, K + 1), dtype=torch.int32, device=dev)

    def sweep(i, x0, bs):
        keys = torch.from_numpy(parallel.chain_keys(fr.PRNGKey(700 + i), chains, 0, 1)).to(dev)
        x0, _, bs, _ = gibbs_kernel(keys, x0, y0_d, None, bs, ts, model.fwd_sampler, None, model.unpack, N,
                                    model.transition_sampler, model.transition_logpdf, model.likelihood_logpdf)
        return x0, bs

    for i in range(3):
        x0, bs = sweep(i, x0, bs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        x0, bs = sweep(3 + i, x0, bs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if not bool(torch.isfinite(x0).all()):
        raise RuntimeError('Gaussian SB Gibbs sweep produced non-finite samples')
    value = chains * N * K / (ms * 1e-3)
    peak = measured_peak()[0]
    achieved = value * (8 * d + 16) / 1e9
    return {'workload': f'configs[2]: Gaussian Schroedinger bridge particle Gibbs (sb/gibbs.py; d={d}, K={K}, N={N}, EM forward '
                        f'sampler m=10), {chains} conditioning targets, device resident', 'value': value,
            'unit': 'particle-steps/s', 'ms_per_sweep': ms, 'steps': steps, 'dtype': 'f32',
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'algorithmic_bytes_per_particle_step': 8 * d + 16}}


def secondary_hbm_kernels(shapes=None, steps=None):
    """The stand-alone resampling kernels and the per-timestep fused step (ancestors + transition / weight + normalise) of
    north_star items (1)-(2), each timed alone with CUDA events against the HBM roofline with the algorithmic bytes of
    SURVEY 8(d): 8 B per particle for resampling, 8 du + 16 B per particle-step for the transition."""
    import torch
    import fbs_b200
    from fbs_b200 import sdes, _native as nat, random as fr
    from fbs_b200._tensor import ptr, stream
    peak = measured_peak()[0]
    dev = torch.device('cuda', torch.cuda.current_device())

    def timeit(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = []
    for B, N in (shapes or ((65536, 100), (262144, 100), (16384, 1024), (1024, 16384))):
        w = torch.rand(B, N, device=dev)
        w /= w.sum(dim=1, keepdim=True)
        keys = fr.split(torch.from_numpy(fr.PRNGKey(1)).to(dev), B)
        idx = torch.empty(B, N, dtype=torch.int32, device=dev)
        for name, scheme in (('stratified', nat.RESAMPLE_STRATIFIED), ('killing', nat.RESAMPLE_KILLING)):
            ms = timeit(lambda: nat.call('fbs_resample_f32', stream(), scheme, ptr(keys), ptr(w), B, N, ptr(idx)), steps or 10)
            gbs = 8 * B * N / ms / 1e6
            out.append({'kernel': f'fbs_resample_f32 {name}', 'chains': B, 'N': N, 'ms': ms, 'achieved': gbs, 'unit': 'GB/s',
                        'frac': gbs / peak, 'algorithmic_bytes_per_particle': 8})
        del w, keys, idx
    for d, B, N in ((10, 256, 16384), (100, 64, 16384)):
        K = 4
        jm, jc, y0 = gp_setup(d)
        ts = np.linspace(0., 1., K + 1)
        model = fbs_b200.AffineGaussianModel.from_linear_sde(sdes.StationaryConstLinearSDE(a=-0.5, b=1.), jm, jc, d, ts, T=1.)
        us = torch.randn(B, N, d, device=dev)
        us2 = torch.empty_like(us)
        lw = torch.full((B, N), -float(np.log(N)), device=dev)
        lw2 = torch.empty_like(lw)
        A = torch.empty(B, N, dtype=torch.int32, device=dev)
        v, vp, ustar = (torch.randn(B, d, device=dev) for _ in range(3))
        b0 = torch.zeros(B, dtype=torch.int32, device=dev)
        keys = fr.split(torch.from_numpy(fr.PRNGKey(2)).to(dev), B)
        ms = timeit(lambda: nat.call('fbs_csmc_step_affine_f32', stream(), model.struct(), 1, nat.RESAMPLE_KILLING, ptr(keys),
                                     ptr(us), ptr(lw), ptr(v), ptr(vp), ptr(ustar), ptr(b0), ptr(b0), B, N, ptr(A), ptr(us2),
                                     ptr(lw2)), steps or 5)
        gbs = (8 * d + 16) * B * N / ms / 1e6
        out.append({'kernel': 'fbs_csmc_step_affine_f32 (ancestors + transition/weight + normalise, 3 launches)', 'd': d, 'chains': B,
                    'N': N, 'ms': ms, 'achieved': gbs, 'unit': 'GB/s', 'frac': gbs / peak,
                    'algorithmic_bytes_per_particle_step': 8 * d + 16, 'particle_steps_per_s': B * N / ms * 1e3})
    return {'bound': 'hbm', 'peak': peak, 'note': 'in-kernel threefry2x32 (jax.random-compatible) bounds these kernels by instruction issue '
            'below the HBM roofline; see profiles/r1_hbm_kernels.md', 'kernels': out}


def secondary_sharded(world, rank, per_rank=512, steps=3, strong_total=1024):
    """configs[4]: CelebA-HQ-shaped (64x64x3, inpaint-32) random-init score U-Net, ONE chain whose particle set is sharded over
    the ranks (fbs_b200/sharded.py): per step one all-gather of the log-weights + the ancestor gather over NVLink peer memory.

    (1) bit-equality of the sharded sweep with the unsharded one, checked IN THIS RUN on every rank at a small particle
        count; (2) weak scaling, `per_rank` particles per GPU (4096 at 8 GPUs), random reference indices as gibbs_kernel
        draws them (gibbs.py:156) and 'gibbs-eb-ef' initialisation, the step size chosen so that conditional killing replaces
        roughly a third of the particles; (3) strong scaling: `strong_total` particles whatever the GPU count."""
    import torch
    import torch.distributed as dist
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    from fbs_b200.sharded import forward_pass_sharded
    from fbs_b200.nn import ScoreNetModel
    from fbs_b200 import random as fr
    shape = (64, 64, 3)
    base_model, _, rect, obs = _score_model(shape, steps)
    net, sde = base_model.unet, base_model.sde
    rng = np.random.default_rng(0)                                     # identical inputs on every rank
    us_star = rng.standard_normal((steps + 1, rect.size, 3)).astype(np.float32)
    row_bytes = rect.size * 3 * 4

    def problem(T_model, N):
        ts = np.linspace(0., T_model, steps + 1)
        model = ScoreNetModel(net, sde, ts, T_model, rect, obs)        # dt = T_model / steps, reverse time T_model - t
        sd = math.sqrt(T_model / steps) * float(sde.dispersion(T_model))
        vs = np.cumsum(sd * np.random.default_rng(1).standard_normal((steps + 1, obs.size, 3)), axis=0).astype(np.float32)
        bs = np.random.default_rng(2).integers(0, N - 1, size=steps + 1).astype(np.int32)   # gibbs.py:156: in [0, nparticles)
        return model, (fr.PRNGKey(5), us_star, bs, vs, model, csmc.NormalInit(model), R.killing, N - 1)

    def run(args, history=False):
        try:
            return forward_pass_sharded(*args, history=history)
        except RuntimeError as e:
            if 'peer-memory' not in str(e):
                raise
            os.environ['FBS_SHARD_EXCHANGE'] = 'nccl'                   # raised on every rank alike: all ranks take the same branch
            return forward_pass_sharded(*args, history=history)

    def replaced_fraction(r):
        A = r['ancestors']
        return float(np.mean([1.0 - torch.unique(A[k]).numel() / A.shape[1] for k in range(A.shape[0])]))

    # (1) sharded == unsharded, bit for bit, in this very run
    n_small = 8 * world
    model_s, args_s = problem(0.02, n_small)
    r = run(args_s, history=True)
    full = csmc.forward_pass_nn(args_s[0], args_s[1], args_s[2], args_s[3], model_s, args_s[5], R.killing.scheme, n_small - 1,
                                history=True)
    same = (torch.equal(r['As'], full['As'][0]) and torch.equal(r['log_wss'], full['log_wss'][0])
            and torch.equal(r['uss'].reshape(steps + 1, -1, rect.size * 3), full['uss'][0][:, r['lo']:r['hi']]))
    flag = torch.tensor([1 if same else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    bit_equal = bool(flag.item())
    moved_small = float(np.mean(r['moved']))

    # (2) pick the step size whose weight spread makes conditional killing replace about a third of the particles
    trials = []
    for T_model in (2.0, 0.2, 0.02, 0.005, 0.001):
        _, a = problem(T_model, 32 * world)
        trials.append((abs(replaced_fraction(run(a)) - 0.35), T_model))
    T_pick = min(trials)[1]

    def timed(N):
        _, a = problem(T_pick, N)
        run(a)                                                          # warm-up (CUDA graph of this batch size, IPC set-up)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rr = run(a)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mv = torch.tensor([float(np.mean(rr['moved']))], device='cuda', dtype=torch.float64)
        dist.all_reduce(mv, op=dist.ReduceOp.SUM)                       # rows crossing NVLink per step, all ranks
        ms = float(t.item())
        return {'n_particles': N, 'particles_per_gpu': N // world, 'value': N * steps / (ms * 1e-3), 'unit': 'particle-steps/s',
                'ms_per_csmc_step': ms / steps, 'moved_rows_per_step': float(mv.item()),
                'nvlink_bytes_per_step': float(mv.item()) * row_bytes,
                'fraction_of_particles_replaced_per_step': replaced_fraction(rr), 'exchange': rr.get('exchange')}

    weak = timed(per_rank * world)
    strong = weak if strong_total == per_rank * world else timed(strong_total)
    return {'workload': 'configs[4]: CelebA-HQ-shaped 64x64x3 inpaint-32, random-init score U-Net, one chain, particle set sharded '
                        f'over {world} GPUs; random reference indices, explicit-final initialisation, {steps} steps',
            'bitwise_equal_to_unsharded_sweep': bit_equal,
            'bitwise_check': f'{n_small} particles, {steps} steps, ancestors + log-weights + every particle row of the history, on every '
                             f'rank; {moved_small:.1f} rows per step crossed GPUs on rank 0 during the check',
            'weak': weak, 'strong': strong, 'value': weak['value'], 'unit': 'particle-steps/s',
            'moved_rows_per_step': weak['moved_rows_per_step'], 'step_size_T_over_K': T_pick / steps,
            'collectives_per_step': 'all_gather(4 N bytes) of the log-weights; parent rows read from the owners over NVLink inside '
                                    'the gather kernel (CUDA IPC peer memory; FBS_SHARD_EXCHANGE=nccl: batch_isend_irecv)',
            'dtype': 'bf16'}


METRIC = 'CSMC particle-steps/sec (Gibbs sweep)'
WORKLOAD_NAME = ('particle-Gibbs sweep of the toy Gaussian sampler scaled out (experiments/toy/gp_gibbs.py; d=100, K=200, N=100, '
                 'conditional killing, explicit backward), independent chains partitioned over GPUs')


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import fbs_b200
    from fbs_b200 import _native as nat, sdes, parallel
    from fbs_b200 import random as fr
    from fbs_b200.samplers import gibbs_kernel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'NONE'      # keep stdout to the ONE JSON line (NCCL prints its version there)
        dist.init_process_group('nccl', device_id=dev)

    d, K, N, C = D_TOY, K_STEPS, N_PART, args.chains
    jm, jc, y0 = gp_setup(d)
    ts = np.linspace(0., 1., K + 1)
    sde = sdes.StationaryConstLinearSDE(a=-0.5, b=1.)
    model = fbs_b200.AffineGaussianModel.from_linear_sde(sde, jm, jc, d, ts, T=1.)
    gk = dict(ts=ts, fwd_sampler=model.fwd_sampler, sde=sde, unpack=model.unpack, nparticles=N,
              transition_sampler=model.transition_sampler, transition_logpdf=model.transition_logpdf,
              likelihood_logpdf=model.likelihood_logpdf)                  # eb=True, ef=False: gibbs_kernel's defaults

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total_chains = world * C                      # weak scaling: C chains per GPU
    master = fr.PRNGKey(2024)
    run_keys = parallel.threefry_split_host(master, 1 << 14)
    y0_d = torch.from_numpy(y0).to(dev)

    def step_keys(i):
        return parallel.chain_keys(run_keys[i], total_chains, rank, world)

    # chain state resident in HBM: x0 [C, du], bs_star [C, K+1] (us_star is an output only: gibbs.py:91-92 ignores the input)
    state = (torch.zeros((C, d), device=dev), torch.zeros((C, K + 1), dtype=torch.int32, device=dev))

    def one_step(i, state):
        keys = torch.from_numpy(step_keys(i)).to(dev)
        x0, us_star, bs, changed = gibbs_kernel(keys, state[0], y0_d, None, state[1], **gk)
        return (x0, bs), (us_star, changed)

    clocks = ClockSampler(local)
    clocks.start()                     # before the warm-up: nvidia-smi takes a few hundred ms to deliver its first sample
    for i in range(args.warmup):
        state, aux = one_step(i, state)
    barrier()

    # ---- device-resident timed region -------------------------------------------------------
    KERNEL = 'fbs_csmc_forward_affine_f32'
    nat.TIMED[KERNEL] = []
    nat.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_clk0 = clocks.mark()
    ev0.record()
    for i in range(args.steps):
        state, aux = one_step(args.warmup + i, state)
    ev1.record()
    barrier()
    t_clk1 = clocks.mark()
    launches = nat.launch_count()
    ms_total = ev0.elapsed_time(ev1)
    kern_ms = [a.elapsed_time(b) for a, b in nat.TIMED.pop(KERNEL)]
    clk = clocks.stop(t_clk0, t_clk1)
    if not bool(torch.isfinite(state[0]).all()):
        raise RuntimeError('the Gibbs sweep produced non-finite samples')
    moved_frac = float(aux[1].float().mean().item())

    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    psteps_per_step = total_chains * N * K
    value = psteps_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end through the host-buffer API ------------------------------------------------
    # numpy / pinned host tensors in, numpy out: gibbs_kernel copies keys, x0, bs_star H2D and x0, us_star, bs_star, changed
    # D2H every step (what the reference driver does per sweep, gp_gibbs.py:185-190), chains chunked over CUDA streams
    h_state = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in state]
    for h, x in zip(h_state, state):
        h.copy_(x)
    h_y0 = torch.from_numpy(y0).pin_memory()
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 5))
    e2e_warm = 0 if args.no_e2e else max(3, min(args.warmup, 5))   # the first host-buffer calls allocate the page-locked staging blocks
    h_keys = [torch.from_numpy(step_keys(10_000 + i)).pin_memory() for i in range(e2e_steps + e2e_warm)]

    def e2e_step(i, hs):
        return gibbs_kernel(h_keys[i], hs[0], h_y0, None, hs[1], **gk)                   # numpy out (D2H inside)

    h2d = sum(x.numel() * x.element_size() for x in h_state) + C * 2 * 4 + h_y0.numel() * 4
    hs = h_state
    o = ()
    for i in range(e2e_warm):
        o = e2e_step(e2e_steps + i, hs)
        hs = [torch.from_numpy(np.ascontiguousarray(o[0])), torch.from_numpy(np.ascontiguousarray(o[2]))]
    d2h = sum(np.asarray(x).nbytes for x in o)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        o = e2e_step(i, hs)
        hs = [torch.from_numpy(np.ascontiguousarray(o[0])), torch.from_numpy(np.ascontiguousarray(o[2]))]
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, e2e_steps) if e2e_steps else float('nan')
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = psteps_per_step / (e2e_ms * 1e-3)

    secondary = {}
    if not args.no_secondary:
        # a secondary workload must never take the headline line (or another rank) down
        if rank == 0:
            for name, fn in (('pmcmc_step', secondary_pmcmc), ('gaussian_sb_gibbs', secondary_sb),
                             ('hbm_kernels', secondary_hbm_kernels), ('mnist_score_net', secondary_mnist)):
                try:
                    secondary[name] = fn()
                except Exception as e:
                    secondary[name] = {'error': repr(e)[:300]}
        if world > 1:
            try:
                dist.barrier()
                sh = secondary_sharded(world, rank)
            except Exception as e:
                sh = {'error': repr(e)[:300]}
            if rank == 0:
                secondary['celeba_particle_sharded'] = sh

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = float(np.mean(kern_ms))
        alg_bytes = C * N * K * ALG_BYTES_PER_PARTICLE_STEP            # per launch (this rank's chains)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        rb = rng_bound(C * N * K * d / (k_ms * 1e-3))                  # the kernel's own normals per second
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(8)
        traffic_bytes = None
        traffic_src = None
        if traffic and traffic.get('dram_bytes_per_launch'):
            if int(traffic.get('chains', -1)) == C:
                traffic_bytes = int(traffic['dram_bytes_per_launch'])
                traffic_src = f"{traffic.get('source')}; captured at this launch's {C} chains"
            else:
                traffic_bytes = int(traffic['dram_bytes_per_launch'] * C / traffic['chains'])
                traffic_src = f"{traffic.get('source')}; captured at {traffic.get('chains')} chains, scaled to this launch's {C}"
        line = {
            'metric': METRIC, 'value': value,
            'unit': 'particle-steps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_NAME, 'chains_per_gpu': C, 'total_chains': total_chains, 'd': d, 'K': K,
                       'N': N, 'resampling': 'conditional killing', 'parallelism': f'chains x{world} (no collective on the data path)',
                       'l2': 'inputs larger than L2: per-step chain state (forward paths us, vs, us_star_next) = '
                             f'{3 * C * (K + 1) * d * 4 / 1e6:.0f} MB >> 126 MB',
                       'fraction_of_reference_indices_changed_last_step': moved_frac},
            'e2e': None if args.no_e2e else {'value': e2e_value, 'unit': 'particle-steps/s', 'h2d_bytes_per_step': int(h2d),
                    'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_ms, 'steps': e2e_steps, 'warmup': e2e_warm,
                    'how': 'numpy in / numpy out through fbs_b200.samplers.gibbs_kernel; chains chunked over 8 CUDA streams so that '
                           'H2D, kernels and D2H (page-locked staging) overlap'},
            'gpu_launches': int(launches),
            'clocks': clk,
            'roofline': {'bound': 'hbm',
                         'kernel': 'sweep_v3_kernel (fbs_csmc_forward_affine_f32; tcgen05 split-TF32 GEMM + in-kernel threefry + '
                                   'conditional killing resampling)',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic_bytes, 'traffic_source': traffic_src,
                         'peak_source': peak_src,
                         'binding': {'bound': 'issue / in-kernel jax.random generator', **(rb or {})},
                         'kernel_ms': k_ms, 'kernel_share_of_step': k_ms / ms_per_step,
                         'algorithmic_bytes_per_particle_step': ALG_BYTES_PER_PARTICLE_STEP,
                         'note': 'bound/achieved/peak/frac: the nominal HBM roofline of SURVEY 8(d) (8 du + 16 B per particle-step). '
                                 'The persistent kernel keeps the particles in shared memory (traffic << algorithmic bytes), so the '
                                 'resource that binds is instruction issue, dominated by the jax.random-compatible threefry + '
                                 'erf_inv: see "binding" (normals/s against the stand-alone generator, profiles/rng_peaks.json)'},
            'cpu_baseline': cpu,
            'secondary': secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        from fbs_b200.sharded import close_peer_buffers
        close_peer_buffers()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='ours', choices=['ours', 'reference'])
    ap.add_argument('--chains', type=int, default=4144, help='chains per GPU (weak scaling); 4144 = 14 full waves of 148 SMs x 2 chains per CTA')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer end-to-end leg (profiling runs only: the launch list '
                                                          'then holds the device-resident steps alone)')
    ap.add_argument('--no-secondary', action='store_true', help='skip the secondary (score network / particle-sharded) workloads')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
