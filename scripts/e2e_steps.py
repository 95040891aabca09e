"""Per-step wall time of the host-buffer pmcmc_kernel call (bench.py's e2e leg), step by step."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, fbs_b200
from fbs_b200 import sdes, parallel, random as fr
from fbs_b200.samplers import pmcmc_kernel, stratified
d, K, N, C = bench.D_TOY, bench.K_STEPS, bench.N_PART, 4096
jm, jc, y0 = bench.gp_setup(d)
ts = np.linspace(0., 1., K + 1)
sde = sdes.StationaryConstLinearSDE(a=-0.5, b=1.)
model = fbs_b200.AffineGaussianModel.from_linear_sde(sde, jm, jc, d, ts, T=1.)
kw = dict(ts=ts, fwd_ys_sampler=model.fwd_ys_sampler, sde=sde, ref_sampler=model.ref_sampler,
          transition_sampler=model.transition_sampler, likelihood_logpdf=model.likelihood_logpdf, resampling=stratified,
          nparticles=N, delta=bench.DELTA)
keys = parallel.chain_keys(fr.PRNGKey(1), C, 0, 1)
ys = model.fwd_ys_sampler(torch.from_numpy(keys).cuda(), torch.from_numpy(y0).cuda()).cpu().numpy()
state = [np.zeros((C, d), np.float32), np.zeros((C,), np.float32), ys]
for i in range(8):
    k = parallel.chain_keys(fr.PRNGKey(100 + i), C, 0, 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    o = pmcmc_kernel(k, state[0], state[1], state[2], y0, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    state = list(o[:3])
    print(f'step {i}: {1e3 * (t1 - t0):7.1f} ms  pinned_in={torch.from_numpy(state[2]).is_pinned()}', flush=True)
