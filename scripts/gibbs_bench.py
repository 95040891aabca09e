"""Gibbs-sweep throughput (BASELINE.json configs[0], experiments/toy/gp_gibbs.py shapes): particle-steps/s of gibbs_kernel,
device resident.  usage: python scripts/gibbs_bench.py d N chains [K]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, fbs_b200
from fbs_b200 import sdes, parallel, random as fr
from fbs_b200.samplers import gibbs_kernel
d, N, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
K = int(sys.argv[4]) if len(sys.argv) > 4 else 200
jm, jc, y0 = bench.gp_setup(d)
ts = np.linspace(0., 1., K + 1)
sde = sdes.StationaryConstLinearSDE(a=-0.5, b=1.)
model = fbs_b200.AffineGaussianModel.from_linear_sde(sde, jm, jc, d, ts, T=1.)
dev = torch.device('cuda')
y0_d = torch.from_numpy(y0).to(dev)
x0 = torch.zeros((C, d), device=dev)
bs = torch.zeros((C, K + 1), dtype=torch.int32, device=dev)
def step(i, x0, bs):
    keys = torch.from_numpy(parallel.chain_keys(fr.PRNGKey(50 + i), C, 0, 1)).to(dev)
    x0, us_star, bs, _ = gibbs_kernel(keys, x0, y0_d, None, bs, ts, model.fwd_sampler, sde, model.unpack, N,
                                      model.transition_sampler, model.transition_logpdf, model.likelihood_logpdf)
    return x0, bs
for i in range(3):
    x0, bs = step(i, x0, bs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for i in range(n):
    x0, bs = step(3 + i, x0, bs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({'d': d, 'N': N, 'chains': C, 'K': K, 'ms_per_sweep': ms, 'particle_steps_per_s': C * N * K / ms * 1e3}))
