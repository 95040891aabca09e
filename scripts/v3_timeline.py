"""Phase timeline of the tcgen05 sweep kernel (sweep_v3.cu) through fbs_debug_v3_timeline: clock64() stamps of CTA 0 for steps
64..67 of its first chain pair, one warp per role.  usage: python scripts/v3_timeline.py [chains] [mode: gibbs|pmcmc]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, fbs_b200
from fbs_b200 import sdes, parallel, _native as nat, random as fr
from fbs_b200._tensor import ptr
from fbs_b200.samplers import gibbs_kernel

C = int(sys.argv[1]) if len(sys.argv) > 1 else 4144
d, K, N = 100, 200, 100
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
    x0, _, bs, _ = gibbs_kernel(keys, x0, y0_d, None, bs, ts, model.fwd_sampler, sde, model.unpack, N, model.transition_sampler,
                                model.transition_logpdf, model.likelihood_logpdf)
    return x0, bs


for i in range(2):
    x0, bs = step(i, x0, bs)
torch.cuda.synchronize()
buf = torch.zeros(7 * 4 * 16, dtype=torch.int64, device=dev)
nat.call('fbs_debug_v3_timeline', ptr(buf))
x0, bs = step(2, x0, bs)
torch.cuda.synchronize()
nat.call('fbs_debug_v3_timeline', None)
t = buf.cpu().numpy().reshape(7, 4, 16)
t0 = t[0, 0, 0]
roles = ['E g0', 'X g0', 'R g0', 'E g1', 'X g1', 'R g1', 'MMA']
noise_names = ['start', 'noise1', 'accum wait', 'bar_E', 'v epilogue', 'noise2', 'u epilogue', 'noise3', 'bar_all', 'gather',
               'bar_noise', 'hi/lo stores', 'bar_noise2']
r_names = {0: 'start', 1: 'keys+uniforms', 2: 'bar_ER (weights ready)', 7: 'resampling', 8: 'bar_all'}
print(f'cycles relative to the start of step 64 of group 0 (E warp); step period = {(t[0, 3, 0] - t[0, 0, 0]) / 3:.0f} cycles')
for r, name in enumerate(roles):
    for s in range(2):
        row = t[r, s]
        if r == 6:
            print(f'{name} step {64 + s}: g0 ready {row[0] - t0}, g0 issued {row[1] - t0}, g1 ready {row[2] - t0}, g1 issued {row[3] - t0}')
            continue
        names = r_names if r % 3 == 2 else dict(enumerate(noise_names))
        prev = None
        parts = []
        for i in sorted(names):
            if row[i] == 0:
                continue
            parts.append(f'{names[i]} @{row[i] - t0}' + (f' (+{row[i] - prev})' if prev is not None else ''))
            prev = row[i]
        print(f'{name} step {64 + s}: ' + ', '.join(parts))
