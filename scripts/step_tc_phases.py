"""Per-phase cycle counts of the tcgen05 per-timestep kernel (CTA 0), through fbs_debug_step_tc_timers.
usage: python scripts/step_tc_phases.py [d] [chains] [N]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, fbs_b200
from fbs_b200 import sdes, _native as nat, random as fr
from fbs_b200._tensor import ptr, stream

d = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
dev = torch.device('cuda')
jm, jc, y0 = bench.gp_setup(d)
ts = np.linspace(0., 1., 5)
model = fbs_b200.AffineGaussianModel.from_linear_sde(sdes.StationaryConstLinearSDE(a=-0.5, b=1.), jm, jc, d, ts, T=1.)
us = torch.randn(B, N, d, device=dev); us2 = torch.empty_like(us)
lw = torch.full((B, N), -float(np.log(N)), device=dev); lw2 = torch.empty_like(lw)
A = torch.empty(B, N, dtype=torch.int32, device=dev)
v, vp, ustar = (torch.randn(B, d, device=dev) for _ in range(3))
b0 = torch.zeros(B, dtype=torch.int32, device=dev)
keys = fr.split(torch.from_numpy(fr.PRNGKey(2)).to(dev), B)
call = lambda: nat.call('fbs_csmc_step_affine_f32', stream(), model.struct(), 1, nat.RESAMPLE_KILLING, ptr(keys), ptr(us), ptr(lw),
                        ptr(v), ptr(vp), ptr(ustar), ptr(b0), ptr(b0), B, N, ptr(A), ptr(us2), ptr(lw2))
call(); torch.cuda.synchronize()
buf = torch.zeros(8, dtype=torch.int64, device=dev)
nat.call('fbs_debug_step_tc_timers', ptr(buf))
call(); torch.cuda.synchronize()
nat.call('fbs_debug_step_tc_timers', None)
names = ['gather+split+constants', 'barrier #1', 'noise', 'worker barrier', 'accumulator wait', 'epilogue', 'barrier #2', 'stores']
c = buf.cpu().numpy()
tiles = -(-B * (-(-(N // 2) // 64)) // 148)
print(f'd={d} chains={B} N={N}: {tiles} tiles per CTA, {c.sum() / tiles:.0f} cycles per tile')
for n, x in zip(names, c):
    print(f'  {n:26s} {x / tiles:9.0f} cycles/tile  {100 * x / c.sum():5.1f} %')
