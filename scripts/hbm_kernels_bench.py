"""Stand-alone resampling and per-timestep transition kernels against the HBM roofline (8 B per particle for
resampling, 8 du + 16 B per particle-step for the transition).  usage: python scripts/hbm_kernels_bench.py"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, fbs_b200
from fbs_b200 import sdes, _native as nat, random as fr
from fbs_b200._tensor import ptr, stream
from fbs_b200.samplers import resampling as R, csmc

PEAK = bench.measured_peak()[0]
dev = torch.device('cuda')


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = []
for B, N in ((65536, 100), (16384, 1024), (1024, 16384)):
    w = torch.rand(B, N, device=dev)
    w /= w.sum(dim=1, keepdim=True)
    keys = fr.split(torch.from_numpy(fr.PRNGKey(1)).to(dev), B)
    idx = torch.empty(B, N, dtype=torch.int32, device=dev)
    for name, scheme in (('stratified', nat.RESAMPLE_STRATIFIED), ('killing', nat.RESAMPLE_KILLING)):
        ms = timeit(lambda: nat.call('fbs_resample_f32', stream(), scheme, ptr(keys), ptr(w), B, N, ptr(idx)))
        gbs = 8 * B * N / ms / 1e6
        out.append({'kernel': f'fbs_resample_f32 {name}', 'B': B, 'N': N, 'ms': ms, 'GB/s': gbs, 'frac_hbm': gbs / PEAK})
for d, B, N in ((10, 256, 16384), (100, 64, 16384)):
    K = 4
    jm, jc, y0 = bench.gp_setup(d)
    ts = np.linspace(0., 1., K + 1)
    model = fbs_b200.AffineGaussianModel.from_linear_sde(sdes.StationaryConstLinearSDE(a=-0.5, b=1.), jm, jc, d, ts, T=1.)
    us = torch.randn(B, N, d, device=dev); us2 = torch.empty_like(us)
    lw = torch.full((B, N), -np.log(N), device=dev); lw2 = torch.empty_like(lw)
    A = torch.empty(B, N, dtype=torch.int32, device=dev)
    v = torch.randn(B, d, device=dev); vp = torch.randn(B, d, device=dev); ustar = torch.randn(B, d, device=dev)
    b0 = torch.zeros(B, dtype=torch.int32, device=dev)
    keys = fr.split(torch.from_numpy(fr.PRNGKey(2)).to(dev), B)
    ms = timeit(lambda: nat.call('fbs_csmc_step_affine_f32', stream(), model.struct(), 1, nat.RESAMPLE_KILLING, ptr(keys), ptr(us),
                                 ptr(lw), ptr(v), ptr(vp), ptr(ustar), ptr(b0), ptr(b0), B, N, ptr(A), ptr(us2), ptr(lw2)), n=5)
    gbs = (8 * d + 16) * B * N / ms / 1e6
    out.append({'kernel': 'fbs_csmc_step_affine_f32 (ancestors + transition/weight + normalise)', 'd': d, 'B': B, 'N': N, 'ms': ms,
                'GB/s': gbs, 'frac_hbm': gbs / PEAK, 'particle_steps_per_s': B * N / ms * 1e3})
for o in out:
    print(json.dumps(o))
