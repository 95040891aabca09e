"""Per-op device time of one eager (no CUDA graph) score-network evaluation, warm caches, CUDA events around every
C-ABI call.  usage: python scripts/unet_op_times.py 28x28x1 101"""
import os, sys, collections
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbs_b200 import _native as nat
from fbs_b200.nn import ScoreUNet
from oracle import unet as ou   # random checkpoint only
shape = tuple(int(a) for a in sys.argv[1].split('x')) if len(sys.argv) > 1 else (28, 28, 1)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 101
net = ScoreUNet(ou.init_unet_params(0, shape[2]), shape, dt=2. / 200)
x = torch.randn(B, *shape, device='cuda')
for _ in range(3):
    net(x, 0.5, use_graph=False)
names = [n for n in nat.SIGNATURES if n.startswith('fbs_nn_')]
tot = collections.Counter(); cnt = collections.Counter()
reps = 5
for _ in range(reps):
    for n in names:
        nat.TIMED[n] = []
    net(x, 0.5, use_graph=False)
    torch.cuda.synchronize()
    for n in names:
        for a, b in nat.TIMED.pop(n):
            tot[n] += a.elapsed_time(b) * 1e3; cnt[n] += 1
s = sum(tot.values()) / reps
print(f'sum of op times {s:.0f} us per evaluation (B={B}, {shape})')
for n, t in tot.most_common():
    print(f'{t / reps / s:6.3f} {t / reps:8.1f} us  n={cnt[n] // reps:3d}  avg {t / cnt[n]:6.1f} us  {n}')
