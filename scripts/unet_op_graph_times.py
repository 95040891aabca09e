"""Device time of every launch of one score-network evaluation, each replayed on its own from a CUDA graph (no host launch
cost, warm L2): where the evaluation's time goes, launch by launch.
usage: python scripts/unet_op_graph_times.py [28x28x1] [B] [reps]"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbs_b200 import _native as nat
from fbs_b200.nn import ScoreUNet
from fbs_b200._tensor import stream
from oracle import unet as ou   # random checkpoint only
shape = tuple(int(a) for a in sys.argv[1].split('x')) if len(sys.argv) > 1 else (28, 28, 1)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 101
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
net = ScoreUNet(ou.init_unet_params(0, shape[2]), shape, dt=2. / 200)
x = torch.randn(B, *shape, device='cuda')
for _ in range(2):
    net(x, 0.5, use_graph=False)
torch.cuda.synchronize()
calls = []
orig = nat._call


def rec(handle, name, args):
    calls.append((name, args))
    return orig(handle, name, args)


nat._call = rec
net(x, 0.5, use_graph=False)
nat._call = orig
torch.cuda.synchronize()


def label(name, args):
    if name == 'fbs_nn_conv_bf16':
        a = args[1]._obj
        return f'conv {a.H}x{a.W} {a.C0}+{a.C1}->{a.Cout} {a.kh}x{a.kw}' + (' shuffle' if a.pixel_shuffle else '') + \
            (' +f32' if a.out_f32 else '') + (' +bf16' if a.out_bf16 else '')
    return name.replace('fbs_nn_', '') + ' ' + ' '.join(str(v) for v in args[2:6] if isinstance(v, int) and v < 10 ** 6)


handle = nat.lib()
rows = []
for name, args in calls:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            orig(handle, name, (stream(),) + tuple(args[1:]))   # the capture stream, not the recorded one
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    rows.append((label(name, args), name, a.elapsed_time(b) * 1e3 / reps))
tot = sum(r[2] for r in rows)
print(f'{len(rows)} launches, sum {tot:.0f} us (B={B}, {shape})')
by = collections.OrderedDict()
for lab, name, us in rows:
    c = by.setdefault(lab, [0, 0.])
    c[0] += 1; c[1] += us
for lab, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f'{us / tot:6.3f} {us:8.1f} us  n={n:2d}  avg {us / n:6.1f} us  {lab}')
