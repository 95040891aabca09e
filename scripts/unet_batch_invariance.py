"""Is the score network batch invariant?  unet(x)[:n] against unet(x[:n]), eager and graph.  usage: python scripts/unet_batch_invariance.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbs_b200.nn import ScoreUNet
from oracle import unet as ou
net = ScoreUNet(ou.init_unet_params(0, 1), (28, 28, 1), dt=2. / 200)
for n, B in ((6, 18), (7, 21), (101, 808)):
    x = torch.randn(B, 28, 28, 1, device='cuda')
    for graph in (False, True):
        big = net(x, 0.5, use_graph=graph).clone()
        small = net(x[:n].contiguous(), 0.5, use_graph=graph).clone()
        mid = net(x[n:2 * n].contiguous(), 0.5, use_graph=graph).clone()
        print(n, B, 'graph' if graph else 'eager', 'first chunk equal:', torch.equal(big[:n], small), float((big[:n] - small).abs().max()),
              'second chunk equal:', torch.equal(big[n:2 * n], mid), float((big[n:2 * n] - mid).abs().max()))
