import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from fbs_b200.nn import ScoreUNet
from oracle import unet as ou
for shape in ((28, 28, 1), (32, 32, 3)):
    H, W, C = shape
    params = ou.init_unet_params(5, C)
    rng = np.random.default_rng(6)
    x = rng.standard_normal((5, H, W, C)).astype(np.float32)
    for flag in (False, True):
        net = ScoreUNet(params, shape, dt=2. / 200)
        net.gn_input_bf16 = flag
        for t in (0.02, 1.3):
            want = ou.unet_forward(params, x, t, 2. / 200)
            got = net(torch.from_numpy(x).cuda(), t).cpu().numpy()
            scale = float(np.abs(want).max()); err = np.abs(got - want)
            print(shape, 'gn_input_bf16', flag, 't', t, 'max', err.max() / scale, 'mean', err.mean() / scale)
import time
net = ScoreUNet(ou.init_unet_params(0, 1), (28, 28, 1), dt=2. / 200)
for flag in (False, True):
    net.gn_input_bf16 = flag; net._graphs.clear(); net._gn_slots.clear()
    x = torch.randn(101, 28, 28, 1, device='cuda')
    for _ in range(5): net(x, 0.5)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): net(x, 0.5)
    b.record(); torch.cuda.synchronize()
    print('gn_input_bf16', flag, a.elapsed_time(b) / 50, 'ms')
