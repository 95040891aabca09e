import sys, time, json
import numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from fbs_b200.nn import ScoreUNet
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import unet as ou   # params only (random checkpoint)
shape = tuple(int(a) for a in sys.argv[1].split('x')) if len(sys.argv) > 1 else (28, 28, 1)
B = int(sys.argv[2]) if len(sys.argv) > 2 else 101
params = ou.init_unet_params(0, shape[2])
net = ScoreUNet(params, shape, dt=2./200)
x = torch.randn(B, *shape, device='cuda')
for _ in range(3): net(x, 0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for i in range(n): net(x, 0.5 + 0.01 * i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flops = {(28,28,1): 2.788e9, (64,64,3): 14.635e9, (32,32,3): 14.635e9/4}.get(shape, 0) * B
print(json.dumps({'shape': shape, 'B': B, 'ms_per_eval': ms, 'evals_per_s': B / ms * 1e3, 'tflops': flops / ms / 1e9}))
