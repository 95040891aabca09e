"""Cost of one dependent kernel node in a CUDA graph on this GPU: a chain of N tiny launches replayed from a graph."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbs_b200.nn import ops
x = torch.zeros(1, 2, 2, 64, device='cuda', dtype=torch.bfloat16)
y = torch.zeros(1, 2, 2, 256, device='cuda', dtype=torch.bfloat16)
for n in (100, 400):
    ops.space_to_depth(x, y); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            ops.space_to_depth(x, y)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    print(f'{n} dependent tiny kernel nodes: {a.elapsed_time(b) * 1e3 / n:.2f} us per node')
