"""Per-shape device time of `fbs_nn_conv_bf16` at the score network's layer shapes (MNIST 28x28 U-Net, dim 64, mults 1-2-4),
back-to-back launches replayed from one CUDA graph, CUDA events around the replay.  usage: python scripts/conv_bench.py [B] [reps] [label filter]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbs_b200.nn import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 101
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
only = sys.argv[3] if len(sys.argv) > 3 else ''     # substring filter on the label
eager = os.environ.get('CONV_BENCH_EAGER') == '1'   # plain launches (for ncu)
dev = 'cuda'
# (label, H, W, Cin0, Cin1, Cout, k, f32 out, bf16 out, count per evaluation)
shapes = [
    ('res 28 64->64 3x3', 28, 28, 64, 0, 64, 3, True, False, 8),
    ('res 28 128->64 3x3 (cat)', 28, 28, 64, 64, 64, 3, True, False, 3),
    ('res 28 128->64 1x1 (cat)', 28, 28, 64, 64, 64, 1, True, False, 3),
    ('qkv 28 64->384 1x1', 28, 28, 64, 0, 384, 1, False, True, 2),
    ('out 28 128->64 1x1', 28, 28, 128, 0, 64, 1, True, False, 2),
    ('res 14 64->64 3x3', 14, 14, 64, 0, 64, 3, True, False, 4),
    ('res 14 128->128 3x3', 14, 14, 128, 0, 128, 3, True, False, 4),
    ('res 14 192->128 3x3 (cat)', 14, 14, 128, 64, 128, 3, True, False, 2),
    ('res 7 128->128 3x3', 7, 7, 128, 0, 128, 3, True, False, 4),
    ('res 7 256->256 3x3', 7, 7, 256, 0, 256, 3, True, False, 9),
    ('res 7 384->256 3x3 (cat)', 7, 7, 256, 128, 256, 3, True, False, 2),
    ('up 7 256->512 3x3 shuffle', 7, 7, 256, 0, 512, 3, False, True, 1),
    ('up 14 128->256 3x3 shuffle', 14, 14, 128, 0, 256, 3, False, True, 1),
]
print(f'B={B}, {reps} back-to-back launches per shape')
tot = 0.
for label, H, W, c0, c1, cout, k, f32o, bfo, cnt in shapes:
    if only not in label:
        continue
    in0 = torch.randn(B, H, W, c0, device=dev).to(torch.bfloat16)
    in1 = torch.randn(B, H, W, c1, device=dev).to(torch.bfloat16) if c1 else None
    w = (torch.randn(cout, k * k * (c0 + c1), device=dev) * 0.05).to(torch.bfloat16)
    bias = torch.zeros(cout, device=dev)
    shuffle = 'shuffle' in label
    of = torch.empty(B, H, W, cout, device=dev) if f32o else None
    ob = torch.empty((B, 2 * H, 2 * W, cout // 4) if shuffle else (B, H, W, cout), device=dev, dtype=torch.bfloat16) if (bfo or True) else None
    res = {}
    for variant in ('f32+bf16', 'bf16 only'):
        kw = dict(in1=in1, bias=bias, pixel_shuffle=shuffle, out_bf16=ob)
        if variant == 'f32+bf16' and f32o:
            kw['out_f32'] = of
        for _ in range(3):
            ops.conv(in0, w, cout, k, k, -1 if k == 3 else 0, H, W, **kw)
        torch.cuda.synchronize()
        if eager:
            res[variant] = float('nan')
            continue
        g = torch.cuda.CUDAGraph()          # graph replay: no host launch cost in the figure
        with torch.cuda.graph(g):
            for _ in range(reps):
                ops.conv(in0, w, cout, k, k, -1 if k == 3 else 0, H, W, **kw)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record(); torch.cuda.synchronize()
        res[variant] = a.elapsed_time(b) * 1e3 / reps
    fl = 2. * B * H * W * cout * k * k * (c0 + c1)
    us = res['f32+bf16']
    tot += us * cnt
    print(f'{label:32s} {us:7.1f} us  ({res["bf16 only"]:6.1f} bf16-only)  {fl / us * 1e-6:7.1f} TFLOP/s  x{cnt}')
print(f'weighted sum {tot:.0f} us')
