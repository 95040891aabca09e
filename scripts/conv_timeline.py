"""Phase timeline of conv_gemm_kernel's CTA 0 (fbs_debug_conv_timeline): clock64 stamps of the TMA producer, the MMA issuer and
one epilogue warp, per tile.  usage: python scripts/conv_timeline.py [B] [label filter]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbs_b200 import _native as nat
from fbs_b200.nn import ops
from fbs_b200._tensor import ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 101
only = sys.argv[2] if len(sys.argv) > 2 else ''
shapes = [
    ('res 28 64->64 3x3', 28, 28, 64, 0, 64, 3, False),
    ('res 28 128->64 3x3 (cat)', 28, 28, 64, 64, 64, 3, False),
    ('res 28 128->64 1x1 (cat)', 28, 28, 64, 64, 64, 1, False),
    ('qkv 28 64->384 1x1', 28, 28, 64, 0, 384, 1, False),
    ('res 14 128->128 3x3', 14, 14, 128, 0, 128, 3, False),
    ('res 7 256->256 3x3', 7, 7, 256, 0, 256, 3, False),
    ('up 14 128->512 3x3 shuffle', 14, 14, 128, 0, 512, 3, True),
]
buf = torch.zeros(256, dtype=torch.int64, device='cuda')
for label, H, W, c0, c1, cout, k, shuffle in shapes:
    if only not in label:
        continue
    in0 = torch.randn(B, H, W, c0, device='cuda').to(torch.bfloat16)
    in1 = torch.randn(B, H, W, c1, device='cuda').to(torch.bfloat16) if c1 else None
    w = (torch.randn(cout, k * k * (c0 + c1), device='cuda') * 0.05).to(torch.bfloat16)
    bias = torch.zeros(cout, device='cuda')
    of = None if shuffle else torch.empty(B, H, W, cout, device='cuda')
    ob = torch.empty((B, 2 * H, 2 * W, cout // 4) if shuffle else (B, H, W, cout), device='cuda', dtype=torch.bfloat16)
    kw = dict(in1=in1, bias=bias, pixel_shuffle=shuffle, out_bf16=ob, out_f32=of)
    for _ in range(3):
        ops.conv(in0, w, cout, k, k, -1 if k == 3 else 0, H, W, **kw)
    buf.zero_()
    nat.call('fbs_debug_conv_timeline', ptr(buf))
    ops.conv(in0, w, cout, k, k, -1 if k == 3 else 0, H, W, **kw)
    torch.cuda.synchronize()
    nat.call('fbs_debug_conv_timeline', None)
    t = buf.cpu().numpy()
    t0 = t[0]
    r = lambda i: int(t[i] - t0) if t[i] else -1
    print(f'== {label} (B={B}); cycles since kernel entry of CTA 0')
    print(f'   prologue done {r(1)}   weights issued {r(2)}   weights landed / MMA starts {r(3)}   exit {r(4)}')
    print('   tile: producer first load | MMA has accumulator | first operands landed | MMAs issued | epilogue start | epilogue end | first chunk in registers | first chunk stored')
    for i in range(8):
        o = 8 + 8 * i
        if not t[o + 1]:
            break
        print(f'   {i:3d}: {r(o):8d} {r(o + 1):8d} {r(o + 2):8d} {r(o + 3):8d} {r(o + 4):8d} {r(o + 5):8d} {r(o + 6):8d} {r(o + 7):8d}')
