"""TEST INFRASTRUCTURE -- float32 CPU restatement of the reference score network (fbs/nn/unet.py, fbs/nn/base.py:44-77,
fbs/nn/utils.py:53-57) with torch CPU ops standing in for flax.linen (third-party, absent from /root/reference:
flax==0.8.2 -- Conv: NHWC, HWIO kernels, symmetric integer padding, bias; GroupNorm eps 1e-6; LayerNorm eps given,
scale only when use_bias=False; gelu = tanh approximation; swish = x * sigmoid(x)).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  Parity unpinned: flax cannot be
imported here and the reference holds no golden vectors for the network; the only pinned piece is PixelShuffle's
channel order (tests/test_nns.py:7-16 compares it with torch.nn.PixelShuffle).

Parameters are a flat dict name -> float32 numpy array; `init_unet_params` draws a random "checkpoint" (there is no
network access for the real ones): conv / dense kernels N(0, 1 / fan_in) as flax's lecun_normal, every bias and norm
parameter perturbed so that no code path is exercised with an identity.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F


def sinusoidal_embedding(t, out_dim=64, max_period=10_000):
    """base.py:44-77."""
    half = out_dim // 2
    fs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / (half - 1))
    embs = t * fs
    return torch.cat([torch.sin(embs), torch.cos(embs)], dim=-1)


def pixel_shuffle(x, scale=2):
    """utils.py:53-57: 'b h w (h2 w2 c) -> b (h h2) (w w2) c'."""
    B, H, W, C4 = x.shape
    C = C4 // (scale * scale)
    x = x.reshape(B, H, W, scale, scale, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H * scale, W * scale, C)


class _P:
    def __init__(self, params):
        self.p = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()}

    def __getitem__(self, k):
        return self.p[k]


def _conv(P, name, x, stride=1, padding=1, bias=True, kernel=None):
    """flax nn.Conv on NHWC input with an HWIO kernel."""
    w = P[name + '.kernel'] if kernel is None else kernel          # [kh, kw, Cin, Cout]
    b = P[name + '.bias'] if bias else None
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, stride=stride, padding=padding)
    return y.permute(0, 2, 3, 1)


def standardize_kernel(w, eps=1e-5):
    """unet.py:110-118: per output filter over (kh, kw, Cin)."""
    mean = w.mean(dim=(0, 1, 2), keepdim=True)
    var = w.var(dim=(0, 1, 2), unbiased=False, keepdim=True)
    return (w - mean) / torch.sqrt(var + eps)


def _ws_conv(P, name, x):
    return _conv(P, name, x, kernel=standardize_kernel(P[name + '.kernel']))


def _group_norm(P, name, x, groups=8, eps=1e-6):
    B, H, W, C = x.shape
    xg = x.reshape(B, H * W, groups, C // groups)
    mean = xg.mean(dim=(1, 3), keepdim=True)
    var = xg.var(dim=(1, 3), unbiased=False, keepdim=True)
    y = ((xg - mean) / torch.sqrt(var + eps)).reshape(B, H, W, C)
    return y * P[name + '.scale'] + P[name + '.bias']


def _layer_norm(P, name, x, eps=1e-5):
    mean = x.mean(dim=-1, keepdim=True)
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * P[name + '.scale']


def _dense(P, name, x):
    return x @ P[name + '.kernel'] + P[name + '.bias']


def _swish(x):
    return x * torch.sigmoid(x)


def _resnet_block(P, name, x, time_emb, dim, groups=8):
    """unet.py:127-172."""
    C = x.shape[-1]
    h = _ws_conv(P, name + '.conv_0', x)
    h = _group_norm(P, name + '.norm_0', h, groups)
    te = _dense(P, name + '.time_mlp.dense_0', _swish(time_emb))[:, None, None, :]
    scale, shift = te[..., :dim], te[..., dim:]
    h = h * (1 + scale) + shift
    h = _swish(h)
    h = _ws_conv(P, name + '.conv_1', h)
    h = _swish(_group_norm(P, name + '.norm_1', h, groups))
    if C != dim:
        x = _conv(P, name + '.res_conv_0', x, padding=0)
    return x + h


def _split_heads(qkv, heads):
    B, H, W, D3 = qkv.shape
    dim = D3 // 3
    q, k, v = qkv[..., :dim], qkv[..., dim:2 * dim], qkv[..., 2 * dim:]
    return [t.reshape(B, H * W, heads, dim // heads) for t in (q, k, v)]


def _linear_attention(P, name, x, heads=4, dim_head=32):
    """unet.py:209-245."""
    B, H, W, C = x.shape
    q, k, v = _split_heads(_conv(P, name + '.to_qkv.conv_0', x, padding=0, bias=False), heads)
    q = torch.softmax(q, dim=-1)
    k = torch.softmax(k, dim=-3)
    q = q / math.sqrt(dim_head)
    v = v / (H * W)
    context = torch.einsum('bnhd,bnhe->bhde', k, v)
    out = torch.einsum('bhde,bnhd->bhen', context, q)
    out = out.permute(0, 3, 1, 2).reshape(B, H, W, heads * dim_head)      # 'b h e (x y) -> b x y (h e)'
    out = _conv(P, name + '.to_out.conv_0', out, padding=0)
    return _layer_norm(P, name + '.to_out.norm_0', out)


def _attention(P, name, x, heads=4, dim_head=32, scale=10):
    """unet.py:175-206."""
    B, H, W, C = x.shape
    q, k, v = _split_heads(_conv(P, name + '.to_qkv.conv_0', x, padding=0, bias=False), heads)

    def l2norm(t):  # axis=1 (the token axis!) as written upstream, unet.py:25-39,192
        return t / torch.clamp(torch.linalg.norm(t, dim=1, keepdim=True), min=1e-12)

    q, k = l2norm(q), l2norm(k)
    sim = torch.einsum('bihd,bjhd->bhij', q, k) * scale
    attn = torch.softmax(sim, dim=-1)
    out = torch.einsum('bhij,bjhd->bhid', attn, v)
    out = out.permute(0, 2, 1, 3).reshape(B, H, W, heads * dim_head)      # 'b h (x y) d -> b x y (h d)'
    return _conv(P, name + '.to_out.conv_0', out, padding=0)


def _attn_block(P, name, x, linear=True):
    """unet.py:248-264."""
    normed = _layer_norm(P, name + '.norm', x)
    out = _linear_attention(P, name + '.attn', normed) if linear else _attention(P, name + '.attn', normed)
    return out + x


def unet_forward(params, x, time, dt, dim=64, dim_mults=(1, 2, 4), groups=8):
    """UNet.__call__ (unet.py:279-368) with upsampling='pixel_shuffle'.  x: [B, H, W, C] float32, time: scalar."""
    P = _P(params)
    x = torch.from_numpy(np.asarray(x, dtype=np.float32))
    B = x.shape[0]
    hs = []
    h = _conv(P, 'init.conv_0', x, padding=3)
    hs.append(h)
    temb = sinusoidal_embedding(torch.tensor(float(time) / dt, dtype=torch.float32), out_dim=dim).expand(B, dim)
    temb = _dense(P, 'time.dense_0', temb)
    temb = _dense(P, 'time.dense_1', F.gelu(temb, approximate='tanh'))
    nres = len(dim_mults)
    for ind in range(nres):
        dim_in = h.shape[-1]
        h = _resnet_block(P, f'down_{ind}.resblock_0', h, temb, dim_in, groups)
        hs.append(h)
        h = _resnet_block(P, f'down_{ind}.resblock_1', h, temb, dim_in, groups)
        h = _attn_block(P, f'down_{ind}.attnblock_0', h)
        hs.append(h)
        if ind < nres - 1:
            h = _conv(P, f'down_{ind}.downsample_0', h, stride=2, padding=1)
    mid_dim = dim * dim_mults[-1]
    h = _conv(P, f'down_{nres - 1}.conv_0', h)
    h = _resnet_block(P, 'mid.resblock_0', h, temb, mid_dim, groups)
    h = _attn_block(P, 'mid.attenblock_0', h, linear=False)
    h = _resnet_block(P, 'mid.resblock_1', h, temb, mid_dim, groups)
    for ind in reversed(range(nres)):
        dim_in = dim * dim_mults[ind]
        dim_out = dim * dim_mults[ind - 1] if ind > 0 else dim
        h = torch.cat([h, hs.pop()], dim=-1)
        h = _resnet_block(P, f'up_{ind}.resblock_0', h, temb, dim_in, groups)
        h = torch.cat([h, hs.pop()], dim=-1)
        h = _resnet_block(P, f'up_{ind}.resblock_1', h, temb, dim_in, groups)
        h = _attn_block(P, f'up_{ind}.attnblock_0', h)
        if ind > 0:
            h = _conv(P, f'up_{ind}.upsample_0.conv_0', h)
            h = pixel_shuffle(h, 2)
            h = _conv(P, f'up_{ind}.upsample_0.conv_1', h)
    h = _conv(P, 'up_0.conv_0', h)
    h = torch.cat([h, hs.pop()], dim=-1)
    out = _resnet_block(P, 'final.resblock_0', h, temb, dim, groups)
    out = _conv(P, 'final.conv_0', out, padding=0)
    return out.numpy()


# ------------------------------------------------------------------------------------------------------------
# parameter shapes / random "checkpoint"
# ------------------------------------------------------------------------------------------------------------
def unet_param_shapes(in_ch, dim=64, dim_mults=(1, 2, 4), heads=4, dim_head=32):
    shapes = {}

    def conv(name, k, cin, cout, bias=True):
        shapes[name + '.kernel'] = (k, k, cin, cout)
        if bias:
            shapes[name + '.bias'] = (cout,)

    def dense(name, cin, cout):
        shapes[name + '.kernel'] = (cin, cout)
        shapes[name + '.bias'] = (cout,)

    def res(name, cin, d):
        conv(name + '.conv_0', 3, cin, d)
        shapes[name + '.norm_0.scale'] = (d,); shapes[name + '.norm_0.bias'] = (d,)
        dense(name + '.time_mlp.dense_0', 4 * dim, 2 * d)
        conv(name + '.conv_1', 3, d, d)
        shapes[name + '.norm_1.scale'] = (d,); shapes[name + '.norm_1.bias'] = (d,)
        if cin != d:
            conv(name + '.res_conv_0', 1, cin, d)

    def attn(name, c, linear=True):
        shapes[name + '.norm.scale'] = (c,)
        conv(name + '.attn.to_qkv.conv_0', 1, c, 3 * heads * dim_head, bias=False)
        conv(name + '.attn.to_out.conv_0', 1, heads * dim_head, c)
        if linear:
            shapes[name + '.attn.to_out.norm_0.scale'] = (c,)

    conv('init.conv_0', 7, in_ch, dim)
    dense('time.dense_0', dim, 4 * dim)
    dense('time.dense_1', 4 * dim, 4 * dim)
    nres = len(dim_mults)
    c = dim
    for ind in range(nres):
        res(f'down_{ind}.resblock_0', c, c)
        res(f'down_{ind}.resblock_1', c, c)
        attn(f'down_{ind}.attnblock_0', c)
        if ind < nres - 1:
            conv(f'down_{ind}.downsample_0', 4, c, dim * dim_mults[ind])
            c = dim * dim_mults[ind]
    mid = dim * dim_mults[-1]
    conv(f'down_{nres - 1}.conv_0', 3, c, mid)
    res('mid.resblock_0', mid, mid)
    attn('mid.attenblock_0', mid, linear=False)
    res('mid.resblock_1', mid, mid)
    for ind in reversed(range(nres)):
        dim_in = dim * dim_mults[ind]
        dim_out = dim * dim_mults[ind - 1] if ind > 0 else dim
        res(f'up_{ind}.resblock_0', dim_in + dim_out, dim_in)
        res(f'up_{ind}.resblock_1', dim_in + dim_out, dim_in)
        attn(f'up_{ind}.attnblock_0', dim_in)
        if ind > 0:
            conv(f'up_{ind}.upsample_0.conv_0', 3, dim_in, 4 * dim_in)
            conv(f'up_{ind}.upsample_0.conv_1', 3, dim_in, dim_out)
    conv('up_0.conv_0', 3, dim, dim)
    res('final.resblock_0', 2 * dim, dim)
    conv('final.conv_0', 1, dim, in_ch)
    return shapes


def init_unet_params(seed, in_ch, dim=64, dim_mults=(1, 2, 4)):
    rng = np.random.default_rng(seed)
    params = {}
    for name, shape in unet_param_shapes(in_ch, dim, dim_mults).items():
        if name.endswith('.kernel'):
            fan_in = int(np.prod(shape[:-1]))
            params[name] = (rng.standard_normal(shape) / math.sqrt(fan_in)).astype(np.float32)
        elif name.endswith('.scale'):
            params[name] = (1. + 0.1 * rng.standard_normal(shape)).astype(np.float32)
        else:
            params[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
    return params
