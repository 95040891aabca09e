"""Score network on the tensor cores (fbs_b200/nn, nn_conv.cu, nn_ops.cu) against plain float32 torch references:
each kernel against the same op in torch (bf16-rounded operands where the kernel takes bf16), the whole U-Net and
the NN-score closures against the oracle restatement of fbs/nn/unet.py + experiments/imgs/inpainting.py:94-147.

Tolerances: the kernels multiply bf16 operands and accumulate in fp32.  Single ops given identical bf16 inputs:
rtol 2e-3 (summation order).  Whole network vs the fp32 oracle: bf16 rounding of every activation and weight,
|err| <= 4e-2 * max|ref| (measured ~1e-2), log-weights atol 2e-2 * sqrt(q c) relative scale (see the closure test).
"""
import math
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import unet as ou          # noqa: E402
from oracle import jax_random as jr    # noqa: E402


def _bf(x):
    return x.to(torch.bfloat16)


def _conv_ref(x, w, bias, pad, stride=1):
    """fp32 CPU conv on NHWC input / HWIO kernel."""
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), bias, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('B,H,W,C0,C1,Cout,k', [
    (3, 28, 28, 64, 0, 64, 3),       # the most common layer
    (5, 14, 14, 128, 64, 128, 3),    # two sources (skip concatenation), tile rows cross the image end
    (4, 7, 7, 256, 128, 256, 3),     # two samples per M tile
    (2, 7, 7, 256, 0, 1024, 3),      # N tiles of 256
    (3, 28, 28, 64, 0, 384, 1),      # qkv projection: N tile 192, no bias
    (2, 16, 16, 64, 64, 64, 1),      # res_conv over a concatenation
    (1, 64, 64, 64, 0, 64, 3),       # CelebA-64 width: two full rows per tile
    (1, 28, 28, 64, 0, 64, 3),       # haloed-box scheme: one sample (7 tiles)
    (2, 28, 28, 64, 64, 64, 3),      # ... two sources, two resident weight chunks, two-stage ring
    (5, 14, 14, 128, 0, 128, 3),     # ... two N tiles per M tile (a CTA keeps its N tile), ragged last row tile
    (3, 32, 32, 64, 0, 192, 3),      # ... three N tiles, three rows of 34 per tile
    (150, 14, 14, 64, 0, 64, 3),     # ... more tiles than SMs: CTAs loop, accumulator buffers and ring wrap
    (200, 28, 28, 64, 0, 64, 3),     # ... ~9.5 tiles per CTA: the four accumulator buffers and both epilogue groups wrap twice
    (200, 28, 28, 64, 0, 64, 1),     # the same for the per-tap scheme (1x1)
    (5, 7, 7, 256, 0, 256, 3),       # ... N tile 32 (four resident weight chunks), one sample with its halo per tile
    (3, 14, 14, 128, 64, 128, 3),    # ... N tile 32, two sources
])
def test_conv_matches_torch(B, H, W, C0, C1, Cout, k):
    from fbs_b200.nn import ops
    g = torch.Generator().manual_seed(1)
    x0 = _bf(torch.randn(B, H, W, C0, generator=g))
    x1 = _bf(torch.randn(B, H, W, C1, generator=g)) if C1 else None
    w = _bf(torch.randn(k, k, C0 + C1, Cout, generator=g) / math.sqrt(k * k * (C0 + C1)))
    bias = torch.randn(Cout, generator=g) if k == 3 else None
    res = torch.randn(B, H, W, Cout, generator=g)
    xin = torch.cat([x0, x1], dim=-1) if C1 else x0
    want = _conv_ref(xin.float(), w.float(), bias, k // 2) + res
    wp = w.float().reshape(k * k * (C0 + C1), Cout).t().contiguous().to(torch.bfloat16).cuda()
    out32 = torch.empty(B, H, W, Cout, device='cuda')
    out16 = torch.empty(B, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    ops.conv(x0.cuda(), wp, Cout, k, k, -(k // 2), H, W, in1=None if x1 is None else x1.cuda(),
             bias=None if bias is None else bias.cuda(), residual=res.cuda(), out_f32=out32, out_bf16=out16)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out32.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(out16.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=2e-2)


@pytest.mark.parametrize('B,H,W,C0,C1,Cout', [(3, 28, 28, 64, 0, 64), (5, 14, 14, 128, 64, 128), (2, 32, 32, 64, 0, 128),
                                              (150, 14, 14, 64, 0, 64), (2, 16, 16, 128, 0, 256)])
def test_conv_hands_groupnorm_its_statistics(B, H, W, C0, C1, Cout):
    """conv -> GroupNorm pair: the statistics come from the convolution's epilogue (`gn_partials`), the normalisation is one
    streaming pass (fp32 or bf16 activations); against the oracle GroupNorm of the fp32 convolution output."""
    from fbs_b200.nn import ops
    g = torch.Generator().manual_seed(7)
    x0 = _bf(torch.randn(B, H, W, C0, generator=g))
    x1 = _bf(torch.randn(B, H, W, C1, generator=g)) if C1 else None
    w = _bf(torch.randn(3, 3, C0 + C1, Cout, generator=g) / math.sqrt(9 * (C0 + C1)))
    bias = torch.randn(Cout, generator=g)
    gamma, beta = torch.randn(Cout, generator=g), torch.randn(Cout, generator=g)
    tss = torch.randn(2 * Cout, generator=g) * 0.3
    res = torch.randn(B, H, W, Cout, generator=g)
    wp = w.float().reshape(9 * (C0 + C1), Cout).t().contiguous().to(torch.bfloat16).cuda()
    kw = dict(in1=None if x1 is None else x1.cuda(), bias=bias.cuda())
    args = (x0.cuda(), wp, Cout, 3, 3, -1, H, W)
    t32 = torch.empty(B, H, W, Cout, device='cuda')
    t16 = torch.empty(B, H, W, Cout, device='cuda', dtype=torch.bfloat16)
    slots = ops.conv_gn_slots(*args, out_f32=t32, **kw)
    assert slots > 0
    part = torch.full((B, slots, Cout // 4, 2), float('nan'), device='cuda')     # every entry must be written
    ops.conv(*args, out_f32=t32, out_bf16=t16, gn_partials=part, **kw)
    torch.cuda.synchronize()
    y = t32.cpu()
    np.testing.assert_allclose(part[..., 0].sum(dim=(1, 2)).cpu().numpy(), y.sum(dim=(1, 2, 3)).numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(part[..., 1].sum(dim=(1, 2)).cpu().numpy(), (y * y).sum(dim=(1, 2, 3)).numpy(), rtol=1e-4)
    Pm = ou._P({'n.scale': gamma.numpy(), 'n.bias': beta.numpy()})
    want = ou._swish(ou._group_norm(Pm, 'n', y) * (1 + tss[:Cout]) + tss[Cout:]) + res
    out = torch.empty_like(t32)
    out16 = torch.empty_like(t16)
    ops.groupnorm_swish_stats(t32, part, gamma.cuda(), beta.cuda(), 8, tss=tss.cuda(), residual=res.cuda(), out_f32=out, out_bf16=out16)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out16.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=2e-2)
    ops.groupnorm_swish_stats(t16, part, gamma.cuda(), beta.cuda(), 8, tss=tss.cuda(), residual=res.cuda(), out_f32=out)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=2e-2, atol=4e-2)   # bf16 activations
    if Cout <= 128:   # ... and the LayerNorm of the attention block that follows, from the same registers
        lg = torch.randn(Cout, generator=g)
        ln16 = torch.empty_like(t16)
        ops.groupnorm_swish_stats(t32, part, gamma.cuda(), beta.cuda(), 8, tss=tss.cuda(), residual=res.cuda(), out_f32=out,
                                  ln_gamma=lg.cuda(), ln_out_bf16=ln16)
        np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
        want_ln = ou._layer_norm(ou._P({'n.scale': lg.numpy()}), 'n', want)
        np.testing.assert_allclose(ln16.float().cpu().numpy(), want_ln.numpy(), rtol=1e-2, atol=2e-2)
    # tilings that pack several samples into a tile, residuals and pixel shuffle cannot: the query says so, the call refuses
    assert ops.conv_gn_slots(_bf(torch.randn(4, 7, 7, 256)).cuda(), _bf(torch.randn(256, 9 * 256)).cuda(), 256, 3, 3, -1, 7, 7,
                             out_f32=torch.empty(4, 7, 7, 256, device='cuda')) == 0
    with pytest.raises(ValueError):
        ops.conv(*args, out_f32=t32, residual=res.cuda(), gn_partials=part, **kw)


def test_conv_pixel_shuffle_and_stride2():
    from fbs_b200.nn import ops
    from fbs_b200.nn.unet import _pack, _pack_stride2
    g = torch.Generator().manual_seed(2)
    # Upsample conv_0 + PixelShuffle (unet.py:68-69)
    B, H, W, C = 3, 7, 7, 128
    x = _bf(torch.randn(B, H, W, C, generator=g))
    w = _bf(torch.randn(3, 3, C, 4 * C, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(4 * C, generator=g)
    want = ou.pixel_shuffle(_conv_ref(x.float(), w.float(), bias, 1), 2)
    out = torch.empty(B, 2 * H, 2 * W, C, device='cuda', dtype=torch.bfloat16)
    ops.conv(x.cuda(), torch.from_numpy(_pack(w.float().numpy())).to(torch.bfloat16).cuda(), 4 * C, 3, 3, -1, H, W, bias=bias.cuda(),
             out_bf16=out, pixel_shuffle=True)
    np.testing.assert_allclose(out.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=2e-2)
    # Downsample: 4x4 stride 2 padding 1 (unet.py:50) as space-to-depth + 2x2
    B, H, W, C, Cout = 3, 28, 28, 64, 128
    x = _bf(torch.randn(B, H, W, C, generator=g))
    w = _bf(torch.randn(4, 4, C, Cout, generator=g) / math.sqrt(16 * C))
    bias = torch.randn(Cout, generator=g)
    want = _conv_ref(x.float(), w.float(), bias, 1, stride=2)
    s2d = torch.empty(B, H // 2 + 1, W // 2 + 1, 4 * C, device='cuda', dtype=torch.bfloat16)
    ops.space_to_depth(x.cuda(), s2d)
    out = torch.empty(B, H // 2, W // 2, Cout, device='cuda')
    ops.conv(s2d, torch.from_numpy(_pack_stride2(w.float().numpy())).to(torch.bfloat16).cuda(), Cout, 2, 2, 0, H // 2, W // 2,
             bias=bias.cuda(), out_f32=out)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize('B,P,C', [(80, 784, 64), (76, 196, 128), (75, 49, 256), (74, 4096, 64), (3, 100, 512)])
def test_norm_kernels_for_full_batches(B, P, C):
    """The one-CTA-per-sample GroupNorm (batches that fill the machine; (74, 4096, 64) does not fit shared memory and re-reads)
    and the 16-byte LayerNorm against the oracle pieces, with and without the optional operands."""
    from fbs_b200.nn import ops
    g = torch.Generator().manual_seed(B + P)
    x = torch.randn(B, P, 1, C, generator=g) * 1.5 - 0.3
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    tss = torch.randn(2 * C, generator=g) * 0.3
    res = torch.randn(B, P, 1, C, generator=g)
    Pm = ou._P({'n.scale': gamma.numpy(), 'n.bias': beta.numpy()})
    out = torch.empty_like(x, device='cuda')
    out16 = torch.empty(x.shape, device='cuda', dtype=torch.bfloat16)
    want = ou._swish(ou._group_norm(Pm, 'n', x) * (1 + tss[:C]) + tss[C:]) + res
    ops.groupnorm_swish(x.cuda(), gamma.cuda(), beta.cuda(), 8, tss=tss.cuda(), residual=res.cuda(), out_f32=out, out_bf16=out16)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=3e-5, atol=3e-5)
    np.testing.assert_allclose(out16.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=1e-2)
    want = ou._swish(ou._group_norm(Pm, 'n', x))
    ops.groupnorm_swish(x.cuda(), gamma.cuda(), beta.cuda(), 8, out_f32=out)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=3e-5, atol=3e-5)
    want = ou._layer_norm(Pm, 'n', x) + res
    ops.layernorm(x.cuda(), gamma.cuda(), residual=res.cuda(), out_f32=out, out_bf16=out16)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=3e-5, atol=3e-5)
    np.testing.assert_allclose(out16.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=1e-2)
    want = ou._layer_norm(Pm, 'n', x)
    ops.layernorm(x.cuda(), gamma.cuda(), out_f32=out)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=3e-5, atol=3e-5)


def test_norms_and_attention_match_oracle_pieces():
    from fbs_b200.nn import ops
    g = torch.Generator().manual_seed(3)
    B, H, W, C = 3, 14, 14, 128
    x = torch.randn(B, H, W, C, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    tss = torch.randn(2 * C, generator=g) * 0.3
    res = torch.randn(B, H, W, C, generator=g)
    P = ou._P({'n.scale': gamma.numpy(), 'n.bias': beta.numpy()})
    want = ou._swish(ou._group_norm(P, 'n', x) * (1 + tss[:C]) + tss[C:]) + res
    out = torch.empty_like(x, device='cuda')
    ops.groupnorm_swish(x.cuda(), gamma.cuda(), beta.cuda(), 8, tss=tss.cuda(), residual=res.cuda(), out_f32=out)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=2e-5, atol=2e-5)
    want = ou._layer_norm(P, 'n', x) + res
    ops.layernorm(x.cuda(), gamma.cuda(), residual=res.cuda(), out_f32=out)
    np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=2e-5, atol=2e-5)
    # attention cores on a given qkv
    for P_, linear in ((196, True), (784, True), (100, True), (49, False), (300, False)):
        qkv = _bf(torch.randn(B, P_, 1, 384, generator=g)).float()
        q, k, v = ou._split_heads(qkv, 4)
        if linear:
            qs, ks = torch.softmax(q, dim=-1) / math.sqrt(32), torch.softmax(k, dim=-3)
            ctx = torch.einsum('bnhd,bnhe->bhde', ks, v / P_)
            want = torch.einsum('bhde,bnhd->bhen', ctx, qs).permute(0, 3, 1, 2).reshape(B, P_, 128)
        else:
            l2 = lambda t: t / torch.clamp(torch.linalg.norm(t, dim=1, keepdim=True), min=1e-12)
            sim = torch.einsum('bihd,bjhd->bhij', l2(q), l2(k)) * 10
            want = torch.einsum('bhij,bjhd->bhid', torch.softmax(sim, dim=-1), v).permute(0, 2, 1, 3).reshape(B, P_, 128)
        o = torch.empty(B, P_, 128, device='cuda', dtype=torch.bfloat16)
        (ops.linear_attention if linear else ops.attention)(qkv.to(torch.bfloat16).cuda().contiguous(), o)
        np.testing.assert_allclose(o.float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=2e-3 * float(want.abs().max()))


def _mnist_setup(seed=0, B=3):
    params = ou.init_unet_params(seed, 1)
    rng = np.random.default_rng(seed + 1)
    x = rng.standard_normal((B, 28, 28, 1)).astype(np.float32)
    return params, x


def test_time_mlp_and_stem_match_oracle():
    from fbs_b200.nn import ScoreUNet
    params, x = _mnist_setup()
    net = ScoreUNet(params, (28, 28, 1), dt=2. / 200)
    net(torch.from_numpy(x).cuda(), 0.731, use_graph=False)
    torch.cuda.synchronize()
    P = ou._P(params)
    temb = ou.sinusoidal_embedding(torch.tensor(0.731 / (2. / 200)), 64)[None]
    temb = ou._dense(P, 'time.dense_1', F.gelu(ou._dense(P, 'time.dense_0', temb), approximate='tanh'))
    for blk, off, d in net._time_blocks:
        want = ou._dense(P, blk + '.time_mlp.dense_0', ou._swish(temb))[0]
        np.testing.assert_allclose(net.table[off:off + 2 * d].cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-4)
    want = ou._conv(P, 'init.conv_0', torch.from_numpy(x), padding=3)
    np.testing.assert_allclose(net._bufs[(3, 'h0_f32')].cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('shape', [(28, 28, 1), (32, 32, 3)])
def test_unet_matches_oracle(shape):
    from fbs_b200.nn import ScoreUNet
    H, W, C = shape
    params = ou.init_unet_params(5, C)
    rng = np.random.default_rng(6)
    B = 5
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    net = ScoreUNet(params, shape, dt=2. / 200)
    for t in (0.02, 1.3):
        want = ou.unet_forward(params, x, t, 2. / 200)
        got_eager = net(torch.from_numpy(x).cuda(), t, use_graph=False).cpu().numpy().copy()
        got_graph = net(torch.from_numpy(x).cuda(), t).cpu().numpy().copy()
        np.testing.assert_array_equal(got_eager, got_graph)      # the CUDA graph replays the same kernels
        scale = float(np.abs(want).max())
        err = np.abs(got_graph - want)
        assert err.max() <= 4e-2 * scale, (err.max(), scale)
        assert err.mean() <= 6e-3 * scale, (err.mean(), scale)
    # batch-size independence: one sample alone gives the same result as inside the batch (bitwise: same tiles' math)
    one = net(torch.from_numpy(x[:1]).cuda(), 1.3).cpu().numpy()
    np.testing.assert_allclose(one[0], got_graph[0], rtol=1e-3, atol=1e-3 * scale)


def test_score_closures_match_oracle():
    """ScoreNetModel.step == transition_sampler + likelihood_logpdf of experiments/imgs/inpainting.py:122-147 restated
    with the fp32 oracle network: noise bit-pinned by the key, means / log-weights within the bf16 network tolerance."""
    from fbs_b200.nn import ScoreUNet, ScoreNetModel
    from fbs_b200 import sdes
    params, _ = _mnist_setup(7)
    H = W = 28
    T, K, N = 2.0, 200, 6
    ts = np.linspace(0., T, K + 1)
    sde = sdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=5., t0=0., T=T)
    # inpaint-15 style mask: a 15x15 square unobserved
    rect = np.array([(i + 4) * W + (j + 4) for i in range(15) for j in range(15)], dtype=np.int32)
    obs = np.setdiff1d(np.arange(H * W, dtype=np.int32), rect)
    net = ScoreUNet(params, (H, W, 1), dt=T / 200)
    model = ScoreNetModel(net, sde, ts, T, rect, obs)
    rng = np.random.default_rng(8)
    us = rng.standard_normal((N, rect.size, 1)).astype(np.float32)
    v_prev = rng.standard_normal((obs.size, 1)).astype(np.float32)
    v_next = (v_prev + 0.05 * rng.standard_normal((obs.size, 1))).astype(np.float32)
    key = jr.PRNGKey(99)
    k = 120
    t_prev = ts[k]
    us_new, lw = model.step(us, v_prev, v_next, t_prev, key)
    us_new, lw = us_new.cpu().numpy(), lw.cpu().numpy()
    # oracle restatement
    dt = T / K
    s = T - t_prev
    img = np.zeros((N, H * W, 1), np.float32)
    img[:, rect] = us
    img[:, obs] = v_prev
    score = ou.unet_forward(params, img.reshape(N, H, W, 1), s, T / 200).reshape(N, H * W, 1)
    a, g = sde.drift_coef(s), sde.dispersion(s)
    rd = -a * img + g ** 2 * score
    sd = math.sqrt(dt) * g
    mean_u = us + rd[:, rect] * dt
    want_us = mean_u + np.float32(sd) * jr.normal(key, us.shape)
    z = (v_next[None] - (v_prev[None] + rd[:, obs] * dt)) / sd
    want_lw = (-0.5 * z ** 2 - math.log(sd) - 0.5 * math.log(2 * math.pi)).sum(axis=(1, 2))
    np.testing.assert_allclose(us_new, want_us, rtol=0, atol=4e-2 * dt * g ** 2 * float(np.abs(score).max()) + 1e-5)
    # log-weights: differences between particles are what the sampler uses
    np.testing.assert_allclose(lw - lw.mean(), want_lw - want_lw.mean(), rtol=0, atol=5e-2 * max(1.0, float(np.ptp(want_lw))))
    np.testing.assert_allclose(lw, want_lw, rtol=2e-3, atol=0.5)
    # the separate closures agree with the fused step
    us2 = model.transition_sampler(us, v_prev, t_prev, key).cpu().numpy()
    lw2 = model.likelihood_logpdf(v_next, us, v_prev, t_prev).cpu().numpy()
    np.testing.assert_array_equal(us2, us_new)
    np.testing.assert_array_equal(lw2, lw)


def _inpaint_problem(K=5, N=7, seed=11):
    from fbs_b200.nn import ScoreUNet, ScoreNetModel
    from fbs_b200 import sdes
    params = ou.init_unet_params(seed, 1)
    H = W = 28
    T = 2.0
    ts = np.linspace(0., T, K + 1)
    sde = sdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=5., t0=0., T=T)
    rect = np.array([(i + 6) * W + (j + 6) for i in range(15) for j in range(15)], dtype=np.int32)
    obs = np.setdiff1d(np.arange(H * W, dtype=np.int32), rect)
    net = ScoreUNet(params, (H, W, 1), dt=T / 200)
    model = ScoreNetModel(net, sde, ts, T, rect, obs)
    return params, model, sde, ts, T, rect, obs


def _oracle_step(params, sde, T, dt, rect, obs, us, v_prev, v_next, t_prev, key):
    """transition_sampler + likelihood_logpdf of inpainting.py:122-147 with the fp32 oracle network."""
    N = us.shape[0]
    s = T - t_prev
    img = np.zeros((N, 784, 1), np.float32)
    img[:, rect] = us
    img[:, obs] = v_prev
    score = ou.unet_forward(params, img.reshape(N, 28, 28, 1), s, T / 200).reshape(N, 784, 1)
    a, g = sde.drift_coef(s), sde.dispersion(s)
    rd = -a * img + g ** 2 * score
    sd = math.sqrt(dt) * g
    new = us + rd[:, rect] * dt + np.float32(sd) * jr.normal(key, us.shape)
    z = (v_next[None] - (v_prev[None] + rd[:, obs] * dt)) / sd
    return new.astype(np.float32), (-0.5 * z ** 2 - math.log(sd) - 0.5 * math.log(2 * math.pi)).sum(axis=(1, 2))


@pytest.mark.parametrize('explicit_final', [False, True])
def test_nn_forward_pass_teacher_forced(explicit_final):
    """csmc.py:132-164 with the score-network closures, step by step from the kernel's own history: ancestors bit-exact
    on identical weights, children / log-weights within the bf16 network tolerance, the reference slot pinned exactly."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    from oracle import cond_resampling as ocr, csmc as ocsmc
    K, N = 5, 7
    params, model, sde, ts, T, rect, obs = _inpaint_problem(K, N)
    rng = np.random.default_rng(3)
    us_star = rng.standard_normal((K + 1, rect.size, 1)).astype(np.float32)
    vs = np.cumsum(0.05 * rng.standard_normal((K + 1, obs.size, 1)), axis=0).astype(np.float32)
    key = jr.PRNGKey(5)
    init = csmc.NormalInit(model) if explicit_final else csmc.DegenerateInit(N)
    Np = N + 1 if explicit_final else N
    bs_star = jr.randint(jr.PRNGKey(6), (K + 1,), 0, N).astype(np.int32)
    As, log_wss, uss = csmc.forward_pass(key, us_star, bs_star, vs, ts, init.sampler, init.likelihood_logpdf,
                                         model.transition_sampler, model.likelihood_logpdf, R.killing, N)
    uss = uss.reshape(K + 1, Np, rect.size, 1)
    assert As.shape == (K, Np) and log_wss.shape == (K + 1, Np)
    key_init, key_scan = jr.split(key, 2)
    if explicit_final:
        u0 = jr.normal(key_init, (Np, rect.size, 1))
        u0[bs_star[0]] = us_star[0]
        np.testing.assert_allclose(uss[0], u0, rtol=0, atol=5e-7)
    else:
        np.testing.assert_array_equal(uss[0], np.broadcast_to(us_star[0], uss[0].shape))
        np.testing.assert_allclose(log_wss[0], -np.log(N), atol=1e-6)
    dt = T / K
    for k, step_key in enumerate(jr.split(key_scan, K)):
        key_res, key_tr = jr.split(step_key, 2)
        A = ocr.killing(key_res, np.exp(log_wss[k]).astype(np.float32), bs_star[k], bs_star[k + 1], True)
        np.testing.assert_array_equal(As[k], A)
        want_us, want_lw = _oracle_step(params, sde, T, dt, rect, obs, uss[k][As[k]], vs[k], vs[k + 1], ts[k], key_tr)
        want_us[bs_star[k + 1]] = us_star[k + 1]
        np.testing.assert_array_equal(uss[k + 1][bs_star[k + 1]], us_star[k + 1])
        # dt = T / K = 0.4 here: the bf16 network error (<= 4e-2 of the score scale) is multiplied by g^2 dt ~ 1
        np.testing.assert_allclose(uss[k + 1], want_us, rtol=0, atol=3e-2 * max(1.0, float(np.abs(want_us).max())))
        want = ocsmc.normalise(want_lw, log_space=True)
        np.testing.assert_allclose(log_wss[k + 1], want, rtol=0, atol=5e-2 * max(1.0, float(np.ptp(want))))


def test_nn_gibbs_kernel_runs():
    """gibbs_kernel (gibbs.py:68-168, method 'gibbs-eb-ef' of experiments/bashes/imgs_gibbs.sh) over the score network."""
    from fbs_b200.samplers import gibbs_kernel
    K, N = 4, 5
    params, model, sde, ts, T, rect, obs = _inpaint_problem(K, N)
    rng = np.random.default_rng(4)
    x0 = rng.standard_normal((rect.size, 1)).astype(np.float32)
    y0 = rng.uniform(size=(obs.size, 1)).astype(np.float32)
    bs_star = np.zeros((K + 1,), np.int32)
    key = jr.PRNGKey(21)
    x0n, us_star, bs_next, changed = gibbs_kernel(key, x0, y0, None, bs_star, ts, model.fwd_sampler, sde, model.unpack, N,
                                                  model.transition_sampler, model.transition_logpdf, model.likelihood_logpdf,
                                                  explicit_backward=True, explicit_final=True)
    assert x0n.shape == (rect.size,) and us_star.shape == (K + 1, rect.size)
    assert np.isfinite(x0n).all() and np.isfinite(us_star).all()
    assert bs_next.shape == (K + 1,) and (bs_next >= 0).all() and (bs_next < N).all()
    np.testing.assert_array_equal(us_star[-1], x0n)
    # bs_star_next is randint(key_csmc_bwd_bs, (K + 1,), 0, nparticles)   (gibbs.py:156)
    kc = jr.split(jr.split(key, 3)[1], 4)
    np.testing.assert_array_equal(bs_next, jr.randint(kc[3], (K + 1,), 0, N))


def test_sharded_sweep_single_rank_equals_unsharded():
    """fbs_b200/sharded.py with a one-rank NCCL group is the unsharded sweep (the 2-GPU equality is checked by
    scripts/sharded_check.py under torchrun; the exchange logic for G > 1 by the gloo tests on CPU)."""
    import torch.distributed as dist
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    from fbs_b200.sharded import forward_pass_sharded
    K, N = 3, 6
    params, model, sde, ts, T, rect, obs = _inpaint_problem(K, N)
    rng = np.random.default_rng(13)
    us_star = rng.standard_normal((K + 1, rect.size, 1)).astype(np.float32)
    vs = np.cumsum(0.05 * rng.standard_normal((K + 1, obs.size, 1)), axis=0).astype(np.float32)
    bs_star = jr.randint(jr.PRNGKey(2), (K + 1,), 0, N).astype(np.int32)
    key = jr.PRNGKey(17)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group('nccl', init_method='tcp://127.0.0.1:29577', rank=0, world_size=1)
    try:
        for init in (csmc.DegenerateInit(N), csmc.NormalInit(model)):
            r = forward_pass_sharded(key, us_star, bs_star, vs, model, init, R.killing, N, history=True)
            full = csmc.forward_pass_nn(key, us_star, bs_star, vs, model, init, R.killing.scheme, N, history=True)
            assert torch.equal(r['As'], full['As'][0])
            assert torch.equal(r['log_wss'], full['log_wss'][0])
            assert torch.equal(r['uss'].reshape(full['uss'].shape[1:]), full['uss'][0])
            assert r['moved'] == [0] * K
    finally:
        if created:
            from fbs_b200.sharded import close_peer_buffers
            close_peer_buffers()
            dist.destroy_process_group()


def test_nn_pmcmc_filter_step_and_bootstrap_filter_one_step():
    """pmcmc_filter_step (smc.py:115-158) and bootstrap_filter (smc.py:58-88) over the score network against the oracle
    closures for K = 1 (one step is not chaotic): evidence, resampled indices on identical weights, particles."""
    from fbs_b200.samplers import pmcmc_filter_step, bootstrap_filter, stratified
    from oracle import resampling as orx
    K, N = 1, 8
    params, model, sde, ts, T, rect, obs = _inpaint_problem(K, N)
    rng = np.random.default_rng(31)
    u0s = rng.standard_normal((N, rect.size, 1)).astype(np.float32)
    vs = np.cumsum(0.05 * rng.standard_normal((K + 1, obs.size, 1)), axis=0).astype(np.float32)
    key = jr.PRNGKey(12)
    uT, log_ell = pmcmc_filter_step(key, vs, u0s, ts, model.transition_sampler, model.likelihood_logpdf, stratified, N)
    key_prop, key_res = jr.split(jr.split(key, K)[0], 2)
    mean, lw, sd = model.mean_and_logw(u0s, vs[0], vs[1], ts[0])
    mean, lw = mean.cpu().numpy(), lw.cpu().numpy().astype(np.float64)
    c = np.log(np.exp(lw - lw.max()).sum()) + lw.max()
    np.testing.assert_allclose(log_ell, c - math.log(N), rtol=1e-5, atol=1e-4)
    inds = orx.stratified(np.exp(lw - c).astype(np.float32), key_res)
    want = mean[inds] + np.float32(sd) * jr.normal(key_prop, u0s.shape)
    np.testing.assert_allclose(uT.reshape(want.shape), want, rtol=1e-5, atol=1e-5)
    # the mean / weights themselves against the fp32 oracle network
    want_us, want_lw = _oracle_step(params, sde, T, T / K, rect, obs, u0s, vs[0], vs[1], ts[0], key_prop)
    np.testing.assert_allclose(lw - lw.mean(), want_lw - want_lw.mean(), rtol=0, atol=5e-2 * max(1.0, float(np.ptp(want_lw))))
    # bootstrap filter: same step, proposal first then resampling of the children
    def init_sampler(key_, v0, n):
        return u0s
    us_bf, nell = bootstrap_filter(model.transition_sampler, model.likelihood_logpdf, vs, ts, init_sampler, key, N, stratified)
    k_init, k_steps = jr.split(key, 2)
    kp, kr = jr.split(jr.split(k_steps, K)[0], 2)
    children = mean + np.float32(sd) * jr.normal(kp, u0s.shape)
    inds = orx.stratified(np.exp(lw - c).astype(np.float32), kr)
    np.testing.assert_allclose(us_bf.reshape(children.shape), children[inds], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(nell, -(c - math.log(N)), rtol=1e-5, atol=1e-4)


def test_nn_pmcmc_kernel_and_gibbs_init_run():
    """pmcmc_kernel (smc.py:171-258) and gibbs_init (gibbs.py:23-65, filter and smoother) over the score network."""
    from fbs_b200.samplers import pmcmc_kernel, gibbs_init, stratified
    K, N = 3, 6
    params, model, sde, ts, T, rect, obs = _inpaint_problem(K, N)
    rng = np.random.default_rng(41)
    y0 = rng.uniform(size=(obs.size, 1)).astype(np.float32)
    key = jr.PRNGKey(3)
    ys = model.fwd_ys_sampler(jr.PRNGKey(4), y0).cpu().numpy()
    uT = rng.standard_normal((rect.size, 1)).astype(np.float32)
    for delta in (None, 0.1):
        uT2, le2, ys2, st = pmcmc_kernel(key, uT, np.float32(-1e30), ys, y0, ts, model.fwd_ys_sampler, sde, model.ref_sampler,
                                         model.transition_sampler, model.likelihood_logpdf, stratified, N, delta=delta)
        assert bool(st.is_accepted) and np.isfinite(le2) and np.isfinite(uT2).all() and ys2.shape == (K + 1, obs.size)
        assert uT2.shape == (rect.size,)
    for method in ('filter', 'smoother'):
        x0, us_star = gibbs_init(key, y0, (rect.size, 1), ts, model.fwd_sampler, sde, model.unpack, model.transition_sampler,
                                 model.transition_logpdf, model.likelihood_logpdf, N, method=method)
        assert x0.shape == (rect.size,) and us_star.shape == (K + 1, rect.size)
        assert np.isfinite(x0).all() and np.isfinite(us_star).all()


def test_nn_forward_pass_batched_chains_equals_single_chains():
    """forward_pass_nn_chains (C conditioning targets through ONE score evaluation per step) gives every chain exactly the
    result of its own forward_pass_nn: the batching is a scheduling decision, not a numerical one."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    K, N, C = 3, 6, 3
    params, model, sde, ts, T, rect, obs = _inpaint_problem(K, N)
    rng = np.random.default_rng(21)
    us_star = rng.standard_normal((C, K + 1, rect.size, 1)).astype(np.float32)
    vs = np.cumsum(0.05 * rng.standard_normal((C, K + 1, obs.size, 1)), axis=1).astype(np.float32)
    bs_star = np.stack([jr.randint(jr.PRNGKey(30 + ci), (K + 1,), 0, N) for ci in range(C)]).astype(np.int32)
    keys = jr.split(jr.PRNGKey(23), C)
    for init in (csmc.DegenerateInit(N), csmc.NormalInit(model)):
        got = csmc.forward_pass_nn_chains(keys, us_star, bs_star, vs, model, init, R.killing.scheme, N)
        n_part = got['N']
        assert tuple(got['us_last'].shape) == (C, n_part, rect.size) and tuple(got['log_ws_last'].shape) == (C, n_part)
        for ci in range(C):
            one = csmc.forward_pass_nn(keys[ci], us_star[ci], bs_star[ci], vs[ci], model, init, R.killing.scheme, N, history=False)
            assert torch.equal(got['us_last'][ci], one['us_last'][0]), f'chain {ci}: particles differ'
            assert torch.equal(got['log_ws_last'][ci], one['log_ws_last'][0]), f'chain {ci}: weights differ'


@pytest.mark.parametrize('explicit_final', [True, False])
def test_nn_gibbs_kernel_batched_targets_equals_single_targets(explicit_final):
    """gibbs_kernel over the score network with keys [C, 2] and one y0 per chain: every chain gets exactly what its own
    single-target call returns (the C chains only share the score evaluations)."""
    from fbs_b200.samplers import gibbs_kernel
    K, N, C = 3, 5, 3
    params, model, sde, ts, T, rect, obs = _inpaint_problem(K, N)
    rng = np.random.default_rng(8)
    x0 = rng.standard_normal((C, rect.size)).astype(np.float32)
    y0 = rng.uniform(size=(C, obs.size)).astype(np.float32)
    bs_star = np.stack([jr.randint(jr.PRNGKey(40 + ci), (K + 1,), 0, N) for ci in range(C)]).astype(np.int32)
    keys = jr.split(jr.PRNGKey(77), C)
    kw = dict(explicit_backward=True, explicit_final=explicit_final)
    args = (ts, model.fwd_sampler, sde, model.unpack, N, model.transition_sampler, model.transition_logpdf, model.likelihood_logpdf)
    x0n, us_next, bs_next, changed = gibbs_kernel(keys, x0, y0, None, bs_star, *args, **kw)
    assert x0n.shape == (C, rect.size) and us_next.shape == (C, K + 1, rect.size) and bs_next.shape == (C, K + 1)
    for ci in range(C):
        a, b, c_, d = gibbs_kernel(keys[ci], x0[ci], y0[ci], None, bs_star[ci], *args, **kw)
        np.testing.assert_array_equal(x0n[ci], a)
        np.testing.assert_array_equal(us_next[ci], b)
        np.testing.assert_array_equal(bs_next[ci], c_)
        np.testing.assert_array_equal(changed[ci], d)


def test_peer_gather_kernel_owner_arithmetic():
    """fbs_gather_rows_peer_f32 with a pointer table of three LOCAL buffers standing in for three ranks (the CUDA-IPC imports
    need several processes: scripts/sharded_check.py): dst[b] = bufs[idx[b] // n][idx[b] % n]."""
    from fbs_b200 import _native as nat
    from fbs_b200._tensor import ptr, stream
    n, G = 5, 3
    for row in (12, 7):                                     # 16-byte vector path and scalar path
        bufs = [torch.randn(n, row, device='cuda') for _ in range(G)]
        table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device='cuda')
        idx = torch.tensor([14, 0, 7, 7, 3, 10, 4, 5], dtype=torch.int32, device='cuda')
        dst = torch.empty(idx.numel(), row, device='cuda')
        nat.call('fbs_gather_rows_peer_f32', stream(), ptr(table), ptr(idx), idx.numel(), row, n, G, ptr(dst))
        want = torch.cat(bufs)[idx.long()]
        assert torch.equal(dst, want)


# ------------------------------------------------------------------------------------------------------------------
# Schroedinger-bridge image closures (experiments/sb_imgs/supr.py:84-137): raw-network reverse drift (param_bwd) and the
# Euler--Maruyama forward sampler whose drift is a second network (param_fwd)
# ------------------------------------------------------------------------------------------------------------------
def _sb_problem(K=4):
    from fbs_b200.nn import ScoreUNet, ScoreNetModel
    from fbs_b200 import sdes
    from oracle import sdes as osdes
    from oracle.sb_images import SBImageModel
    H = W = 28
    T = 0.5                                                            # supr.py:45
    ts = np.linspace(0., T, K + 1)
    param_b, param_f = ou.init_unet_params(21, 1), ou.init_unet_params(22, 1)
    obs = np.array([i * W + j for i in range(0, H, 4) for j in range(0, W, 4)], dtype=np.int32)   # supr-4: 49 observed pixels
    unobs = np.setdiff1d(np.arange(H * W, dtype=np.int32), obs)
    net_dt = 0.5 / 200                                                 # supr.py:67
    sde = sdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=5., t0=0., T=T)
    osde = osdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=5., t0=0., T=T)
    model = ScoreNetModel(ScoreUNet(param_b, (H, W, 1), dt=net_dt), sde, ts, T, unobs, obs, drift_mode=True,
                          fwd_unet=ScoreUNet(param_f, (H, W, 1), dt=net_dt))
    omodel = SBImageModel(param_b, param_f, osde, ts, T, (H, W, 1), unobs, obs, net_dt)
    return model, omodel, sde, ts, T, unobs, obs


def test_sb_image_closures_match_oracle():
    """transition_sampler / likelihood_logpdf / transition_logpdf in drift mode (supr.py:104-129): noise bit-pinned, drift
    within the bf16 network tolerance (4e-2 of the drift scale, multiplied by dt)."""
    K, N = 4, 6
    model, om, sde, ts, T, unobs, obs = _sb_problem(K)
    rng = np.random.default_rng(1)
    us = rng.standard_normal((N, unobs.size, 1)).astype(np.float32)
    v_prev = rng.uniform(size=(obs.size, 1)).astype(np.float32)
    v_next = (v_prev + 0.05 * rng.standard_normal(v_prev.shape)).astype(np.float32)
    u_eval = rng.standard_normal((unobs.size, 1)).astype(np.float32)
    key = jr.PRNGKey(3)
    k = 2
    t_prev = ts[k]
    drift = om.reverse_drift(om.concat(us, v_prev), t_prev)
    scale = float(np.abs(drift).max())
    dt = T / K
    got_us = model.transition_sampler(us, v_prev, t_prev, key).cpu().numpy()
    np.testing.assert_allclose(got_us, om.transition_sampler(us, v_prev, t_prev, key), rtol=0, atol=4e-2 * dt * scale + 1e-5)
    # the noise is exactly normal(key, us.shape) scaled by sqrt(dt) dispersion(T - t): remove the oracle mean and compare
    sd = np.float32(math.sqrt(dt)) * np.float32(om.reverse_dispersion(t_prev))
    np.testing.assert_allclose(got_us - om.transition_mean(us, v_prev, t_prev), sd * jr.normal(key, us.shape), rtol=0,
                               atol=4e-2 * dt * scale + 1e-5)
    lw = model.likelihood_logpdf(v_next, us, v_prev, t_prev).cpu().numpy()
    want_lw = om.likelihood_logpdf(v_next, us, v_prev, t_prev)
    np.testing.assert_allclose(lw - lw.mean(), want_lw - want_lw.mean(), rtol=0, atol=5e-2 * max(1.0, float(np.ptp(want_lw))))
    tlp = model.transition_logpdf(u_eval, us, v_prev, t_prev).cpu().numpy()
    want_tlp = om.transition_logpdf(u_eval, us, v_prev, t_prev)
    np.testing.assert_allclose(tlp, want_tlp, rtol=2e-3, atol=5e-2 * max(1.0, float(np.ptp(want_tlp))))
    # one fused evaluation == the separate closures
    us2, lw2 = model.step(us, v_prev, v_next, t_prev, key)
    np.testing.assert_array_equal(us2.cpu().numpy(), got_us)
    np.testing.assert_array_equal(lw2.cpu().numpy(), lw)
    with pytest.raises(NotImplementedError):
        model.fwd_ys_sampler(key, v_prev)


def test_sb_image_forward_sampler_teacher_forced():
    """fwd_sampler = euler_maruyama(key, concat(x0, y0), ts, nn_drift(., t, param_fwd), sde.dispersion, 1, return_path=True)
    (supr.py:132-137, simulators.py:81-92), every interval re-derived by the oracle from the kernel's own state."""
    K = 5
    model, om, sde, ts, T, unobs, obs = _sb_problem(K)
    rng = np.random.default_rng(2)
    x0 = rng.uniform(size=(unobs.size, 1)).astype(np.float32)
    y0 = rng.uniform(size=(obs.size, 1)).astype(np.float32)
    key = jr.PRNGKey(9)
    path = model.fwd_sampler(key, x0, y0).cpu().numpy()
    assert path.shape == (K + 1, 28, 28, 1)
    np.testing.assert_array_equal(path[0], om.concat(x0[None], y0)[0])
    keys = jr.split(key, K)
    ts32 = ts.astype(np.float32)
    for k in range(K):
        ddt = np.float32(ts32[k + 1] - ts32[k])
        drift = ou.unet_forward(om.param_fwd, path[k][None], float(ts32[k]), om.net_dt)[0]
        want = path[k] + drift * ddt + np.float32(om.sde.dispersion(ts32[k])) * np.sqrt(ddt) * jr.normal(keys[k], (1, 28, 28, 1))[0]
        np.testing.assert_allclose(path[k + 1], want, rtol=0, atol=4e-2 * float(ddt) * float(np.abs(drift).max()) + 1e-5)
    # reversed / unpacked as gibbs_kernel consumes it (gibbs.py:127-130)
    us, vs = model.fwd_sampler_reversed(key, x0, y0)
    np.testing.assert_array_equal(us.cpu().numpy()[0], path[::-1].reshape(K + 1, 784)[:, unobs])
    np.testing.assert_array_equal(vs.cpu().numpy()[0], path[::-1].reshape(K + 1, 784)[:, obs])


def test_sb_image_forward_pass_teacher_forced_and_gibbs_kernel():
    """csmc.py:132-164 and gibbs.py:68-168 (explicit_backward = explicit_final = True as in supr.py:171-176) over the
    drift-mode closures."""
    from fbs_b200.samplers import gibbs_kernel
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    from oracle import cond_resampling as ocr, csmc as ocsmc
    K, N = 4, 5
    model, om, sde, ts, T, unobs, obs = _sb_problem(K)
    rng = np.random.default_rng(5)
    us_star = rng.standard_normal((K + 1, unobs.size, 1)).astype(np.float32)
    vs = np.cumsum(0.05 * rng.standard_normal((K + 1, obs.size, 1)), axis=0).astype(np.float32)
    key = jr.PRNGKey(8)
    init = csmc.NormalInit(model)
    Np = N + 1
    bs_star = jr.randint(jr.PRNGKey(6), (K + 1,), 0, N).astype(np.int32)
    As, log_wss, uss = csmc.forward_pass(key, us_star, bs_star, vs, ts, init.sampler, init.likelihood_logpdf,
                                         model.transition_sampler, model.likelihood_logpdf, R.killing, N)
    uss = uss.reshape(K + 1, Np, unobs.size, 1)
    key_init, key_scan = jr.split(key, 2)
    u0 = jr.normal(key_init, (Np, unobs.size, 1))
    u0[bs_star[0]] = us_star[0]
    np.testing.assert_allclose(uss[0], u0, rtol=0, atol=5e-7)
    for k, step_key in enumerate(jr.split(key_scan, K)):
        key_res, key_tr = jr.split(step_key, 2)
        A = ocr.killing(key_res, np.exp(log_wss[k]).astype(np.float32), bs_star[k], bs_star[k + 1], True)
        np.testing.assert_array_equal(As[k], A)
        parents = uss[k][As[k]]
        want_us = om.transition_sampler(parents, vs[k], ts[k], key_tr)
        want_us[bs_star[k + 1]] = us_star[k + 1]
        np.testing.assert_array_equal(uss[k + 1][bs_star[k + 1]], us_star[k + 1])
        np.testing.assert_allclose(uss[k + 1], want_us, rtol=0, atol=3e-2 * max(1.0, float(np.abs(want_us).max())))
        want = ocsmc.normalise(om.likelihood_logpdf(vs[k + 1], parents, vs[k], ts[k]).astype(np.float32), log_space=True)
        np.testing.assert_allclose(log_wss[k + 1], want, rtol=0, atol=5e-2 * max(1.0, float(np.ptp(want))))
    # the Gibbs kernel of supr.py:171-176
    x0 = rng.uniform(size=(unobs.size, 1)).astype(np.float32)
    y0 = rng.uniform(size=(obs.size, 1)).astype(np.float32)
    gkey = jr.PRNGKey(21)
    x0n, us_next, bs_next, changed = gibbs_kernel(gkey, x0, y0, None, bs_star, ts, model.fwd_sampler, sde, model.unpack, N,
                                                  model.transition_sampler, model.transition_logpdf, model.likelihood_logpdf,
                                                  explicit_backward=True, explicit_final=True)
    assert x0n.shape == (unobs.size,) and us_next.shape == (K + 1, unobs.size) and np.isfinite(us_next).all()
    np.testing.assert_array_equal(us_next[-1], x0n)
    kc = jr.split(jr.split(gkey, 3)[1], 4)
    np.testing.assert_array_equal(bs_next, jr.randint(kc[3], (K + 1,), 0, N))
    # us_star_next is the reversed forward path started at the selected x0 (gibbs.py:155)
    want_path = model.fwd_sampler(kc[2], x0n.reshape(unobs.size, 1), y0).cpu().numpy()
    np.testing.assert_array_equal(us_next, want_path[::-1].reshape(K + 1, 784)[:, unobs])
