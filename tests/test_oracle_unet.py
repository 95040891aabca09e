"""CPU checks of the oracle restatement of the score network (oracle/unet.py) and of the product's host-side weight
preparation (no GPU): the one pinned piece (PixelShuffle, /root/reference/tests/test_nns.py:7-16), the parameter census of
SURVEY App. D, and the algebra the tensor-core path relies on (weight packing, stride-2 -> space-to-depth 2x2)."""
import math
import numpy as np
import torch
import torch.nn.functional as F
from oracle import unet as ou


def test_pixel_shuffle_matches_torch_as_in_reference_test():
    # the reference's own test: torch.nn.PixelShuffle(2) on NCHW == PixelShuffle(2) on NHWC (tests/test_nns.py:7-16)
    torch.manual_seed(666)
    img = torch.randn(3, 4, 2, 2)
    want = torch.nn.PixelShuffle(2)(img).permute(0, 2, 3, 1)
    got = ou.pixel_shuffle(img.permute(0, 2, 3, 1).contiguous(), 2)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    # and the channel order 'b h w (h2 w2 c)' itself (fbs/nn/utils.py:53-57) with c > 1
    x = torch.arange(2 * 3 * 8, dtype=torch.float32).reshape(1, 2, 3, 8)
    y = ou.pixel_shuffle(x, 2)
    for h2 in range(2):
        for w2 in range(2):
            for c in range(2):
                assert y[0, 2 * 1 + h2, 2 * 2 + w2, c] == x[0, 1, 2, (h2 * 2 + w2) * 2 + c]


def test_parameter_census_matches_survey():
    shapes = ou.unet_param_shapes(1)
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n - 12.99e6) < 0.01e6, n                     # 12.99 M parameters
    from fbs_b200.nn.unet import unet_param_shapes
    assert unet_param_shapes(3) == ou.unet_param_shapes(3)  # the parameter names / shapes are the interface


def test_forward_shapes_and_time_dependence():
    params = ou.init_unet_params(0, 1)
    x = np.random.default_rng(0).standard_normal((2, 28, 28, 1)).astype(np.float32)
    a = ou.unet_forward(params, x, 0.1, 0.01)
    b = ou.unet_forward(params, x, 1.7, 0.01)
    assert a.shape == x.shape and np.isfinite(a).all()
    assert np.abs(a - b).max() > 1e-3                        # the time embedding reaches the output
    one = ou.unet_forward(params, x[:1], 0.1, 0.01)
    np.testing.assert_allclose(one[0], a[0], rtol=1e-4, atol=1e-5)   # samples are independent


def test_weight_packing_algebra():
    from fbs_b200.nn.unet import _pack, _pack_stride2, _standardize
    rng = np.random.default_rng(1)
    # K ordering (ty, tx, channel): the GEMM with im2col rows reproduces the convolution
    x = torch.from_numpy(rng.standard_normal((2, 6, 5, 8)).astype(np.float32))
    w = rng.standard_normal((3, 3, 8, 4)).astype(np.float32)
    want = ou._conv(ou._P({'c.kernel': w}), 'c', x, bias=False)
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    cols = torch.stack([xp[:, ty:ty + 6, tx:tx + 5, :] for ty in range(3) for tx in range(3)], dim=3).reshape(2, 6, 5, 72)
    got = cols @ torch.from_numpy(_pack(w)).t()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    # 4x4 stride-2 padding-1 convolution == 2x2 convolution over the shifted space-to-depth copy
    x = torch.from_numpy(rng.standard_normal((2, 8, 6, 4)).astype(np.float32))
    w = rng.standard_normal((4, 4, 4, 3)).astype(np.float32)
    want = ou._conv(ou._P({'c.kernel': w}), 'c', x, stride=2, padding=1, bias=False)
    Ho, Wo = 8 // 2 + 1, 6 // 2 + 1
    s2d = torch.zeros(2, Ho, Wo, 4, 4)
    for i in range(Ho):
        for j in range(Wo):
            for r in range(2):
                for s in range(2):
                    hh, ww = 2 * i - 1 + r, 2 * j - 1 + s
                    if 0 <= hh < 8 and 0 <= ww < 6:
                        s2d[:, i, j, 2 * r + s] = x[:, hh, ww]
    s2d = s2d.reshape(2, Ho, Wo, 16)
    cols = torch.stack([s2d[:, ty:ty + 4, tx:tx + 3, :] for ty in range(2) for tx in range(2)], dim=3).reshape(2, 4, 3, 64)
    got = cols @ torch.from_numpy(_pack_stride2(w)).t()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    # weight standardisation: float32 restatement on both sides
    np.testing.assert_allclose(_standardize(w), ou.standardize_kernel(torch.from_numpy(w)).numpy(), rtol=1e-5, atol=1e-6)
