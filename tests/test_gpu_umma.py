"""tcgen05 plumbing: the split-TF32 UMMA GEMM (descriptors, K-major core-matrix layout, TMEM read-back) against
float64 numpy.  hi*hi + lo*hi + hi*lo with tf32-rounded hi keeps float32-level accuracy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('du,dv', [(8, 8), (100, 100), (16, 40), (124, 124), (12, 4)])
def test_umma_split_tf32_gemm(du, dv):
    import torch
    from fbs_b200 import _native as nat
    from fbs_b200.models import pack_umma_image
    from fbs_b200._tensor import ptr, stream
    rng = np.random.default_rng(du * 1000 + dv)
    D = du + dv
    du8, dv8 = (du + 7) // 8 * 8, (dv + 7) // 8 * 8
    nout = du8 + dv8 + (8 if (du8 + dv8) % 16 else 0)
    Mu = rng.normal(size=(1, D, du)).astype(np.float32)              # rows: outputs, cols: u inputs
    img = pack_umma_image(Mu, du, dv)
    assert img.shape == (1, du8 // 8, 2, 2, nout // 8, 8, 4)
    A = np.zeros((128, du8), np.float32)
    A[:, :du] = (rng.normal(size=(128, du)) * 3).astype(np.float32)
    dev = torch.device('cuda')
    A_d, B_d = torch.from_numpy(A).to(dev), torch.from_numpy(img).to(dev)
    out = torch.full((128, nout), float('nan'), device=dev)
    nat.call('fbs_debug_umma_gemm', stream(), ptr(A_d), ptr(B_d), du8, nout, ptr(out))
    got = out.cpu().numpy()
    want = A[:, :du].astype(np.float64) @ Mu[0].astype(np.float64).T  # [128, D]
    scale = np.abs(A[:, :du]).astype(np.float64) @ np.abs(Mu[0]).astype(np.float64).T
    err_u = np.abs(got[:, :du] - want[:, :du]) / scale[:, :du]
    err_v = np.abs(got[:, du8:du8 + dv] - want[:, du:]) / scale[:, du:]
    # float32 sequential dot products of this length sit at ~1e-7 relative to sum |a||b|; plain TF32 would be ~5e-4
    assert max(err_u.max(), err_v.max()) < 1.5e-6, (err_u.max(), err_v.max())
    assert np.all(got[:, du:du8] == 0) and np.all(got[:, du8 + dv:] == 0)   # padded outputs are exact zeros
