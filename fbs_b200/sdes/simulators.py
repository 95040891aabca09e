"""Euler--Maruyama for affine drifts (``fbs/sdes/simulators.py:53-106``) on the GPU.

The reference takes an arbitrary ``drift(x, t)`` closure.  A fused kernel cannot call Python, so
the drift must expose ``drift.affine(t) -> (A, a)`` (``AffineDrift`` or the drift returned by
``make_gaussian_bw_sb``); anything else raises -- there is no interpreted fallback.
Kernel: ``fbs_em_affine_path_f32`` (fbs_b200/csrc/sde_kernels.cu).
"""
import numpy as np
import torch
from .. import _native as nat
from .._tensor import dev, empty, ptr, stream, out, is_host


class AffineDrift:
    """``drift(x, t) = A(t) x + a(t)`` given as a Python callable ``affine(t) -> (A [D,D], a [D])``."""

    def __init__(self, affine):
        self.affine = affine

    def __call__(self, x, t):
        A, a = self.affine(t)
        return np.asarray(x) @ np.asarray(A).T + np.asarray(a)


def em_tables(ts, drift, dispersion, integration_nsteps):
    """Host precompute of the per-sub-step tables the kernel streams (float32 time grid as in the reference)."""
    if not hasattr(drift, 'affine'):
        raise TypeError('euler_maruyama: the drift must expose .affine(t) -> (A, a); opaque closures cannot be fused '
                        'and fbs_b200 has no interpreted fallback')
    ts32 = np.asarray(ts.detach().cpu().numpy() if isinstance(ts, torch.Tensor) else ts, dtype=np.float32)
    K, m = ts32.shape[0] - 1, int(integration_nsteps)
    AT, avec, disp, ddts = [], [], [], []
    for k in range(K):
        t, t_next = ts32[k], ts32[k + 1]
        ddt = np.float32(np.abs(t_next - t) / np.float32(m))            # simulators.py:90
        grid = np.linspace(t, t_next - ddt, m, dtype=np.float32)        # simulators.py:92
        ddts.append(ddt)
        for t_ in grid:
            A, a = drift.affine(float(t_))
            AT.append(np.asarray(A, dtype=np.float64).T)
            avec.append(np.asarray(a, dtype=np.float64))
            disp.append(float(dispersion(float(t_))))
    return (np.stack(AT).astype(np.float32), np.stack(avec).astype(np.float32), np.asarray(ddts, dtype=np.float32),
            np.asarray(disp, dtype=np.float32), K, m)


def em_path(key, x0, tables, du=None, rev=False):
    AT, avec, ddt, disp, K, m = tables
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B = k.shape[0]
    x = dev(x0, torch.float32)
    batched = x.dim() == 2
    D = x.shape[-1]
    args = [dev(a, torch.float32) for a in (AT, avec, ddt, disp)]
    if not rev:
        path = empty((B, K + 1, D), torch.float32)
        nat.call('fbs_em_affine_path_f32', stream(), ptr(k), ptr(x), int(batched), *[ptr(a) for a in args], B, K, m, D,
                 D, 0, ptr(path), None)
        return out(path[0] if single else path, host)
    us = empty((B, K + 1, du), torch.float32)
    vs = empty((B, K + 1, D - du), torch.float32)
    nat.call('fbs_em_affine_path_f32', stream(), ptr(k), ptr(x), int(batched), *[ptr(a) for a in args], B, K, m, D,
             int(du), 1, ptr(us), ptr(vs))
    if single:
        us, vs = us[0], vs[0]
    return out(us, host), out(vs, host)


def euler_maruyama(key, x0, ts, drift, dispersion, integration_nsteps: int = 1, return_path: bool = False):
    """``fbs.sdes.euler_maruyama`` for affine drifts; same argument meaning and return values."""
    tables = em_tables(ts, drift, dispersion, integration_nsteps)
    path = em_path(key, x0, tables)
    if return_path:
        return path
    return path[..., -1, :]
