"""Conditional resamplers ``(key, weights, i, j, conditional) -> indices`` -- API of
``fbs/samplers/csmc/resamplings.py:10-125``.

weights ``(N,)`` or ``(B, N)``; ``i``/``j`` Python ints or ``(B,)`` arrays.  Kernel:
``fbs_cond_resample_f32`` (fbs_b200/csrc/resample_kernels.cu).  Like upstream
(resamplings.py:129), conditional systematic resampling raises ``NotImplementedError``.
"""
import numpy as np
import torch
from ... import _native as nat
from ..._tensor import dev, empty, ptr, stream, out, is_host


def _run(scheme, key, weights, i, j, conditional):
    host = is_host(weights)
    w = dev(weights, torch.float32)
    k = dev(key, torch.uint32)
    single = w.dim() == 1
    w = w.reshape(-1, w.shape[-1])
    k = k.reshape(-1, 2)
    B, N = w.shape
    if k.shape[0] != B:
        raise ValueError('one key per weight vector is required')

    def ivec(x):
        if isinstance(x, torch.Tensor):
            t = dev(x, torch.int32).reshape(-1)
        else:
            t = dev(np.asarray(x, dtype=np.int32).reshape(-1), torch.int32)
        return t.expand(B).contiguous() if t.shape[0] == 1 and B > 1 else t

    iv, jv = (ivec(i), ivec(j)) if conditional else (None, None)
    idx = empty((B, N), torch.int32)
    nat.call('fbs_cond_resample_f32', stream(), scheme, ptr(k), ptr(w), ptr(iv), ptr(jv), int(bool(conditional)), B, N,
             ptr(idx))
    return out(idx[0] if single else idx, host)


def multinomial(key, weights, i=0, j=0, conditional=True):
    return _run(nat.RESAMPLE_MULTINOMIAL, key, weights, i, j, conditional)


def killing(key, weights, i=0, j=0, conditional=True):
    return _run(nat.RESAMPLE_KILLING, key, weights, i, j, conditional)


def systematic(key, weights, i=0, j=0, conditional=True):
    if conditional:
        raise NotImplementedError('Not implemented, not used.')
    return _run(nat.RESAMPLE_SYSTEMATIC, key, weights, i, j, False)


multinomial.scheme, killing.scheme, systematic.scheme = (nat.RESAMPLE_MULTINOMIAL, nat.RESAMPLE_KILLING,
                                                         nat.RESAMPLE_SYSTEMATIC)
for _f in (multinomial, killing, systematic):
    _f.family = 'conditional'
