"""Unconditional resamplers ``(weights, key) -> indices`` -- API of ``fbs/samplers/resampling.py:43-101``.

weights ``(N,)`` or ``(B, N)`` (with keys ``(B, 2)``).  Kernel: ``fbs_resample_f32``
(fbs_b200/csrc/resample_kernels.cu); cumulative sums are sequential float32.
"""
import torch
from .. import _native as nat
from .._tensor import dev, empty, ptr, stream, out, is_host


def _run(scheme, weights, key):
    host = is_host(weights)
    w = dev(weights, torch.float32)
    k = dev(key, torch.uint32)
    single = w.dim() == 1
    w = w.reshape(-1, w.shape[-1])
    k = k.reshape(-1, 2)
    if k.shape[0] != w.shape[0]:
        raise ValueError('one key per weight vector is required')
    idx = empty(w.shape, torch.int32)
    nat.call('fbs_resample_f32', stream(), scheme, ptr(k), ptr(w), w.shape[0], w.shape[1], ptr(idx))
    return out(idx[0] if single else idx, host)


def systematic(weights, key):
    return _run(nat.RESAMPLE_SYSTEMATIC, weights, key)


def stratified(weights, key):
    return _run(nat.RESAMPLE_STRATIFIED, weights, key)


def multinomial(weights, key):
    """Sorted-uniform multinomial ("Not tested." upstream, resampling.py:62-68)."""
    return _run(nat.RESAMPLE_MULTINOMIAL, weights, key)


def killing(weights, key):
    return _run(nat.RESAMPLE_KILLING, weights, key)


systematic.scheme, stratified.scheme = nat.RESAMPLE_SYSTEMATIC, nat.RESAMPLE_STRATIFIED
multinomial.scheme, killing.scheme = nat.RESAMPLE_MULTINOMIAL, nat.RESAMPLE_KILLING
for _f in (systematic, stratified, multinomial, killing):
    _f.family = 'unconditional'
