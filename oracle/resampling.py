"""Unconditional resamplers ``(weights, key) -> indices``.  TEST INFRASTRUCTURE.

Restates ``/root/reference/fbs/samplers/resampling.py``:
``_sorted_uniforms`` :36-40, ``_systematic_or_stratified`` :43-51, ``systematic`` :54-55,
``stratified`` :58-59, ``multinomial`` :62-68, ``killing`` :71-101.
"""
import numpy as np
from . import jax_random as jr


def _sorted_uniforms(n, key):
    # resampling.py:36-40
    us = jr.uniform(key, (n + 1,))
    z = jr.seq_cumsum(-np.log(us).astype(np.float32))
    return (z[:-1] / z[-1]).astype(np.float32)


def _systematic_or_stratified(weights, key, is_systematic):
    # resampling.py:43-51
    weights = np.asarray(weights, dtype=np.float32)
    n = weights.shape[0]
    u = jr.uniform(key, ()) if is_systematic else jr.uniform(key, (n,))
    pts = ((np.arange(n, dtype=np.float32) + u) / np.float32(n)).astype(np.float32)
    idx = np.searchsorted(jr.seq_cumsum(weights), pts, side='left')
    return np.clip(idx, 0, n - 1).astype(np.int32)


def systematic(weights, key):
    return _systematic_or_stratified(weights, key, True)


def stratified(weights, key):
    return _systematic_or_stratified(weights, key, False)


def multinomial(weights, key):
    # resampling.py:62-68 ("Not tested." upstream)
    weights = np.asarray(weights, dtype=np.float32)
    n = weights.shape[0]
    idx = np.searchsorted(jr.seq_cumsum(weights), _sorted_uniforms(n, key), side='left')
    return np.clip(idx, 0, n - 1).astype(np.int32)


def killing(weights, key):
    # resampling.py:71-101
    weights = np.asarray(weights, dtype=np.float32)
    key_1, key_2, _ = jr.split(key, 3)
    n = weights.shape[0]
    w_max = weights.max()
    killed = (jr.uniform(key_1, (n,)) * w_max) >= weights
    idx = np.arange(n, dtype=np.int32)
    return np.where(~killed, idx, jr.choice(key_2, n, (n,), p=weights)).astype(np.int32)
