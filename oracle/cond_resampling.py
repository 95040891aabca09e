"""Conditional resamplers ``(key, weights, i, j, conditional) -> indices``.  TEST INFRASTRUCTURE.

Restates ``/root/reference/fbs/samplers/csmc/resamplings.py``: ``multinomial`` :10-37,
``killing`` :40-88 (the one the Gibbs kernel hard-codes, gibbs.py:149,164), ``systematic``
:91-125.  The conditional systematic branch raises upstream (:129) and raises here.
"""
import numpy as np
from . import jax_random as jr


def multinomial(key, weights, i=0, j=0, conditional=True):
    weights = np.asarray(weights, dtype=np.float32)
    n = weights.shape[0]
    idx = jr.choice(key, n, (n,), p=weights)           # resamplings.py:34
    if conditional:
        idx = idx.copy()
        idx[j] = i                                     # :36
    return idx.astype(np.int32)


def killing(key, weights, i=0, j=0, conditional=True, return_parts=False):
    weights = np.asarray(weights, dtype=np.float32)
    key_1, key_2, key_3 = jr.split(key, 3)             # :66
    n = weights.shape[0]
    w_max = weights.max()                              # :69
    u1 = jr.uniform(key_1, (n,))
    killed = (u1 * w_max) >= weights                   # :71
    idx = np.arange(n, dtype=np.int32)
    ch = jr.choice(key_2, n, (n,), p=weights)
    idx = np.where(~killed, idx, ch).astype(np.int32)  # :73-74
    if not conditional:
        return idx
    f = np.float32
    j_prob = ((f(1.) - weights / w_max) / f(n)).astype(np.float32)   # :79
    j_prob[i] = f(0.)                                                # :80
    j_prob_i = np.maximum(f(1.) - jr.seq_sum(j_prob), f(0.))         # :81
    j_prob[i] = j_prob_i                                             # :82
    J = int(jr.choice(key_3, n, (), p=j_prob))                       # :84
    idx = np.roll(idx, j - J)                                        # :85
    idx[j] = i                                                       # :86
    if return_parts:
        return idx.astype(np.int32), dict(J=J, killed=killed, choice=ch, j_prob=j_prob)
    return idx.astype(np.int32)


def _standard_systematic(key, weights):
    # :120-125 -- note: no clip, arange in default int then promoted to float32
    weights = np.asarray(weights, dtype=np.float32)
    n = weights.shape[0]
    u = ((np.arange(n).astype(np.float32) + jr.uniform(key, ())) / np.float32(n)).astype(np.float32)
    return np.searchsorted(jr.seq_cumsum(weights), u, side='left').astype(np.int32)


def systematic(key, weights, i=0, j=0, conditional=True):
    if conditional:
        raise NotImplementedError('Not implemented, not used.')      # :129
    return _standard_systematic(key, weights)
