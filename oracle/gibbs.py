"""Particle-Gibbs kernel.  TEST INFRASTRUCTURE.

Restates ``/root/reference/fbs/samplers/gibbs.py``: ``gibbs_init`` :23-65, ``gibbs_kernel``
:68-168, ``force_move`` :171-214.  ``marg_y=True`` (Doob bridge, "Not used in our paper",
gibbs.py:115) is not restated.
"""
import math
import numpy as np
from . import jax_random as jr
from .csmc import forward_pass as csmc_fwd, csmc_kernel
from .cond_resampling import killing
from .resampling import stratified
from .smc import bootstrap_filter, bootstrap_backward_smoother


def force_move(key, weights, k):
    weights = np.asarray(weights)
    f = weights.dtype.type
    M = weights.shape[0]
    key_1, key_2 = jr.split(key, 2)                                                  # :197
    w_k = weights[k]
    temp = f(1) - w_k
    rest = weights.copy(); rest[k] = 0                                               # :202
    threshold = np.maximum(f(1) - np.exp(f(-M)), f(1 - 1e-12))                       # :203 (== 1.0 in fp32)
    rest = (rest / temp).astype(weights.dtype) if w_k < threshold else np.full((M,), f(1) / f(M), dtype=weights.dtype)
    i = int(jr.choice(key_1, M, (), p=rest))                                         # :207
    u = jr.uniform(key_2, ()) if weights.dtype == np.float32 else jr.uniform64(key_2, ())
    accept = u * (f(1) - weights[i]) < temp                                          # :209
    with np.errstate(divide='ignore', invalid='ignore'):
        alpha = np.nansum(temp * rest / (f(1) - weights))                            # :211
    return (i if accept else int(k)), float(np.clip(alpha, 0, 1.))


def gibbs_kernel(key, x0, y0, us_star, bs_star, ts, fwd_sampler, sde, unpack, nparticles,
                 transition_sampler, transition_logpdf, likelihood_logpdf,
                 marg_y=False, explicit_backward=True, explicit_final=False, return_aux=False, **kwargs):
    if marg_y:
        raise NotImplementedError('marg_y=True is not part of the restated path')
    key_fwd, key_csmc, key_bridge = jr.split(key, 3)                                 # :126
    path_xy = fwd_sampler(key_fwd, x0, y0, **kwargs)                                 # :127
    path_x, path_y = unpack(path_xy, **kwargs)
    us = path_x[::-1]
    vs = path_y[::-1]
    dtype = us.dtype

    if explicit_final:                                                               # :132-137
        def init_sampler(key_, n_samples):
            shape = (n_samples, *us.shape[1:])
            return jr.normal(key_, shape) if dtype == np.float32 else jr.normal64(key_, shape)

        def init_likelihood_logpdf(v0, u0s, v1, **kw):
            return likelihood_logpdf(v0, u0s, v1, ts[0], **kw)
    else:                                                                            # :139-144
        def init_sampler(*_):
            return us[0] * np.ones((nparticles, *us.shape[1:]), dtype=dtype)

        def init_likelihood_logpdf(*_, **__):
            return (-math.log(nparticles) * np.ones(nparticles)).astype(dtype)

    aux = {}
    if explicit_backward:
        key_csmc_fwd, key_csmc_x0, key_csmc_bwd_us, key_csmc_bwd_bs = jr.split(key_csmc, 4)   # :147
        As, log_ws, uss = csmc_fwd(key_csmc_fwd, us, bs_star, vs, ts, init_sampler, init_likelihood_logpdf,
                                   transition_sampler, likelihood_logpdf, killing, nparticles, **kwargs)
        idx, _ = force_move(key_csmc_x0, np.exp(log_ws[-1]).astype(dtype), int(bs_star[-1]))  # :152
        x0 = uss[-1, idx]                                                            # :154
        us_star_next = unpack(fwd_sampler(key_csmc_bwd_us, x0, y0, **kwargs), **kwargs)[0][::-1]  # :155
        bs_star_next = jr.randint(key_csmc_bwd_bs, (us.shape[0],), 0, nparticles)    # :156
        aux = dict(As=As, log_ws=log_ws, uss=uss, idx=idx, us=us, vs=vs)
    else:
        us_star_next, bs_star_next = csmc_kernel(key_csmc, us, bs_star, vs, ts, init_sampler, init_likelihood_logpdf,
                                                 transition_sampler, transition_logpdf, likelihood_logpdf,
                                                 killing, nparticles, backward=False, **kwargs)   # :158-166
    x0_next = us_star_next[-1]                                                       # :167
    out = (x0_next, us_star_next, bs_star_next, bs_star_next != np.asarray(bs_star))
    return out + (aux,) if return_aux else out


def gibbs_init(key, y0, x0_shape, ts, fwd_sampler, sde, unpack, transition_sampler, transition_logpdf,
               likelihood_logpdf, nparticles, method='smoother', marg_y=False, x0=None, **kwargs):
    """gibbs.py:23-65 with ``marg_y=False`` (the bridge branch is not restated)."""
    if marg_y:
        raise NotImplementedError('marg_y=True is not part of the restated path')
    y0 = np.asarray(y0)
    dtype = y0.dtype
    if x0 is None:
        x0 = np.zeros(x0_shape, dtype=dtype)
    key_fwd, key_bridge, key_u0, key_bf, key_fwd2, key_bwd = jr.split(key, 6)        # :39
    path_xy = fwd_sampler(key_fwd, x0, y0, **kwargs)
    _, path_y = unpack(path_xy, **kwargs)
    vs = path_y[::-1]

    def init_sampler(*_):                                                            # :46-48
        shape = (nparticles, *x0_shape)
        return jr.normal(key_u0, shape) if dtype == np.float32 else jr.normal64(key_u0, shape)

    if method == 'filter':
        approx_x0 = bootstrap_filter(transition_sampler, likelihood_logpdf, vs, ts, init_sampler, key_bf, nparticles,
                                     stratified, log=True, return_last=True, **kwargs)[0][0]
        approx_us_star = unpack(fwd_sampler(key_fwd2, approx_x0, y0, **kwargs), **kwargs)[0][::-1]
    elif method == 'smoother':
        uss = bootstrap_filter(transition_sampler, likelihood_logpdf, vs, ts, init_sampler, key_bf, nparticles,
                               stratified, log=True, return_last=False, **kwargs)[0]
        approx_x0 = uss[-1, 0]
        approx_us_star = bootstrap_backward_smoother(key_bwd, uss, vs, ts, transition_logpdf, **kwargs)
    else:
        raise ValueError(f"Unknown method {method}")
    return approx_x0, approx_us_star
