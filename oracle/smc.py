"""Bootstrap filter / smoother and the pseudo-marginal kernel.  TEST INFRASTRUCTURE.

Restates ``/root/reference/fbs/samplers/smc.py``: ``bootstrap_filter`` :9-88,
``bootstrap_backward_smoother`` :91-112, ``pmcmc_filter_step`` :115-158, ``pcn_proposal``
:161-168, ``pmcmc_kernel`` :171-258; ``MCMCState`` from ``fbs/samplers/common.py:5-9``.
"""
import math
from typing import NamedTuple
import numpy as np
from . import jax_random as jr
from .csmc import logsumexp


class MCMCState(NamedTuple):
    acceptance_prob: float
    is_accepted: bool
    prop_log_ell: float
    log_ell: float


def bootstrap_filter(transition_sampler, measurement_cond_pdf, vs, ts, init_sampler, key, nparticles, resampling,
                     log=True, return_last=True, **kwargs):
    nsteps = vs.shape[0] - 1
    key_init, key_steps = jr.split(key)                                              # :77
    us = np.asarray(init_sampler(key_init, vs[0], nparticles))                       # :78
    keys = jr.split(key_steps, nsteps)
    dtype = us.dtype
    log_nell = dtype.type(0.)
    hist = [us]
    for k in range(nsteps):
        v, v_prev, t_prev = vs[k + 1], vs[k], ts[k]
        key_proposal, key_resampling = jr.split(keys[k])                             # :61
        us_new = transition_sampler(us, v_prev, t_prev, key_proposal, **kwargs)      # :63
        log_w = measurement_cond_pdf(v, us, v_prev, t_prev, **kwargs)                # :65
        _c = logsumexp(log_w)
        log_nell = dtype.type(log_nell - (_c - dtype.type(math.log(nparticles))))    # :67
        inds = resampling(np.exp(log_w - _c).astype(dtype), key_resampling)          # :68-69
        us = us_new[inds, ...]                                                       # :72
        hist.append(us)
    if return_last:
        return us, log_nell
    return np.stack(hist), log_nell


def bootstrap_backward_smoother(key, filter_us, vs, ts, transition_logpdf, *args, **kwargs):
    nsteps = filter_us.shape[0] - 1
    key_last, key_smoother = jr.split(key, 2)                                        # :108
    n = filter_us.shape[1]
    uT = filter_us[-1][int(jr.choice(key, n, ()))]                                   # :109 (unsplit key, as written)
    keys = jr.split(key_smoother, nsteps)
    u = uT
    traj = []
    for q, k in enumerate(range(nsteps - 1, -1, -1)):                                # :110-111
        log_ws = transition_logpdf(u, filter_us[k], vs[k], ts[k], *args, **kwargs)   # :102
        log_ws = log_ws - logsumexp(log_ws)
        u = filter_us[k][int(jr.choice(keys[q], n, (), p=np.exp(log_ws).astype(log_ws.dtype)))]  # :104
        traj.append(u)
    return np.concatenate([np.stack(traj[::-1]), uT[None]], axis=0)


def pmcmc_filter_step(key, vs_bridge, u0s, ts, transition_sampler, likelihood_logpdf, resampling, nparticles,
                      return_hist=False, **kwargs):
    nsteps = ts.shape[0] - 1
    keys = jr.split(key, nsteps)                                                     # :154
    us = np.asarray(u0s)
    dtype = us.dtype
    log_ell = dtype.type(0.)
    hist = []
    for k in range(nsteps):
        v, v_prev, t_prev = vs_bridge[k + 1], vs_bridge[k], ts[k]
        key_proposal, key_resampling = jr.split(keys[k])                             # :142
        log_ws = likelihood_logpdf(v, us, v_prev, t_prev, **kwargs)                  # :144
        _c = logsumexp(log_ws)
        log_ell = dtype.type(dtype.type(log_ell - dtype.type(math.log(nparticles))) + _c)   # :146
        inds = resampling(np.exp(log_ws - _c).astype(dtype), key_resampling)         # :147-148
        us_prev = us[inds, ...]                                                      # :149
        us = np.asarray(transition_sampler(us_prev, v_prev, t_prev, key_proposal, **kwargs))  # :150
        if return_hist:
            hist.append(dict(log_ws=log_ws, inds=inds, us=us, log_ell=log_ell))
    if return_hist:
        return us, log_ell, hist
    return us, log_ell


def pcn_proposal(key, delta, x, mean, sampler):
    beta = 2 / (2 + delta)                                                           # :164
    key_rnds = jr.split(key, 2)
    r0, r1 = sampler(key_rnds[0]), sampler(key_rnds[1])                              # :166
    p = x + math.sqrt(delta / 2) * (r0 - mean)
    return (beta * p + (1 - beta) * mean + math.sqrt(1 - beta) * (r1 - mean)).astype(x.dtype)   # :167-168


def pmcmc_kernel(key, uT, log_ell, ys, y0, ts, fwd_ys_sampler, sde, ref_sampler, transition_sampler,
                 likelihood_logpdf, resampling, nparticles, delta=None, which_u=0, **kwargs):
    key_prop, key_u0, key_filter, key_mh = jr.split(key, 4)                          # :231
    if delta is None:
        prop_ys = fwd_ys_sampler(key_prop, y0)
    else:
        mean = np.stack([sde.mean(t, ts[0], y0) for t in ts]).astype(ys.dtype)       # :236
        prop_ys = pcn_proposal(key_prop, delta, ys, mean, lambda key_: fwd_ys_sampler(key_, y0))
    vs = prop_ys[::-1]
    u0s = ref_sampler(key_u0, vs[0], nparticles)                                     # :241
    prop_uTs, prop_log_ell = pmcmc_filter_step(key_filter, vs, u0s, ts, transition_sampler, likelihood_logpdf,
                                               resampling, nparticles, **kwargs)
    prop_uT = prop_uTs[which_u]
    dtype = prop_uTs.dtype
    log_acc_prob = np.minimum(dtype.type(0.), dtype.type(prop_log_ell - log_ell))    # :246
    z = jr.uniform(key_mh, ()) if dtype == np.float32 else jr.uniform64(key_mh, ())  # :248
    with np.errstate(divide='ignore'):
        acc_flag = bool(np.log(z) < log_acc_prob)                                    # :249
    state = MCMCState(acceptance_prob=float(np.exp(log_acc_prob)), is_accepted=acc_flag,
                      prop_log_ell=float(prop_log_ell), log_ell=float(log_ell))
    if acc_flag:
        return prop_uT, prop_log_ell, prop_ys, state                                 # :255-258
    return uT, log_ell, ys, state


def twisted_smc(key, y, ts, init_sampler, transition_logpdf, twisting_logpdf, twisting_prop_sampler, twisting_prop_logpdf,
                resampling, nparticles, return_hist=False, **kwargs):
    """smc.py:261-309 (Algorithm 1 of arXiv 2306.17775).  The scan walks ``ts[1:]`` (:305), so "t_prev" below is ts[k + 1]."""
    nsteps = ts.shape[0] - 1
    key_init, key_filter = jr.split(key, 2)                                          # :296
    keys = jr.split(key_filter, nsteps)
    xs = np.asarray(init_sampler(key_init, nparticles))                              # :299
    log_ps = twisting_logpdf(y, xs, ts[0], **kwargs)                                 # :300
    log_ws = (log_ps - logsumexp(log_ps)).astype(log_ps.dtype)                       # :301
    hist = []
    for k in range(nsteps):
        t_prev = ts[k + 1]
        key_resampling, key_prop = jr.split(keys[k])                                 # :280
        inds = resampling(np.exp(log_ws).astype(log_ws.dtype), key_resampling)       # :283
        xs_prev = xs[inds, ...]                                                      # :284
        log_ps_prev = log_ps[inds, ...]                                              # :285
        xs = np.asarray(twisting_prop_sampler(key_prop, xs_prev, t_prev, y, **kwargs))   # :288
        log_ps = twisting_logpdf(y, xs, t_prev, **kwargs)                            # :291
        log_ws = (transition_logpdf(xs, xs_prev, t_prev) + log_ps
                  - twisting_prop_logpdf(xs, xs_prev, t_prev, y, **kwargs) - log_ps_prev)   # :292-293
        log_ws = (log_ws - logsumexp(log_ws)).astype(log_ps.dtype)                   # :294
        if return_hist:
            hist.append(dict(inds=inds, xs=xs, log_ws=log_ws))
    return (xs, log_ws, hist) if return_hist else (xs, log_ws)
