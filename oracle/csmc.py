"""Conditional SMC kernel.  TEST INFRASTRUCTURE.

Restates ``/root/reference/fbs/samplers/csmc/csmc.py``: ``csmc_kernel`` :14-77,
``forward_pass`` :80-164 (step body :132-148), ``backward_sampling_pass`` :167-227,
``backward_scanning_pass`` :230-270, ``normalise`` :273-292, ``barker_move`` :295-297.
"""
import numpy as np
from . import jax_random as jr


def logsumexp(a):
    """``jax.scipy.special.logsumexp`` for a 1-D array (max-shifted, sequential sum)."""
    a = np.asarray(a)
    amax = a.max()
    if not np.isfinite(amax):
        amax = a.dtype.type(0)
    return (np.log(jr.seq_sum(np.exp(a - amax).astype(a.dtype))) + amax).astype(a.dtype)


def normalise(log_weights, log_space=False):
    log_weights = (log_weights - logsumexp(log_weights)).astype(log_weights.dtype)   # :289
    return log_weights if log_space else np.exp(log_weights).astype(log_weights.dtype)


def barker_move(key, ws):
    return int(jr.choice(key, ws.shape[0], (), p=ws))                                # :295-297


def forward_pass(key, us_star, bs_star, vs, ts, init_sampler, init_likelihood_logpdf,
                 transition_sampler, likelihood_logpdf, cond_resampling, nsamples, **kwargs):
    nsteps = us_star.shape[0] - 1
    key_init, key_scan = jr.split(key, 2)                                            # :150
    us0 = np.array(init_sampler(key_init, nsamples + 1))                             # :151
    us0[bs_star[0]] = us_star[0]                                                     # :152
    log_ws0 = normalise(init_likelihood_logpdf(vs[0], us0, vs[1], **kwargs), log_space=True)  # :154-155
    keys = jr.split(key_scan, nsteps)                                                # :157

    log_ws, us_prev = log_ws0, us0
    log_wss, As, uss = [log_ws0], [], [us0]
    for k in range(nsteps):
        v, v_prev, t_prev = vs[k + 1], vs[k], ts[k]
        b_star_prev, b_star, u_star = int(bs_star[k]), int(bs_star[k + 1]), us_star[k + 1]
        key_resampling, key_transition = jr.split(keys[k], 2)                        # :136
        A = cond_resampling(key_resampling, np.exp(log_ws).astype(log_ws.dtype), b_star_prev, b_star, True)  # :139
        us_prev = np.take(us_prev, A, axis=0)                                        # :140
        us = np.array(transition_sampler(us_prev, v_prev, t_prev, key_transition, **kwargs))  # :142
        us[b_star] = u_star                                                          # :143
        log_ws = normalise(likelihood_logpdf(v, us_prev, v_prev, t_prev, **kwargs), log_space=True)  # :145-146
        log_wss.append(log_ws); As.append(A); uss.append(us)
        us_prev = us
    return np.stack(As), np.stack(log_wss), np.stack(uss)


def backward_scanning_pass(key, As, xss, log_w_T):
    B = barker_move(key, normalise(log_w_T))                                         # :257
    K = As.shape[0]
    xs = [xss[-1, B]]
    Bs = [B]
    for t in range(K - 1, -1, -1):                                                   # :260-267
        B = int(As[t][B])
        xs.append(xss[t, B]); Bs.append(B)
    return np.stack(xs[::-1]), np.array(Bs[::-1], dtype=np.int32)


def backward_sampling_pass(key, transition_logpdf, vs, ts, uss, log_ws, *args, **kwargs):
    K_plus_one = uss.shape[0]
    keys = jr.split(key, K_plus_one)                                                 # :194
    B = barker_move(keys[-1], normalise(log_ws[-1]))                                 # :200-201
    x = uss[-1, B]
    xs, Bs = [x], [B]
    # :217 -- keys[:-1] are consumed in order while time runs backwards
    for q, t in enumerate(range(K_plus_one - 2, -1, -1)):
        G = transition_logpdf(x, uss[t], vs[t], ts[t], *args, **kwargs)              # :206
        G = G - G.max()
        w = normalise((G + log_ws[t]).astype(G.dtype))                               # :208-209
        B = int(jr.choice(keys[q], w.shape[0], (), p=w))                             # :210
        x = uss[t, B]
        xs.append(x); Bs.append(B)
    return np.stack(xs[::-1]), np.array(Bs[::-1], dtype=np.int32)


def csmc_kernel(key, us_star, bs_star, vs, ts, init_sampler, init_likelihood_logpdf, transition_sampler,
                transition_logpdf, measurement_cond_logpdf, cond_resampling, nsamples, backward=False, **kwargs):
    key_fwd, key_bwd = jr.split(key, 2)                                              # :65
    As, log_ws, xss = forward_pass(key_fwd, us_star, bs_star, vs, ts, init_sampler, init_likelihood_logpdf,
                                   transition_sampler, measurement_cond_logpdf, cond_resampling, nsamples, **kwargs)
    if backward:
        return backward_sampling_pass(key_bwd, transition_logpdf, vs, ts, xss, log_ws, **kwargs)
    return backward_scanning_pass(key_bwd, As, xss, log_ws[-1])
