"""Bootstrap filter / smoother and the pseudo-marginal kernel -- API of ``fbs/samplers/smc.py``
(``bootstrap_filter`` :9-88, ``bootstrap_backward_smoother`` :91-112, ``pmcmc_filter_step`` :115-158,
``pcn_proposal`` :161-168, ``pmcmc_kernel`` :171-258).

``pmcmc_filter_step`` is ONE persistent kernel (``fbs_pmcmc_filter_affine_f32``).  ``pmcmc_kernel`` adds the
forward-noising, pCN, reference-sampling and MH kernels around it -- seven launches per MCMC step for any
number of chains.  ``bootstrap_filter`` / ``bootstrap_backward_smoother`` (chain initialisation only) are
composed step by step from the closure kernels.
"""
import math
import numpy as np
import torch
from .. import _native as nat
from .._tensor import dev, empty, ptr, stream, out, is_host
from .. import random as frandom
from .common import MCMCState
from .csmc.csmc import _model_of, _scheme_of
from ..nn.unet import ScoreNetModel
from ..nn import ops as nnops


def pmcmc_filter_step(key, vs_bridge, u0s, ts, transition_sampler, likelihood_logpdf, resampling, nparticles,
                      return_history=False, **kwargs):
    """smc.py:115-158 -> ``(uT [.., N, du], log_ell [..])``."""
    model = _model_of(transition_sampler, likelihood_logpdf)
    scheme = _scheme_of(resampling, 'unconditional')
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B, K, N = k.shape[0], model.K, int(nparticles)
    if isinstance(model, ScoreNetModel):
        if B != 1 or return_history:
            raise NotImplementedError('score-network chains run one at a time, without history')
        uT, log_ell = _pmcmc_filter_step_nn(k[0], vs_bridge, u0s, model, resampling, N)
        res = (uT.reshape(1, N, model.du), log_ell.reshape(1))
        if single:
            res = tuple(t[0] for t in res)
        return tuple(out(t, host) for t in res)
    v = dev(vs_bridge, torch.float32).reshape(B, K + 1, model.dv)
    u0 = dev(u0s, torch.float32).reshape(B, N, model.du)
    uT = empty((B, N, model.du), torch.float32)
    log_ell = empty((B,), torch.float32)
    inds = lwh = ush = None
    if return_history:
        inds = empty((B, K, N), torch.int32)
        lwh = empty((B, K, N), torch.float32)
        ush = empty((B, K, N, model.du), torch.float32)
    ws, ws_bytes = model.workspace(B)
    nat.call('fbs_pmcmc_filter_affine_f32', stream(), model.struct(), ptr(k), ptr(v), ptr(u0), scheme, B, N, ptr(uT),
             ptr(log_ell), ptr(inds), ptr(lwh), ptr(ush), ptr(ws), ws_bytes)
    res = (uT, log_ell) + ((inds, lwh, ush) if return_history else ())
    if single:
        res = tuple(t[0] for t in res)
    return tuple(out(t, host) for t in res)


def _pmcmc_filter_step_nn(key, vs_bridge, u0s, model, resampling, N):
    """smc.py:115-158 over a ScoreNetModel, one chain: per step ONE score evaluation gives the weights of the current
    particles and their transition means; the resampled particles' transition is the gathered mean plus fresh noise."""
    K, p, c = model.K, model.p, model.c
    v = dev(vs_bridge, torch.float32).reshape(K + 1, model.q, c)
    us = dev(u0s, torch.float32).reshape(N, p, c).contiguous()
    sk = frandom.split(frandom.split(key, K), 2).contiguous()       # [K, (proposal, resampling), 2]   smc.py:142,154
    log_ell = torch.zeros((), dtype=torch.float32, device=us.device)
    logN = np.float32(math.log(N))
    gathered = torch.empty_like(us)
    for kk in range(K):
        mean, lw, sd = model.mean_and_logw(us, v[kk], v[kk + 1], model.ts[kk])          # smc.py:144
        c_ = torch.logsumexp(lw, dim=0)
        log_ell = log_ell - logN + c_                                                   # smc.py:146
        inds = resampling(torch.exp(lw - c_).contiguous(), sk[kk, 1].contiguous())      # smc.py:147-148
        nnops.gather_rows(mean, inds.reshape(N).to(torch.int32), gathered)              # smc.py:149-150
        us = gathered + sd * frandom.normal(sk[kk, 0].contiguous(), (N, p, c))
    return us, log_ell


def pcn_proposal(key, delta: float, x, mean, sampler):
    """smc.py:161-168.  ``sampler(key)`` is called on the two split keys (batched when ``key`` is)."""
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    keys = frandom.split(k.reshape(-1, 2), 2)
    r0 = dev(sampler(keys[:, 0].contiguous() if not single else keys[0, 0].contiguous()), torch.float32)
    r1 = dev(sampler(keys[:, 1].contiguous() if not single else keys[0, 1].contiguous()), torch.float32)
    xt = dev(x, torch.float32)
    B = keys.shape[0]
    n = xt.numel() // B
    mt = dev(mean, torch.float32).reshape(-1)
    if mt.numel() != n:
        raise ValueError('mean must be shared across chains (shape of one path)')
    o = torch.empty_like(xt)
    nat.call('fbs_pcn_combine_f32', stream(), float(delta), ptr(xt), ptr(mt), ptr(r0), ptr(r1), B, n, ptr(o))
    return out(o, host)


PIPELINE_MIN_CHAINS = 512     # host-buffer calls with at least this many chains are chunked over CUDA streams
PIPELINE_CHUNKS = int(__import__('os').environ.get('FBS_PIPELINE_CHUNKS', '4'))   # measured: 8 -> 59.8, 4 -> 56.6, 2 -> 57.3 ms per host-buffer Gibbs sweep (4144 chains)
_streams = []


def _chunk_bounds(B: int):
    """Chunks of chains for the host-buffer pipeline.  The sweep kernels run two chains per CTA on every SM, so a chunk
    that is not a multiple of 2 x (number of SMs) chains ends in a partly empty wave; chunk sizes are rounded up to whole waves
    (4144 chains on 148 SMs: chunks of 1184 = four waves, the last one 592)."""
    wave = 2 * torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    nchunks = max(1, min(PIPELINE_CHUNKS, B // (PIPELINE_MIN_CHAINS // 2)))
    size = -(-B // nchunks)
    if size >= wave:
        size = -(-size // wave) * wave
    return [(lo, min(lo + size, B)) for lo in range(0, B, size)]


def _pmcmc_kernel_pipelined(key, uT, log_ell, ys, y0, ts, fwd_ys_sampler, sde, ref_sampler, transition_sampler,
                            likelihood_logpdf, resampling, nparticles, delta, which_u, kwargs):
    """Host-buffer call on many chains: the chains are independent, so they are cut into chunks that run on separate CUDA
    streams -- chunk c's kernels overlap chunk c + 1's host-to-device copies and chunk c - 1's device-to-host copies
    (both PCIe directions busy while the SMs work).  Results land in page-locked buffers returned as numpy views."""
    k_all = np.asarray(key) if not isinstance(key, torch.Tensor) else key
    B = k_all.shape[0]
    bounds = _chunk_bounds(B)
    nchunks = len(bounds)
    while len(_streams) < nchunks:
        _streams.append(torch.cuda.Stream())

    def host_t(x):
        return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))

    k_h, uT_h, le_h, ys_h = host_t(key), host_t(uT), host_t(log_ell), host_t(ys)
    y0_d = dev(y0, torch.float32)
    outs = None
    cur = torch.cuda.current_stream()
    for c, (lo, hi) in enumerate(bounds):
        st = _streams[c]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            args = [t[lo:hi].to(y0_d.device, non_blocking=True) for t in (k_h, uT_h, le_h, ys_h)]
            r = pmcmc_kernel(args[0], args[1], args[2], args[3], y0_d, ts, fwd_ys_sampler, sde, ref_sampler, transition_sampler,
                             likelihood_logpdf, resampling, nparticles, delta=delta, which_u=which_u, **kwargs)
            flat = list(r[:3]) + list(r[3])
            if outs is None:
                outs = [torch.empty((B,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True) for t in flat]
            for o, t in zip(outs, flat):
                o[lo:hi].copy_(t, non_blocking=True)
    for st in _streams[:nchunks]:
        st.synchronize()
    res = [o.numpy() for o in outs]
    return res[0], res[1], res[2], MCMCState(*res[3:])


def pmcmc_kernel(key, uT, log_ell, ys, y0, ts, fwd_ys_sampler, sde, ref_sampler, transition_sampler,
                 likelihood_logpdf, resampling, nparticles, delta=None, which_u=0, **kwargs):
    """One pseudo-marginal MCMC step targeting p(u_T | v_T = y0); same arguments as the reference (smc.py:171-184).

    Returns ``(uT, log_ell, ys, MCMCState)``.  With keys ``[B, 2]`` every array carries the chain axis.
    """
    model = _model_of(transition_sampler, likelihood_logpdf)
    host = is_host(key)
    if host and np.ndim(key) == 2 and np.shape(key)[0] >= PIPELINE_MIN_CHAINS and getattr(model, '_pipeline_warm', False):
        # (the first call runs unchunked: it creates the model's cached device arrays on one stream)
        return _pmcmc_kernel_pipelined(key, uT, log_ell, ys, y0, ts, fwd_ys_sampler, sde, ref_sampler, transition_sampler,
                                       likelihood_logpdf, resampling, nparticles, delta, which_u, kwargs)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B, K, N = k.shape[0], model.K, int(nparticles)
    keys = frandom.split(k, 4)                                                     # smc.py:231
    key_prop, key_u0, key_filter, key_mh = (keys[:, i].contiguous() for i in range(4))
    ys_d = dev(ys, torch.float32).reshape(B, K + 1, model.dv).clone()
    uT_d = dev(uT, torch.float32).reshape(B, model.du).clone()
    le_d = dev(log_ell, torch.float32).reshape(B).clone()
    y0_d = dev(y0, torch.float32).reshape(-1, model.dv)
    if y0_d.shape[0] != 1:
        raise ValueError('y0 is shared across chains (in_axes=None in the reference driver, gp_pmcmc.py:161)')
    if delta is None:
        prop_ys = dev(fwd_ys_sampler(key_prop, y0_d[0]), torch.float32).reshape(B, K + 1, model.dv)   # smc.py:234
    else:
        tsh = np.asarray(ts.detach().cpu().numpy() if isinstance(ts, torch.Tensor) else ts, dtype=np.float32)
        coef = np.asarray(sde.transition(tsh, tsh[0], dtype=np.float32)[0], dtype=np.float32)         # sde.mean, smc.py:236
        mean = dev(coef, torch.float32)[:, None] * y0_d                                              # [K+1, dv]
        prop_ys = pcn_proposal(key_prop, delta, ys_d, mean.contiguous(), lambda key_: fwd_ys_sampler(key_, y0_d[0]))
        prop_ys = prop_ys.reshape(B, K + 1, model.dv)
    vs = torch.flip(prop_ys, dims=[1]).contiguous()                                # smc.py:239
    u0s = dev(ref_sampler(key_u0, vs[:, 0].contiguous(), N), torch.float32)        # smc.py:241
    prop_uTs, prop_log_ell = pmcmc_filter_step(key_filter, vs, u0s, ts, transition_sampler, likelihood_logpdf,
                                               resampling, N, **kwargs)
    acc_prob = empty((B,), torch.float32)
    is_acc = empty((B,), torch.uint8)
    old_log_ell = le_d.clone()
    nat.call('fbs_mh_accept_f32', stream(), ptr(key_mh), ptr(prop_uTs), ptr(prop_log_ell), ptr(prop_ys), B, N, model.du,
             (K + 1) * model.dv, int(which_u), ptr(uT_d), ptr(le_d), ptr(ys_d), ptr(acc_prob), ptr(is_acc))
    state = MCMCState(acceptance_prob=acc_prob, is_accepted=is_acc.bool(), prop_log_ell=prop_log_ell,
                      log_ell=old_log_ell)
    res = [uT_d, le_d, ys_d]
    model._pipeline_warm = True
    if single:
        res = [t[0] for t in res]
        state = MCMCState(*[t[0] for t in state])
    return (*[out(t, host) for t in res], MCMCState(*[out(t, host) for t in state]))


def bootstrap_filter(transition_sampler, measurement_cond_pdf, vs, ts, init_sampler, key, nparticles, resampling,
                     log: bool = True, return_last: bool = True, return_history: bool = False, **kwargs):
    """smc.py:9-88 -> (samples, negative log-likelihood).  Affine models: the whole K-step scan is ONE launch
    (``fbs_bootstrap_filter_affine_f32``); ``return_history=True`` (tests) appends the per-step resampling indices and
    unnormalised log-weights.  Score-network models: one score evaluation per step."""
    if not log:
        raise NotImplementedError('only the log-domain filter is used by the reference drivers')
    model = _model_of(transition_sampler, measurement_cond_pdf)
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B, K, N = k.shape[0], model.K, int(nparticles)
    if isinstance(model, ScoreNetModel):
        if B != 1:
            raise NotImplementedError('score-network chains run one at a time')
        res = _bootstrap_filter_nn(k[0], vs, model, init_sampler, resampling, N, return_last)
        res = tuple(t.unsqueeze(0) for t in res)
        if single:
            res = tuple(t[0] for t in res)
        return tuple(out(t, host) for t in res)
    scheme = _scheme_of(resampling, 'unconditional')
    v = dev(vs, torch.float32).reshape(B, K + 1, model.dv)
    ks = frandom.split(k, 2)                                                       # smc.py:77
    key_init, key_steps = ks[:, 0].contiguous(), ks[:, 1].contiguous()
    u0 = dev(init_sampler(key_init, v[:, 0].contiguous(), N), torch.float32).reshape(B, N, model.du).contiguous()   # smc.py:78
    uT = empty((B, N, model.du), torch.float32)
    log_nell = empty((B,), torch.float32)
    # the whole scan (smc.py:58-74,79-84) in one launch; the history is written straight behind the initial set
    hist = None if return_last else empty((B, K + 1, N, model.du), torch.float32)
    inds = lwh = None
    if return_history:
        inds = empty((B, K, N), torch.int32)
        lwh = empty((B, K, N), torch.float32)
    ush = None
    if hist is not None:
        hist[:, 0].copy_(u0)
        ush = empty((B, K, N, model.du), torch.float32)
    nat.call('fbs_bootstrap_filter_affine_f32', stream(), model.struct(), ptr(key_steps), ptr(v), ptr(u0), scheme, B, N,
             ptr(uT), ptr(log_nell), ptr(inds), ptr(lwh), ptr(ush))
    if hist is not None:
        hist[:, 1:].copy_(ush)
    res = (uT, log_nell) if return_last else (hist, log_nell)
    if return_history:
        res = res + (inds, lwh)
    if single:
        res = tuple(t[0] for t in res)
    return tuple(out(t, host) for t in res)


def _bootstrap_filter_nn(key, vs, model, init_sampler, resampling, N, return_last):
    """smc.py:58-88 over a ScoreNetModel, one chain: ONE score evaluation per step feeds the proposal and the weights."""
    K, p, c = model.K, model.p, model.c
    v = dev(vs, torch.float32).reshape(K + 1, model.q, c)
    ks = frandom.split(key, 2)                                                     # smc.py:77
    us = dev(init_sampler(ks[0].contiguous(), v[0], N), torch.float32).reshape(N, p, c).contiguous()
    sk = frandom.split(frandom.split(ks[1].contiguous(), K), 2).contiguous()       # [K, (proposal, resampling), 2]
    log_nell = torch.zeros((), dtype=torch.float32, device=us.device)
    logN = np.float32(math.log(N))
    hist = [us]
    for kk in range(K):
        us_new, lw = model.step(us, v[kk], v[kk + 1], model.ts[kk], sk[kk, 0].contiguous())      # smc.py:63,65
        c_ = torch.logsumexp(lw, dim=0)
        log_nell = log_nell - (c_ - logN)                                          # smc.py:67
        inds = resampling(torch.exp(lw - c_).contiguous(), sk[kk, 1].contiguous())  # smc.py:68-69
        us = torch.empty_like(us_new)
        nnops.gather_rows(us_new, inds.reshape(N).to(torch.int32), us)             # smc.py:72
        if not return_last:
            hist.append(us)
    if return_last:
        return us.reshape(N, model.du), log_nell
    return torch.stack(hist, dim=0).reshape(K + 1, N, model.du), log_nell


def bootstrap_backward_smoother(key, filter_us, vs, ts, transition_logpdf, *args, **kwargs):
    """smc.py:91-112 (the unsplit ``key`` draws u_T, as written upstream)."""
    model = _model_of(transition_logpdf)
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B = k.shape[0]
    fu = dev(filter_us, torch.float32)
    N, du = fu.shape[-2], fu.shape[-1]
    K = model.K
    if not isinstance(model, ScoreNetModel):
        # the whole backward recursion in one launch (mode 1 of fbs_backward_sample_affine_f32).  A history without a chain
        # axis under batched keys is shared by all of them (the reference vmaps over the keys only, test_filters.py:138-141)
        shared = 1 if (B > 1 and fu.numel() == (K + 1) * N * du) else 0
        nb = 1 if shared else B
        fu = fu.reshape(nb, K + 1, N, du).contiguous()
        v = dev(vs, torch.float32).reshape(nb, K + 1, model.dv).contiguous()
        xs = empty((B, K + 1, du), torch.float32)
        nat.call('fbs_backward_sample_affine_f32', stream(), model.struct(), 1, ptr(k), ptr(v), ptr(fu), None, shared, B, N,
                 ptr(xs), None)
        return out(xs[0] if single else xs, host)
    fu = fu.reshape(B, K + 1, N, du)
    v = dev(vs, torch.float32).reshape(B, K + 1, model.dv)
    ks = frandom.split(k, 2)                                                       # smc.py:108
    key_smoother = ks[:, 1].contiguous()
    ar = torch.arange(B, device=fu.device)
    uT = fu[ar, -1, frandom.randint(k, (), 0, N).reshape(B).long()]               # smc.py:109
    skeys = frandom.split(key_smoother, K)
    u = uT
    traj = []
    for q, t in enumerate(range(K - 1, -1, -1)):                                   # smc.py:110-111
        lw = model.transition_logpdf(u[0], fu[0, t], v[0, t], model.ts[t]).reshape(1, N)
        w = torch.exp(lw - torch.logsumexp(lw, dim=-1, keepdim=True)).contiguous()
        idx = frandom.choice(skeys[:, q].contiguous(), N, (), p=w).reshape(B).long()
        u = fu[ar, t, idx]
        traj.append(u)
    res = torch.cat([torch.stack(traj[::-1], dim=1), uT[:, None]], dim=1)
    return out(res[0] if single else res, host)


def twisted_smc(key, y, ts, init_sampler, transition_logpdf, twisting_logpdf, twisting_prop_sampler, twisting_prop_logpdf,
                resampling, nparticles, return_history: bool = False, **kwargs):
    """Twisted SMC, same arguments and returns as ``fbs.samplers.twisted_smc`` (smc.py:261-309): ``(samples [.., N, d],
    log_weights [.., N])``.  The four closures must be the bound methods of ONE :class:`fbs_b200.TwistedAffineModel`; the
    whole K-step scan is one launch (``fbs_twisted_smc_affine_f32``).  Keys ``[B, 2]`` run B independent samplers (``y`` shared
    or ``[B, d]``).  ``return_history=True`` (tests) appends the per-step resampling indices, particles and normalised log-weights."""
    from ..models import TwistedAffineModel
    owners = {id(getattr(f, '__self__', None)) for f in (transition_logpdf, twisting_logpdf, twisting_prop_sampler,
                                                         twisting_prop_logpdf)}
    model = getattr(transition_logpdf, '__self__', None)
    if len(owners) != 1 or not isinstance(model, TwistedAffineModel):
        raise TypeError('twisted_smc fuses the scan into a CUDA kernel and cannot call opaque Python closures: pass the bound '
                        'methods of one fbs_b200.TwistedAffineModel (no interpreted fallback exists)')
    scheme = _scheme_of(resampling, 'unconditional')
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B, K, N, d = k.shape[0], model.K, int(nparticles), model.d
    ks = frandom.split(k, 2)                                                       # smc.py:296
    key_init, key_filter = ks[:, 0].contiguous(), ks[:, 1].contiguous()
    x0 = dev(init_sampler(key_init if not single else key_init[0], N), torch.float32).reshape(B, N, d).contiguous()   # smc.py:299
    yd = dev(y, torch.float32).reshape(-1, d).contiguous()
    if yd.shape[0] not in (1, B):
        raise ValueError('y must be one observation or one per key')
    a = model.device_arrays()
    samples, log_ws = empty((B, N, d), torch.float32), empty((B, N), torch.float32)
    inds = xs_hist = lw_hist = None
    if return_history:
        inds, xs_hist, lw_hist = empty((B, K, N), torch.int32), empty((B, K, N, d), torch.float32), empty((B, K, N), torch.float32)
    nat.call('fbs_twisted_smc_affine_f32', stream(), ptr(a['MT']), ptr(a['Mr']), ptr(a['m']), ptr(a['sd']), ptr(a['g2']),
             float(model.dt), float(model.obs_var), K, d, ptr(key_filter), ptr(yd), int(yd.shape[0] == B and B > 1), ptr(x0), scheme,
             B, N, ptr(samples), ptr(log_ws), ptr(inds), ptr(xs_hist), ptr(lw_hist))
    res = (samples, log_ws) + ((inds, xs_hist, lw_hist) if return_history else ())
    if single:
        res = tuple(t[0] for t in res)
    return tuple(out(t, host) for t in res)
