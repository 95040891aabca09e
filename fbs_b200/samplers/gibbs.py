"""Particle-Gibbs kernel -- API of ``fbs/samplers/gibbs.py`` (``gibbs_init`` :23-65, ``gibbs_kernel`` :68-168,
``force_move`` :171-214).

One Gibbs sweep = 6 launches for any number of chains: key splits, forward noising (written reversed and
split into (us, vs) by the kernel), the persistent CSMC sweep, forced move + x0 selection, the second
forward noising, and the ``randint`` for the next reference indices.
``marg_y=True`` (Doob bridge, "Not used in our paper", gibbs.py:115) is not built.
"""
import numpy as np
import torch
from .. import _native as nat
from .._tensor import dev, empty, ptr, stream, out, is_host
from .. import random as frandom
from ..models import AffineGaussianModel
from ..nn.unet import ScoreNetModel
from .csmc.csmc import (DegenerateInit, NormalInit, forward_pass_device, backward_scanning_pass, _model_of,
                        _check_reference_indices)
from .csmc.resamplings import killing
from .resampling import stratified
from .smc import bootstrap_filter, bootstrap_backward_smoother


def force_move(key, weights, k):
    """Forced-move trajectory selection (gibbs.py:171-214) -> ``(index, alpha)``.  weights normalised."""
    host = is_host(weights)
    w = dev(weights, torch.float32)
    single = w.dim() == 1
    w = w.reshape(-1, w.shape[-1])
    B, N = w.shape
    kk = dev(key, torch.uint32).reshape(-1, 2)
    kv = dev(np.asarray(k, dtype=np.int32).reshape(-1) if not isinstance(k, torch.Tensor) else k, torch.int32).reshape(-1)
    if kv.shape[0] == 1 and B > 1:
        kv = kv.expand(B).contiguous()
    idx = empty((B,), torch.int32)
    alpha = empty((B,), torch.float32)
    nat.call('fbs_force_move_f32', stream(), ptr(kk), ptr(w), 0, None, ptr(kv), B, N, 0, ptr(idx), ptr(alpha), None)
    if single:
        idx, alpha = idx[0], alpha[0]
    return out(idx, host), out(alpha, host)


def _fwd_reversed(model, fwd_sampler, unpack, key, x0, y0, kwargs):
    """(us, vs) = reversed, unpacked forward path.  Fast path when fwd_sampler/unpack are the model's own."""
    if getattr(fwd_sampler, '__self__', None) is model and getattr(unpack, '__self__', None) is model:
        return model.fwd_sampler_reversed(key, x0, y0)
    path_xy = fwd_sampler(key, x0, y0, **kwargs)                                   # gibbs.py:127
    path_x, path_y = unpack(path_xy, **kwargs)
    px, py = dev(path_x, torch.float32), dev(path_y, torch.float32)
    tdim = px.dim() - 2
    return torch.flip(px, dims=[tdim]).contiguous(), torch.flip(py, dims=[tdim]).contiguous()   # gibbs.py:129-130


def _gibbs_kernel_pipelined(key, x0, y0, bs_star, ts, fwd_sampler, sde, unpack, nparticles, transition_sampler,
                            transition_logpdf, likelihood_logpdf, explicit_backward, explicit_final, kwargs):
    """Host-buffer call on many chains (the reference driver's per-sweep ``np`` copies, gp_gibbs.py:185-190): the chains are
    independent, so they are cut into chunks that run on separate CUDA streams -- chunk c's kernels overlap chunk c + 1's
    host-to-device copies and chunk c - 1's device-to-host copies.  Results land in page-locked buffers returned as numpy
    views.  Chain b's numbers do not depend on the chunking (tests/test_gpu_csmc.py)."""
    from . import smc as _smc
    B = np.shape(key)[0]
    _check_reference_indices(bs_star, int(nparticles) + 1 if explicit_final else int(nparticles))
    bounds = _smc._chunk_bounds(B)
    nchunks = len(bounds)
    while len(_smc._streams) < nchunks:
        _smc._streams.append(torch.cuda.Stream())

    def host_t(x, dtype):
        if isinstance(x, torch.Tensor):
            return x
        np_dt = {torch.float32: np.float32, torch.int32: np.int32, torch.uint32: np.uint32}[dtype]
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np_dt, copy=False)))

    k_h, x0_h, bs_h = host_t(key, torch.uint32), host_t(x0, torch.float32), host_t(bs_star, torch.int32)
    y0_d = dev(y0, torch.float32)
    per_chain_y0 = y0_d.dim() == 2 and y0_d.shape[0] == B
    outs = None
    cur = torch.cuda.current_stream()
    d = y0_d.device
    for c, (lo, hi) in enumerate(bounds):
        st = _smc._streams[c]
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            args = [t[lo:hi].to(d, non_blocking=True) for t in (k_h, x0_h, bs_h)]
            r = gibbs_kernel(args[0], args[1], y0_d[lo:hi] if per_chain_y0 else y0_d, None, args[2], ts, fwd_sampler, sde, unpack,
                             nparticles, transition_sampler, transition_logpdf, likelihood_logpdf, marg_y=False,
                             explicit_backward=explicit_backward, explicit_final=explicit_final, **kwargs)
            if outs is None:
                outs = [torch.empty((B,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True) for t in r]
            for o, t in zip(outs, r):
                o[lo:hi].copy_(t, non_blocking=True)
    for st in _smc._streams[:nchunks]:
        st.synchronize()
    return tuple(o.numpy() for o in outs)


def gibbs_kernel(key, x0, y0, us_star, bs_star, ts, fwd_sampler, sde, unpack, nparticles, transition_sampler,
                 transition_logpdf, likelihood_logpdf, marg_y: bool = False, explicit_backward: bool = True,
                 explicit_final: bool = False, **kwargs):
    """Gibbs kernel of the forward-backward conditional sampler; same arguments and returns as the reference
    (gibbs.py:68-168): ``(x0, us_star, bs_star, bs_star_next != bs_star)``.

    ``us_star`` is ignored, as upstream (gibbs.py:91-92).  With keys ``[B, 2]``: ``x0 [B, du]``,
    ``bs_star [B, K+1]``, ``y0 [dv]`` shared or ``[B, dv]`` (one conditioning target per chain; with a score-network model the
    B chains share every score evaluation -- the reference loops over test images, inpainting.py:205-210).
    """
    if marg_y:
        raise NotImplementedError('marg_y=True (Doob bridge for y) is not part of the accelerated path')
    model = _model_of(transition_sampler, likelihood_logpdf)
    host = is_host(key)
    from . import smc as _smc
    if (host and isinstance(model, AffineGaussianModel) and np.ndim(key) == 2 and np.shape(key)[0] >= _smc.PIPELINE_MIN_CHAINS
            and getattr(model, '_gibbs_pipeline_warm', False)):
        # (the first call runs unchunked: it creates the model's cached device arrays on one stream)
        return _gibbs_kernel_pipelined(key, x0, y0, bs_star, ts, fwd_sampler, sde, unpack, nparticles, transition_sampler,
                                       transition_logpdf, likelihood_logpdf, explicit_backward, explicit_final, kwargs)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B, K, N = k.shape[0], model.K, int(nparticles)
    if isinstance(model, ScoreNetModel) and B != 1 and not explicit_backward:
        raise NotImplementedError('several image chains at once need explicit_backward=True (the batched sweep keeps no history)')
    x0_d = dev(x0, torch.float32).reshape(B, model.du)
    y0_d = dev(y0, torch.float32).reshape(-1, model.dv)
    _check_reference_indices(bs_star, N + 1 if explicit_final else N)
    bs = dev(bs_star, torch.int32).reshape(B, K + 1)

    ks = frandom.split(k, 3)                                                       # gibbs.py:126
    key_fwd, key_csmc = ks[:, 0].contiguous(), ks[:, 1].contiguous()
    us, vs = _fwd_reversed(model, fwd_sampler, unpack, key_fwd, x0_d, y0_d, kwargs)
    us, vs = us.reshape(B, K + 1, model.du), vs.reshape(B, K + 1, model.dv)
    init = NormalInit(model) if explicit_final else DegenerateInit(N)

    if explicit_backward:
        kc = frandom.split(key_csmc, 4)                                            # gibbs.py:147
        key_csmc_fwd, key_csmc_x0, key_csmc_bwd_us, key_csmc_bwd_bs = (kc[:, i].contiguous() for i in range(4))
        r = forward_pass_device(key_csmc_fwd, us, bs, vs, model, init, killing.scheme, N, history=False)
        idx = empty((B,), torch.int32)
        x0_new = empty((B, model.du), torch.float32)
        b_last = bs[:, -1].contiguous()                                            # kept referenced until the launch
        nat.call('fbs_force_move_f32', stream(), ptr(key_csmc_x0), ptr(r['log_ws_last']), 1, ptr(r['us_last']),
                 ptr(b_last), B, r['N'], model.du, ptr(idx), None, ptr(x0_new))   # gibbs.py:152-154
        us_star_next, _ = _fwd_reversed(model, fwd_sampler, unpack, key_csmc_bwd_us, x0_new, y0_d, kwargs)  # :155
        us_star_next = us_star_next.reshape(B, K + 1, model.du)
        bs_star_next = frandom.randint(key_csmc_bwd_bs, (K + 1,), 0, N)            # gibbs.py:156
    else:
        kc = frandom.split(key_csmc, 2)                                            # csmc.py:65
        r = forward_pass_device(kc[:, 0].contiguous(), us, bs, vs, model, init, killing.scheme, N, history=True)
        us_star_next, bs_star_next = backward_scanning_pass(kc[:, 1].contiguous(), r['As'], r['uss'],
                                                            r['log_wss'][:, -1].contiguous())
    x0_next = us_star_next[:, -1].contiguous()                                     # gibbs.py:167
    model._gibbs_pipeline_warm = True
    changed = bs_star_next != bs
    res = (x0_next, us_star_next, bs_star_next, changed)
    if single:
        res = tuple(t[0] for t in res)
    return tuple(out(t, host) for t in res)


def gibbs_init(key, y0, x0_shape, ts, fwd_sampler, sde, unpack, transition_sampler, transition_logpdf,
               likelihood_logpdf, nparticles, method: str = 'smoother', marg_y: bool = False, x0=None, **kwargs):
    """Initialise the Gibbs chain with a bootstrap filter / smoother draw (gibbs.py:23-65, ``marg_y=False``)."""
    if marg_y:
        raise NotImplementedError('marg_y=True (Doob bridge for y) is not part of the accelerated path')
    model = _model_of(transition_sampler, likelihood_logpdf)
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B, N = k.shape[0], int(nparticles)
    y0_d = dev(y0, torch.float32).reshape(-1, model.dv)
    x0_d = torch.zeros((B, model.du), dtype=torch.float32, device=k.device) if x0 is None else \
        dev(x0, torch.float32).reshape(B, model.du)
    ks = frandom.split(k, 6)                                                       # gibbs.py:39
    key_fwd, _, key_u0, key_bf, key_fwd2, key_bwd = (ks[:, i].contiguous() for i in range(6))
    _, vs = _fwd_reversed(model, fwd_sampler, unpack, key_fwd, x0_d, y0_d, kwargs)

    def init_sampler(*_):                                                          # gibbs.py:46-48
        return frandom.normal(key_u0, (N, model.du))

    if method == 'filter':
        approx_x0 = bootstrap_filter(transition_sampler, likelihood_logpdf, vs, ts, init_sampler, key_bf, N, stratified,
                                     log=True, return_last=True, **kwargs)[0][:, 0]
        approx_us_star, _ = _fwd_reversed(model, fwd_sampler, unpack, key_fwd2, approx_x0.contiguous(), y0_d, kwargs)
    elif method == 'smoother':
        uss = bootstrap_filter(transition_sampler, likelihood_logpdf, vs, ts, init_sampler, key_bf, N, stratified,
                               log=True, return_last=False, **kwargs)[0]
        approx_x0 = uss[:, -1, 0]
        approx_us_star = bootstrap_backward_smoother(key_bwd, uss, vs, ts, transition_logpdf, **kwargs)
    else:
        raise ValueError(f"Unknown method {method}")
    res = (approx_x0, approx_us_star)
    if single:
        res = tuple(t[0] for t in res)
    return tuple(out(t, host) for t in res)
