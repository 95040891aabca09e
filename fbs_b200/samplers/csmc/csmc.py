"""Conditional SMC -- API of ``fbs/samplers/csmc/csmc.py`` (``csmc_kernel`` :14-77, ``forward_pass`` :80-164,
``backward_scanning_pass`` :230-270, ``backward_sampling_pass`` :167-227, ``normalise`` :273-292,
``barker_move`` :295-297).

``forward_pass`` runs the whole K-step sweep in ONE persistent kernel (``fbs_csmc_forward_affine_f32``)
when the callables it is given are bound methods of an ``AffineGaussianModel`` plus one of the two init
objects below.  Anything else raises ``TypeError``: there is no interpreted or CPU fallback.

Batching: the reference ``vmap``s over chains; here every array may carry a leading chain axis ``B``
(keys ``[B, 2]``) and the kernel's grid runs over it.
"""
import math
import numpy as np
import torch
from ... import _native as nat
from ..._tensor import dev, empty, ptr, stream, out, is_host
from ...models import AffineGaussianModel
from ...nn.unet import ScoreNetModel
from ...nn import ops as nnops
from ... import random as frandom


class DegenerateInit:
    """``explicit_final=False`` initialisation (gibbs.py:139-144): N copies of the reference start, uniform weights."""

    def __init__(self, nparticles: int):
        self.nparticles = int(nparticles)

    def sampler(self, *_):
        raise TypeError('DegenerateInit is consumed by the fused forward_pass; it is not called')

    def likelihood_logpdf(self, *_, **__):
        raise TypeError('DegenerateInit is consumed by the fused forward_pass; it is not called')

    @property
    def init_log_w(self) -> float:
        return float(np.float32(-math.log(self.nparticles)))


class NormalInit:
    """``explicit_final=True`` initialisation (gibbs.py:132-137): N(0, I) particles, weights from the likelihood at ts[0]."""

    def __init__(self, model: AffineGaussianModel):
        self.model = model

    def sampler(self, key, n):
        return frandom.normal(key, (n, self.model.du))

    def likelihood_logpdf(self, v0, u0s, v1, **kwargs):
        return self.model.likelihood_logpdf(v0, u0s, v1, self.model.ts[0], **kwargs)


def _model_of(*fns):
    models = {id(getattr(f, '__self__', None)): getattr(f, '__self__', None) for f in fns}
    if len(models) != 1:
        raise TypeError('transition_sampler / likelihood_logpdf must be bound methods of ONE model object')
    model = next(iter(models.values()))
    if not isinstance(model, (AffineGaussianModel, ScoreNetModel)):
        raise TypeError('fbs_b200 fuses the sweep into a CUDA kernel and cannot call opaque Python closures: pass '
                        'bound methods of an fbs_b200.AffineGaussianModel or fbs_b200.nn.ScoreNetModel (no interpreted fallback exists)')
    return model


def _check_reference_indices(bs_star, n):
    """``0 <= bs_star < n``, checked when the indices arrive as host data (no device synchronisation is spent on device
    tensors: the kernels clamp what they load, as a JAX gather would)."""
    if bs_star is None or (isinstance(bs_star, torch.Tensor) and bs_star.is_cuda):
        return
    b = np.asarray(bs_star)
    if b.size and (b.min() < 0 or b.max() >= n):
        raise ValueError(f'bs_star must lie in [0, {n}): got [{int(b.min())}, {int(b.max())}] (a reference-particle index from a '
                         f'run with a different nparticles / explicit_final?)')


def _scheme_of(resampling, family):
    if getattr(resampling, 'family', None) != family or not hasattr(resampling, 'scheme'):
        raise TypeError(f'{family} resampling must be one of the fbs_b200 resampling functions')
    return resampling.scheme


def forward_pass_nn(key, us_star, bs_star, vs, model, init, scheme, nsamples, history=True):
    """csmc.py:132-164 for ONE chain with a :class:`ScoreNetModel`: a host loop over the K steps, each step = conditional
    resampling kernel, ancestor gather, ONE score-network evaluation (CUDA-graph replay) feeding both the transition
    and the weight (the reference evaluates the network twice, SURVEY finding 6c), pinning and normalisation.
    Shapes: us_star [K + 1, p, c], vs [K + 1, q, c], bs_star [K + 1]."""
    k = dev(key, torch.uint32).reshape(2)
    K, p, q, c = model.K, model.p, model.q, model.c
    us_star = dev(us_star, torch.float32).reshape(K + 1, p, c)
    v = dev(vs, torch.float32).reshape(K + 1, q, c)
    bs = dev(bs_star, torch.int32).reshape(K + 1)
    bsl = bs.long()
    ks = frandom.split(k, 2)                                                       # csmc.py:150
    key_init, key_scan = ks[0].contiguous(), ks[1].contiguous()
    sk = frandom.split(frandom.split(key_scan, K), 2).contiguous()                 # csmc.py:157,136: [K, (resampling, transition), 2]
    ts = model.ts
    if isinstance(init, NormalInit):
        N = int(nsamples) + 1                                                      # csmc.py:151
        us = frandom.normal(key_init, (N, p, c)).contiguous()
        us.index_copy_(0, bsl[0:1], us_star[0:1])                                  # csmc.py:152
        lw = model.likelihood_logpdf(v[0], us, v[1], ts[0])                        # gibbs.py:136-137
    elif isinstance(init, DegenerateInit):
        N = init.nparticles
        us = us_star[0:1].expand(N, p, c).contiguous()
        lw = torch.full((N,), init.init_log_w, dtype=torch.float32, device=us.device)
    else:
        raise TypeError('init_sampler / init_likelihood_logpdf must come from DegenerateInit or NormalInit')
    # per step: normalise + exp (one launch), conditional resampling, ancestor gather, image assembly, the score network
    # (one CUDA-graph replay), Euler--Maruyama step + weights + reference pin (one launch) -- no eager tensor arithmetic
    log_w, w = empty((N,), torch.float32), empty((N,), torch.float32)
    nnops.normalise_logw(lw.contiguous(), log_w, w)                                # csmc.py:155 (and the exp of :139)
    As = log_wss = uss = None
    if history:
        As = empty((K, N), torch.int32)
        log_wss = empty((K + 1, N), torch.float32)
        uss = empty((K + 1, N, p, c), torch.float32)
        log_wss[0].copy_(log_w)
        uss[0].copy_(us)
    A = empty((1, N), torch.int32)
    us_prev = torch.empty_like(us)
    for kk in range(K):
        nat.call('fbs_cond_resample_f32', stream(), scheme, ptr(sk[kk, 0]), ptr(w), ptr(bs[kk:kk + 1]), ptr(bs[kk + 1:kk + 2]), 1,
                 1, N, ptr(A))                                                     # csmc.py:139
        nnops.gather_rows(us, A.reshape(N), us_prev)                               # csmc.py:140
        us, lw = model.step(us_prev, v[kk], v[kk + 1], ts[kk], sk[kk, 1], pin_row=bs[kk + 1:kk + 2],
                            pin_value=us_star[kk + 1])                             # csmc.py:142,143,145
        nnops.normalise_logw(lw, log_w, w)                                         # csmc.py:146
        if history:
            As[kk].copy_(A[0])
            log_wss[kk + 1].copy_(log_w)
            uss[kk + 1].copy_(us)
    ex = (lambda t: None if t is None else t.unsqueeze(0))
    return dict(N=N, As=ex(As), log_wss=ex(log_wss), uss=None if uss is None else uss.reshape(1, K + 1, N, p * c),
                log_ws_last=log_w.unsqueeze(0).contiguous(), us_last=us.reshape(1, N, p * c).contiguous())


def forward_pass_nn_chains(keys, us_star, bs_star, vs, model, init, scheme, nsamples):
    """:func:`forward_pass_nn` for C independent chains / conditioning targets in one go (``keys [C, 2]``,
    ``us_star [C, K + 1, p, c]``, ``vs [C, K + 1, q, c]``, ``bs_star [C, K + 1]``): per step ONE batched conditional-resampling
    launch, ONE ancestor gather and ONE score-network evaluation over the C x N particles.  Chain c's result equals
    ``forward_pass_nn(keys[c], us_star[c], bs_star[c], vs[c], ...)`` bit for bit.  Returns the last step (no history)."""
    k = dev(keys, torch.uint32).reshape(-1, 2)
    C = k.shape[0]
    K, p, q, c = model.K, model.p, model.q, model.c
    us_star = dev(us_star, torch.float32).reshape(C, K + 1, p, c)
    v = dev(vs, torch.float32).reshape(C, K + 1, q, c)
    bs = dev(bs_star, torch.int32).reshape(C, K + 1).contiguous()
    ks = frandom.split(k, 2)                                                       # [C, 2, 2]   csmc.py:150
    key_scan = ks[:, 1].contiguous()
    sk = frandom.split(frandom.split(key_scan, K).reshape(C * K, 2), 2).reshape(C, K, 2, 2)   # csmc.py:157,136
    ts = model.ts
    if isinstance(init, NormalInit):
        N = int(nsamples) + 1
        us = torch.stack([frandom.normal(ks[ci, 0].contiguous(), (N, p, c)) for ci in range(C)]).contiguous()
        for ci in range(C):
            us[ci].index_copy_(0, bs[ci, 0:1].long(), us_star[ci, 0:1])
        lw = torch.stack([model.likelihood_logpdf(v[ci, 0], us[ci], v[ci, 1], ts[0]) for ci in range(C)])
    elif isinstance(init, DegenerateInit):
        N = init.nparticles
        us = us_star[:, 0:1].expand(C, N, p, c).contiguous()
        lw = torch.full((C, N), init.init_log_w, dtype=torch.float32, device=us.device)
    else:
        raise TypeError('init_sampler / init_likelihood_logpdf must come from DegenerateInit or NormalInit')
    log_w, w = empty((C, N), torch.float32), empty((C, N), torch.float32)
    nnops.normalise_logw(lw.contiguous(), log_w, w)
    A = empty((C, N), torch.int32)
    us_prev = torch.empty_like(us)
    base = (torch.arange(C, device=us.device, dtype=torch.int32) * N).reshape(C, 1)
    bsT = bs.t().contiguous()                                                      # [K + 1, C]: a step's indices are one row
    skr = sk[:, :, 0].transpose(0, 1).contiguous()                                 # [K, C, 2] resampling keys
    skt = sk[:, :, 1].transpose(0, 1).contiguous()                                 # [K, C, 2] transition keys
    for kk in range(K):
        nat.call('fbs_cond_resample_f32', stream(), scheme, ptr(skr[kk]), ptr(w), ptr(bsT[kk]), ptr(bsT[kk + 1]), 1, C, N,
                 ptr(A))                                                           # csmc.py:139, all chains
        a_glob = (A + base).reshape(C * N)
        nnops.gather_rows(us.reshape(C * N, p * c), a_glob, us_prev.reshape(C * N, p * c))
        us, lw = model.step_chains(us_prev, v[:, kk], v[:, kk + 1], ts[kk], skt[kk], pin_rows=bsT[kk + 1],
                                   pin_values=us_star[:, kk + 1])                  # csmc.py:142,143,145
        nnops.normalise_logw(lw, log_w, w)                                         # csmc.py:146
    return dict(N=N, log_ws_last=log_w.contiguous(), us_last=us.reshape(C, N, p * c).contiguous())


def forward_pass_device(key, us_star, bs_star, vs, model, init, scheme, nsamples, history=True):
    """Device-level forward pass on batched device tensors.  Returns a dict of device tensors."""
    if isinstance(model, ScoreNetModel):
        if dev(key, torch.uint32).numel() > 2:                                     # several conditioning targets
            if history:
                raise NotImplementedError('the batched score-network sweep keeps no history (use explicit_backward=True)')
            return forward_pass_nn_chains(key, us_star, bs_star, vs, model, init, scheme, nsamples)
        return forward_pass_nn(key, us_star, bs_star, vs, model, init, scheme, nsamples, history)
    k = dev(key, torch.uint32).reshape(-1, 2)
    B = k.shape[0]
    K, du, dv = model.K, model.du, model.dv
    us = dev(us_star, torch.float32).reshape(B, K + 1, du)
    bs = dev(bs_star, torch.int32).reshape(B, K + 1)
    v = dev(vs, torch.float32).reshape(B, K + 1, dv)
    if isinstance(init, NormalInit):
        init_mode, N, init_log_w = nat.INIT_NORMAL, int(nsamples) + 1, 0.0      # csmc.py:151: nsamples + 1
    elif isinstance(init, DegenerateInit):
        init_mode, N, init_log_w = nat.INIT_DEGENERATE, init.nparticles, init.init_log_w
    else:
        raise TypeError('init_sampler / init_likelihood_logpdf must come from DegenerateInit or NormalInit')
    _check_reference_indices(bs_star, N)
    res = dict(N=N)
    As = log_wss = uss = None
    if history:
        As = empty((B, K, N), torch.int32)
        log_wss = empty((B, K + 1, N), torch.float32)
        uss = empty((B, K + 1, N, du), torch.float32)
    lw_last = empty((B, N), torch.float32)
    us_last = empty((B, N, du), torch.float32)
    ws, ws_bytes = model.workspace(B)
    nat.call('fbs_csmc_forward_affine_f32', stream(), model.struct(), ptr(k), ptr(us), ptr(bs), ptr(v), init_mode,
             init_log_w, scheme, B, N, ptr(As), ptr(log_wss), ptr(uss), ptr(lw_last), ptr(us_last), ptr(ws), ws_bytes)
    res.update(As=As, log_wss=log_wss, uss=uss, log_ws_last=lw_last, us_last=us_last)
    return res


def forward_pass(key, us_star, bs_star, vs, ts, init_sampler, init_likelihood_logpdf, transition_sampler,
                 likelihood_logpdf, cond_resampling, nsamples, **kwargs):
    """Forward pass of the CSMC kernel (Algorithm 1 of the paper) -> ``(As, log_wss, uss)``.

    Same arguments as ``fbs.samplers.csmc.csmc.forward_pass``; ``ts`` must be the model's grid.
    """
    model = _model_of(transition_sampler, likelihood_logpdf)
    init = getattr(init_sampler, '__self__', None)
    if init is None or init is not getattr(init_likelihood_logpdf, '__self__', None):
        raise TypeError('init_sampler and init_likelihood_logpdf must be the methods of one DegenerateInit/NormalInit')
    scheme = _scheme_of(cond_resampling, 'conditional')
    host = is_host(key)
    single = np.ndim(key) == 1 if host else key.dim() == 1
    r = forward_pass_device(key, us_star, bs_star, vs, model, init, scheme, nsamples, history=True)
    res = (r['As'], r['log_wss'], r['uss'])
    if single:
        res = tuple(t[0] for t in res)
    return tuple(out(t, host) for t in res)



def csmc_step(model, k, step_keys, us_prev, log_ws, vs_k1, vs_k, u_star_k1, b_star_k, b_star_k1, cond_resampling):
    """One CSMC step (the scan body csmc.py:132-148) for particle sets held in global memory -- the per-timestep fused
    kernels (``fbs_csmc_step_affine_f32``: ancestors, fused transition + weight, normalise).  Batched over chains:
    ``step_keys [B, 2]`` (= keys[k] of csmc.py:157), ``us_prev [B, N, du]``, ``log_ws [B, N]`` normalised, ``vs_k1`` / ``vs_k``
    ``[B, dv]`` (v and v_prev), ``u_star_k1 [B, du]``, ``b_star_k`` / ``b_star_k1 [B]``.  Returns ``(A, us, log_ws)``."""
    if not isinstance(model, AffineGaussianModel):
        raise TypeError('csmc_step needs an AffineGaussianModel')
    scheme = _scheme_of(cond_resampling, 'conditional')
    host = is_host(step_keys)
    kk = dev(step_keys, torch.uint32).reshape(-1, 2)
    B = kk.shape[0]
    up = dev(us_prev, torch.float32).reshape(B, -1, model.du)
    N = up.shape[1]
    lw = dev(log_ws, torch.float32).reshape(B, N)
    v1, v0 = dev(vs_k1, torch.float32).reshape(B, model.dv), dev(vs_k, torch.float32).reshape(B, model.dv)
    ustar = dev(u_star_k1, torch.float32).reshape(B, model.du)
    _check_reference_indices(b_star_k, N)
    _check_reference_indices(b_star_k1, N)
    b0, b1 = dev(b_star_k, torch.int32).reshape(B), dev(b_star_k1, torch.int32).reshape(B)
    A = empty((B, N), torch.int32)
    us = empty((B, N, model.du), torch.float32)
    lw_out = empty((B, N), torch.float32)
    nat.call('fbs_csmc_step_affine_f32', stream(), model.struct(), int(k), scheme, ptr(kk), ptr(up), ptr(lw), ptr(v1), ptr(v0),
             ptr(ustar), ptr(b0), ptr(b1), B, N, ptr(A), ptr(us), ptr(lw_out))
    return tuple(out(t, host) for t in (A, us, lw_out))


def normalise(log_weights, log_space=False):
    """csmc.py:273-292 on device tensors / numpy (tiny helper, not on the fused path)."""
    host = is_host(log_weights)
    lw = dev(log_weights, torch.float32)
    lw = lw - torch.logsumexp(lw, dim=-1, keepdim=True)
    return out(lw if log_space else torch.exp(lw), host)


def barker_move(key, ws):
    return frandom.choice(key, ws.shape[-1], (), p=ws)


def backward_scanning_pass(key, As, xss, log_w_T):
    """csmc.py:230-270 -> ``(xs_star [.., K+1, du], bs_star [.., K+1])``; kernel ``fbs_backward_scan_f32``."""
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B = k.shape[0]
    A = dev(As, torch.int32)
    K, N = A.shape[-2], A.shape[-1]
    A = A.reshape(B, K, N)
    x = dev(xss, torch.float32)
    du = x.shape[-1]
    x = x.reshape(B, K + 1, N, du)
    lw = dev(log_w_T, torch.float32).reshape(B, N)
    xs = empty((B, K + 1, du), torch.float32)
    bs = empty((B, K + 1), torch.int32)
    nat.call('fbs_backward_scan_f32', stream(), ptr(k), ptr(A), ptr(x), ptr(lw), B, K, N, du, ptr(xs), ptr(bs))
    if single:
        xs, bs = xs[0], bs[0]
    return out(xs, host), out(bs, host)


def backward_sampling_pass(key, transition_logpdf, vs, ts, uss, log_ws, *args, **kwargs):
    """csmc.py:167-227.  Affine models: the whole recursion is one launch (``fbs_backward_sample_affine_f32``, mode 0);
    score-network models: one score evaluation per step."""
    model = _model_of(transition_logpdf)
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B = k.shape[0]
    x = dev(uss, torch.float32)
    du = x.shape[-1]
    K1, N = (x.shape[-3], x.shape[-2])
    x = x.reshape(B, K1, N, du)
    lw = dev(log_ws, torch.float32).reshape(B, K1, N)
    v = dev(vs, torch.float32).reshape(B, K1, model.dv)
    if not isinstance(model, ScoreNetModel):
        xs = empty((B, K1, du), torch.float32)
        Bs = empty((B, K1), torch.int32)
        vc, xc, lwc = v.contiguous(), x.contiguous(), lw.contiguous()             # referenced until the launch
        nat.call('fbs_backward_sample_affine_f32', stream(), model.struct(), 0, ptr(k), ptr(vc), ptr(xc), ptr(lwc), 0, B, N,
                 ptr(xs), ptr(Bs))
        if single:
            xs, Bs = xs[0], Bs[0]
        return out(xs, host), out(Bs, host)
    keys = frandom.split(k, K1)                                                    # csmc.py:194  [B, K1, 2]
    W_T = torch.exp(lw[:, -1] - torch.logsumexp(lw[:, -1], dim=-1, keepdim=True))  # csmc.py:200
    B_t = frandom.choice(keys[:, -1].contiguous(), N, (), p=W_T).reshape(B).long()
    ar = torch.arange(B, device=x.device)
    x_t = x[ar, -1, B_t]
    xs, Bs = [x_t], [B_t]
    for q, t in enumerate(range(K1 - 2, -1, -1)):                                  # csmc.py:217
        G = model.transition_logpdf(x_t[0], x[0, t], v[0, t], model.ts[t]).reshape(1, N)
        G = G - G.max(dim=-1, keepdim=True).values                                 # csmc.py:207
        lwt = G + lw[:, t]
        w = torch.exp(lwt - torch.logsumexp(lwt, dim=-1, keepdim=True))           # csmc.py:208-209
        B_t = frandom.choice(keys[:, q].contiguous(), N, (), p=w.contiguous()).reshape(B).long()
        x_t = x[ar, t, B_t]
        xs.append(x_t)
        Bs.append(B_t)
    xs = torch.stack(xs[::-1], dim=1)
    Bs = torch.stack(Bs[::-1], dim=1).to(torch.int32)
    if single:
        xs, Bs = xs[0], Bs[0]
    return out(xs, host), out(Bs, host)


def csmc_kernel(key, us_star, bs_star, vs, ts, init_sampler, init_likelihood_logpdf, transition_sampler,
                transition_logpdf, measurement_cond_logpdf, cond_resampling, nsamples, backward=False, **kwargs):
    """Generic cSMC kernel -> ``(xs_star, bs_star)``; same arguments as ``fbs.samplers.csmc.csmc.csmc_kernel``."""
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    keys = frandom.split(k.reshape(-1, 2), 2)                                      # csmc.py:65
    key_fwd, key_bwd = keys[:, 0].contiguous(), keys[:, 1].contiguous()
    model = _model_of(transition_sampler, measurement_cond_logpdf)
    init = getattr(init_sampler, '__self__', None)
    scheme = _scheme_of(cond_resampling, 'conditional')
    r = forward_pass_device(key_fwd, us_star, bs_star, vs, model, init, scheme, nsamples, history=True)
    if backward:
        xs, bs = backward_sampling_pass(key_bwd, transition_logpdf, dev(vs, torch.float32), ts, r['uss'], r['log_wss'])
    else:
        xs, bs = backward_scanning_pass(key_bwd, r['As'], r['uss'], r['log_wss'][:, -1].contiguous())
    if single:
        xs, bs = xs[0], bs[0]
    return out(xs, host), out(bs, host)
