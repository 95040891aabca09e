"""The fused sweep kernels vs the oracle.

A particle system is chaotic: one ancestor index that differs changes everything downstream.  The
sweep is therefore checked TEACHER-FORCED: the kernel returns its full history and every step k is
re-derived by the oracle from the kernel's own state at k-1 (csmc.py:132-148 / smc.py:138-152), so
each step is an independent parity check:
  * ancestors: exact, except where the float32 weight the kernel exponentiated differs from
    numpy's exp by an ULP (counted, must stay below 2e-4 of all draws);
  * particles: rtol 1e-5 / atol 2e-5 against the float64 drift + the bit-pinned float32 noise;
  * normalised log-weights: atol 2e-4.
"""
import numpy as np
import pytest
from oracle import jax_random as jr
from oracle import cond_resampling as ocr
from oracle import resampling as orx
from oracle import csmc as ocsmc
from oracle import gibbs as ogibbs
from helpers import gp_problem, oracle_model, product_model

pytestmark = pytest.mark.gpu

RTOL, ATOL, LW_ATOL = 1e-5, 2e-5, 2e-4


def _inputs(p, om, B, N, seed=0):
    K, d = p['K'], p['d']
    keys = jr.split(jr.PRNGKey(100 + seed), B)
    us_star = np.zeros((B, K + 1, d), np.float32)
    vs = np.zeros((B, K + 1, d), np.float32)
    x0s = jr.normal(jr.PRNGKey(200 + seed), (B, d))
    pk = jr.split(jr.PRNGKey(300 + seed), B)
    for b in range(B):
        path = om.fwd_sampler(pk[b], x0s[b], p['y0'])
        us_star[b], vs[b] = path[::-1, :d], path[::-1, d:]
    bs_star = np.stack([jr.randint(k, (K + 1,), 0, N) for k in jr.split(jr.PRNGKey(400 + seed), B)]).astype(np.int32)
    return keys, us_star, bs_star, vs


def _check_forward_history(p, om64, keys, us_star, bs_star, vs, As, log_wss, uss, scheme, init_normal):
    K, d = p['K'], p['d']
    B, N = As.shape[0], As.shape[2]
    ts = om64.ts
    resample = getattr(ocr, scheme)
    mismatches = 0
    for b in range(B):
        key_init, key_scan = jr.split(keys[b], 2)
        step_keys = jr.split(key_scan, K)
        # initial particles / weights (csmc.py:150-155)
        if init_normal:
            u0 = jr.normal(key_init, (N, d))
            u0[bs_star[b, 0]] = us_star[b, 0]
            np.testing.assert_allclose(uss[b, 0], u0, rtol=0, atol=5e-7)
            lw0 = ocsmc.normalise(om64.likelihood_logpdf(vs[b, 0].astype(np.float64), uss[b, 0].astype(np.float64),
                                                         vs[b, 1].astype(np.float64), ts[0]), log_space=True)
            np.testing.assert_allclose(log_wss[b, 0], lw0, atol=LW_ATOL)
        else:
            np.testing.assert_array_equal(uss[b, 0], np.broadcast_to(us_star[b, 0], (N, d)))
            np.testing.assert_allclose(log_wss[b, 0], -np.log(N), atol=1e-6)
        for k in range(K):
            key_res, key_tr = jr.split(step_keys[k], 2)
            w = np.exp(log_wss[b, k]).astype(np.float32)
            A_or = resample(key_res, w, int(bs_star[b, k]), int(bs_star[b, k + 1]), True)
            mismatches += int((A_or != As[b, k]).sum())
            assert As[b, k, bs_star[b, k + 1]] == bs_star[b, k]                  # csmc.py:139 / resamplings.py:86
            assert As[b, k].min() >= 0 and As[b, k].max() < N
            parents = uss[b, k][As[b, k]].astype(np.float64)
            mean = om64.transition_mean(parents, vs[b, k].astype(np.float64), ts[k])
            want = mean + np.float64(om64.transition_sd(ts[k])) * jr.normal(key_tr, (N, d)).astype(np.float64)
            want[bs_star[b, k + 1]] = us_star[b, k + 1]
            np.testing.assert_allclose(uss[b, k + 1], want, rtol=RTOL, atol=ATOL, err_msg=f'particles b={b} k={k}')
            np.testing.assert_array_equal(uss[b, k + 1, bs_star[b, k + 1]], us_star[b, k + 1])   # csmc.py:143
            lw = ocsmc.normalise(om64.likelihood_logpdf(vs[b, k + 1].astype(np.float64), parents,
                                                        vs[b, k].astype(np.float64), ts[k]), log_space=True)
            np.testing.assert_allclose(log_wss[b, k + 1], lw, atol=LW_ATOL, err_msg=f'log-weights b={b} k={k}')
    total = B * K * N
    assert mismatches <= max(1, int(2e-4 * total)), f'{mismatches} ancestor mismatches out of {total}'
    return mismatches


@pytest.mark.parametrize('d,N,K,B', [(1, 10, 12, 5), (3, 7, 10, 9), (10, 10, 20, 14), (10, 100, 8, 3), (16, 33, 6, 2),
                                     (100, 100, 4, 2), (12, 4, 9, 70), (5, 2, 6, 3)])
@pytest.mark.parametrize('scheme', ['killing', 'multinomial'])
def test_forward_pass_teacher_forced(d, N, K, B, scheme):
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    p = gp_problem(d, K=K)
    om32, om64 = oracle_model(p, np.float32), oracle_model(p, np.float64)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N)
    init = csmc.DegenerateInit(N)
    As, log_wss, uss = csmc.forward_pass(keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf,
                                         pm.transition_sampler, pm.likelihood_logpdf, getattr(R, scheme), N)
    assert As.shape == (B, K, N) and log_wss.shape == (B, K + 1, N) and uss.shape == (B, K + 1, N, d)
    _check_forward_history(p, om64, keys, us_star, bs_star, vs, As, log_wss, uss, scheme, False)
    # normalised weights
    np.testing.assert_allclose(np.log(np.exp(log_wss.astype(np.float64)).sum(-1)), 0., atol=1e-5)
    # unbatched call == row of the batched call (vmap semantics)
    A1, l1, u1 = csmc.forward_pass(keys[1], us_star[1], bs_star[1], vs[1], p['ts'], init.sampler,
                                   init.likelihood_logpdf, pm.transition_sampler, pm.likelihood_logpdf,
                                   getattr(R, scheme), N)
    np.testing.assert_array_equal(A1, As[1]); np.testing.assert_array_equal(u1, uss[1]); np.testing.assert_array_equal(l1, log_wss[1])


@pytest.mark.parametrize('d,N,K,B', [(2, 10, 10, 6), (10, 10, 12, 4)])
def test_forward_pass_explicit_final(d, N, K, B):
    """explicit_final=True: nparticles + 1 particles, N(0, I) start, likelihood weights with (v, v_prev) = (vs[0], vs[1])."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    p = gp_problem(d, K=K, sde_kind='lin')
    om32, om64 = oracle_model(p, np.float32), oracle_model(p, np.float64)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N + 1, seed=5)
    init = csmc.NormalInit(pm)
    As, log_wss, uss = csmc.forward_pass(keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf,
                                         pm.transition_sampler, pm.likelihood_logpdf, R.killing, N)
    assert As.shape == (B, K, N + 1)
    _check_forward_history(p, om64, keys, us_star, bs_star, vs, As, log_wss, uss, 'killing', True)


def test_forward_pass_vs_float32_reference_closures():
    """Against the reference-faithful float32 closures (Cholesky solve per call, gp_gibbs.py:78-81) the
    agreement is limited by THEIR float32 solve (cond(cov) ~ 1e3): rtol 2e-3 on the transition means."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    d, N, K, B = 10, 10, 10, 3
    p = gp_problem(d, K=K)
    om32 = oracle_model(p, np.float32)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N, seed=9)
    init = csmc.DegenerateInit(N)
    As, log_wss, uss = csmc.forward_pass(keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf,
                                         pm.transition_sampler, pm.likelihood_logpdf, R.killing, N)
    for b in range(B):
        step_keys = jr.split(jr.split(keys[b], 2)[1], K)
        for k in range(K):
            _, key_tr = jr.split(step_keys[k], 2)
            parents = uss[b, k][As[b, k]]
            want = om32.transition_sampler(parents, vs[b, k], om32.ts[k], key_tr)
            want[bs_star[b, k + 1]] = us_star[b, k + 1]
            np.testing.assert_allclose(uss[b, k + 1], want, rtol=2e-3, atol=2e-3)


def test_closures_match_oracle():
    """The three closures on their own (gp_gibbs.py:120-135) against the float64 oracle."""
    d, N = 10, 17
    p = gp_problem(d, K=20)
    om64 = oracle_model(p, np.float64)
    pm, _ = product_model(p)
    us = jr.normal(jr.PRNGKey(1), (N, d))
    v, vp, u = jr.normal(jr.PRNGKey(2), (d,)), jr.normal(jr.PRNGKey(3), (d,)), jr.normal(jr.PRNGKey(4), (d,))
    key = jr.PRNGKey(5)
    for k in (0, 7, 19):
        t = p['ts'][k]
        got = pm.transition_sampler(us, vp, t, key)
        want = om64.transition_mean(us.astype(np.float64), vp.astype(np.float64), om64.ts[k]) \
            + np.float64(om64.transition_sd(om64.ts[k])) * jr.normal(key, (N, d))
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(pm.likelihood_logpdf(v, us, vp, t),
                                   om64.likelihood_logpdf(v.astype(np.float64), us.astype(np.float64), vp.astype(np.float64), om64.ts[k]),
                                   rtol=2e-5, atol=1e-3)
        np.testing.assert_allclose(pm.transition_logpdf(u, us, vp, t),
                                   om64.transition_logpdf(u.astype(np.float64), us.astype(np.float64), vp.astype(np.float64), om64.ts[k]),
                                   rtol=2e-5, atol=1e-3)
    with pytest.raises(ValueError):
        pm.transition_sampler(us, vp, 0.123456, key)


@pytest.mark.parametrize('d,N,K,B', [(1, 10, 10, 4), (10, 100, 10, 3), (10, 25, 16, 7), (100, 100, 3, 2), (6, 8, 8, 150)])
@pytest.mark.parametrize('scheme', ['stratified', 'systematic', 'killing'])
def test_pmcmc_filter_step_teacher_forced(d, N, K, B, scheme):
    _check_pmcmc_filter(d, N, K, B, scheme)


@pytest.mark.parametrize('mode', ['csmc-killing', 'pmcmc-stratified'])
def test_bench_shape_teacher_forced_on_tensor_cores(mode, monkeypatch):
    """The benchmarked configuration at its own size -- d = 100, K = 200, N = 100 (bench.py; gp_gibbs.py / gp_pmcmc.py) --
    through the tcgen05 kernel PINNED (sweep_impl = 3 fails instead of falling back), an odd number of chains (the second
    warp group of the last CTA idles), every one of the 200 steps of every chain re-derived by the oracle: K = 200 cycles
    the TMA ring and the mbarrier phase bits ~50x more often than the K <= 4 shapes above."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v3')
    d, N, K, B = 100, 100, 200, 5
    if mode == 'pmcmc-stratified':
        _check_pmcmc_filter(d, N, K, B, 'stratified')
        return
    p = gp_problem(d, K=K)
    om32, om64 = oracle_model(p, np.float32), oracle_model(p, np.float64)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N, seed=77)
    init = csmc.DegenerateInit(N)
    As, log_wss, uss = csmc.forward_pass(keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf,
                                         pm.transition_sampler, pm.likelihood_logpdf, R.killing, N)
    assert As.shape == (B, K, N)
    _check_forward_history(p, om64, keys, us_star, bs_star, vs, As, log_wss, uss, 'killing', False)


def test_pinned_tensor_core_kernel_refuses_ineligible_shapes(monkeypatch):
    """sweep_impl = 3 must not silently fall back: an odd particle count is not eligible for the tcgen05 kernel."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v3')
    d, N, K, B = 8, 130, 3, 1
    p = gp_problem(d, K=K)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, oracle_model(p, np.float32), B, N)
    init = csmc.DegenerateInit(N)
    with pytest.raises(NotImplementedError):
        csmc.forward_pass(keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf, pm.transition_sampler,
                          pm.likelihood_logpdf, R.killing, N)


def _check_pmcmc_filter(d, N, K, B, scheme):
    from fbs_b200.samplers import smc, resampling as R
    p = gp_problem(d, K=K)
    om32, om64 = oracle_model(p, np.float32), oracle_model(p, np.float64)
    pm, _ = product_model(p)
    keys, _, _, vs = _inputs(p, om32, B, N, seed=2)
    u0s = np.stack([om32.ref_sampler(k, vs[b, 0], N) for b, k in enumerate(jr.split(jr.PRNGKey(55), B))])
    uT, log_ell, inds, lwh, ush = smc.pmcmc_filter_step(keys, vs, u0s, p['ts'], pm.transition_sampler,
                                                        pm.likelihood_logpdf, getattr(R, scheme), N,
                                                        return_history=True)
    ts = om64.ts
    mism = 0
    for b in range(B):
        step_keys = jr.split(keys[b], K)
        prev = u0s[b]
        acc = np.float32(0.)
        for k in range(K):
            key_prop, key_res = jr.split(step_keys[k], 2)                         # smc.py:142
            lw = om64.likelihood_logpdf(vs[b, k + 1].astype(np.float64), prev.astype(np.float64),
                                        vs[b, k].astype(np.float64), ts[k])
            np.testing.assert_allclose(lwh[b, k], lw, rtol=2e-5, atol=2e-3)
            c = ocsmc.logsumexp(lwh[b, k])
            acc = np.float32(np.float32(acc - np.float32(np.log(N))) + c)        # smc.py:146
            w = np.exp(lwh[b, k] - c).astype(np.float32)
            want_inds = getattr(orx, scheme)(w, key_res)
            mism += int((want_inds != inds[b, k]).sum())
            parents = prev[inds[b, k]].astype(np.float64)
            want = om64.transition_mean(parents, vs[b, k].astype(np.float64), ts[k]) \
                + np.float64(om64.transition_sd(ts[k])) * jr.normal(key_prop, (N, d))
            np.testing.assert_allclose(ush[b, k], want, rtol=RTOL, atol=ATOL)
            prev = ush[b, k]
        np.testing.assert_allclose(log_ell[b], acc, rtol=1e-5, atol=1e-3)
        np.testing.assert_array_equal(uT[b], ush[b, -1])
    assert mism <= max(1, int(2e-4 * B * K * N)), mism


def test_force_move_matches_oracle():
    from fbs_b200.samplers import gibbs
    rng = np.random.default_rng(3)
    for N in (2, 10, 101):
        B = 64
        w = rng.random((B, N)).astype(np.float32) ** 2
        w[0] = 0.; w[0, 1 % N] = 1.                      # degenerate: all mass on one particle
        w = (w / w.sum(1, keepdims=True, dtype=np.float32)).astype(np.float32)
        k = rng.integers(0, N, B).astype(np.int32)
        k[0] = 1 % N                                       # w_k == 1 -> uniform fallback branch (gibbs.py:204-205)
        keys = jr.split(jr.PRNGKey(N), B)
        idx, alpha = gibbs.force_move(keys, w, k)
        for b in range(B):
            i_or, a_or = ogibbs.force_move(keys[b], w[b], int(k[b]))
            assert idx[b] == i_or, (N, b)
            np.testing.assert_allclose(alpha[b], a_or, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('explicit_final', [False, True])
def test_gibbs_kernel_composition(explicit_final):
    """gibbs_kernel (gibbs.py:68-168) piece by piece against the oracle under the same keys."""
    from fbs_b200.samplers import gibbs_kernel
    from fbs_b200.samplers.csmc import csmc
    from fbs_b200.samplers.csmc.resamplings import killing
    from fbs_b200 import random as fr
    d, N, K, B = 4, 10, 15, 6
    p = gp_problem(d, K=K)
    om32 = oracle_model(p, np.float32)
    pm, sde = product_model(p)
    keys = jr.split(jr.PRNGKey(2024), B)
    x0 = jr.normal(jr.PRNGKey(1), (B, d))
    bs_star = np.stack([jr.randint(k, (K + 1,), 0, N) for k in jr.split(jr.PRNGKey(2), B)]).astype(np.int32)
    x0n, usn, bsn, changed = gibbs_kernel(keys, x0, p['y0'], None, bs_star, p['ts'], pm.fwd_sampler, sde, pm.unpack, N,
                                          pm.transition_sampler, pm.transition_logpdf, pm.likelihood_logpdf,
                                          explicit_backward=True, explicit_final=explicit_final)
    assert x0n.shape == (B, d) and usn.shape == (B, K + 1, d) and bsn.shape == (B, K + 1)
    np.testing.assert_array_equal(x0n, usn[:, -1])                                # gibbs.py:167
    np.testing.assert_array_equal(changed, bsn != bs_star)                        # gibbs.py:168
    for b in range(B):
        key_fwd, key_csmc, _ = jr.split(keys[b], 3)
        k_fwd2, k_x0, k_us, k_bs = jr.split(key_csmc, 4)
        np.testing.assert_array_equal(bsn[b], jr.randint(k_bs, (K + 1,), 0, N))  # gibbs.py:156
        # forward noising + reversal (gibbs.py:127-130): the kernel's path is 1e-5-close to the oracle's ...
        path = om32.fwd_sampler(key_fwd, x0[b], p['y0'])
        us, vs = pm.fwd_sampler_reversed(key_fwd, x0[b], p['y0'])
        np.testing.assert_allclose(us, path[::-1, :d], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(vs, path[::-1, d:], rtol=2e-5, atol=2e-6)
        # ... and the sweep + forced move on that path reproduce x0 exactly (gibbs.py:148-154)
        init = csmc.NormalInit(pm) if explicit_final else csmc.DegenerateInit(N)
        r = csmc.forward_pass_device(k_fwd2, us, bs_star[b], vs, pm, init, killing.scheme, N, history=False)
        lw_last, us_last = r['log_ws_last'][0].cpu().numpy(), r['us_last'][0].cpu().numpy()
        idx, _ = ogibbs.force_move(k_x0, np.exp(lw_last).astype(np.float32), int(bs_star[b, -1]))   # gibbs.py:152
        np.testing.assert_array_equal(x0n[b], us_last[idx])
        want_us = om32.fwd_sampler(k_us, x0n[b], p['y0'])[::-1, :d]              # gibbs.py:155
        np.testing.assert_allclose(usn[b], want_us, rtol=2e-5, atol=2e-6)


def test_gibbs_kernel_generic_fwd_sampler_equals_fast_path():
    """A user closure built from simulate_cond_forward + slicing gives the same sweep as the model's own sampler."""
    from fbs_b200.samplers import gibbs_kernel
    from fbs_b200 import sdes
    import torch
    d, N, K, B = 3, 10, 10, 5
    p = gp_problem(d, K=K)
    pm, sde = product_model(p)
    _, _, sim = sdes.make_linear_sde(sde)
    keys = jr.split(jr.PRNGKey(5), B)
    x0 = jr.normal(jr.PRNGKey(6), (B, d))
    bs = np.zeros((B, K + 1), np.int32)

    def fwd_sampler(key_, x0_, y0_):
        y = y0_.expand(x0_.shape[0], -1) if y0_.dim() == 2 and y0_.shape[0] == 1 else y0_
        return sim(key_, torch.cat([x0_, y], dim=-1), p['ts'])

    def unpack(xy):
        return xy[..., :d], xy[..., d:]

    a = gibbs_kernel(keys, x0, p['y0'], None, bs, p['ts'], fwd_sampler, sde, unpack, N, pm.transition_sampler,
                     pm.transition_logpdf, pm.likelihood_logpdf)
    b = gibbs_kernel(keys, x0, p['y0'], None, bs, p['ts'], pm.fwd_sampler, sde, pm.unpack, N, pm.transition_sampler,
                     pm.transition_logpdf, pm.likelihood_logpdf)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_opaque_closures_are_rejected():
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    p = gp_problem(2, K=4)
    pm, _ = product_model(p)
    init = csmc.DegenerateInit(4)
    args = (jr.PRNGKey(0), np.zeros((5, 2), np.float32), np.zeros(5, np.int32), np.zeros((5, 2), np.float32), p['ts'],
            init.sampler, init.likelihood_logpdf)
    with pytest.raises(TypeError):
        csmc.forward_pass(*args, lambda *a: None, pm.likelihood_logpdf, R.killing, 4)
    with pytest.raises(TypeError):
        csmc.forward_pass(*args, pm.transition_sampler, pm.likelihood_logpdf, lambda *a: None, 4)


def test_csmc_kernel_backward_scanning():
    """csmc_kernel(backward=False): ancestor tracing B_{t-1} = A_t[B_t] (csmc.py:260-264) against the oracle on
    the kernel's own forward history."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    d, N, K, B = 3, 10, 12, 8
    p = gp_problem(d, K=K)
    om32 = oracle_model(p, np.float32)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N, seed=4)
    init = csmc.DegenerateInit(N)
    xs, bs = csmc.csmc_kernel(keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf,
                              pm.transition_sampler, pm.transition_logpdf, pm.likelihood_logpdf, R.killing, N,
                              backward=False)
    kf = np.stack([jr.split(k, 2)[0] for k in keys])
    As, log_wss, uss = csmc.forward_pass(kf, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf,
                                         pm.transition_sampler, pm.likelihood_logpdf, R.killing, N)
    for b in range(B):
        key_bwd = jr.split(keys[b], 2)[1]
        xs_or, bs_or = ocsmc.backward_scanning_pass(key_bwd, As[b], uss[b], log_wss[b, -1])
        np.testing.assert_array_equal(bs[b], bs_or)
        np.testing.assert_array_equal(xs[b], xs_or)


@pytest.mark.parametrize('d,N,K,B', [(10, 10, 12, 9), (100, 100, 3, 3)])
def test_general_and_tiled_sweep_kernels_agree(d, N, K, B, monkeypatch):
    """The general kernel (csmc_kernels.cu) and the tiled TMA kernel (sweep_v2.cu) are two implementations of the
    same sweep: identical random streams, ancestors equal, floats equal up to summation order."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    p = gp_problem(d, K=K)
    om32 = oracle_model(p, np.float32)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N, seed=11)
    init = csmc.DegenerateInit(N)
    args = (keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf, pm.transition_sampler,
            pm.likelihood_logpdf, R.killing, N)
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v1')
    A1, l1, u1 = csmc.forward_pass(*args)
    monkeypatch.delenv('FBS_SWEEP_IMPL')
    A2, l2, u2 = csmc.forward_pass(*args)
    # step 1 is computed from identical inputs by both kernels
    np.testing.assert_array_equal(A1[:, 0], A2[:, 0])
    np.testing.assert_allclose(u1[:, 1], u2[:, 1], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(l1[:, 1], l2[:, 1], atol=1e-4)
    assert (A1 == A2).mean() > 0.99


@pytest.mark.parametrize('delta', [None, 0.3])
def test_pmcmc_kernel_composition(delta):
    """pmcmc_kernel (smc.py:171-258) piece by piece against the oracle under the same keys: proposal path (fresh or
    pCN), reference draw, filter, Metropolis--Hastings decision and state selection."""
    from fbs_b200.samplers import pmcmc_kernel, pmcmc_filter_step, stratified
    from oracle import smc as osmc
    d, N, K, B = 3, 16, 12, 24
    p = gp_problem(d, K=K)
    om32 = oracle_model(p, np.float32)
    pm, sde = product_model(p)
    keys = jr.split(jr.PRNGKey(77), B)
    uT = jr.normal(jr.PRNGKey(1), (B, d))
    ys = np.stack([om32.fwd_ys_sampler(k, p['y0']) for k in jr.split(jr.PRNGKey(2), B)])
    # half the chains start from a huge log_ell (always reject), half from a tiny one (always accept), rest in between
    log_ell = np.concatenate([np.full(B // 3, 1e30), np.full(B // 3, -1e30), np.zeros(B - 2 * (B // 3))]).astype(np.float32)
    uT2, le2, ys2, st = pmcmc_kernel(keys, uT, log_ell, ys, p['y0'], p['ts'], pm.fwd_ys_sampler, sde, pm.ref_sampler,
                                     pm.transition_sampler, pm.likelihood_logpdf, stratified, N, delta=delta)
    for b in range(B):
        key_prop, key_u0, key_filter, key_mh = jr.split(keys[b], 4)
        if delta is None:
            prop_ys = om32.fwd_ys_sampler(key_prop, p['y0'])
        else:
            mean = np.stack([om32.sde.mean(t, om32.ts[0], p['y0']) for t in om32.ts]).astype(np.float32)
            prop_ys = osmc.pcn_proposal(key_prop, delta, ys[b], mean, lambda k_: om32.fwd_ys_sampler(k_, p['y0']))
        vs = prop_ys[::-1]
        u0s = om32.ref_sampler(key_u0, vs[0], N)
        # the filter is chaotic w.r.t. 1e-6 input differences, so run OUR filter on the oracle's inputs for the evidence
        puT, ple = pmcmc_filter_step(key_filter, vs.copy(), u0s, p['ts'], pm.transition_sampler, pm.likelihood_logpdf,
                                     stratified, N)
        np.testing.assert_allclose(st.prop_log_ell[b], ple, rtol=2e-4, atol=2e-2)
        z = jr.uniform(key_mh, ())
        log_acc = min(0., float(st.prop_log_ell[b]) - float(log_ell[b]))
        acc = bool(np.log(z) < log_acc)
        assert bool(st.is_accepted[b]) == acc
        np.testing.assert_allclose(st.acceptance_prob[b], np.exp(log_acc), rtol=1e-5, atol=1e-30)
        assert st.log_ell[b] == log_ell[b]
        if acc:
            np.testing.assert_allclose(ys2[b], prop_ys, rtol=2e-5, atol=2e-5)
            assert le2[b] == st.prop_log_ell[b]
        else:
            np.testing.assert_array_equal(ys2[b], ys[b])
            np.testing.assert_array_equal(uT2[b], uT[b])
            assert le2[b] == log_ell[b]
    assert 0 < st.is_accepted.sum() < B


def test_mh_accept_rejects_nan_evidence():
    """smc.py:246-249: ``jnp.minimum(0, nan)`` is nan and ``log z < nan`` is False -- a proposal whose evidence is NaN (all
    weights -inf, a non-finite score) is REJECTED and the chain keeps its last finite state.  (CUDA's fminf would drop the NaN
    and accept every such proposal.)"""
    import torch
    from fbs_b200 import _native as nat
    from fbs_b200._tensor import ptr, stream
    B, N, du, ny = 6, 4, 3, 10
    keys = torch.from_numpy(jr.split(jr.PRNGKey(9), B)).cuda()
    prop_uTs = torch.randn(B, N, du, device='cuda')
    prop_ys = torch.randn(B, ny, device='cuda')
    ple = torch.tensor([float('nan'), 5., float('nan'), float('inf'), -float('inf'), float('nan')], device='cuda')
    le0 = torch.tensor([0., 0., float('nan'), 0., 0., -float('inf')], device='cuda')
    uT, ys, le = torch.zeros(B, du, device='cuda'), torch.zeros(B, ny, device='cuda'), le0.clone()
    acc_prob = torch.empty(B, device='cuda')
    is_acc = torch.empty(B, dtype=torch.uint8, device='cuda')
    nat.call('fbs_mh_accept_f32', stream(), ptr(keys), ptr(prop_uTs), ptr(ple), ptr(prop_ys), B, N, du, ny, 0, ptr(uT), ptr(le),
             ptr(ys), ptr(acc_prob), ptr(is_acc))
    acc = is_acc.cpu().numpy().astype(bool)
    np.testing.assert_array_equal(acc, [False, True, False, True, False, False])
    for b in (0, 2, 4, 5):                                  # rejected: state untouched
        assert torch.equal(uT[b], torch.zeros(du, device='cuda')) and torch.equal(ys[b], torch.zeros(ny, device='cuda'))
        assert torch.equal(le[b:b + 1], le0[b:b + 1]) or (torch.isnan(le[b]) and torch.isnan(le0[b]))
    assert torch.isnan(acc_prob[[0, 2, 5]]).all()
    assert torch.equal(uT[1], prop_uTs[1, 0]) and float(le[1]) == 5.


def test_pmcmc_kernel_host_pipeline_equals_unchunked(monkeypatch):
    """Host-buffer pmcmc_kernel on many chains is chunked over CUDA streams (H2D / kernels / D2H overlapped); chains are
    independent, so the result must equal the unchunked call bit for bit."""
    from fbs_b200.samplers import pmcmc_kernel, stratified
    from fbs_b200.samplers import smc as psmc
    d, N, K, B = 4, 16, 10, 96
    p = gp_problem(d, K=K)
    om32 = oracle_model(p, np.float32)
    pm, sde = product_model(p)
    keys = jr.split(jr.PRNGKey(5), B)
    uT = jr.normal(jr.PRNGKey(1), (B, d))
    ys = np.stack([om32.fwd_ys_sampler(k, p['y0']) for k in jr.split(jr.PRNGKey(2), B)])
    log_ell = np.zeros(B, np.float32)
    kw = dict(ts=p['ts'], fwd_ys_sampler=pm.fwd_ys_sampler, sde=sde, ref_sampler=pm.ref_sampler,
              transition_sampler=pm.transition_sampler, likelihood_logpdf=pm.likelihood_logpdf, resampling=stratified,
              nparticles=N, delta=0.3)
    a = pmcmc_kernel(keys, uT, log_ell, ys, p['y0'], **kw)            # unchunked (also warms the model)
    monkeypatch.setattr(psmc, 'PIPELINE_MIN_CHAINS', 32)
    b = pmcmc_kernel(keys, uT, log_ell, ys, p['y0'], **kw)            # 3 chunks of 32 chains
    for x, y in zip(list(a[:3]) + list(a[3]), list(b[:3]) + list(b[3])):
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))


@pytest.mark.parametrize('d,N,K,B', [(8, 128, 3, 3), (4, 2, 2, 1), (100, 64, 3, 5), (52, 100, 3, 2), (124, 32, 2, 2),
                                     (104, 96, 2, 3), (20, 30, 5, 301)])
def test_forward_pass_tensor_core_kernel_shapes(d, N, K, B, monkeypatch):
    """Edge shapes of the tcgen05 sweep kernel (sweep_v3.cu): full 128 MMA rows, a single tiny chain, odd chain counts
    (the second warp group runs one chain fewer), padded K / N dimensions, the widest accumulator (2 x 128 columns; at
    d = 124 the operand tiles of N = 100 particles no longer fit shared memory, so N = 32), the widest state whose noise tasks still fit
    the 12 noise warps (d = 104 at N = 96), and more chain pairs than SMs.  The kernel is PINNED (an ineligible shape fails instead of falling
    back).  Teacher-forced against the oracle, and equal ancestors to the general kernel."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    p = gp_problem(d, K=K)
    om32, om64 = oracle_model(p, np.float32), oracle_model(p, np.float64)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N, seed=21)
    init = csmc.DegenerateInit(N)
    args = (keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf, pm.transition_sampler,
            pm.likelihood_logpdf, R.killing, N)
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v3')
    As, log_wss, uss = csmc.forward_pass(*args)
    nb = min(B, 4)
    _check_forward_history(p, om64, keys[:nb], us_star[:nb], bs_star[:nb], vs[:nb], As[:nb], log_wss[:nb], uss[:nb], 'killing', False)
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v1')
    A1, l1, u1 = csmc.forward_pass(*args)
    np.testing.assert_array_equal(A1[:, 0], As[:, 0])
    np.testing.assert_allclose(u1[:, 1], uss[:, 1], rtol=1e-5, atol=1e-5)


def test_forward_pass_explicit_final_on_tensor_cores(monkeypatch):
    """explicit_final=True with an even particle count (nparticles = 9 -> 10 rows) takes the tcgen05 kernel, including
    its extra GEMM for the initial weights."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    d, N, K, B = 8, 9, 6, 5
    p = gp_problem(d, K=K, sde_kind='lin')
    om32, om64 = oracle_model(p, np.float32), oracle_model(p, np.float64)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N + 1, seed=8)
    init = csmc.NormalInit(pm)
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v3')
    As, log_wss, uss = csmc.forward_pass(keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf,
                                         pm.transition_sampler, pm.likelihood_logpdf, R.killing, N)
    assert As.shape == (B, K, N + 1)
    _check_forward_history(p, om64, keys, us_star, bs_star, vs, As, log_wss, uss, 'killing', True)


def test_empty_batch_is_a_no_op():
    """Zero chains: every batched entry point returns empty outputs and launches nothing that faults."""
    import torch
    from fbs_b200.samplers import pmcmc_filter_step, stratified
    from fbs_b200.samplers.csmc import resamplings as R
    from fbs_b200 import random as fr
    p = gp_problem(4, K=5)
    pm, _ = product_model(p)
    keys = torch.empty((0, 2), dtype=torch.uint32, device='cuda')
    uT, le = pmcmc_filter_step(keys, torch.empty((0, 6, 4), device='cuda'), torch.empty((0, 8, 4), device='cuda'), p['ts'],
                               pm.transition_sampler, pm.likelihood_logpdf, stratified, 8)
    assert tuple(uT.shape) == (0, 8, 4) and tuple(le.shape) == (0,)
    idx = R.killing(keys, torch.empty((0, 16), device='cuda'), 0, 0, True)
    assert tuple(idx.shape) == (0, 16)
    assert tuple(fr.normal(keys, (3,)).shape) == (0, 3)
    torch.cuda.synchronize()


@pytest.mark.parametrize('d,N,B', [(10, 64, 5), (10, 1000, 3), (6, 16, 4), (2, 2, 2), (20, 32, 3), (10, 33, 2),
                                   (100, 256, 3), (100, 100, 2), (32, 130, 2), (40, 64, 3), (100, 1026, 2), (32, 2, 1), (128, 66, 2),
                                   (36, 4, 150)])
@pytest.mark.parametrize('scheme', ['killing', 'multinomial'])
def test_per_timestep_step_kernels(d, N, B, scheme):
    """fbs_csmc_step_affine_f32 (the per-timestep kernels for particle sets in global memory; register-resident
    transition kernel for du, dv <= 16, the tcgen05 kernel for du >= 32 and even N, the general one otherwise) against
    the oracle's scan body (csmc.py:132-148)."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    K = 6
    p = gp_problem(d, K=K)
    om64 = oracle_model(p, np.float64)
    pm, _ = product_model(p)
    rng = np.random.default_rng(d * 1000 + N)
    k = 3
    step_keys = jr.split(jr.PRNGKey(N), B)
    us_prev = rng.standard_normal((B, N, d)).astype(np.float32)
    lw = rng.standard_normal((B, N)).astype(np.float32)
    lw = (lw - np.log(np.exp(lw.astype(np.float64)).sum(-1, keepdims=True))).astype(np.float32)
    v0, v1 = rng.standard_normal((B, d)).astype(np.float32), rng.standard_normal((B, d)).astype(np.float32)
    ustar = rng.standard_normal((B, d)).astype(np.float32)
    b0 = rng.integers(0, N, size=B).astype(np.int32)
    b1 = rng.integers(0, N, size=B).astype(np.int32)
    A, us, lw_out = csmc.csmc_step(pm, k, step_keys, us_prev, lw, v1, v0, ustar, b0, b1, getattr(R, scheme))
    for b in range(B):
        key_res, key_tr = jr.split(step_keys[b], 2)
        want_A = getattr(ocr, scheme)(key_res, np.exp(lw[b]).astype(np.float32), b0[b], b1[b], True)
        assert (A[b] == want_A).mean() >= 1 - 2e-3            # expf ULP ties, as in the sweep tests
        parents = us_prev[b][A[b]].astype(np.float64)
        want = om64.transition_mean(parents, v0[b].astype(np.float64), om64.ts[k]) \
            + np.float64(om64.transition_sd(om64.ts[k])) * jr.normal(key_tr, (N, d))
        want[b1[b]] = ustar[b]
        np.testing.assert_allclose(us[b], want, rtol=1e-5, atol=2e-5)
        np.testing.assert_array_equal(us[b][b1[b]], ustar[b])
        want_lw = ocsmc.normalise(om64.likelihood_logpdf(v1[b].astype(np.float64), parents, v0[b].astype(np.float64), om64.ts[k]),
                                  log_space=True)
        np.testing.assert_allclose(lw_out[b], want_lw, atol=LW_ATOL)


@pytest.mark.parametrize('d,N,B', [(100, 256, 4), (64, 100, 3), (36, 2050, 2)])
def test_step_tensor_core_vs_cuda_core(d, N, B, monkeypatch):
    """The tcgen05 per-timestep kernel (split-TF32 GEMM in TMEM) against the CUDA-core kernel of the same entry point:
    identical ancestors and noise, drift within float32 rounding."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    p = gp_problem(d, K=4)
    pm, _ = product_model(p)
    rng = np.random.default_rng(d + N)
    step_keys = jr.split(jr.PRNGKey(N + 1), B)
    us_prev = rng.standard_normal((B, N, d)).astype(np.float32)
    lw = np.full((B, N), -np.log(N), np.float32)
    v0, v1 = rng.standard_normal((B, d)).astype(np.float32), rng.standard_normal((B, d)).astype(np.float32)
    ustar = rng.standard_normal((B, d)).astype(np.float32)
    b0 = rng.integers(0, N, size=B).astype(np.int32)
    b1 = rng.integers(0, N, size=B).astype(np.int32)
    got = csmc.csmc_step(pm, 2, step_keys, us_prev, lw, v1, v0, ustar, b0, b1, R.killing)
    monkeypatch.setenv('FBS_STEP_IMPL', 'cuda')
    want = csmc.csmc_step(pm, 2, step_keys, us_prev, lw, v1, v0, ustar, b0, b1, R.killing)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_allclose(got[1], want[1], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(got[2], want[2], atol=2e-4)


@pytest.mark.parametrize('d,N,B', [(12, 10, 80), (100, 100, 70), (8, 7, 300), (10, 16, 90)])
def test_ref_sampler_tiled_and_plain_match_oracle(d, N, B):
    """ref_sampler (gp_gibbs.py:138-141; fbs_gaussian_ref_sample_f32): the register-tiled kernel (du % 4 == 0, >= 64 chains)
    and the plain one (d = 10) against the oracle: noise bit-pinned, Cholesky product within float32 rounding."""
    p = gp_problem(d, K=4)
    om32 = oracle_model(p, np.float32)
    pm, _ = product_model(p)
    rng = np.random.default_rng(d + N)
    keys = jr.split(jr.PRNGKey(d * 7 + N), B)
    yT = rng.standard_normal((B, d)).astype(np.float32)
    got = pm.ref_sampler(keys, yT, N)
    assert got.shape == (B, N, d)
    for b in (0, 1, B // 2, B - 1):
        np.testing.assert_allclose(got[b], om32.ref_sampler(keys[b], yT[b], N), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize('d,N,K,B', [(10, 64, 12, 37), (4, 64, 5, 3), (16, 64, 4, 2), (10, 128, 6, 5), (12, 128, 3, 2), (7, 64, 6, 20),
                                     (10, 64, 260, 2)])
def test_forward_pass_warp_per_chain_kernel(d, N, K, B, monkeypatch):
    """sweep_warp.cu (one warp per chain, particles in registers; the Gaussian-SB shape d = 10, N = 64): pinned, teacher-forced
    against the oracle in both CSMC initialisations, and equal ancestors / particles to the general kernel.  K = 260 takes the
    variant that reads the step matrices through L1 instead of shared memory."""
    from fbs_b200.samplers.csmc import csmc, resamplings as R
    p = gp_problem(d, K=K)
    om32, om64 = oracle_model(p, np.float32), oracle_model(p, np.float64)
    pm, _ = product_model(p)
    keys, us_star, bs_star, vs = _inputs(p, om32, B, N, seed=41)
    init = csmc.DegenerateInit(N)
    args = (keys, us_star, bs_star, vs, p['ts'], init.sampler, init.likelihood_logpdf, pm.transition_sampler,
            pm.likelihood_logpdf, R.killing, N)
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v4')
    As, log_wss, uss = csmc.forward_pass(*args)
    nb = min(B, 3) if K < 100 else 1
    _check_forward_history(p, om64, keys[:nb], us_star[:nb], bs_star[:nb], vs[:nb], As[:nb], log_wss[:nb], uss[:nb], 'killing', False)
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v1')
    A1, l1, u1 = csmc.forward_pass(*args)
    np.testing.assert_array_equal(A1[:, 0], As[:, 0])
    np.testing.assert_allclose(u1[:, 1], uss[:, 1], rtol=1e-5, atol=1e-5)
    assert (A1 == As).mean() > 0.99
    if K > 100:
        return
    # explicit_final (N - 1 = nparticles; the sweep carries N rows) and the multinomial scheme
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v4')
    ninit = csmc.NormalInit(pm)
    As2, lw2, us2 = csmc.forward_pass(keys, us_star, bs_star, vs, p['ts'], ninit.sampler, ninit.likelihood_logpdf,
                                      pm.transition_sampler, pm.likelihood_logpdf, R.multinomial, N - 1)
    assert As2.shape == (B, K, N)
    _check_forward_history(p, om64, keys[:2], us_star[:2], bs_star[:2], vs[:2], As2[:2], lw2[:2], us2[:2], 'multinomial', True)


@pytest.mark.parametrize('scheme', ['stratified', 'killing'])
def test_pmcmc_filter_warp_per_chain_kernel(scheme, monkeypatch):
    monkeypatch.setenv('FBS_SWEEP_IMPL', 'v4')
    _check_pmcmc_filter(10, 64, 9, 21, scheme)
    _check_pmcmc_filter(8, 128, 4, 3, scheme)


def test_gibbs_kernel_host_pipeline_equals_unchunked(monkeypatch):
    """Host-buffer gibbs_kernel on many chains is chunked over CUDA streams (H2D / kernels / D2H overlapped, chunk sizes
    rounded to whole waves of the sweep kernel); chains are independent, so the result equals the unchunked call bit for bit."""
    from fbs_b200.samplers import gibbs_kernel
    from fbs_b200.samplers import smc as psmc
    d, N, K, B = 4, 16, 10, 100
    p = gp_problem(d, K=K)
    pm, sde = product_model(p)
    keys = jr.split(jr.PRNGKey(5), B)
    x0 = jr.normal(jr.PRNGKey(1), (B, d))
    bs = np.stack([jr.randint(k, (K + 1,), 0, N) for k in jr.split(jr.PRNGKey(2), B)]).astype(np.int32)
    args = (keys, x0, p['y0'], None, bs, p['ts'], pm.fwd_sampler, sde, pm.unpack, N, pm.transition_sampler, pm.transition_logpdf,
            pm.likelihood_logpdf)
    a = gibbs_kernel(*args)                                           # unchunked (also warms the model)
    monkeypatch.setattr(psmc, 'PIPELINE_MIN_CHAINS', 32)
    assert len(psmc._chunk_bounds(B)) > 1
    b = gibbs_kernel(*args)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    with pytest.raises(ValueError):
        gibbs_kernel(keys, x0, p['y0'], None, bs + N, p['ts'], pm.fwd_sampler, sde, pm.unpack, N, pm.transition_sampler,
                     pm.transition_logpdf, pm.likelihood_logpdf)   # stale reference indices are refused on the host
