"""twisted_smc (fbs/samplers/smc.py:261-309) with the closures of experiments/toy/gp_twisted.py:66-129 on the CUDA path:
teacher-forced against the oracle restatement (oracle/smc.py twisted_smc + oracle/models.py TwistedGaussianModel, float64
closures; indices exact up to float32 exp ties, particles rtol 1e-5, normalised log-weights atol 2e-3 + rtol 1e-5) and a statistical
check of the conditional samples against the GP-regression posterior the toy driver stores next to them (gp_twisted.py:50-53,151-152)."""
import numpy as np
import pytest
from oracle import jax_random as jr
from oracle import models as om, sdes as osd, resampling as orx, csmc as ocsmc

pytestmark = pytest.mark.gpu


def _problem(d, K, kind='const'):
    import fbs_b200
    from fbs_b200 import sdes
    cov_mat, _, _ = om.gp_regression_setup(d)
    _, y0 = om.gp_draw_y0(jr.PRNGKey(5), d, cov_mat)
    T = 1.
    ts = np.linspace(0., T, K + 1)
    if kind == 'lin':
        osde, psde = osd.StationaryLinLinearSDE(0.02, 4., 0., T), sdes.StationaryLinLinearSDE(0.02, 4., 0., T)
    else:
        osde, psde = osd.StationaryConstLinearSDE(a=-0.5, b=1.), sdes.StationaryConstLinearSDE(a=-0.5, b=1.)
    omod = om.TwistedGaussianModel(osde, np.zeros(d), cov_mat, 1.0, ts, T, np.float64)
    pmod = fbs_b200.TwistedAffineModel(psde, np.zeros(d), cov_mat, 1.0, ts, T)
    return omod, pmod, y0, ts, cov_mat


@pytest.mark.parametrize('d,N,K,B,kind', [(1, 10, 8, 4, 'const'), (3, 50, 12, 5, 'const'), (10, 100, 6, 3, 'lin'), (100, 100, 20, 2, 'const'),
                                          (7, 33, 5, 150, 'const')])
@pytest.mark.parametrize('scheme', ['stratified', 'killing'])
def test_twisted_smc_teacher_forced(d, N, K, B, kind, scheme):
    from fbs_b200.samplers import twisted_smc, resampling as R
    omod, pmod, y0, ts, _ = _problem(d, K, kind)
    keys = jr.split(jr.PRNGKey(17 + d), B)
    xs, lw, inds, xh, lwh = twisted_smc(keys, y0, ts, pmod.init_sampler, pmod.transition_logpdf, pmod.twisting_logpdf,
                                        pmod.twisting_prop_sampler, pmod.twisting_prop_logpdf, getattr(R, scheme), N,
                                        return_history=True)
    assert xs.shape == (B, N, d) and lw.shape == (B, N) and inds.shape == (B, K, N)
    np.testing.assert_array_equal(xs, xh[:, -1])
    np.testing.assert_array_equal(lw, lwh[:, -1])
    y64 = y0.astype(np.float64)
    mism = 0
    for b in range(min(B, 4)):
        key_init, key_filter = jr.split(keys[b])                                  # smc.py:296
        # smc.py:299 -- the kernel starts from the float32 init of the product model: take ITS particles for the teacher forcing
        x_prev = np.asarray(pmod.init_sampler(key_init, N), np.float64)
        np.testing.assert_allclose(x_prev, omod.init_sampler(key_init, N), rtol=2e-5, atol=2e-5)
        lps_prev = omod.twisting_logpdf(y64, x_prev, omod.ts[0])
        lw_prev = ocsmc.normalise(lps_prev.astype(np.float32), log_space=True)
        for k, key_step in enumerate(jr.split(key_filter, K)):
            t = omod.ts[k + 1]
            key_res, key_prop = jr.split(key_step)                                # smc.py:280
            want_inds = getattr(orx, scheme)(np.exp(lw_prev).astype(np.float32), key_res)
            mism += int((want_inds != inds[b, k]).sum())
            a = inds[b, k]
            xp, lpp = x_prev[a], lps_prev[a]                                      # smc.py:284-285
            x_new = omod.twisting_prop_sampler(key_prop, xp, t, y64)              # smc.py:288
            np.testing.assert_allclose(xh[b, k], x_new, rtol=1e-5, atol=2e-5, err_msg=f'particles b={b} k={k}')
            x_new = xh[b, k].astype(np.float64)                                   # teacher forcing: the kernel's own particles
            lps = omod.twisting_logpdf(y64, x_new, t)
            lws = omod.transition_logpdf(x_new, xp, t) + lps - omod.twisting_prop_logpdf(x_new, xp, t, y64) - lpp   # :291-293
            want_lw = lws - ocsmc.logsumexp(lws)
            # (float32 evaluation of terms of size max|lws|: their rounding, 2^-24 relative each, is the floor of the agreement)
            np.testing.assert_allclose(lwh[b, k], want_lw, rtol=1e-5, atol=2e-3 + 4e-6 * float(np.abs(lws).max()),
                                       err_msg=f'log-weights b={b} k={k}')
            x_prev, lps_prev, lw_prev = x_new, lps, lwh[b, k]
    assert mism <= max(1, int(3e-4 * min(B, 4) * K * N)), mism
    # one key == row of the batch
    x1, l1 = twisted_smc(keys[1], y0, ts, pmod.init_sampler, pmod.transition_logpdf, pmod.twisting_logpdf,
                         pmod.twisting_prop_sampler, pmod.twisting_prop_logpdf, getattr(R, scheme), N)
    np.testing.assert_array_equal(x1, xs[1])
    np.testing.assert_array_equal(l1, lw[1])


def test_twisted_smc_conditional_samples_match_gp_posterior():
    """gp_twisted.py:134-147: one conditional sample per sampler run (choice over the final weights); 4096 runs in one batch,
    d = 5, K = 200, 100 particles.  The twisted sampler is consistent, not exact: mean within 0.1, marginal variances within 20 %."""
    from fbs_b200.samplers import twisted_smc, stratified
    from fbs_b200 import random as fr
    d, K, N, B = 5, 200, 100, 4096
    omod, pmod, y0, ts, cov_mat = _problem(d, K)
    gp_mean, gp_cov = om.gp_posterior(cov_mat, y0)
    keys = fr.split(fr.PRNGKey(3), B)
    kk = fr.split(keys, 2)                                                        # key_filter, key_select (:135)
    xs, lw = twisted_smc(np.ascontiguousarray(kk[:, 0]), y0, ts, pmod.init_sampler, pmod.transition_logpdf, pmod.twisting_logpdf,
                         pmod.twisting_prop_sampler, pmod.twisting_prop_logpdf, stratified, N)
    pick = fr.choice(np.ascontiguousarray(kk[:, 1]), N, (), p=np.exp(lw))
    S = xs[np.arange(B), pick]
    np.testing.assert_allclose(S.mean(0), gp_mean, atol=1e-1)
    np.testing.assert_allclose(S.var(0), np.diag(gp_cov), rtol=2e-1)


def test_twisted_smc_rejects_opaque_closures():
    from fbs_b200.samplers import twisted_smc, stratified
    _, pmod, y0, ts, _ = _problem(2, 4)
    with pytest.raises(TypeError):
        twisted_smc(jr.PRNGKey(0), y0, ts, pmod.init_sampler, lambda *a: None, pmod.twisting_logpdf, pmod.twisting_prop_sampler,
                    pmod.twisting_prop_logpdf, stratified, 8)
