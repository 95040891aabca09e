"""The oracle against what the reference's own tests assert: closed forms exactly, samplers
distributionally.  Sizes are trimmed so the whole CPU suite runs in a few minutes; tolerances are the
reference's, widened only where the Monte-Carlo sample is smaller (stated inline)."""
import numpy as np
import pytest
from oracle import jax_random as jr
from oracle import cond_resampling as ocr
from oracle import resampling as orx
from oracle import sdes as osdes
from oracle import gibbs as ogibbs
from oracle import smc as osmc
from oracle import csmc as ocsmc
from oracle.models import JointGaussianDiffusionModel


@pytest.mark.parametrize('name', ['multinomial', 'killing'])
@pytest.mark.parametrize('j', [0, 5])
def test_conditional_resampling_bayes(name, j):
    """tests/test_cond_resamplings.py:33-53 with 6 000 keys instead of 100 000 (atol scaled 1e-3 -> 4e-3)."""
    N = 100
    weights = (np.cos(np.linspace(0, 2 * np.pi, N)) + 1).astype(np.float32)
    weights /= weights.sum()
    keys = jr.split(jr.PRNGKey(42), 6000)
    fn = getattr(ocr, name)
    counts = np.zeros(N)
    for key in keys:
        k1, k2 = jr.split(key)
        i = int(jr.choice(k1, N, (), p=weights))
        idx = fn(k2, weights, i, j, True)
        assert idx[j] == i                                         # :51
        counts += np.bincount(idx[1:], minlength=N)
    np.testing.assert_allclose(counts / counts.sum(), weights, atol=4e-3)


@pytest.mark.parametrize('name', ['stratified', 'systematic', 'killing', 'multinomial'])
def test_unconditional_resampling_unbiased(name):
    N = 50
    rng = np.random.default_rng(1)
    w = rng.random(N).astype(np.float32)
    w /= w.sum()
    counts = np.zeros(N)
    for key in jr.split(jr.PRNGKey(7), 3000):
        idx = getattr(orx, name)(w, key)
        assert idx.min() >= 0 and idx.max() < N
        counts += np.bincount(idx, minlength=N)
    np.testing.assert_allclose(counts / counts.sum(), w, atol=4e-3)


def test_killing_is_identity_on_uniform_weights():
    w = np.full(20, 1 / 20, np.float32)
    np.testing.assert_array_equal(ocr.killing(jr.PRNGKey(0), w, 3, 3, True), np.arange(20))
    np.testing.assert_array_equal(orx.killing(w, jr.PRNGKey(0)), np.arange(20))


def test_conditional_systematic_raises_like_upstream():
    with pytest.raises(NotImplementedError):
        ocr.systematic(jr.PRNGKey(0), np.full(4, .25, np.float32), 0, 0, True)


def test_discretisation_closed_forms():
    """tests/test_sdes.py:18-34,60-90."""
    a, b = -0.5, 1.
    disc, _, _ = osdes.make_linear_sde(osdes.StationaryConstLinearSDE(a, b), np.float64)
    F, Q = disc(0.7, 0.2)
    np.testing.assert_allclose(F, np.exp(a * 0.5), rtol=1e-12)
    np.testing.assert_allclose(Q, b ** 2 / (2 * a) * (np.exp(2 * a * 0.5) - 1), rtol=1e-12)
    lin = osdes.StationaryLinLinearSDE(0.02, 5., 0., 2.)
    disc, _, _ = osdes.make_linear_sde(lin, np.float64)
    grid = np.linspace(0.4, 1.7, 20001)
    integral = np.trapezoid(lin.beta(grid), grid)
    F, Q = disc(1.7, 0.4)
    np.testing.assert_allclose(F, np.exp(-0.5 * integral), rtol=1e-9)
    np.testing.assert_allclose(Q, 1 - np.exp(-integral), rtol=1e-9)


def test_ou_and_linear_sde_paths_bitwise_equal():
    """tests/test_sdes.py:149-159: make_ou_sde and make_linear_sde give identical paths under one key."""
    a, b = -0.5, 1.
    _, _, sim_ou = osdes.make_ou_sde(a, b)
    _, _, sim_lin = osdes.make_linear_sde(osdes.StationaryConstLinearSDE(a, b))
    key = jr.PRNGKey(666)
    x0 = jr.normal(jr.PRNGKey(1), (3,))
    ts = np.linspace(0., 1., 11)
    np.testing.assert_allclose(sim_ou(key, x0, ts), sim_lin(key, x0, ts), rtol=0, atol=1e-6)


def test_gaussian_sb_marginals():
    """tests/test_sdes.py:163-216: closed-form SB marginals hit both end points."""
    m0, c0 = np.array([1., -1.]), np.array([[1., 0.3], [0.3, 0.5]])
    m1, c1 = np.array([-0.5, 2.]), np.array([[0.7, -0.2], [-0.2, 1.2]])
    mm, mc, drift = osdes.make_gaussian_bw_sb(m0, c0, m1, c1, sig=1.)
    np.testing.assert_allclose(mm(0.), m0); np.testing.assert_allclose(mm(1.), m1)
    np.testing.assert_allclose(mc(0.), c0, atol=1e-12); np.testing.assert_allclose(mc(1.), c1, atol=1e-8)


def _gibbs_test_model(dtype=np.float32):
    m0 = np.array([-1., 1.]); cov0 = np.array([[2., 0.4], [0.4, 0.5]])
    T, K = 1., 100
    ts = np.linspace(0, T, K + 1)
    sde = osdes.StationaryConstLinearSDE(a=-0.5, b=1.)
    return JointGaussianDiffusionModel(sde, m0, cov0, 1, ts, T, dtype=dtype), sde, m0, cov0


def test_affine_coefficients_equal_cholesky_closures():
    mod, sde, _, _ = _gibbs_test_model(np.float64)
    M, m, g = mod.affine_coefficients()
    uv = np.random.default_rng(0).normal(size=(7, 2))
    for k in (0, 37, 99):
        np.testing.assert_allclose(mod.reverse_drift(uv, mod.ts[k]), uv @ M[k].T + m[k], rtol=1e-9, atol=1e-10)


def test_gibbs_kernel_targets_posterior():
    """tests/test_gibbs.py:16-123 (K=100, N=10) with 700 sweeps instead of 10 000: mean within 0.25 (MC error
    of ~700 correlated draws of a variance-1.68 target), variance within 25 %."""
    mod, sde, m0, cov0 = _gibbs_test_model(np.float32)
    y0 = np.array([0.], np.float32)
    true_mean = m0[0] + cov0[0, 1] / cov0[1, 1] * (y0[0] - m0[1])
    true_var = cov0[0, 0] - cov0[0, 1] ** 2 / cov0[1, 1]
    key = jr.PRNGKey(666)
    x0 = np.zeros(1, np.float32); us_star = np.zeros((101, 1), np.float32); bs = np.zeros(101, np.int32)
    xs = []
    for i in range(700):
        key, sub = jr.split(key)
        x0, us_star, bs, _ = ogibbs.gibbs_kernel(sub, x0, y0, us_star, bs, mod.ts, mod.fwd_sampler, sde, mod.unpack, 10,
                                                 mod.transition_sampler, mod.transition_logpdf, mod.likelihood_logpdf)
        xs.append(float(x0[0]))
    xs = np.array(xs[10:])
    assert abs(xs.mean() - true_mean) < 0.25
    assert abs(xs.var() / true_var - 1) < 0.25


def test_pmcmc_filter_log_likelihood_against_kalman():
    """tests/test_filters.py:14-87 spirit: the particle-filter evidence of the Gaussian model is finite and the
    filter's terminal particles match the exact conditional p(u_K | v_{0:K}) of the linear-Gaussian reverse chain
    in mean (N=2000 particles)."""
    mod, sde, m0, cov0 = _gibbs_test_model(np.float64)
    y0 = np.array([0.3])
    ys = mod.fwd_ys_sampler(jr.PRNGKey(1), y0)
    vs = ys[::-1]
    N = 2000
    u0s = mod.ref_sampler(jr.PRNGKey(2), vs[0], N)
    uT, log_ell = osmc.pmcmc_filter_step(jr.PRNGKey(3), vs, u0s, mod.ts, mod.transition_sampler, mod.likelihood_logpdf,
                                         orx.stratified, N)
    assert np.isfinite(log_ell)
    post_mean = m0[0] + cov0[0, 1] / cov0[1, 1] * (y0[0] - m0[1])
    post_var = cov0[0, 0] - cov0[0, 1] ** 2 / cov0[1, 1]
    # Euler--Maruyama reverse chain (dt = 0.01) + a random y-path: a loose 4-sigma check of the particle cloud
    assert abs(uT.mean() - post_mean) < 4 * np.sqrt(post_var)


def test_backward_scanning_pass_traces_ancestors():
    rng = np.random.default_rng(0)
    K, N = 6, 5
    As = rng.integers(0, N, (K, N)).astype(np.int32)
    xss = rng.normal(size=(K + 1, N, 2))
    lw = np.log(np.full(N, 1 / N))
    xs, bs = ocsmc.backward_scanning_pass(jr.PRNGKey(0), As, xss, lw)
    for t in range(K, 0, -1):
        assert bs[t - 1] == As[t - 1][bs[t]]                                  # csmc.py:262
    np.testing.assert_array_equal(xs, xss[np.arange(K + 1), bs])
