"""Forward-noising kernels vs the oracle (fbs/sdes/linear.py:190-225, fbs/sdes/simulators.py:53-106) and the
closed forms the reference asserts in tests/test_sdes.py."""
import numpy as np
import pytest
from oracle import jax_random as jr
from oracle import sdes as osdes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind', ['const', 'lin'])
@pytest.mark.parametrize('K,D', [(1, 1), (7, 3), (100, 20), (200, 200)])
def test_simulate_cond_forward_matches_oracle(kind, K, D):
    from fbs_b200 import sdes
    T = 1.
    ts = np.linspace(0., T, K + 1)
    if kind == 'lin':
        psde, osde = sdes.StationaryLinLinearSDE(0.02, 4., 0., T), osdes.StationaryLinLinearSDE(0.02, 4., 0., T)
    else:
        psde, osde = sdes.StationaryConstLinearSDE(-0.5, 1.), osdes.StationaryConstLinearSDE(-0.5, 1.)
    _, _, sim = sdes.make_linear_sde(psde)
    _, _, osim = osdes.make_linear_sde(osde, np.float32)
    B = 5
    keys = jr.split(jr.PRNGKey(K * 31 + D), B)
    x0 = jr.normal(jr.PRNGKey(3), (B, D))
    got = sim(keys, x0, ts)
    assert got.shape == (B, K + 1, D) and got.dtype == np.float32
    for b in range(B):
        want = osim(keys[b], x0[b], ts)
        np.testing.assert_allclose(got[b], want, rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(got[:, 0], x0)
    # single (unbatched) call, as the reference uses it
    np.testing.assert_array_equal(sim(keys[0], x0[0], ts), got[0])


def test_discretisation_closed_forms():
    """tests/test_sdes.py:18-34,60-90: exact OU / linear-beta discretisations."""
    from fbs_b200 import sdes
    a, b = -0.5, 1.
    disc, _, _ = sdes.make_linear_sde(sdes.StationaryConstLinearSDE(a, b))
    for (t, s) in [(0.3, 0.1), (1., 0.), (2.5, 2.4)]:
        F, Q = disc(t, s)
        np.testing.assert_allclose(F, np.exp(a * (t - s)), rtol=1e-6)
        np.testing.assert_allclose(Q, b ** 2 / (2 * a) * (np.exp(2 * a * (t - s)) - 1), rtol=1e-6)
    lin = sdes.StationaryLinLinearSDE(0.02, 5., 0., 2.)
    disc, _, _ = sdes.make_linear_sde(lin)
    grid = np.linspace(0.4, 1.7, 20001)
    integral = np.trapezoid(lin.beta(grid), grid)
    F, Q = disc(1.7, 0.4)
    np.testing.assert_allclose(F, np.exp(-0.5 * integral), rtol=1e-6)
    np.testing.assert_allclose(Q, 1 - np.exp(-integral), rtol=1e-6)


def test_forward_marginal_moments():
    """Terminal law of the forward sampler: N(F x0, Q) (tests/test_sdes.py:93-116 spirit)."""
    from fbs_b200 import sdes, random as fr
    sde = sdes.StationaryConstLinearSDE(-0.5, 1.)
    disc, _, sim = sdes.make_linear_sde(sde)
    ts = np.linspace(0., 1., 51)
    B = 20000
    keys = fr.split(fr.PRNGKey(8), B)
    x0 = np.tile(np.array([[1.5, -0.7]], np.float32), (B, 1))
    xT = sim(keys, x0, ts)[:, -1]
    F, Q = disc(1., 0.)
    np.testing.assert_allclose(xT.mean(0), F * x0[0], atol=3e-2)
    np.testing.assert_allclose(xT.var(0), [Q, Q], rtol=5e-2)


@pytest.mark.parametrize('m', [1, 10])
def test_euler_maruyama_affine_matches_oracle(m):
    from fbs_b200 import sdes
    rng = np.random.default_rng(0)
    D, K = 6, 12
    ts = np.linspace(0., 1., K + 1)
    A0 = rng.normal(size=(D, D)) * 0.4
    a0 = rng.normal(size=(D,))

    def affine(t):
        return A0 * (1. + t), a0 * np.cos(t)

    drift = sdes.AffineDrift(affine)
    dispersion = lambda t: 0.7 + 0.1 * t
    B = 4
    keys = jr.split(jr.PRNGKey(19), B)
    x0 = jr.normal(jr.PRNGKey(20), (B, D))
    got = sdes.euler_maruyama(keys, x0, ts, drift, dispersion, integration_nsteps=m, return_path=True)
    for b in range(B):
        want = osdes.euler_maruyama(keys[b], x0[b], ts,
                                    lambda x, t: (x @ (A0 * (1. + t)).T + a0 * np.cos(t)).astype(np.float32),
                                    lambda t: np.float32(0.7 + 0.1 * t), integration_nsteps=m, return_path=True)
        np.testing.assert_allclose(got[b], want, rtol=2e-4, atol=2e-5)
    term = sdes.euler_maruyama(keys, x0, ts, drift, dispersion, integration_nsteps=m, return_path=False)
    np.testing.assert_array_equal(term, got[:, -1])
    with pytest.raises(TypeError):
        sdes.euler_maruyama(keys, x0, ts, lambda x, t: x, dispersion)


@pytest.mark.parametrize('D,m', [(20, 10), (8, 10), (6, 1), (5, 3), (32, 2), (2, 1), (12, 3), (14, 2)])
def test_euler_maruyama_thread_per_chain_kernel(D, m, monkeypatch):
    """>= 1024 chains of a small system run one thread per chain, or one chain over 2 / 4 lanes when D allows (registers +
    shared-memory noise stash): same paths as the CTA-per-chain kernel on every chain, and the oracle's on a few."""
    from fbs_b200 import sdes
    rng = np.random.default_rng(D * 10 + m)
    K = 7
    ts = np.linspace(0., 1., K + 1)
    A0 = rng.normal(size=(D, D)) * (0.8 / np.sqrt(D))
    a0 = rng.normal(size=(D,))
    drift = sdes.AffineDrift(lambda t: (A0 * (1. + t), a0 * np.cos(t)))
    dispersion = lambda t: 0.7 + 0.1 * t
    B = 1100
    keys = jr.split(jr.PRNGKey(D + m), B)
    x0 = rng.normal(size=(B, D)).astype(np.float32)
    got = sdes.euler_maruyama(keys, x0, ts, drift, dispersion, integration_nsteps=m, return_path=True)   # lane-split if D allows
    monkeypatch.setenv('FBS_EM_IMPL', 'tpc')
    got_tpc = sdes.euler_maruyama(keys, x0, ts, drift, dispersion, integration_nsteps=m, return_path=True)
    monkeypatch.setenv('FBS_EM_IMPL', 'cta')
    ref = sdes.euler_maruyama(keys, x0, ts, drift, dispersion, integration_nsteps=m, return_path=True)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got_tpc, ref, rtol=1e-5, atol=1e-6)
    for b in (0, 517, B - 1):
        want = osdes.euler_maruyama(keys[b], x0[b], ts,
                                    lambda x, t: (x @ (A0 * (1. + t)).T + a0 * np.cos(t)).astype(np.float32),
                                    lambda t: np.float32(0.7 + 0.1 * t), integration_nsteps=m, return_path=True)
        np.testing.assert_allclose(got[b], want, rtol=2e-4, atol=2e-5)


def test_gaussian_sb_marginals():
    """tests/test_sdes.py:163-216 spirit: the closed-form SB drift transports N(m0, c0) to N(m1, c1)."""
    from fbs_b200 import sdes, random as fr
    m0, c0 = np.array([1., -1.]), np.array([[1., 0.3], [0.3, 0.5]])
    m1, c1 = np.array([-0.5, 2.]), np.array([[0.7, -0.2], [-0.2, 1.2]])
    mm, mc, drift = sdes.make_gaussian_bw_sb(m0, c0, m1, c1, sig=1.)
    np.testing.assert_allclose(mm(0.), m0); np.testing.assert_allclose(mm(1.), m1)
    np.testing.assert_allclose(mc(0.), c0, atol=1e-12); np.testing.assert_allclose(mc(1.), c1, atol=1e-10)
    B = 20000
    keys = fr.split(fr.PRNGKey(4), B)
    x0 = (m0 + fr.normal(fr.PRNGKey(5), (B, 2)) @ np.linalg.cholesky(c0).T).astype(np.float32)
    ts = np.linspace(0., 1., 101)
    xT = sdes.euler_maruyama(keys, x0, ts, drift, lambda t: 1., integration_nsteps=2, return_path=False)
    np.testing.assert_allclose(xT.mean(0), m1, atol=5e-2)
    np.testing.assert_allclose(np.cov(xT.T), c1, atol=8e-2)


def test_ou_and_linear_sde_agree_bitwise_on_the_kernels():
    """tests/test_sdes.py:149-159 on the CUDA path: ``make_ou_sde(a, b)`` and ``make_linear_sde(StationaryConstLinearSDE(a, b))``
    give EQUAL discretisations, conditional scores and -- under one key -- bitwise equal forward paths (``assert_array_equal``
    upstream, :159), batched and unbatched."""
    from fbs_b200 import sdes
    a, b = -0.5, 1.
    ou_disc, ou_score, ou_sim = sdes.make_ou_sde(a, b)
    lin_disc, lin_score, lin_sim = sdes.make_linear_sde(sdes.StationaryConstLinearSDE(a=a, b=b))
    F_ou, Q_ou = ou_disc(1.)
    F_l, Q_l = lin_disc(1., 0.)
    np.testing.assert_equal(F_ou, F_l)                                           # :149-150
    np.testing.assert_equal(Q_ou, Q_l)
    np.testing.assert_equal(ou_score(2.2, 1.1, 1.5), lin_score(2.2, 1.1, 1.5, 0.))   # :152-154
    key = jr.PRNGKey(666)
    ts = np.linspace(0., 2., 20)
    path_ou = ou_sim(key, np.array([2.5], np.float32), ts)                       # :156-159
    path_lin = lin_sim(key, np.array([2.5], np.float32), ts)
    assert path_ou.shape == (20, 1)
    np.testing.assert_array_equal(path_ou, path_lin)
    keys = jr.split(key, 7)
    x0 = jr.normal(jr.PRNGKey(1), (7, 5))
    np.testing.assert_array_equal(ou_sim(keys, x0, ts), lin_sim(keys, x0, ts))
    # and the path is the oracle's (float32 restatement of linear.py:190-225)
    _, _, osim = osdes.make_ou_sde(a, b)
    np.testing.assert_allclose(path_ou, osim(key, np.array([2.5], np.float32), ts), rtol=2e-5, atol=2e-6)
