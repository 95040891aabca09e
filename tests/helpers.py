"""Shared fixtures: the same toy models built twice -- once as the oracle's closures, once as the
product's AffineGaussianModel -- from the same numbers."""
import numpy as np
from oracle import jax_random as jr
from oracle import models as omodels
from oracle import sdes as osdes


def gp_problem(d, K=20, T=1., sde_kind='const', obs_var=1., seed=7):
    """experiments/toy/gp_gibbs.py setting at dimension d with K steps."""
    cov_mat, jm, jc = omodels.gp_regression_setup(d, obs_var=obs_var)
    _, y0 = omodels.gp_draw_y0(jr.PRNGKey(seed), d, cov_mat, obs_var)
    ts = np.linspace(0., T, K + 1)
    if sde_kind == 'lin':
        osde = osdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=4., t0=0., T=T)
    else:
        osde = osdes.StationaryConstLinearSDE(a=-0.5, b=1.)
    return dict(d=d, K=K, T=T, ts=ts, jm=jm, jc=jc, y0=y0, osde=osde, cov_mat=cov_mat, sde_kind=sde_kind, obs_var=obs_var)


def oracle_model(p, dtype=np.float32):
    return omodels.JointGaussianDiffusionModel(p['osde'], p['jm'], p['jc'], p['d'], p['ts'], p['T'], dtype=dtype)


def product_model(p):
    import fbs_b200
    from fbs_b200 import sdes
    if p['sde_kind'] == 'lin':
        sde = sdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=4., t0=0., T=p['T'])
    else:
        sde = sdes.StationaryConstLinearSDE(a=-0.5, b=1.)
    return fbs_b200.AffineGaussianModel.from_linear_sde(sde, p['jm'], p['jc'], p['d'], p['ts'], T=p['T']), sde


def ulp_diff(a, b):
    """Distance in float32 units-in-the-last-place between two float32 arrays."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2 ** 31) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2 ** 31) - ib, ib)
    return np.abs(ia - ib)


class LinearGaussianSSM:
    """A time-homogeneous linear-Gaussian state-space model in the closure protocol of the samplers, as the reference's
    own sampler tests build inline (tests/test_filters.py:14-143, tests/test_csmc.py:39-132):

        transition   u' = A u + Buv v_prev + sd_u eps
        likelihood   v ~ N(Hm u + Cm v_prev, sd_v)

    The accelerated closures (fbs_b200.AffineGaussianModel) share ONE step standard deviation between the transition and
    the likelihood, so the observations are carried in scaled units v' = c v with c = sd_u / sd_v: the quadratic forms, hence
    the normalised weights, every resampling decision and the particles, are unchanged; only the constant
    dv log(sd_v / sd_u) per step moves between the likelihood normaliser and the data.  ``scale(ys)`` maps observations
    into those units; both the numpy closures below (float64 oracle side) and ``product()`` consume scaled observations.
    """

    def __init__(self, A, Buv, Hm, Cm, sd_u, sd_v, K):
        self.A, self.Buv = np.atleast_2d(np.asarray(A, np.float64)), np.atleast_2d(np.asarray(Buv, np.float64))
        self.Hm, self.Cm = np.atleast_2d(np.asarray(Hm, np.float64)), np.atleast_2d(np.asarray(Cm, np.float64))
        self.sd_u, self.sd_v, self.K = float(sd_u), float(sd_v), int(K)
        self.du, self.dv = self.A.shape[0], self.Hm.shape[0]
        self.c = self.sd_u / self.sd_v
        self.ts = np.linspace(0., float(K), K + 1)            # unit steps: the affine model's dt is 1

    def scale(self, ys):
        return np.asarray(ys, np.float64) * self.c

    # ---- float64 closures in scaled observation units (oracle side)
    def transition_mean(self, us_prev, v_prev, t_prev=None):
        return us_prev @ self.A.T + (np.asarray(v_prev) / self.c) @ self.Buv.T

    def transition_sampler(self, us_prev, v_prev, t_prev, key):
        return self.transition_mean(us_prev, v_prev) + self.sd_u * jr.normal(key, us_prev.shape).astype(np.float64)

    def transition_logpdf(self, u, us_prev, v_prev, t_prev):
        return omodels.norm_logpdf(np.asarray(u, np.float64), self.transition_mean(us_prev, v_prev), self.sd_u).sum(-1)

    def likelihood_logpdf(self, v, us_prev, v_prev, t_prev):
        mean = self.c * (us_prev @ self.Hm.T) + np.asarray(v_prev) @ self.Cm.T
        return omodels.norm_logpdf(np.asarray(v, np.float64), mean, self.sd_u).sum(-1)

    # ---- the same model as structured data for the kernels
    def product(self):
        import fbs_b200
        du, dv, D = self.du, self.dv, self.du + self.dv
        M = np.zeros((D, D))
        M[:du, :du] = self.A - np.eye(du)
        M[:du, du:] = self.Buv / self.c
        M[du:, :du] = self.c * self.Hm
        M[du:, du:] = self.Cm - np.eye(dv)
        return fbs_b200.AffineGaussianModel(np.tile(M, (self.K, 1, 1)), np.zeros((self.K, D)), np.full((self.K,), self.sd_u), 1.0,
                                            du, self.ts)
