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
