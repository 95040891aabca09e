"""Model closures of the reference's experiment drivers, batched over particles.  TEST INFRASTRUCTURE.

Restates the closures of ``/root/reference/experiments/toy/gp_gibbs.py:32-150`` (identical in
``experiments/toy/gp_pmcmc.py:28-151`` and, on a 2-D Gaussian, ``tests/test_gibbs.py:23-96``)
and of ``experiments/sb/gibbs.py:28-164``.

The reference evaluates the joint reverse drift through a Cholesky factorisation of the
marginal covariance at every call (``gp_gibbs.py:78-81``) and wraps per-particle functions in
``jax.vmap``.  Here the same arithmetic is written batched: one factorisation per call, the
solve applied to all particles.  ``dtype=float32`` mirrors the experiments; ``float64`` gives
the exact-arithmetic answer the fp32 kernel is compared with at tight tolerance.
"""
import math
import numpy as np
import scipy.linalg as sla
from . import jax_random as jr
from .sdes import (make_linear_sde, StationaryConstLinearSDE, StationaryLinLinearSDE,
                   make_gaussian_bw_sb, euler_maruyama)


def _normal(key, shape, dtype):
    return jr.normal(key, shape) if dtype == np.float32 else jr.normal64(key, shape)


def norm_logpdf(x, loc, scale):
    """``jax.scipy.stats.norm.logpdf``: ``-(log(2 pi scale^2) + (x-loc)^2/scale^2) / 2``."""
    dtype = np.result_type(x, loc)
    scale = np.asarray(scale, dtype=dtype)
    s2 = scale * scale
    log_norm = np.log(dtype.type(2 * np.pi) * s2)
    quad = (x - loc) ** 2 / s2
    return ((log_norm + quad) / dtype.type(-2)).astype(dtype)


class JointGaussianDiffusionModel:
    """Reverse-diffusion closures for a jointly Gaussian (X, Y) under a scalar linear SDE.

    ``joint_mean (D,)``, ``joint_cov (D, D)`` with ``D = du + dv``; ``X`` is the first ``du``
    coordinates (``unpack``, gp_gibbs.py:89-90).
    """

    def __init__(self, sde, joint_mean, joint_cov, du, ts, T, dt=None, dtype=np.float32):
        self.sde, self.du, self.T, self.dtype = sde, du, T, dtype
        self.ts = np.asarray(ts, dtype=dtype)
        self.dt = (T / (len(ts) - 1)) if dt is None else dt           # python float, gp_gibbs.py:63
        self.joint_mean = np.asarray(joint_mean, dtype=dtype)
        self.joint_cov = np.asarray(joint_cov, dtype=dtype)
        self.D = self.joint_mean.shape[0]
        self.dv = self.D - du
        self.discretise, _, self.simulate_cond_forward = make_linear_sde(sde, dtype)

    # -- gp_gibbs.py:73-81
    def forward_m_cov(self, t):
        F_, Q_ = self.discretise(t, self.ts[0])
        return F_ * self.joint_mean, (F_ ** 2 * self.joint_cov + Q_ * np.eye(self.D, dtype=self.dtype)).astype(self.dtype)

    def score(self, z, t):
        mt, covt = self.forward_m_cov(t)
        chol = sla.cho_factor(covt)
        return (-sla.cho_solve(chol, (z - mt).T).T).astype(self.dtype)

    def unpack(self, xy):
        return xy[..., :self.du], xy[..., self.du:]

    # -- gp_gibbs.py:94-109
    def reverse_drift(self, uv, t):
        dtype = self.dtype
        tt = dtype(self.T) - dtype(t)
        g = dtype(self.sde.dispersion(tt))
        return (-self.sde.drift(uv, tt) + g ** 2 * self.score(uv, tt)).astype(dtype)

    def reverse_dispersion(self, t):
        return self.dtype(self.sde.dispersion(self.dtype(self.T) - self.dtype(t)))

    def _uv(self, us, v):
        return np.concatenate([us, np.broadcast_to(v, (us.shape[0], self.dv))], axis=1).astype(self.dtype)

    def transition_mean(self, us_prev, v_prev, t_prev):
        """Mean of the Euler--Maruyama transition (the deterministic part of gp_gibbs.py:120-122)."""
        drift_u = self.reverse_drift(self._uv(us_prev, v_prev), t_prev)[:, :self.du]
        return (us_prev + drift_u * self.dtype(self.dt)).astype(self.dtype)

    def transition_sd(self, t_prev):
        return self.dtype(math.sqrt(self.dt)) * self.reverse_dispersion(t_prev)

    # -- gp_gibbs.py:120-135
    def transition_sampler(self, us_prev, v_prev, t_prev, key_):
        dtype = self.dtype
        drift_u = self.reverse_drift(self._uv(us_prev, v_prev), t_prev)[:, :self.du]
        noise = _normal(key_, us_prev.shape, dtype)
        return (us_prev + drift_u * dtype(self.dt)
                + dtype(math.sqrt(self.dt)) * self.reverse_dispersion(t_prev) * noise).astype(dtype)

    def transition_logpdf(self, u, us_prev, v_prev, t_prev):
        dtype = self.dtype
        drift_u = self.reverse_drift(self._uv(us_prev, v_prev), t_prev)[:, :self.du]
        return np.sum(norm_logpdf(u, us_prev + drift_u * dtype(self.dt),
                                  dtype(math.sqrt(self.dt)) * self.reverse_dispersion(t_prev)), axis=-1).astype(dtype)

    def likelihood_logpdf(self, v, us_prev, v_prev, t_prev):
        dtype = self.dtype
        drift_v = self.reverse_drift(self._uv(us_prev, v_prev), t_prev)[:, self.du:]
        cond_m = v_prev + drift_v * dtype(self.dt)
        return np.sum(norm_logpdf(v, cond_m, dtype(math.sqrt(self.dt)) * self.reverse_dispersion(t_prev)),
                      axis=-1).astype(dtype)

    # -- gp_gibbs.py:84-86,138-149
    def ref_sampler(self, key_, yT, nsamples_):
        dtype, d = self.dtype, self.du
        m_ref, cov_ref = self.forward_m_cov(self.dtype(self.T))
        chol_ref = sla.cho_factor(cov_ref[d:, d:])
        m_ = m_ref[:d] + cov_ref[:d, d:] @ sla.cho_solve(chol_ref, yT - m_ref[d:])
        cov_ = cov_ref[:d, :d] - cov_ref[:d, d:] @ sla.cho_solve(chol_ref, cov_ref[d:, :d])
        return (m_ + _normal(key_, (nsamples_, d), dtype) @ np.linalg.cholesky(cov_)).astype(dtype)

    def fwd_sampler(self, key_, x0_, y0_):
        return self.simulate_cond_forward(key_, np.concatenate([x0_, y0_]).astype(self.dtype), self.ts)

    def fwd_ys_sampler(self, key_, y0_):
        return self.simulate_cond_forward(key_, np.asarray(y0_, dtype=self.dtype), self.ts)

    # -- exact affine form of the same drift, float64: drift(uv, t_k) = M_k uv + m_k
    def affine_coefficients(self):
        """(M (K, D, D), m (K, D), g (K,)) in float64 for the K step times ``ts[:-1]``."""
        K = len(self.ts) - 1
        jm, jc = self.joint_mean.astype(np.float64), self.joint_cov.astype(np.float64)
        disc64, _, _ = make_linear_sde(self.sde, np.float64)
        M = np.zeros((K, self.D, self.D)); m = np.zeros((K, self.D)); g = np.zeros((K,))
        eye = np.eye(self.D)
        for k in range(K):
            tt = float(self.T) - float(self.ts[k])
            F_, Q_ = disc64(tt, float(self.ts[0]))
            cov = F_ ** 2 * jc + Q_ * eye
            prec = np.linalg.inv(cov)
            gk = float(self.sde.dispersion(tt))
            a_lin = float(self.sde.drift(1.0, tt))           # drift is a(t) x
            M[k] = -a_lin * eye - gk ** 2 * prec
            m[k] = gk ** 2 * prec @ (F_ * jm)
            g[k] = gk
        return M, m, g


def gp_regression_setup(d, obs_var=1., ell=1., sigma=1., dtype=np.float64):
    """gp_gibbs.py:32-58: exponential-kernel GP prior on linspace(0, 5, d) + iid noise."""
    zs = np.linspace(0., 5., d)
    cov_mat = sigma ** 2 * np.exp(-np.abs(zs[None, :] - zs[:, None]) / ell)
    joint_mean = np.zeros((2 * d,))
    joint_cov = np.block([[cov_mat, cov_mat], [cov_mat, cov_mat + obs_var * np.eye(d)]])
    return cov_mat.astype(dtype), joint_mean.astype(dtype), joint_cov.astype(dtype)


def gp_draw_y0(key, d, cov_mat, obs_var=1.):
    """gp_gibbs.py:44-47 (float32 stream)."""
    key, subkey = jr.split(key)
    fs = np.linalg.cholesky(cov_mat.astype(np.float32)) @ jr.normal(subkey, (d,))
    key, subkey = jr.split(key)
    y0 = fs + np.float32(math.sqrt(obs_var)) * jr.normal(subkey, (d,))
    return key, y0.astype(np.float32)


def gp_posterior(cov_mat, y0, obs_var=1.):
    """gp_gibbs.py:50-53, float64."""
    cov_mat = cov_mat.astype(np.float64)
    d = cov_mat.shape[0]
    chol = sla.cho_factor(cov_mat + obs_var * np.eye(d))
    mean = cov_mat @ sla.cho_solve(chol, y0.astype(np.float64))
    cov = cov_mat - cov_mat @ sla.cho_solve(chol, cov_mat)
    return mean, cov


class GaussianSBModel:
    """Closures of experiments/sb/gibbs.py:62-150 (Gaussian Schroedinger bridge, sigma=1)."""

    def __init__(self, joint_mean, joint_cov, ref_m, ref_cov, du, ts, T=1., dtype=np.float32, em_nsteps=10):
        self.dtype, self.du, self.T = dtype, du, T
        self.ts = np.asarray(ts, dtype=dtype)
        self.dt = T / (len(ts) - 1)
        self.D = joint_mean.shape[0]
        self.dv = self.D - du
        self.ref_m, self.ref_cov = np.asarray(ref_m, dtype), np.asarray(ref_cov, dtype)
        self.em_nsteps = em_nsteps
        self.marginal_mean, self.marginal_cov, self.drift = make_gaussian_bw_sb(
            np.asarray(joint_mean, dtype), np.asarray(joint_cov, dtype), self.ref_m, self.ref_cov, sig=1.)

    def dispersion(self, _):
        return self.dtype(1.)

    def score(self, z, t):
        mt, covt = self.marginal_mean(t), self.marginal_cov(t)
        chol = sla.cho_factor(covt)
        return (-sla.cho_solve(chol, (z - mt).T).T).astype(self.dtype)

    def unpack(self, xy):
        return xy[..., :self.du], xy[..., self.du:]

    def reverse_drift(self, uv, t):
        tt = self.dtype(self.T) - self.dtype(t)
        return (-self.drift(uv, tt) + self.score(uv, tt)).astype(self.dtype)     # sb/gibbs.py:93-94 (dispersion 1)

    def _uv(self, us, v):
        return np.concatenate([us, np.broadcast_to(v, (us.shape[0], self.dv))], axis=1).astype(self.dtype)

    def transition_sampler(self, us_prev, v_prev, t_prev, key_):
        dtype = self.dtype
        drift_u = self.reverse_drift(self._uv(us_prev, v_prev), t_prev)[:, :self.du]
        return (us_prev + drift_u * dtype(self.dt)
                + dtype(math.sqrt(self.dt)) * _normal(key_, us_prev.shape, dtype)).astype(dtype)

    def transition_logpdf(self, u, us_prev, v_prev, t_prev):
        dtype = self.dtype
        drift_u = self.reverse_drift(self._uv(us_prev, v_prev), t_prev)[:, :self.du]
        return np.sum(norm_logpdf(u, us_prev + drift_u * dtype(self.dt), dtype(math.sqrt(self.dt))), axis=-1).astype(dtype)

    def likelihood_logpdf(self, v, us_prev, v_prev, t_prev):
        dtype = self.dtype
        drift_v = self.reverse_drift(self._uv(us_prev, v_prev), t_prev)[:, self.du:]
        return np.sum(norm_logpdf(v, v_prev + drift_v * dtype(self.dt), dtype(math.sqrt(self.dt))), axis=-1).astype(dtype)

    def fwd_sampler(self, key_, x0_, y0_):
        xy0 = np.concatenate([x0_, y0_]).astype(self.dtype)
        return euler_maruyama(key_, xy0, self.ts, lambda x, t: self.drift(x[None], t)[0], self.dispersion,
                              integration_nsteps=self.em_nsteps, return_path=True, dtype=self.dtype)


class TwistedGaussianModel:
    """Closures of experiments/toy/gp_twisted.py:66-129: the reverse diffusion of the X-MARGINAL of the GP regression model,
    the Gaussian twisting function p~(y | u, t) = N(y; u + reverse_drift(u, t) dt, obs_var) (:114-116) and the proposal that
    adds ``g^2 grad_u log p~`` to the drift (:88-90, :122-129).  ``jax.grad`` of the twisting function is written out
    analytically: with reverse_drift(u, t) = M_t u + m_t the denoising estimate is affine, ``(I + dt M_t) u + dt m_t``, so the
    gradient is ``(I + dt M_t)^T (y - estimate) / obs_var`` -- exactly what autodiff returns.
    """

    def __init__(self, sde, mean_x, cov_x, obs_var, ts, T, dtype=np.float64):
        self.sde, self.T, self.dtype, self.obs_var = sde, T, dtype, float(obs_var)
        self.ts = np.asarray(ts, dtype=dtype)
        self.K = self.ts.shape[0] - 1
        self.dt = T / self.K                                           # gp_twisted.py:57 (python float)
        self.mean_x, self.cov_x = np.asarray(mean_x, dtype), np.asarray(cov_x, dtype)
        self.d = self.mean_x.shape[0]
        self.discretise, _, _ = make_linear_sde(sde, dtype)

    def forward_m_cov(self, t):                                        # :66-68
        F_, Q_ = self.discretise(t, self.ts[0])
        return F_ * self.mean_x, (F_ ** 2 * self.cov_x + Q_ * np.eye(self.d, dtype=self.dtype)).astype(self.dtype)

    def affine(self, t):
        """(M_t, m_t, g_t) with reverse_drift(u, t) = M_t u + m_t (:71-85), float64."""
        tt = float(self.T) - float(t)
        mt, covt = self.forward_m_cov(self.dtype(tt))
        prec = np.linalg.inv(covt.astype(np.float64))
        g = float(self.sde.dispersion(tt))
        a_lin = float(self.sde.drift(1.0, tt))
        return -a_lin * np.eye(self.d) - g * g * prec, g * g * prec @ mt.astype(np.float64), g

    def reverse_drift(self, u, t):                                     # :84-85
        M, m, _ = self.affine(t)
        return (u @ M.T + m).astype(self.dtype)

    def init_sampler(self, key_, n):                                   # :108-111
        m_ref, cov_ref = self.forward_m_cov(self.dtype(self.T))
        # (the float32 random stream of the experiment -- jax_enable_x64 is off, gp_twisted.py:21 -- whatever `dtype` the
        #  arithmetic of this restatement runs in)
        return (m_ref + _normal(key_, (n, self.d), np.float32).astype(self.dtype) @ np.linalg.cholesky(cov_ref).T).astype(self.dtype)

    def transition_logpdf(self, u, u_prev, t_prev):                    # :99-105
        _, _, g = self.affine(t_prev)
        return norm_logpdf(u, u_prev + self.reverse_drift(u_prev, t_prev) * self.dtype(self.dt),
                           self.dtype(math.sqrt(self.dt) * g)).sum(-1)

    def twisting_logpdf(self, y, u, t):                                # :114-116
        est = u + self.reverse_drift(u, t) * self.dtype(self.dt)
        return norm_logpdf(np.asarray(y, self.dtype), est, self.dtype(math.sqrt(self.obs_var))).sum(-1)

    def reverse_cond_drift(self, u, t, y):                             # :88-90
        M, m, g = self.affine(t)
        G = np.eye(self.d) + self.dt * M
        est = u + (u @ M.T + m) * self.dt
        grad = ((np.asarray(y, np.float64) - est) / self.obs_var) @ G
        return (u @ M.T + m + g * g * grad).astype(self.dtype)

    def twisting_prop_mean(self, us, t, y):
        return us + self.reverse_cond_drift(us, t, y) * self.dtype(self.dt)

    def twisting_prop_sampler(self, key_, us, t, y):                   # :122-124
        _, _, g = self.affine(t)
        return (self.twisting_prop_mean(us, t, y)
                + self.dtype(math.sqrt(self.dt) * g) * _normal(key_, us.shape, np.float32).astype(self.dtype)).astype(self.dtype)

    def twisting_prop_logpdf(self, u, u_prev, t, y):                   # :127-129
        _, _, g = self.affine(t)
        return norm_logpdf(u, self.twisting_prop_mean(u_prev, t, y), self.dtype(math.sqrt(self.dt) * g)).sum(-1)
