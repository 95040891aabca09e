"""Linear SDEs, exact discretisation, forward-noising and Euler--Maruyama.  TEST INFRASTRUCTURE.

Restates ``/root/reference/fbs/sdes/linear.py`` (classes :13-112, ``make_ou_sde`` :115-162,
``make_linear_sde`` :165-227, ``make_gaussian_bw_sb`` :397-457), ``fbs/sdes/simulators.py``
(``euler_maruyama`` :53-106) and ``fbs/utils.py:21-28`` (``sqrtm``).

``dtype`` selects float32 (the experiments, ``jax_enable_x64`` off) or float64 (the
reference's x64 tests).  Random streams: float32 uses the pinned ``jax_random.normal``;
float64 uses ``normal64`` (stream not pinned, statistical tests only).
"""
import math
import numpy as np
from . import jax_random as jr


def _normal(key, shape, dtype):
    return jr.normal(key, shape) if dtype == np.float32 else jr.normal64(key, shape)


class LinearSDE:
    pass


class StationaryConstLinearSDE(LinearSDE):
    """dX = a X dt + b dW.  linear.py:13-45."""

    def __init__(self, a, b):
        self.a, self.b = a, b

    def drift(self, x, t):
        return self.a * x

    def dispersion(self, t):
        return self.b

    def mean(self, t, s, m0):
        return m0 * np.exp(self.a * (t - s))

    def variance(self, t, s):
        return self.b ** 2 / (2 * self.a) * (np.exp(2 * self.a * (t - s)) - 1)


class StationaryLinLinearSDE(LinearSDE):
    """dX = -0.5 beta(t) X dt + sqrt(beta(t)) dW, beta linear in t.  linear.py:48-92."""

    def __init__(self, beta_min, beta_max, t0, T):
        self.beta_min, self.beta_max, self.t0, self.T = beta_min, beta_max, t0, T

    def beta(self, t):
        bmin, bmax, t0, T = self.beta_min, self.beta_max, self.t0, self.T
        return (bmax - bmin) / (T - t0) * t + (bmin * T - bmax * t0) / (T - t0)

    def beta_integral(self, t, s):
        bmin, bmax, t0, T = self.beta_min, self.beta_max, self.t0, self.T
        return 0.5 * (t - s) * ((bmax - bmin) / (T - t0) * (t + s) + 2 * (bmin * T - bmax * t0) / (T - t0))

    def drift(self, x, t):
        return -0.5 * self.beta(t) * x

    def dispersion(self, t):
        return np.sqrt(self.beta(t))

    def mean(self, t, s, m0):
        return m0 * np.exp(-0.5 * self.beta_integral(t, s))

    def variance(self, t, s):
        return 1 - np.exp(-self.beta_integral(t, s))


class StationaryExpLinearSDE(LinearSDE):
    """linear.py:95-112."""

    def __init__(self, a, b, c, z):
        self.a, self.b, self.c, self.z = a, b, c, z

    def drift(self, x, t):
        return self.a * np.exp(self.c * (t - self.z)) * x

    def dispersion(self, t):
        return self.b * np.exp(self.c * (t - self.z) / 2)


def make_linear_sde(sde, dtype=np.float32):
    """linear.py:165-227.  Returns (discretise_linear_sde, cond_score_t_0, simulate_cond_forward)."""

    def discretise_linear_sde(t, s):
        t = np.asarray(t, dtype=dtype)
        s = np.asarray(s, dtype=dtype)
        if isinstance(sde, StationaryLinLinearSDE):                    # :172-174
            r = sde.beta_integral(t, s).astype(dtype)
            return np.exp(dtype(-0.5) * r).astype(dtype), (dtype(1) - np.exp(-r)).astype(dtype)
        elif isinstance(sde, StationaryConstLinearSDE):                # :175-177
            a, b = sde.a, sde.b
            return (np.exp(dtype(a) * (t - s)).astype(dtype),
                    (dtype(b ** 2 / (2 * a)) * (np.exp(dtype(2 * a) * (t - s)) - dtype(1))).astype(dtype))
        elif isinstance(sde, StationaryExpLinearSDE):                  # :178-182
            a, b, c, z = sde.a, sde.b, sde.c, sde.z
            stationary_variance = -b ** 2 / (2 * a)
            r = (dtype(a) * (np.exp(dtype(c) * (t - dtype(z))) - np.exp(dtype(c) * (s - dtype(z)))) / dtype(c)).astype(dtype)
            return np.exp(r).astype(dtype), (dtype(stationary_variance) * (dtype(1) - np.exp(dtype(2) * r))).astype(dtype)
        raise NotImplementedError('...')

    def cond_score_t_0(x, t, x0, s):
        F, Q = discretise_linear_sde(t, s)
        return -(x - F * x0) / Q

    def simulate_cond_forward(key, x0, ts, t0=None, keep_path=True):
        # :190-225
        x0 = np.asarray(x0, dtype=dtype)
        ts = np.asarray(ts, dtype=dtype)
        if keep_path:
            rnds = _normal(key, (ts.shape[0] - 1, *x0.shape), dtype)
            Fs, Qs = discretise_linear_sde(ts[1:], ts[:-1])
            sq = np.sqrt(Qs).astype(dtype)
            path = np.empty((ts.shape[0], *x0.shape), dtype=dtype)
            path[0] = x0
            x = x0
            for k in range(ts.shape[0] - 1):
                x = (Fs[k] * x + sq[k] * rnds[k]).astype(dtype)        # :216
                path[k + 1] = x
            return path
        Fs, Qs = discretise_linear_sde(ts, t0)
        rnds = _normal(key, (*ts.shape, *x0.shape), dtype)
        ex = (slice(None),) + (None,) * x0.ndim
        return (Fs[ex] * x0 + np.sqrt(Qs)[ex] * rnds).astype(dtype)

    return discretise_linear_sde, cond_score_t_0, simulate_cond_forward


def make_ou_sde(a, b, dtype=np.float32):
    """linear.py:115-162 (independent OU, time-homogeneous)."""
    sde = StationaryConstLinearSDE(a, b)
    disc, _, sim = make_linear_sde(sde, dtype)

    def discretise_ou_sde(t):
        return disc(t, 0.)

    def cond_score_t_0(x, t, x0):
        F, Q = discretise_ou_sde(t)
        return -(x - F * x0) / Q

    def simulate_cond_forward(key, x0, ts, keep_path=True):
        x0 = np.asarray(x0, dtype=dtype)
        ts = np.asarray(ts, dtype=dtype)
        if keep_path:
            dts = np.diff(ts)
            rnds = _normal(key, (dts.shape[0], x0.shape[0]), dtype)
            path = [x0]
            x = x0
            for k in range(dts.shape[0]):
                F, Q = discretise_ou_sde(dts[k])
                x = (F * x + np.sqrt(Q) * rnds[k]).astype(dtype)
                path.append(x)
            return np.stack(path)
        Fs, Qs = discretise_ou_sde(ts)
        rnds = _normal(key, (ts.shape[0], x0.shape[0]), dtype)
        return (Fs[:, None] * x0[None, :] + np.sqrt(Qs)[:, None] * rnds).astype(dtype)

    return discretise_ou_sde, cond_score_t_0, simulate_cond_forward


def euler_maruyama(key, x0, ts, drift, dispersion, integration_nsteps=1, return_path=False, dtype=np.float32):
    """simulators.py:53-106.  One key per interval; ``(m, *shape)`` normals per interval."""
    x0 = np.asarray(x0, dtype=dtype)
    ts = np.asarray(ts, dtype=dtype)
    nint = ts.shape[0] - 1
    keys = jr.split(key, nint)                                         # :81
    x = x0
    path = [x0]
    m = integration_nsteps
    for k in range(nint):
        t, t_next = ts[k], ts[k + 1]
        ddt = (np.abs(t_next - t) / dtype(m)).astype(dtype)            # :90
        rnds = _normal(keys[k], (m, *x0.shape), dtype)                 # :91
        tgrid = np.linspace(t, t_next - ddt, m, dtype=dtype)           # :92
        sq = np.sqrt(ddt).astype(dtype)
        for q in range(m):
            t_ = tgrid[q]
            x = (x + drift(x, t_) * ddt + dispersion(t_) * sq * rnds[q]).astype(dtype)  # :87
        path.append(x)
    return np.stack(path) if return_path else x


def sqrtm(mat):
    """fbs/utils.py:21-28 (eigh branch)."""
    w, v = np.linalg.eigh(mat)
    return (v @ np.diag(np.sqrt(w)) @ v.T).astype(mat.dtype)


def make_gaussian_bw_sb(mean0, cov0, mean1, cov1, sig=1.):
    """linear.py:397-457.  Gaussian Schroedinger bridge, Brownian reference on [0, 1]."""
    dtype = np.asarray(cov0).dtype.type
    d = mean0.shape[0]
    eye = np.eye(d, dtype=dtype)
    sqrt0 = sqrtm(cov0)
    D_sig = sqrtm(dtype(4) * sqrt0 @ cov1 @ sqrt0 + dtype(sig ** 4) * eye)            # :425
    C_sig = dtype(0.5) * (sqrt0 @ np.linalg.solve(sqrt0.T, D_sig.T).T - dtype(sig ** 2) * eye)  # :426

    def marginal_mean(t):
        return ((1 - t) * mean0 + t * mean1).astype(dtype)

    def marginal_cov(t):
        return ((1 - t) ** 2 * cov0 + t ** 2 * cov1 + t * (1 - t) * (C_sig + C_sig.T)
                + (t * sig ** 2) * (1 - t) * eye).astype(dtype)                        # :444-445

    def s(t):
        pt = t * cov1 + (1 - t) * C_sig
        qt = (1 - t) * cov0 + t * C_sig
        return (pt - qt.T - sig ** 2 * t * eye).astype(dtype)                           # :447-450

    def drift(x, t):
        import scipy.linalg as sla
        mt = marginal_mean(t)
        chol = sla.cho_factor(marginal_cov(t))
        return (s(t).T @ sla.cho_solve(chol, (x - mt).T)).T.astype(dtype) - mean0 + mean1  # :452-455

    drift.s = s
    return marginal_mean, marginal_cov, drift
