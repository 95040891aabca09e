"""Scalar linear SDEs, their exact discretisation and the forward-noising sampler.

API of ``fbs/sdes/linear.py`` (classes :13-112, ``make_linear_sde`` :165-227, ``make_ou_sde`` :115-162).
The SDE objects do host-side scalar maths (coefficients of K steps); the sampling itself runs in
``fbs_ou_forward_path_f32`` (fbs_b200/csrc/sde_kernels.cu).
"""
import numpy as np
import torch
from .. import _native as nat
from .._tensor import dev, empty, ptr, stream, out, is_host


class LinearSDE:
    """dX = a(t) X dt + b(t) dW with scalar a, b."""

    def drift_coef(self, t):
        raise NotImplementedError

    def drift(self, x, t):
        return self.drift_coef(t) * x

    def transition(self, t, s, dtype=np.float64):
        """(F, Q) with X_t | X_s ~ N(F X_s, Q I).  ``dtype=float32`` reproduces the rounding of the
        float32 reference (linear.py:172-182 evaluated with x64 off), cancellation included."""
        raise NotImplementedError

    def mean(self, t, s, m0):
        return m0 * self.transition(t, s)[0]

    def variance(self, t, s):
        return self.transition(t, s)[1]


class StationaryConstLinearSDE(LinearSDE):
    """dX = a X dt + b dW (fbs/sdes/linear.py:13-45)."""

    def __init__(self, a, b):
        self.a, self.b = a, b

    def drift_coef(self, t):
        return self.a

    def dispersion(self, t):
        return self.b

    def transition(self, t, s, dtype=np.float64):
        f = np.dtype(dtype).type
        lag = np.asarray(t, dtype=dtype) - np.asarray(s, dtype=dtype)
        if dtype == np.float64:
            return np.exp(self.a * lag), self.b ** 2 / (2 * self.a) * np.expm1(2 * self.a * lag)
        return np.exp(f(self.a) * lag), f(self.b ** 2 / (2 * self.a)) * (np.exp(f(2 * self.a) * lag) - f(1))


class StationaryLinLinearSDE(LinearSDE):
    """dX = -beta(t)/2 X dt + sqrt(beta(t)) dW, beta linear from beta_min at t0 to beta_max at T (linear.py:48-92)."""

    def __init__(self, beta_min, beta_max, t0, T):
        self.beta_min, self.beta_max, self.t0, self.T = beta_min, beta_max, t0, T

    def beta(self, t):
        slope = (self.beta_max - self.beta_min) / (self.T - self.t0)
        return self.beta_min + slope * (np.asarray(t, dtype=np.float64) - self.t0)

    def beta_integral(self, t, s):
        t = np.asarray(t, dtype=np.float64)
        s = np.asarray(s, dtype=np.float64)
        return 0.5 * (t - s) * (self.beta(t) + self.beta(s))  # trapezoid is exact for a linear beta

    def drift_coef(self, t):
        return -0.5 * self.beta(t)

    def dispersion(self, t):
        return np.sqrt(self.beta(t))

    def transition(self, t, s, dtype=np.float64):
        if dtype == np.float64:
            r = self.beta_integral(t, s)
            return np.exp(-0.5 * r), -np.expm1(-r)
        f = np.dtype(dtype).type
        t, s = np.asarray(t, dtype=dtype), np.asarray(s, dtype=dtype)
        slope = f((self.beta_max - self.beta_min) / (self.T - self.t0))
        icpt2 = f(2 * (self.beta_min * self.T - self.beta_max * self.t0) / (self.T - self.t0))
        r = f(0.5) * (t - s) * (slope * (t + s) + icpt2)  # linear.py:64-67 in float32
        return np.exp(f(-0.5) * r), f(1) - np.exp(-r)


class StationaryExpLinearSDE(LinearSDE):
    """a(t) = a exp(c (t - z)), b(t) = b exp(c (t - z) / 2) (linear.py:95-112)."""

    def __init__(self, a, b, c, z):
        self.a, self.b, self.c, self.z = a, b, c, z

    def drift_coef(self, t):
        return self.a * np.exp(self.c * (np.asarray(t, dtype=np.float64) - self.z))

    def dispersion(self, t):
        return self.b * np.exp(self.c * (np.asarray(t, dtype=np.float64) - self.z) / 2)

    def transition(self, t, s, dtype=np.float64):
        f = np.dtype(dtype).type
        t = np.asarray(t, dtype=dtype)
        s = np.asarray(s, dtype=dtype)
        r = f(self.a) * (np.exp(f(self.c) * (t - f(self.z))) - np.exp(f(self.c) * (s - f(self.z)))) / f(self.c)
        if dtype == np.float64:
            return np.exp(r), self.b ** 2 / (2 * self.a) * np.expm1(2 * r)
        return np.exp(r), f(-self.b ** 2 / (2 * self.a)) * (f(1) - np.exp(f(2) * r))


def _ts_host(ts):
    if isinstance(ts, torch.Tensor):
        ts = ts.detach().cpu().numpy()
    return np.asarray(ts, dtype=np.float64)


def step_coefficients(sde: LinearSDE, ts):
    """float32 (F_k, sqrt(Q_k)) of the K intervals of ``ts`` -- what linear.py:215-216 evaluates per step.

    Evaluated in float32 exactly as the float32 reference does (x64 off), so the rounding of
    ``Q = c (exp(2 a dt) - 1)`` -- a cancellation worth ~1e-5 relative at dt = 0.005 -- is reproduced.
    """
    ts32 = np.asarray(_ts_host(ts), dtype=np.float32)
    F, Q = sde.transition(ts32[1:], ts32[:-1], dtype=np.float32)
    return np.asarray(F, dtype=np.float32), np.sqrt(np.asarray(Q, dtype=np.float32)).astype(np.float32)


def forward_path(key, x0, F, sqrtQ, du=None, rev=False):
    """Launch the OU forward-noising kernel.  x0 [D] or [B, D]; key [2] or [B, 2].

    rev=False -> path [B, K+1, D].  rev=True -> (us [B, K+1, du], vs [B, K+1, D-du]) time-reversed.
    """
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    k = k.reshape(-1, 2)
    B = k.shape[0]
    x = dev(x0, torch.float32)
    batched = x.dim() == 2
    if batched and x.shape[0] != B:
        raise ValueError('x0 batch does not match the number of keys')
    D = x.shape[-1]
    Fd, Sd = dev(F, torch.float32), dev(sqrtQ, torch.float32)
    K = Fd.shape[0]
    if not rev:
        path = empty((B, K + 1, D), torch.float32)
        nat.call('fbs_ou_forward_path_f32', stream(), ptr(k), ptr(x), int(batched), ptr(Fd), ptr(Sd), B, K, D, D, 0,
                 ptr(path), None)
        return out(path[0] if single else path, host)
    us = empty((B, K + 1, du), torch.float32)
    vs = empty((B, K + 1, D - du), torch.float32)
    nat.call('fbs_ou_forward_path_f32', stream(), ptr(k), ptr(x), int(batched), ptr(Fd), ptr(Sd), B, K, D, int(du), 1,
             ptr(us), ptr(vs))
    if single:
        us, vs = us[0], vs[0]
    return out(us, host), out(vs, host)


def make_linear_sde(sde: LinearSDE):
    """``fbs.sdes.make_linear_sde``: (discretise_linear_sde, cond_score_t_0, simulate_cond_forward)."""

    def discretise_linear_sde(t, s):
        F, Q = sde.transition(_ts_host(t), _ts_host(s))
        return np.float32(F) if np.ndim(F) == 0 else F.astype(np.float32), \
            np.float32(Q) if np.ndim(Q) == 0 else Q.astype(np.float32)

    def cond_score_t_0(x, t, x0, s):
        F, Q = discretise_linear_sde(t, s)
        return -(x - float(F) * x0) / float(Q)

    def simulate_cond_forward(key, x0, ts, t0=None, keep_path=True):
        """linear.py:190-225.  With a batch of keys ``[B, 2]``, ``x0`` is ``[B, ...]`` (what the
        reference gets from ``vmap``) or a shared 1-D vector."""
        if not keep_path:
            raise NotImplementedError('keep_path=False is not on the CSMC hot path')
        x = x0 if isinstance(x0, torch.Tensor) else np.asarray(x0, dtype=np.float32)
        batched_key = np.ndim(key) == 2 if not isinstance(key, torch.Tensor) else key.dim() == 2
        batched_x = batched_key and x.ndim >= 2
        feat = tuple(x.shape[1:]) if batched_x else tuple(x.shape)
        flat = x.reshape((x.shape[0], -1)) if batched_x else x.reshape(-1)
        F, sq = step_coefficients(sde, ts)
        path = forward_path(key, flat, F, sq)
        return path.reshape(tuple(path.shape[:-1]) + feat)

    return discretise_linear_sde, cond_score_t_0, simulate_cond_forward


def make_ou_sde(a, b):
    """``fbs.sdes.make_ou_sde`` (linear.py:115-162): time-homogeneous OU, same kernels."""
    sde = StationaryConstLinearSDE(a, b)
    disc, _, sim = make_linear_sde(sde)

    def discretise_ou_sde(t):
        return disc(t, 0.)

    def cond_score_t_0(x, t, x0):
        F, Q = discretise_ou_sde(t)
        return -(x - float(F) * x0) / float(Q)

    def simulate_cond_forward(key, x0, ts, keep_path=True):
        return sim(key, x0, ts, keep_path=keep_path)

    return discretise_ou_sde, cond_score_t_0, simulate_cond_forward
