"""Closures of the Schroedinger-bridge image runs.  TEST INFRASTRUCTURE.

Restates ``/root/reference/experiments/sb_imgs/supr.py:84-137`` over the float32 NumPy network of ``oracle/unet.py``:
``reverse_drift`` :84-85 (the RAW output of the backward network at ``T - t``; no ``-a x + g^2 score`` wrapping),
``reverse_drift_u`` / ``_v`` :88-97, ``reverse_dispersion`` :100-101, ``transition_sampler`` :104-110,
``transition_logpdf`` :113-121, ``likelihood_logpdf`` :124-129 and ``fwd_sampler`` :132-137 (Euler--Maruyama with the
FORWARD network as drift, ``integration_nsteps = 1``).  ``concat`` / ``unpack`` are the mask scatter / gather of
``fbs/data/images.py:333-363`` (oracle/images.py).
"""
import math
import numpy as np
from . import jax_random as jr
from . import unet as ou
from .models import norm_logpdf
from .sdes import euler_maruyama


class SBImageModel:
    def __init__(self, param_bwd, param_fwd, sde, ts, T, image_shape, unobs, obs, net_dt):
        self.param_bwd, self.param_fwd, self.sde, self.T = param_bwd, param_fwd, sde, float(T)
        self.ts = np.asarray(ts, dtype=np.float32)
        self.K = self.ts.shape[0] - 1
        self.dt = self.T / self.K                                      # supr.py:46 (python float)
        self.shape, self.unobs, self.obs, self.net_dt = tuple(image_shape), np.asarray(unobs), np.asarray(obs), net_dt

    def concat(self, us, v):
        """dataset.concat for a batch of particles sharing one observed part: us [n, p, c], v [q, c] -> [n, H, W, c]."""
        H, W, C = self.shape
        img = np.zeros((us.shape[0], H * W, C), np.float32)
        img[:, self.unobs] = us
        img[:, self.obs] = v
        return img.reshape(us.shape[0], H, W, C)

    def unpack(self, xy):
        H, W, C = self.shape
        flat = xy.reshape(*xy.shape[:-3], H * W, C)
        return flat[..., self.unobs, :], flat[..., self.obs, :]

    def reverse_drift(self, uv, t):                                    # :84-85
        return ou.unet_forward(self.param_bwd, uv, self.T - float(t), self.net_dt)

    def reverse_dispersion(self, t):                                   # :100-101
        return float(self.sde.dispersion(self.T - float(t)))

    def transition_mean(self, us_prev, v_prev, t_prev):
        rdu, _ = self.unpack(self.reverse_drift(self.concat(us_prev, v_prev), t_prev))
        return us_prev + rdu * np.float32(self.dt)

    def transition_sampler(self, us_prev, v_prev, t_prev, key_):       # :104-110
        sd = np.float32(math.sqrt(self.dt)) * np.float32(self.reverse_dispersion(t_prev))
        return (self.transition_mean(us_prev, v_prev, t_prev) + sd * jr.normal(key_, us_prev.shape)).astype(np.float32)

    def transition_logpdf(self, u, us_prev, v_prev, t_prev):           # :113-121
        sd = np.float32(math.sqrt(self.dt)) * np.float32(self.reverse_dispersion(t_prev))
        return norm_logpdf(u[None], self.transition_mean(us_prev, v_prev, t_prev), sd).sum(axis=(1, 2))

    def likelihood_logpdf(self, v, us_prev, v_prev, t_prev):           # :124-129
        _, rdv = self.unpack(self.reverse_drift(self.concat(us_prev, v_prev), t_prev))
        sd = np.float32(math.sqrt(self.dt)) * np.float32(self.reverse_dispersion(t_prev))
        return norm_logpdf(v[None], v_prev[None] + rdv * np.float32(self.dt), sd).sum(axis=(1, 2))

    def fwd_sampler(self, key_, x0_, y0_):                             # :132-137
        xy0 = self.concat(x0_[None], y0_)[0]

        def fwd_drift(x, t):
            return ou.unet_forward(self.param_fwd, x[None], float(t), self.net_dt)[0]

        return euler_maruyama(key_, xy0, self.ts, fwd_drift, self.sde.dispersion, integration_nsteps=1, return_path=True)
