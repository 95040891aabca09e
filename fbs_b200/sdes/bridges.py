"""Closed-form Gaussian Schroedinger bridge (``fbs/sdes/linear.py:397-457``), host-side float64.

Returns the same three callables as the reference; ``drift`` additionally exposes
``drift.affine(t) -> (A, a)`` with ``drift(x, t) = A x + a`` so the samplers can hand it to the
affine kernels instead of calling it per particle.
"""
import numpy as np


def _sym_sqrt(mat):
    w, v = np.linalg.eigh(mat)
    return (v * np.sqrt(np.clip(w, 0., None))) @ v.T


def make_gaussian_bw_sb(mean0, cov0, mean1, cov1, sig: float = 1.):
    mean0, cov0, mean1, cov1 = (np.asarray(a, dtype=np.float64) for a in (mean0, cov0, mean1, cov1))
    d = mean0.shape[0]
    eye = np.eye(d)
    root0 = _sym_sqrt(cov0)
    d_sig = _sym_sqrt(4. * root0 @ cov1 @ root0 + sig ** 4 * eye)
    c_sig = 0.5 * (root0 @ np.linalg.solve(root0.T, d_sig.T).T - sig ** 2 * eye)

    def marginal_mean(t):
        return (1. - t) * mean0 + t * mean1

    def marginal_cov(t):
        return ((1. - t) ** 2 * cov0 + t ** 2 * cov1 + t * (1. - t) * (c_sig + c_sig.T)
                + sig ** 2 * t * (1. - t) * eye)

    def _s(t):
        return (t * cov1 + (1. - t) * c_sig) - ((1. - t) * cov0 + t * c_sig).T - sig ** 2 * t * eye

    def affine(t):
        t = float(t)
        A = _s(t).T @ np.linalg.inv(marginal_cov(t))
        return A, -A @ marginal_mean(t) - mean0 + mean1

    def drift(x, t):
        A, a = affine(t)
        return np.asarray(x, dtype=np.float64) @ A.T + a

    drift.affine = affine
    return marginal_mean, marginal_cov, drift
