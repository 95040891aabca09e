"""Model objects that stand in for the reference's Python closures.

The reference's samplers are higher-order functions: the experiment scripts hand them closures
(``transition_sampler``, ``likelihood_logpdf``, ``transition_logpdf``, ``fwd_sampler``, ``unpack``,
``ref_sampler``; ``experiments/toy/gp_gibbs.py:78-150``, ``experiments/sb/gibbs.py:82-164``).  A CUDA
kernel cannot call back into Python, so this package provides objects whose *bound methods* satisfy
the same callable protocol and additionally carry the structured data the fused kernels need.  The
samplers recognise those bound methods and dispatch a whole sweep to one C-ABI call; an opaque
closure is an error (no interpreted fallback).

``AffineGaussianModel``: the joint reverse drift is affine, ``drift(uv, t_k) = M_k uv + m_k`` -- true
for every Gaussian toy / Schroedinger-bridge configuration of the reference.
"""
import ctypes
import math
import numpy as np
import torch
from . import _native as nat
from ._tensor import dev, empty, ptr, stream, out, is_host
from .sdes.linear import LinearSDE, step_coefficients, forward_path
from .sdes.simulators import em_tables, em_path


def tf32_round(x):
    """float32 -> nearest tf32 (10-bit mantissa), ties away from zero: what ``cvt.rna.tf32.f32`` computes."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def pack_umma_image(Mu, du, dv):
    """Tensor-core image of the step matrices (include/fbs_b200.h: ``MTc``).

    ``Mu [K, D, du]``: rows = outputs (u then v), columns = u inputs.  Returns float32
    ``[K, nkb, 2, 2, nout // 8, 8, 4]`` -- per step and 8-input K-block: (hi | lo) x 2 k-chunks x 8-row x 16-byte core
    matrices, the no-swizzle K-major layout a UMMA shared-memory descriptor with LBO = nout/8 * 128, SBO = 128 reads.
    """
    K = Mu.shape[0]
    du8, dv8 = (du + 7) // 8 * 8, (dv + 7) // 8 * 8
    nout = du8 + dv8
    if nout % 16:
        nout += 8
    B = np.zeros((K, nout, du8), dtype=np.float32)
    B[:, :du, :du] = Mu[:, :du, :]
    B[:, du8:du8 + dv, :du] = Mu[:, du:, :]
    hi = tf32_round(B)
    lo = (B - hi).astype(np.float32)
    nkb = du8 // 8
    img = np.empty((K, nkb, 2, 2, nout // 8, 8, 4), dtype=np.float32)
    for part, X in enumerate((hi, lo)):
        # X[k, o, j] with o = rg * 8 + r8, j = kb * 8 + c * 4 + kk  ->  img[k, kb, part, c, rg, r8, kk]
        img[:, :, part] = X.reshape(K, nout // 8, 8, nkb, 2, 4).transpose(0, 3, 4, 1, 2, 5)
    return img


def _host64(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


class AffineGaussianModel:
    """Reverse-diffusion closures with an affine joint drift.

    Parameters (host, float64): ``M [K, D, D]``, ``m [K, D]`` with ``drift(uv, ts[k]) = M[k] uv + m[k]``;
    ``g [K]`` the reverse dispersion at ``ts[k]``; ``dt`` the constant the closures multiply the drift
    with (the reference uses the Python float ``T / nsteps``, gp_gibbs.py:63,121); ``du`` the size of the
    unobserved block (``unpack``: first ``du`` coordinates, gp_gibbs.py:89-90).
    """

    def __init__(self, M, m, g, dt, du, ts, forward=None, ref=None, sde=None):
        M, m, g = _host64(M), _host64(m), _host64(g)
        self.K, self.D = M.shape[0], M.shape[1]
        self.du, self.dv = int(du), self.D - int(du)
        self.ts = np.asarray(_host64(ts), dtype=np.float32)
        if self.ts.shape[0] != self.K + 1:
            raise ValueError('ts must have K + 1 entries')
        self.dt = float(dt)
        self.sde = sde
        self._forward = forward      # ('ou', F, sqrtQ) | ('em', tables)
        self._ref = ref              # (a, Bm, c, L) of the Gaussian terminal conditional
        # float32 constants exactly as the float32 reference forms them: sqrt(dt) (python float) * g (f32)
        sd32 = (np.float32(math.sqrt(self.dt)) * g.astype(np.float32)).astype(np.float32)
        self.host = dict(
            MT=np.ascontiguousarray(np.transpose(M, (0, 2, 1))).astype(np.float32),
            m=m.astype(np.float32),
            dt=np.full((self.K,), self.dt, dtype=np.float32),
            sd=sd32,
            lognorm=(self.dv * np.log(2. * np.pi * sd32.astype(np.float64) ** 2)).astype(np.float32))
        # u-input rows re-packed for the tiled sweep kernel: [K][du][dup | dvp], zero padded (fbs_b200.h: MTp)
        dup, dvp = (self.du + 3) // 4 * 4, (self.dv + 3) // 4 * 4
        MTp = np.zeros((self.K, self.du, dup + dvp), dtype=np.float32)
        MTp[:, :, :self.du] = self.host['MT'][:, :self.du, :self.du]
        MTp[:, :, dup:dup + self.dv] = self.host['MT'][:, :self.du, self.du:]
        self.host['MTp'] = MTp
        # tensor-core image for the tcgen05 sweep kernel (only when the GEMM fits one UMMA tile: nout <= 256)
        du8, dv8 = (self.du + 7) // 8 * 8, (self.dv + 7) // 8 * 8
        if du8 + dv8 + 8 <= 256 + 8 and self.du % 4 == 0:
            self.host['MTc'] = pack_umma_image(np.ascontiguousarray(M[:, :, :self.du]).astype(np.float32), self.du, self.dv)
        self._dev = None
        self._struct = None
        self._ws = None

    # ---------------------------------------------------------------- construction helpers
    @classmethod
    def from_linear_sde(cls, sde: LinearSDE, joint_mean, joint_cov, du, ts, T=None, dt=None):
        """Jointly Gaussian data (X, Y) ~ N(joint_mean, joint_cov) noised by a scalar linear SDE
        (gp_gibbs.py:55-109): marginal at forward time s is N(F_s mu, F_s^2 Sigma + Q_s I), the score is
        linear, hence the reverse drift ``-a(s) z + g(s)^2 score(z, s)`` at ``s = T - t`` is affine."""
        ts64 = _host64(ts)
        K = ts64.shape[0] - 1
        T = float(ts64[-1]) if T is None else float(T)
        dt = T / K if dt is None else float(dt)
        mu, Sigma = _host64(joint_mean), _host64(joint_cov)
        D = mu.shape[0]
        eye = np.eye(D)
        ts32 = ts64.astype(np.float32).astype(np.float64)
        M = np.empty((K, D, D))
        m = np.empty((K, D))
        g = np.empty((K,))
        for k in range(K):
            s = float(np.float32(T) - np.float32(ts32[k]))           # T - t_prev as the f32 closures see it
            F, Q = sde.transition(s, ts32[0])
            prec = np.linalg.solve(F * F * Sigma + Q * eye, eye)
            gk = float(sde.dispersion(s))
            M[k] = -float(sde.drift_coef(s)) * eye - gk * gk * prec
            m[k] = gk * gk * (prec @ (F * mu))
            g[k] = gk
        # terminal conditional p(x_T | y_T) of the noised joint (gp_gibbs.py:84-86,138-141)
        FT, QT = sde.transition(T, ts32[0])
        mT, cT = FT * mu, FT * FT * Sigma + QT * eye
        Bm = np.linalg.solve(cT[du:, du:], cT[du:, :du]).T
        cov_c = cT[:du, :du] - Bm @ cT[du:, :du]
        ref = (mT[:du], Bm, mT[du:], np.linalg.cholesky(cov_c))
        F32, sq32 = step_coefficients(sde, ts64)
        return cls(M, m, g, dt, du, ts64, forward=('ou', F32, sq32), ref=ref, sde=sde)

    @classmethod
    def from_gaussian_sb(cls, joint_mean, joint_cov, ref_mean, ref_cov, du, ts, sig=1., em_nsteps=10):
        """Gaussian Schroedinger bridge between the data joint and a Gaussian reference
        (experiments/sb/gibbs.py:62-150): reverse drift ``-drift_sb(z, T - t) + score(z, T - t)``, unit dispersion."""
        from .sdes.bridges import make_gaussian_bw_sb
        ts64 = _host64(ts)
        K = ts64.shape[0] - 1
        T = float(ts64[-1])
        mm, mc, drift = make_gaussian_bw_sb(joint_mean, joint_cov, ref_mean, ref_cov, sig=sig)
        D = np.asarray(joint_mean).shape[0]
        ts32 = ts64.astype(np.float32)
        M = np.empty((K, D, D))
        m = np.empty((K, D))
        for k in range(K):
            s = float(np.float32(T) - ts32[k])
            A, a = drift.affine(s)
            prec = np.linalg.inv(mc(s))
            M[k] = -A - sig ** 2 * prec
            m[k] = -a + sig ** 2 * prec @ mm(s)
        rm, rc = _host64(ref_mean), _host64(ref_cov)
        Bm = np.linalg.solve(rc[du:, du:], rc[du:, :du]).T
        ref = (rm[:du], Bm, rm[du:], np.linalg.cholesky(rc[:du, :du] - Bm @ rc[du:, :du]))
        tables = em_tables(ts64, drift, lambda _: sig, em_nsteps)
        return cls(M, m, np.full((K,), float(sig)), T / K, du, ts64, forward=('em', tables), ref=ref)

    # ---------------------------------------------------------------- device residency
    def device_arrays(self):
        if self._dev is None:
            self._dev = {k: dev(v, torch.float32) for k, v in self.host.items()}
            st = nat.AffineModelStruct()
            st.K, st.du, st.dv, st.reserved = self.K, self.du, self.dv, 0
            for name in ('MT', 'm', 'dt', 'sd', 'lognorm', 'MTp'):
                setattr(st, name, self._dev[name].data_ptr())
            st.MTc = self._dev['MTc'].data_ptr() if 'MTc' in self._dev else None
            self._struct = st
        return self._dev

    def struct(self):
        self.device_arrays()
        return ctypes.byref(self._struct)

    def workspace(self, B: int):
        """(tensor, nbytes) scratch for the tiled sweep kernel: per-chain step vectors of all K steps.  One buffer per
        CUDA stream: chunks of chains pipelined on different streams (samplers/smc.py) must not share scratch."""
        self.device_arrays()
        nbytes = int(nat.lib().fbs_sweep_workspace_bytes(ctypes.byref(self._struct), int(B)))
        if not isinstance(self._ws, dict):
            self._ws = {}
        sid = torch.cuda.current_stream().cuda_stream
        ws = self._ws.get(sid)
        if ws is None or ws.numel() < nbytes:
            ws = self._ws[sid] = torch.empty((nbytes,), dtype=torch.uint8, device=self._dev['MT'].device)
        return ws, nbytes

    def step_index(self, t_prev) -> int:
        t = float(t_prev.item() if isinstance(t_prev, torch.Tensor) else t_prev)
        k = int(np.argmin(np.abs(self.ts[:-1].astype(np.float64) - t)))
        if abs(float(self.ts[k]) - t) > 1e-6 * max(1., abs(t)):
            raise ValueError(f't_prev={t} is not one of the model step times ts[:-1]')
        return k

    # ---------------------------------------------------------------- the closure protocol
    def unpack(self, xy, **kwargs):
        return xy[..., :self.du], xy[..., self.du:]

    def _eval(self, k, key, us_prev, v, v_prev, u_eval, want):
        host = is_host(us_prev)
        up = dev(us_prev, torch.float32)
        single = up.dim() == 2
        if single:
            up = up.unsqueeze(0)
        B, N = up.shape[0], up.shape[1]

        def vec(x, d):
            if x is None:
                return None
            t = dev(x, torch.float32).reshape(-1, d)
            return t.expand(B, d).contiguous() if t.shape[0] == 1 and B > 1 else t

        vp, vv, ue = vec(v_prev, self.dv), vec(v, self.dv), vec(u_eval, self.du)
        kk = None if key is None else dev(key, torch.uint32).reshape(-1, 2)
        us_out = empty((B, N, self.du), torch.float32) if want == 'us' else None
        lw = empty((B, N), torch.float32) if want == 'lw' else None
        tlp = empty((B, N), torch.float32) if want == 'tlp' else None
        nat.call('fbs_affine_eval_f32', stream(), self.struct(), int(k), ptr(kk), ptr(up), ptr(vv), ptr(vp), ptr(ue), B,
                 N, ptr(us_out), ptr(lw), ptr(tlp))
        res = {'us': us_out, 'lw': lw, 'tlp': tlp}[want]
        return out(res[0] if single else res, host)

    def transition_sampler(self, us_prev, v_prev, t_prev, key, **kwargs):
        """(n, du), (dv,), float, key -> (n, du)   -- gp_gibbs.py:120-122."""
        return self._eval(self.step_index(t_prev), key, us_prev, None, v_prev, None, 'us')

    def likelihood_logpdf(self, v, us_prev, v_prev, t_prev, **kwargs):
        """(dv,), (n, du), (dv,), float -> (n,)    -- gp_gibbs.py:132-135."""
        return self._eval(self.step_index(t_prev), None, us_prev, v, v_prev, None, 'lw')

    def transition_logpdf(self, u, us_prev, v_prev, t_prev, **kwargs):
        """(du,), (n, du), (dv,), float -> (n,)    -- gp_gibbs.py:125-129."""
        return self._eval(self.step_index(t_prev), None, us_prev, None, v_prev, u, 'tlp')

    def fwd_sampler(self, key, x0, y0, **kwargs):
        """Forward-noise the joint (x0, y0) over ``ts`` -> path [.., K+1, D] (gp_gibbs.py:144-145; sb/gibbs.py:141-143)."""
        xy0 = self._joint0(key, x0, y0)
        if self._forward[0] == 'ou':
            return forward_path(key, xy0, self._forward[1], self._forward[2])
        return em_path(key, xy0, self._forward[1])

    def fwd_sampler_reversed(self, key, x0, y0):
        """Same draw as ``fwd_sampler`` but returned as ``(us, vs) = (path_x[::-1], path_y[::-1])`` (gibbs.py:128-130),
        written directly in that layout by the kernel."""
        xy0 = self._joint0(key, x0, y0)
        if self._forward[0] == 'ou':
            return forward_path(key, xy0, self._forward[1], self._forward[2], du=self.du, rev=True)
        return em_path(key, xy0, self._forward[1], du=self.du, rev=True)

    def fwd_ys_sampler(self, key, y0, **kwargs):
        """Forward-noise y alone (separable forward process; gp_gibbs.py:148-149)."""
        if self._forward[0] != 'ou':
            raise NotImplementedError('fwd_ys_sampler needs a separable (scalar linear SDE) forward process')
        return forward_path(key, y0, self._forward[1], self._forward[2])

    def ref_sampler(self, key, yT, nsamples, **kwargs):
        """Draw ``nsamples`` from the terminal conditional p(x_T | y_T) (gp_gibbs.py:138-141)."""
        host = is_host(key)
        kk = dev(key, torch.uint32)
        single = kk.dim() == 1
        kk = kk.reshape(-1, 2)
        B = kk.shape[0]
        yt = dev(yT, torch.float32).reshape(-1, self.dv)
        if yt.shape[0] == 1 and B > 1:
            yt = yt.expand(B, self.dv).contiguous()
        if not hasattr(self, '_ref_dev'):
            self._ref_dev = [dev(np.ascontiguousarray(a), torch.float32) for a in self._ref]
        a, Bm, c, L = self._ref_dev
        o = empty((B, int(nsamples), self.du), torch.float32)
        nat.call('fbs_gaussian_ref_sample_f32', stream(), ptr(kk), ptr(yt), ptr(a), ptr(Bm), ptr(c), ptr(L), B,
                 int(nsamples), self.du, self.dv, ptr(o))
        return out(o[0] if single else o, host)

    def _joint0(self, key, x0, y0):
        host_key = is_host(key)
        kk_batched = (np.ndim(key) == 2) if host_key else (key.dim() == 2)
        x = dev(x0, torch.float32)
        y = dev(y0, torch.float32)
        if kk_batched:
            B = key.shape[0]
            x = x.reshape(-1, self.du)
            y = y.reshape(-1, self.dv)
            if x.shape[0] == 1 and B > 1:
                x = x.expand(B, -1)
            if y.shape[0] == 1 and B > 1:
                y = y.expand(B, -1)
            return torch.cat([x, y], dim=1).contiguous()
        return torch.cat([x.reshape(-1), y.reshape(-1)]).contiguous()


class TwistedAffineModel:
    """The closures of the twisted-SMC comparison sampler of the toy experiments (experiments/toy/gp_twisted.py:66-129): the
    reverse diffusion of the X-marginal of a Gaussian ``N(mean_x, cov_x)`` under a scalar linear SDE (affine drift
    ``M_t u + m_t``), the Gaussian twisting function ``p~(y | u, t) = N(y; u + reverse_drift(u, t) dt, obs_var)`` and the
    proposal whose drift adds ``g^2 grad_u log p~`` (analytic here: the denoising estimate is affine in ``u``).

    The bound methods satisfy the callable protocol of ``twisted_smc`` (fbs/samplers/smc.py:261-268) and carry the per-time
    coefficient tables (time indices 0..K: the scan walks ``ts[1:]``, the initial twisting uses ``ts[0]``) the one-launch
    kernel ``fbs_twisted_smc_affine_f32`` reads.  Called on their own they raise: the sampler fuses them.
    """

    def __init__(self, sde: LinearSDE, mean_x, cov_x, obs_var, ts, T=None):
        ts64 = _host64(ts)
        self.K = ts64.shape[0] - 1
        self.T = float(ts64[-1]) if T is None else float(T)
        self.dt = self.T / self.K                                     # gp_twisted.py:57
        self.obs_var = float(obs_var)
        self.ts = ts64.astype(np.float32)
        mu, Sigma = _host64(mean_x), _host64(cov_x)
        self.d = d = mu.shape[0]
        eye = np.eye(d)
        M = np.empty((self.K + 1, d, d))
        m = np.empty((self.K + 1, d))
        g = np.empty((self.K + 1,))
        ts32 = self.ts.astype(np.float64)
        for k in range(self.K + 1):
            s = float(np.float32(self.T) - np.float32(ts32[k]))       # T - t as the float32 closures see it
            F, Q = sde.transition(s, ts32[0])
            prec = np.linalg.solve(F * F * Sigma + Q * eye, eye)
            gk = float(sde.dispersion(s))
            M[k] = -float(sde.drift_coef(s)) * eye - gk * gk * prec  # reverse_drift, gp_twisted.py:84-85
            m[k] = gk * gk * (prec @ (F * mu))
            g[k] = gk
        FT, QT = sde.transition(self.T, ts32[0])
        self._ref = (FT * mu, np.linalg.cholesky(FT * FT * Sigma + QT * eye))   # terminal reference, gp_twisted.py:80,108-111
        self.host = dict(MT=np.ascontiguousarray(np.transpose(M, (0, 2, 1))).astype(np.float32), Mr=M.astype(np.float32),
                         m=m.astype(np.float32), g2=(g * g).astype(np.float32),
                         sd=(np.float32(math.sqrt(self.dt)) * g.astype(np.float32)).astype(np.float32))
        self._dev = None

    def device_arrays(self):
        if self._dev is None:
            self._dev = {k: dev(v, torch.float32) for k, v in self.host.items()}
            self._dev['ref_m'] = dev(self._ref[0], torch.float32)
            self._dev['ref_LT'] = dev(np.ascontiguousarray(self._ref[1].T), torch.float32)
        return self._dev

    def init_sampler(self, key, nparticles):
        """gp_twisted.py:108-111: ``m_ref + normal(key, (n, d)) @ cholesky(cov_ref)^T`` (batched over keys ``[B, 2]``), by the
        Gaussian reference-sampling kernel (``fbs_gaussian_ref_sample_f32`` with an empty conditioning part)."""
        host = is_host(key)
        kk = dev(key, torch.uint32)
        single = kk.dim() == 1
        kk = kk.reshape(-1, 2)
        B = kk.shape[0]
        a = self.device_arrays()
        if 'zero1' not in a:
            a['zero1'] = torch.zeros((max(B, 1),), dtype=torch.float32, device=kk.device)
            a['zeroB'] = torch.zeros((self.d,), dtype=torch.float32, device=kk.device)
        if a['zero1'].numel() < B:
            a['zero1'] = torch.zeros((B,), dtype=torch.float32, device=kk.device)
        o = empty((B, int(nparticles), self.d), torch.float32)
        nat.call('fbs_gaussian_ref_sample_f32', stream(), ptr(kk), ptr(a['zero1']), ptr(a['ref_m']), ptr(a['zeroB']), ptr(a['zero1']),
                 ptr(a['ref_LT']), B, int(nparticles), self.d, 1, ptr(o))
        return out(o[0] if single else o, host)

    def _fused(self, *_, **__):
        raise TypeError('the closures of a TwistedAffineModel are consumed by the fused twisted_smc kernel; they are not called')

    transition_logpdf = twisting_logpdf = twisting_prop_sampler = twisting_prop_logpdf = _fused
