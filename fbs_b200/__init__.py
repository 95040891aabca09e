"""fbs_b200 -- B200-native (sm_100a) implementation of the fbs CSMC / particle-Gibbs / pMCMC hot path.

Mirrors the call surface of ``fbs.samplers`` and ``fbs.sdes`` (zgbkdlm/fbs); every numerical
operation runs in hand-written CUDA kernels behind the C ABI of ``include/fbs_b200.h``.  There is
no CPU fallback: importing is cheap, but the first call needs ``fbs_b200/_lib/libfbs_b200.so``
(``python -m fbs_b200.build``) and a CUDA device.
"""
from . import random, sdes, samplers  # noqa: F401
from .models import AffineGaussianModel, TwistedAffineModel  # noqa: F401

__version__ = '0.1.0'
