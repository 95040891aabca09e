"""``jax.random`` work-alikes on the GPU (threefry2x32, jax 0.4.26 non-partitionable bit layout).

Keys are ``uint32[2]`` (or ``[B, 2]`` for a batch of chains).  Passing numpy keys makes the call a
host-buffer call (inputs copied H2D, results returned as numpy); passing CUDA tensors keeps
everything on the device.  Kernels: fbs_b200/csrc/random_kernels.cu.
"""
import numpy as np
import torch
from . import _native as nat
from ._tensor import dev, empty, ptr, stream, out, is_host


def PRNGKey(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey`` with x64 disabled: ``[0, seed]``."""
    seed = int(seed)
    if not 0 <= seed < 2 ** 32:
        raise ValueError('seed must be in [0, 2**32)')
    return np.array([0, seed], dtype=np.uint32)


def _keys(key):
    host = is_host(key)
    k = dev(key, torch.uint32)
    single = k.dim() == 1
    if single:
        k = k.reshape(1, 2)
    if k.dim() != 2 or k.shape[1] != 2:
        raise ValueError(f'key must have shape (2,) or (B, 2), got {tuple(k.shape)}')
    return k, single, host


def _shape(shape):
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(int(s) for s in shape)


def split(key, num: int = 2):
    k, single, host = _keys(key)
    o = empty((k.shape[0], num, 2), torch.uint32)
    nat.call('fbs_random_split', stream(), ptr(k), k.shape[0], int(num), ptr(o))
    return out(o[0] if single else o, host)


def random_bits(key, shape):
    k, single, host = _keys(key)
    shape = _shape(shape)
    n = int(np.prod(shape)) if shape else 1
    o = empty((k.shape[0], n), torch.uint32)
    nat.call('fbs_random_bits_u32', stream(), ptr(k), k.shape[0], n, ptr(o))
    o = o.reshape((k.shape[0],) + shape)
    return out(o[0] if single else o, host)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    k, single, host = _keys(key)
    shape = _shape(shape)
    n = int(np.prod(shape)) if shape else 1
    o = empty((k.shape[0], n), torch.float32)
    nat.call('fbs_random_uniform_f32', stream(), ptr(k), k.shape[0], n, float(np.float32(minval)),
             float(np.float32(maxval)), ptr(o))
    o = o.reshape((k.shape[0],) + shape)
    return out(o[0] if single else o, host)


def normal(key, shape=()):
    k, single, host = _keys(key)
    shape = _shape(shape)
    n = int(np.prod(shape)) if shape else 1
    o = empty((k.shape[0], n), torch.float32)
    nat.call('fbs_random_normal_f32', stream(), ptr(k), k.shape[0], n, ptr(o))
    o = o.reshape((k.shape[0],) + shape)
    return out(o[0] if single else o, host)


def randint(key, shape, minval: int, maxval: int):
    k, single, host = _keys(key)
    shape = _shape(shape)
    n = int(np.prod(shape)) if shape else 1
    o = empty((k.shape[0], n), torch.int32)
    nat.call('fbs_random_randint_i32', stream(), ptr(k), k.shape[0], n, int(minval), int(maxval), ptr(o))
    o = o.reshape((k.shape[0],) + shape)
    return out(o[0] if single else o, host)


def choice(key, n: int, shape=(), p=None):
    """``jax.random.choice(key, n, shape, replace=True, p=p)`` for integer ``n``."""
    if p is None:
        return randint(key, shape, 0, n)
    k, single, host = _keys(key)
    shape = _shape(shape)
    nd = int(np.prod(shape)) if shape else 1
    pt = dev(p, torch.float32)
    if pt.dim() == 1:
        pt = pt.reshape(1, -1).expand(k.shape[0], -1).contiguous()
    if pt.shape != (k.shape[0], n):
        raise ValueError(f'p must have shape ({n},) or (B, {n})')
    o = empty((k.shape[0], nd), torch.int32)
    nat.call('fbs_random_choice_f32', stream(), ptr(k), ptr(pt), k.shape[0], int(n), nd, ptr(o))
    o = o.reshape((k.shape[0],) + shape)
    return out(o[0] if single else o, host)
