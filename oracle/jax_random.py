"""NumPy restatement of the ``jax.random`` pieces the fbs hot path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The source restated here is NOT under
/root/reference: it is ``jax==0.4.26`` (pinned in /root/reference/requirements_freeze.txt),
default ``jax_threefry_partitionable=False``, ``jax_enable_x64=False``.  The algorithm is
the published Threefry-2x32-20 (Salmon et al., "Parallel random numbers: as easy as 1, 2,
3", SC'11; Random123) and JAX's public documentation of how bits become samples.

Reference call sites that depend on these semantics:
``fbs/samplers/csmc/csmc.py:65,136,150,157,194,210,297``;
``fbs/samplers/csmc/resamplings.py:34,66,71,74,84,122``;
``fbs/samplers/resampling.py:38,46-48``; ``fbs/samplers/gibbs.py:126,134,147,156,197,207-208``;
``fbs/samplers/smc.py:61,77,79,104,109,142,154,165,231,248``; ``fbs/sdes/linear.py:220,224``;
``fbs/sdes/simulators.py:81,91``.

Pinned by tests/test_oracle_random.py against the Random123 KATs and the three values
published in JAX's docs; ``randint`` and ``choice`` streams are parity-unpinned.
"""
import numpy as np

_U32 = np.uint32
_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)


def PRNGKey(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey`` with x64 off: ``uint32[2] = [0, seed]`` for 0 <= seed < 2**32."""
    seed = int(seed)
    if not 0 <= seed < 2 ** 32:
        raise ValueError('oracle PRNGKey restates the 0 <= seed < 2**32 case only')
    return np.array([0, seed], dtype=_U32)


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  All arguments broadcastable uint32 arrays.

    Key schedule ``ks = [k0, k1, k0 ^ k1 ^ 0x1BD11BDA]``; five groups of four rounds with
    rotation sets alternating (13,15,26,6)/(17,29,16,24); after group g (0-based) inject
    ``x0 += ks[(g+1)%3]; x1 += ks[(g+2)%3] + (g+1)``.
    """
    with np.errstate(over='ignore'):
        k0 = np.asarray(k0, dtype=_U32)
        k1 = np.asarray(k1, dtype=_U32)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = np.asarray(x0, dtype=_U32) + ks[0]
        x1 = np.asarray(x1, dtype=_U32) + ks[1]
        for g in range(5):
            for r in (_ROT_A if g % 2 == 0 else _ROT_B):
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + _U32(g + 1)
    return x0, x1


def random_bits(key, n: int) -> np.ndarray:
    """``_threefry_random_bits_original`` for 32-bit words, flat output of length ``n``.

    Counters ``iota(n)`` zero-padded to even length ``2h``; the first half feeds ``x0`` and
    the second half ``x1`` of the same block, so outputs ``i`` and ``i + h`` share a block.
    """
    key = np.asarray(key, dtype=_U32)
    n = int(n)
    if n == 0:
        return np.zeros((0,), dtype=_U32)
    h = (n + 1) // 2
    c = np.arange(2 * h, dtype=np.uint64)
    c[n:] = 0  # odd n: one zero pad
    c = c.astype(_U32)
    y0, y1 = threefry2x32(key[0], key[1], c[:h], c[h:])
    return np.concatenate([y0, y1])[:n]


def split(key, num: int = 2) -> np.ndarray:
    """``jax.random.split``: ``random_bits`` over ``2*num`` counters reshaped ``(num, 2)``."""
    return random_bits(key, 2 * int(num)).reshape(int(num), 2)


def _bits_to_unit_float(bits: np.ndarray) -> np.ndarray:
    fb = (bits >> _U32(9)) | _U32(0x3F800000)
    return fb.view(np.float32) - np.float32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    """``jax.random.uniform`` (float32): mantissa-fill, ``max(lo, f*(hi-lo)+lo)``."""
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
    n = int(np.prod(shape)) if len(shape) else 1
    f = _bits_to_unit_float(random_bits(key, n))
    lo = np.float32(minval)
    hi = np.float32(maxval)
    out = np.maximum(lo, f * np.float32(hi - lo) + lo)
    return out.reshape(shape).astype(np.float32)


# XLA's float32 erf_inv (Giles, "Approximating the erfinv function"), Horner, fp32.
_ERFINV_CENTRAL = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
                   -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
_ERFINV_TAIL = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
                -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)


def erf_inv(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        w = -np.log1p(-(x * x)).astype(np.float32)
        small = w < np.float32(5.0)
        wc = w - np.float32(2.5)
        wt = np.sqrt(w).astype(np.float32) - np.float32(3.0)
        ww = np.where(small, wc, wt).astype(np.float32)
        p = np.where(small, np.float32(_ERFINV_CENTRAL[0]), np.float32(_ERFINV_TAIL[0])).astype(np.float32)
        for cc, ct in zip(_ERFINV_CENTRAL[1:], _ERFINV_TAIL[1:]):
            p = (np.where(small, np.float32(cc), np.float32(ct)).astype(np.float32) + p * ww).astype(np.float32)
        out = (p * x).astype(np.float32)
        out = np.where(np.abs(x) == np.float32(1.0), np.float32(np.inf) * x, out)
    return out.astype(np.float32)


_NORMAL_LO = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
_SQRT2 = np.float32(np.sqrt(2))


def normal(key, shape=()) -> np.ndarray:
    """``jax.random.normal`` (float32): ``sqrt(2) * erf_inv(uniform(nextafter(-1, 0), 1))``."""
    u = uniform(key, shape, minval=_NORMAL_LO, maxval=np.float32(1.0))
    return (_SQRT2 * erf_inv(u)).astype(np.float32)


CHUNKED_MIN_N = 1024   # rows at least this long use the chunked order below (kChunkedMinN in fbs_resample.cuh)


def seq_cumsum(w: np.ndarray) -> np.ndarray:
    """Cumulative sum in the array's own dtype, in the summation order of the oracle convention (see __init__):
    sequential ``c[i] = fl(c[i-1] + w[i])`` for rows shorter than CHUNKED_MIN_N; for longer rows chunks of 8 consecutive
    elements, ``local_c`` = sequential sums inside chunk c, ``P_0 = 0``, ``P_{c+1} = fl(P_c + local_c[-1])``,
    ``c[8 c + t] = fl(P_c + local_c[t])`` -- the serial chain is n / 8 additions, which is what lets a 16384-particle row
    (BASELINE.json configs[4]) be scanned in microseconds.  XLA fixes no order; this is the contract both sides follow."""
    w = np.asarray(w)
    n = w.shape[0]
    if w.ndim != 1 or n < CHUNKED_MIN_N:
        return np.cumsum(w, dtype=w.dtype)
    pad = (-n) % 8
    x = np.concatenate([w, np.zeros(pad, dtype=w.dtype)]).reshape(-1, 8)
    local = np.cumsum(x, axis=1, dtype=w.dtype)
    P = np.concatenate([np.zeros(1, dtype=w.dtype), np.cumsum(local[:, -1], dtype=w.dtype)[:-1]])
    return (P[:, None] + local).astype(w.dtype).reshape(-1)[:n]


def seq_sum(w: np.ndarray):
    w = np.asarray(w)
    if w.shape[0] == 0:
        return w.dtype.type(0)
    return seq_cumsum(w)[-1]


def randint(key, shape, minval: int, maxval: int) -> np.ndarray:
    """``jax.random.randint`` (int32).  Two 32-bit draws combined modulo ``span``."""
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    n = int(np.prod(shape)) if len(shape) else 1
    k1, k2 = split(key, 2)
    hi_bits = random_bits(k1, n).astype(np.uint64)
    lo_bits = random_bits(k2, n).astype(np.uint64)
    span = np.uint64(max(int(maxval) - int(minval), 1) if maxval > minval else 1)
    mult = np.uint64((2 ** 16) % int(span))
    mult = np.uint64((int(mult) * int(mult)) % (2 ** 32)) % span  # uint32 wrap, then rem
    off = (((hi_bits % span) * mult) % np.uint64(2 ** 32) + (lo_bits % span)) % np.uint64(2 ** 32)
    off = off % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32).reshape(shape)


def choice(key, n: int, shape=(), p=None) -> np.ndarray:
    """``jax.random.choice(key, n, shape, replace=True, p=p)`` for integer ``n``.

    ``c = cumsum(p); r = c[-1] * (1 - uniform(key, shape)); searchsorted(c, r)`` (side left).
    """
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    if p is None:
        return randint(key, shape, 0, n)
    p = np.asarray(p)
    c = seq_cumsum(p)
    u = uniform(key, shape).astype(p.dtype) if p.dtype == np.float32 else uniform64(key, shape)
    r = c[-1] * (p.dtype.type(1) - u)
    return np.searchsorted(c, r, side='left').astype(np.int32)


def uniform64(key, shape=()) -> np.ndarray:
    """float64 uniform under ``jax_enable_x64`` (two 32-bit words per draw, 52 mantissa bits).

    Only used by the float64 statistical tests, which do not depend on the stream.
    """
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    n = int(np.prod(shape)) if len(shape) else 1
    k1, k2 = split(key, 2)
    hi = random_bits(k1, n).astype(np.uint64)
    lo = random_bits(k2, n).astype(np.uint64)
    bits = (hi << np.uint64(32)) | lo
    fb = (bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)
    return (fb.view(np.float64) - 1.0).reshape(shape)


def normal64(key, shape=()) -> np.ndarray:
    """float64 normal under ``jax_enable_x64`` (statistical tests only; stream not pinned)."""
    from scipy.special import erfinv
    lo = np.nextafter(-1.0, 0.0)
    u = np.maximum(lo, uniform64(key, shape) * (1.0 - lo) + lo)
    return np.sqrt(2.0) * erfinv(u)
