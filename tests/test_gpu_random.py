"""CUDA jax.random kernels vs the oracle: bits / split / uniform / randint / choice bit-exact, normals within
2 ULP (the kernel uses CUDA's log1pf and fused multiply-adds, the oracle numpy's log1p and separate ops)."""
import numpy as np
import pytest
from oracle import jax_random as jr
from helpers import ulp_diff

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n', [1, 2, 3, 8, 101, 4096, 100001])
def test_random_bits_exact(n):
    from fbs_b200 import random as fr
    keys = np.stack([jr.PRNGKey(s) for s in (0, 1, 666)] + [jr.split(jr.PRNGKey(9), 2)[1]])
    got = fr.random_bits(keys, (n,))
    for b, k in enumerate(keys):
        np.testing.assert_array_equal(got[b], jr.random_bits(k, n))


def test_split_exact_and_published_value():
    from fbs_b200 import random as fr
    assert fr.split(fr.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    keys = jr.split(jr.PRNGKey(5), 7)
    for num in (2, 3, 4, 6, 200):
        got = fr.split(keys, num)
        for b in range(keys.shape[0]):
            np.testing.assert_array_equal(got[b], jr.split(keys[b], num))


def test_uniform_exact():
    from fbs_b200 import random as fr
    assert np.float32(fr.uniform(fr.PRNGKey(0))) == np.float32(0.41845703)
    key = jr.PRNGKey(11)
    for n in (1, 5, 1000, 33333):
        np.testing.assert_array_equal(fr.uniform(key, (n,)), jr.uniform(key, (n,)))
    np.testing.assert_array_equal(fr.uniform(key, (3, 5), minval=-2., maxval=3.5), jr.uniform(key, (3, 5), -2., 3.5))


def test_normal_within_ulps():
    from fbs_b200 import random as fr
    assert abs(float(fr.normal(fr.PRNGKey(0), (1,))[0]) - (-0.20584226)) < 2e-7
    key = jr.PRNGKey(123)
    n = 1 << 20
    got, want = fr.normal(key, (n,)), jr.normal(key, (n,))
    d = ulp_diff(got, want)
    # near zero the result is tiny and ULP distance is not meaningful: use an absolute floor too
    bad = (d > 4) & (np.abs(got - want) > 2e-7)
    assert not bad.any(), (int(d.max()), float(np.abs(got - want).max()))
    assert (d == 0).mean() > 0.5
    assert np.isfinite(got).all()


def test_randint_exact():
    from fbs_b200 import random as fr
    keys = jr.split(jr.PRNGKey(77), 5)
    for (lo, hi, n) in [(0, 10, 201), (0, 100, 1001), (-3, 4, 17), (0, 1, 5), (0, 101, 64)]:
        got = fr.randint(keys, (n,), lo, hi)
        for b in range(5):
            np.testing.assert_array_equal(got[b], jr.randint(keys[b], (n,), lo, hi))


def test_choice_exact():
    from fbs_b200 import random as fr
    rng = np.random.default_rng(0)
    keys = jr.split(jr.PRNGKey(31), 6)
    for N in (1, 2, 10, 100, 1000, 1024, 3001):              # >= 1024: the chunked summation order of the contract
        p = rng.random((6, N)).astype(np.float32)
        p /= p.sum(axis=1, keepdims=True)
        p[0, N // 2:] = 0.  # zero tail
        got = fr.choice(keys, N, (257,), p=p)
        for b in range(6):
            np.testing.assert_array_equal(got[b], jr.choice(keys[b], N, (257,), p=p[b]))
    assert fr.choice(keys[0], 5, (), p=np.array([0, 0, 1, 0, 0], np.float32)) == 2
