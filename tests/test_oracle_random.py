"""Pin the oracle's jax.random restatement against every known answer available offline."""
import numpy as np
from oracle import jax_random as jr


def _tf(k0, k1, x0, x1):
    a, b = jr.threefry2x32(np.uint32(k0), np.uint32(k1), np.array([x0], np.uint32), np.array([x1], np.uint32))
    return int(a[0]), int(b[0])


def test_threefry_random123_known_answers():
    # Random123 kat_vectors, threefry2x32 20 rounds
    assert _tf(0, 0, 0, 0) == (0x6b200159, 0x99ba4efe)
    assert _tf(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == (0x1cb996fc, 0xbb002be7)
    assert _tf(0x13198a2e, 0x03707344, 0x243f6a88, 0x85a308d3) == (0xc4923a9c, 0x483df7a0)


def test_published_jax_values():
    key = jr.PRNGKey(0)
    assert key.tolist() == [0, 0]
    # jax docs: jax.random.split(jax.random.PRNGKey(0))
    assert jr.split(key).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    # jax docs: jax.random.uniform(PRNGKey(0)) -> 0.41845703 ; normal(PRNGKey(0), (1,)) -> -0.20584226
    assert np.float32(jr.uniform(key)) == np.float32(0.41845703)
    assert np.float32(jr.normal(key, (1,))[0]) == np.float32(-0.20584226)


def test_random_bits_layout_odd_even():
    key = jr.PRNGKey(42)
    for n in (1, 2, 3, 7, 8, 101):
        bits = jr.random_bits(key, n)
        h = (n + 1) // 2
        for e in (0, n // 2, n - 1):
            b = e if e < h else e - h
            x1 = b + h if b + h < n else 0
            y0, y1 = jr.threefry2x32(key[0], key[1], np.array([b], np.uint32), np.array([x1], np.uint32))
            assert bits[e] == (y0[0] if e < h else y1[0])


def test_uniform_normal_ranges_and_moments():
    u = jr.uniform(jr.PRNGKey(1), (200000,))
    assert u.dtype == np.float32 and u.min() >= 0. and u.max() < 1.
    assert abs(u.mean() - 0.5) < 3e-3
    z = jr.normal(jr.PRNGKey(2), (200000,))
    assert np.isfinite(z).all()
    assert abs(z.mean()) < 1e-2 and abs(z.var() - 1.) < 2e-2


def test_randint_and_choice():
    r = jr.randint(jr.PRNGKey(3), (100000,), 0, 10)
    assert r.min() == 0 and r.max() == 9
    assert np.abs(np.bincount(r, minlength=10) / 1e5 - 0.1).max() < 5e-3
    p = np.array([0.1, 0.2, 0.3, 0.4], dtype=np.float32)
    c = jr.choice(jr.PRNGKey(4), 4, (100000,), p=p)
    assert np.abs(np.bincount(c, minlength=4) / 1e5 - p).max() < 5e-3
    # zero-weight tail is never selected, u = 0 maps to the last positive-weight index
    c2 = jr.choice(jr.PRNGKey(5), 4, (1000,), p=np.array([0.5, 0.5, 0., 0.], np.float32))
    assert c2.max() <= 1


def test_chunked_summation_contract_matches_its_definition():
    """seq_cumsum / seq_sum for rows of >= 1024 elements (the summation-order contract the CUDA kernels follow,
    fbs_resample.cuh kChunkedMinN): chunks of 8, sequential inside a chunk, sequential over the chunk totals -- against a
    literal scalar-loop evaluation of that definition; shorter rows stay the plain sequential sum."""
    rng = np.random.default_rng(0)
    for n in (1023, 1024, 1029, 4096):
        w = rng.random(n).astype(np.float32)
        w /= w.sum()
        got = jr.seq_cumsum(w)
        if n < jr.CHUNKED_MIN_N:
            np.testing.assert_array_equal(got, np.cumsum(w, dtype=np.float32))
            continue
        P = np.float32(0)
        want = np.zeros(n, np.float32)
        for c0 in range(0, n, 8):
            acc = np.float32(0)
            for t in range(c0, min(c0 + 8, n)):
                acc = np.float32(acc + w[t])
                want[t] = np.float32(P + acc)
            P = np.float32(P + acc)
        np.testing.assert_array_equal(got, want)
        assert jr.seq_sum(w) == want[-1]
