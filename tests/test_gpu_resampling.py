"""Resampling kernels vs the oracle: ancestor indices bit-exact on identical weights and keys (the same float32 summation
order on both sides: sequential below 1024 particles, the chunked order of oracle.jax_random.seq_cumsum from there on), plus
the reference's own acceptance criteria (tests/test_cond_resamplings.py:15-53) on the CUDA kernels."""
import numpy as np
import pytest
from oracle import jax_random as jr
from oracle import cond_resampling as ocr
from oracle import resampling as orx

pytestmark = pytest.mark.gpu


def _weights(rng, B, N, kind):
    if kind == 'random':
        w = rng.random((B, N)).astype(np.float32) ** 3
    elif kind == 'uniform':
        w = np.ones((B, N), np.float32)
    elif kind == 'onehot':
        w = np.zeros((B, N), np.float32)
        w[np.arange(B), rng.integers(0, N, B)] = 1.
    elif kind == 'peaked':
        w = np.exp(-rng.random((B, N)) * 30).astype(np.float32)
    else:  # zero tail
        w = rng.random((B, N)).astype(np.float32)
        w[:, N // 2 + 1:] = 0.
    return (w / w.sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)


@pytest.mark.parametrize('N', [1, 2, 3, 10, 33, 100, 101, 1000, 1023, 1024, 1029, 4096, 12001, 16384])
@pytest.mark.parametrize('kind', ['random', 'uniform', 'onehot', 'peaked', 'zerotail'])
def test_conditional_indices_exact(N, kind):
    from fbs_b200.samplers.csmc import resamplings as R
    rng = np.random.default_rng(N * 7 + len(kind))
    B = 8 if N < 4096 else 3
    w = _weights(rng, B, N, kind)
    keys = jr.split(jr.PRNGKey(N + 3), B)
    i = rng.integers(0, N, B).astype(np.int32)
    j = rng.integers(0, N, B).astype(np.int32)
    for name in ('killing', 'multinomial'):
        got = getattr(R, name)(keys, w, i, j, True)
        got_u = getattr(R, name)(keys, w, conditional=False)
        for b in range(B):
            np.testing.assert_array_equal(got[b], getattr(ocr, name)(keys[b], w[b], int(i[b]), int(j[b]), True),
                                          err_msg=f'{name} cond b={b}')
            np.testing.assert_array_equal(got_u[b], getattr(ocr, name)(keys[b], w[b], conditional=False),
                                          err_msg=f'{name} uncond b={b}')
        assert (got[np.arange(B), j] == i).all()           # pinned slot (resamplings.py:36,86)
    got_s = R.systematic(keys, w, conditional=False)
    for b in range(B):
        np.testing.assert_array_equal(got_s[b], ocr.systematic(keys[b], w[b], conditional=False))
    with pytest.raises(NotImplementedError):
        R.systematic(keys, w, i, j, True)                    # resamplings.py:129


@pytest.mark.parametrize('N', [1, 2, 10, 100, 101, 1000, 1024, 1500, 16384])
@pytest.mark.parametrize('kind', ['random', 'uniform', 'onehot', 'peaked', 'zerotail'])
def test_unconditional_indices_exact(N, kind):
    from fbs_b200.samplers import resampling as R
    rng = np.random.default_rng(N * 13 + len(kind))
    B = 8 if N < 4096 else 3
    w = _weights(rng, B, N, kind)
    keys = jr.split(jr.PRNGKey(N + 5), B)
    for name in ('stratified', 'systematic', 'killing'):
        got = getattr(R, name)(w, keys)
        for b in range(B):
            np.testing.assert_array_equal(got[b], getattr(orx, name)(w[b], keys[b]), err_msg=f'{name} b={b}')
        assert got.min() >= 0 and got.max() <= N - 1
    # sorted-uniform multinomial ("Not tested." upstream): -log(u) differs between libms at the ULP level, so
    # demand monotone output and near-total agreement rather than bit equality
    got = R.multinomial(w, keys)
    agree = np.mean([np.mean(got[b] == orx.multinomial(w[b], keys[b])) for b in range(B)])
    assert agree > 0.99
    assert (np.diff(got, axis=1) >= 0).all()


@pytest.mark.parametrize('N', [2, 10, 33, 100, 101, 257, 1000, 1400])
def test_many_chains_exact(N):
    """Enough chains that the tiled kernel holds many chains per CTA (one thread per chain for the sequential sums,
    items of several chains inside one warp) and a ragged last tile."""
    from fbs_b200.samplers.csmc import resamplings as RC
    from fbs_b200.samplers import resampling as R
    rng = np.random.default_rng(N)
    B = 148 * 4 * 3 + 37 if N <= 101 else 700
    w = np.concatenate([_weights(rng, B - B // 2, N, 'random'), _weights(rng, B // 2, N, 'peaked')])
    keys = jr.split(jr.PRNGKey(N + 11), B)
    i = rng.integers(0, N, B).astype(np.int32)
    j = rng.integers(0, N, B).astype(np.int32)
    check = rng.choice(B, 48, replace=False).tolist() + [0, B - 1]
    for name in ('killing', 'multinomial'):
        got = getattr(RC, name)(keys, w, i, j, True)
        got_u = getattr(RC, name)(keys, w, conditional=False)
        for b in check:
            np.testing.assert_array_equal(got[b], getattr(ocr, name)(keys[b], w[b], int(i[b]), int(j[b]), True),
                                          err_msg=f'{name} cond b={b}')
            np.testing.assert_array_equal(got_u[b], getattr(ocr, name)(keys[b], w[b], conditional=False),
                                          err_msg=f'{name} uncond b={b}')
        assert (got[np.arange(B), j] == i).all()
    got_s = RC.systematic(keys, w, conditional=False)
    for b in check:
        np.testing.assert_array_equal(got_s[b], ocr.systematic(keys[b], w[b], conditional=False))
    for name in ('stratified', 'systematic', 'killing'):
        got = getattr(R, name)(w, keys)
        for b in check:
            np.testing.assert_array_equal(got[b], getattr(orx, name)(w[b], keys[b]), err_msg=f'{name} b={b}')


def test_single_vector_and_killing_identity():
    from fbs_b200.samplers.csmc import resamplings as R
    key = jr.PRNGKey(1)
    w = np.full((50,), 1 / 50, np.float32)
    got = R.killing(key, w, 7, 7, True)
    np.testing.assert_array_equal(got, np.arange(50))       # uniform weights: nothing is killed, J = i
    got = R.killing(key, w, 3, 10, True)
    np.testing.assert_array_equal(got, ocr.killing(key, w, 3, 10, True))
    assert got[10] == 3


@pytest.mark.parametrize('name', ['multinomial', 'killing'])
@pytest.mark.parametrize('seed', [42, 666])
def test_reference_unconditional_criterion(name, seed):
    """tests/test_cond_resamplings.py:15-29 (N=1000, 100 000 keys) on the CUDA kernels."""
    from fbs_b200.samplers.csmc import resamplings as R
    from fbs_b200 import random as fr
    keys = fr.split(fr.PRNGKey(seed), 100_000)
    weights = np.cos(np.linspace(0, 2 * np.pi, 1000, dtype=np.float32)) + 1
    weights = (weights / weights.sum()).astype(np.float32)
    idx = getattr(R, name)(keys, np.broadcast_to(weights, (100_000, 1000)).copy(), conditional=False)
    bincount = np.bincount(idx[:, 1:].ravel(), minlength=1000)
    np.testing.assert_allclose(bincount / bincount.sum(), weights, atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize('name', ['multinomial', 'killing'])
@pytest.mark.parametrize('j', [0, 5, 50])
def test_reference_conditional_bayes_criterion(name, j):
    """tests/test_cond_resamplings.py:33-53 (N=100, 100 000 keys) on the CUDA kernels."""
    from fbs_b200.samplers.csmc import resamplings as R
    from fbs_b200 import random as fr
    N = 100
    keys = fr.split(fr.PRNGKey(42), 100_000)
    weights = np.cos(np.linspace(0, 2 * np.pi, N, dtype=np.float32)) + 1
    weights = (weights / weights.sum()).astype(np.float32)
    k12 = fr.split(keys, 2)
    W = np.broadcast_to(weights, (100_000, N)).copy()
    pivot = fr.choice(np.ascontiguousarray(k12[:, 0]), N, (), p=W)
    idx = getattr(R, name)(np.ascontiguousarray(k12[:, 1]), W, pivot.astype(np.int32), np.full(100_000, j, np.int32), True)
    np.testing.assert_allclose(idx[:, j], pivot, atol=1e-3)
    bincount = np.bincount(idx[:, 1:].ravel(), minlength=N)
    np.testing.assert_allclose(bincount / bincount.sum(), weights, atol=1e-3, rtol=1e-3)
