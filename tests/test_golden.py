"""tests/golden/oracle_golden.npz (written by tests/golden/make_golden.py from the oracle; see its docstring for what the
fixtures are and are not): the oracle still reproduces them (CPU), and the CUDA kernels reproduce them (GPU)."""
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'oracle_golden.npz'))
sys.path.insert(0, os.path.join(HERE, 'golden'))


def test_oracle_reproduces_golden():
    import make_golden
    now = make_golden.build()
    assert sorted(now) == sorted(GOLD.files)
    for k in GOLD.files:
        np.testing.assert_array_equal(now[k], GOLD[k], err_msg=k)


@pytest.mark.gpu
def test_cuda_random_reproduces_golden():
    from fbs_b200 import random as fr
    key = GOLD['key']
    np.testing.assert_array_equal(fr.split(key, 5), GOLD['split_5'])
    np.testing.assert_array_equal(fr.random_bits(key, (101,)), GOLD['bits_101'])
    np.testing.assert_array_equal(fr.uniform(key, (64,)), GOLD['uniform_64'])
    np.testing.assert_array_equal(fr.randint(key, (50,), 3, 17), GOLD['randint_50'])
    np.testing.assert_allclose(fr.normal(key, (7, 9)), GOLD['normal_7x9'], rtol=0, atol=5e-7)   # <= 4 ULP contract


@pytest.mark.gpu
@pytest.mark.parametrize('n', [10, 100, 257])
def test_cuda_resampling_reproduces_golden(n):
    from fbs_b200.samplers.csmc import resamplings as RC
    from fbs_b200.samplers import resampling as R
    from oracle import jax_random as jr
    w, k = GOLD[f'w_{n}'], jr.PRNGKey(1000 + n)
    i, j = np.int32(3 % n), np.int32(7 % n)
    np.testing.assert_array_equal(RC.killing(k, w, i, j, True), GOLD[f'cond_killing_{n}'])
    np.testing.assert_array_equal(RC.multinomial(k, w, i, j, True), GOLD[f'cond_multinomial_{n}'])
    np.testing.assert_array_equal(RC.systematic(k, w, conditional=False), GOLD[f'cond_systematic_{n}'])
    np.testing.assert_array_equal(R.stratified(w, k), GOLD[f'stratified_{n}'])
    np.testing.assert_array_equal(R.systematic(w, k), GOLD[f'systematic_{n}'])
    np.testing.assert_array_equal(R.killing(w, k), GOLD[f'killing_{n}'])
    # the same vectors inside a batch large enough for the tiled kernel's many-chain tiles
    B = 700
    ks = np.tile(k[None], (B, 1))
    W = np.tile(w[None], (B, 1))
    got = RC.killing(ks, W, np.full(B, i, np.int32), np.full(B, j, np.int32), True)
    assert (got == GOLD[f'cond_killing_{n}'][None]).all()
    assert (R.stratified(W, ks) == GOLD[f'stratified_{n}'][None]).all()


@pytest.mark.gpu
def test_cuda_masks_reproduce_golden():
    from fbs_b200.data import ImageRestore
    from oracle import jax_random as jr
    m = ImageRestore('inpaint-15', (28, 28, 1)).gen_mask(jr.PRNGKey(5))
    assert m.shift == int(GOLD['inpaint_shift'])
    np.testing.assert_array_equal(m.unobs_inds_ravelled, GOLD['inpaint_rect'])
    s = ImageRestore('supr-4', (28, 28, 1)).gen_mask(jr.PRNGKey(6))
    np.testing.assert_array_equal(s.obs_inds_ravelled, GOLD['supr_obs'])
