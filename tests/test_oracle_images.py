"""Mask logic of fbs/data/images.py:254-361 (oracle restatement) and the result layouts of the reference drivers."""
import numpy as np
import pytest
from oracle import images as oi
from oracle import jax_random as jr


@pytest.mark.parametrize('shape,size', [((28, 28, 1), 15), ((64, 64, 3), 32), ((32, 32, 3), 32)])
def test_inpaint_mask_partitions_the_image(shape, size):
    w, h, c = shape
    for seed in range(5):
        shift, rect, obs = oi.gen_inpaint_mask(jr.PRNGKey(seed), shape, size, size)
        assert len(rect) == min(size, w) * min(size, h) and len(obs) == w * h - len(rect)     # du / dv of SURVEY App. B
        assert np.array_equal(np.sort(np.concatenate([rect, obs])), np.arange(w * h))
        assert 0 <= shift <= max(min(w, h) - size - 1, 0)
        rows, cols = np.divmod(rect, h)
        assert rows.min() == shift and cols.min() == shift and rows.max() == shift + min(size, w) - 1


@pytest.mark.parametrize('rate', [2, 4])
@pytest.mark.parametrize('random', [True, False])
def test_supr_mask_observes_one_pixel_per_block(rate, random):
    shape = (28, 28, 1)
    unobs, obs = oi.gen_supr_mask(jr.PRNGKey(7), shape, rate, random)
    assert len(obs) == (28 // rate) ** 2 and len(unobs) == 28 * 28 - len(obs)
    rows, cols = np.divmod(obs, 28)
    blocks = (rows // rate) * (28 // rate) + cols // rate
    assert np.array_equal(np.sort(blocks), np.arange(len(obs)))                               # one per block
    if not random:
        assert (rows % rate == rate // 2).all() and (cols % rate == rate // 2).all()


def test_unpack_concat_round_trip():
    shape = (28, 28, 1)
    _, rect, obs = oi.gen_inpaint_mask(jr.PRNGKey(1), shape, 15, 15)
    img = np.random.default_rng(0).random((3, 28, 28, 1)).astype(np.float32)
    x, y = oi.unpack(img, shape, rect, obs)
    assert x.shape == (3, 225, 1) and y.shape == (3, 559, 1)
    np.testing.assert_array_equal(oi.concat(x, y, shape, rect, obs), img)


def test_result_layouts(tmp_path):
    """np.save / np.savez layouts of inpainting.py:249-251 and gp_gibbs.py:193-195 (what the tabulators load)."""
    from fbs_b200.data import results
    imgs = np.random.default_rng(1).random((4, 28, 28, 1)).astype(np.float32)
    path = results.save_restored_images(str(tmp_path / 'mnist-inpaint-15-3'), imgs, 'gibbs', explicit_backward=True)
    assert path.endswith('mnist-inpaint-15-3-gibbs-eb.npy')
    np.testing.assert_array_equal(np.load(path), imgs)
    samples = np.zeros((2, 5, 10))
    p2 = results.save_chain_samples(str(tmp_path / 'gibbs-eb-const-10-0'), samples, np.zeros(10), np.eye(10))
    z = np.load(p2)
    assert sorted(z.files) == ['gp_cov', 'gp_mean', 'samples'] and z['samples'].shape == (2, 5, 10)
    with pytest.raises(ValueError):
        results.save_restored_images(str(tmp_path / 'x'), imgs[0])


def test_product_masks_host_paths_match_oracle():
    """The parts of fbs_b200.data.ImageRestore that need no random draw (fixed super-resolution mask, unpack / concat on numpy
    arrays) against the oracle -- on CPU; the random masks are checked on the GPU (tests/test_gpu_images.py)."""
    from fbs_b200.data import ImageRestore
    shape = (28, 28, 1)
    ds = ImageRestore('supr-4', shape, sr_random=False)
    assert ds.unobs_shape == (735, 1)
    m = ds.gen_mask(jr.PRNGKey(0))
    unobs, obs = oi.gen_supr_mask(jr.PRNGKey(0), shape, 4, False)
    np.testing.assert_array_equal(m.obs_inds_ravelled, obs)
    np.testing.assert_array_equal(m.unobs_inds_ravelled, unobs)
    img = np.random.default_rng(2).random((2, 28, 28, 1)).astype(np.float32)
    x, y = ds.unpack(img, m)
    wx, wy = oi.unpack(img, shape, unobs, obs)
    np.testing.assert_array_equal(x, wx)
    np.testing.assert_array_equal(y, wy)
    np.testing.assert_array_equal(ds.concat(x, y, m), img)
    assert ImageRestore('inpaint-15', shape).unobs_shape == (225, 1)
    with pytest.raises(ValueError):
        ImageRestore('denoise-3', shape)
