"""fbs_b200.data.ImageRestore (masks drawn with the CUDA jax.random work-alike, unpack / concat on device tensors)
against the oracle restatement of fbs/data/images.py:254-361."""
import numpy as np
import pytest
from oracle import images as oi
from oracle import jax_random as jr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('task,shape', [('inpaint-15', (28, 28, 1)), ('inpaint-32', (64, 64, 3)), ('supr-4', (28, 28, 1)),
                                        ('supr-2', (64, 64, 3))])
def test_masks_match_the_oracle(task, shape):
    from fbs_b200.data import ImageRestore
    ds = ImageRestore(task, shape)
    s = int(task.split('-')[-1])
    for seed in (0, 1, 42):
        key = jr.PRNGKey(seed)
        mask = ds.gen_mask(key)
        if 'inpaint' in task:
            shift, rect, obs = oi.gen_inpaint_mask(key, shape, s, s)
            assert mask.shift == shift and (mask.width, mask.height) == (s, s)
        else:
            rect, obs = oi.gen_supr_mask(key, shape, s, True)
        np.testing.assert_array_equal(mask.unobs_inds_ravelled, rect)
        np.testing.assert_array_equal(mask.obs_inds_ravelled, obs)
        assert (len(mask.unobs_inds_ravelled), shape[2]) == ds.unobs_shape
    fixed = ImageRestore('supr-4', (28, 28, 1), sr_random=False).gen_mask(jr.PRNGKey(0))
    np.testing.assert_array_equal(fixed.obs_inds_ravelled, oi.gen_supr_mask(jr.PRNGKey(0), (28, 28, 1), 4, False)[1])


def test_unpack_concat_on_device():
    import torch
    from fbs_b200.data import ImageRestore
    shape = (28, 28, 1)
    ds = ImageRestore('inpaint-15', shape)
    mask = ds.gen_mask(jr.PRNGKey(3))
    img = np.random.default_rng(0).random((5, 28, 28, 1)).astype(np.float32)
    x, y = ds.unpack(torch.from_numpy(img).cuda(), mask)
    wx, wy = oi.unpack(img, shape, mask.unobs_inds_ravelled, mask.obs_inds_ravelled)
    np.testing.assert_array_equal(x.cpu().numpy(), wx)
    np.testing.assert_array_equal(y.cpu().numpy(), wy)
    np.testing.assert_array_equal(ds.concat(x, y, mask).cpu().numpy(), img)
    np.testing.assert_array_equal(ds.concat(x, y[0], mask).cpu().numpy()[1],
                                  oi.concat(wx[1], wy[0], shape, mask.unobs_inds_ravelled, mask.obs_inds_ravelled))
    np.testing.assert_array_equal(ds.concat(wx, wy, mask), img)                       # numpy in, numpy out
