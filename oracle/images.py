"""TEST INFRASTRUCTURE (see oracle/__init__.py).  NumPy restatement of the mask logic of fbs/data/images.py:254-361,
written as the reference writes it (itertools.product index lists, ravel_multi_index with mode='clip', setdiff1d)."""
import itertools
import numpy as np
from . import jax_random as jr


def gen_supr_mask(key, image_shape, rate, random=True):
    """images.py:254-279 -> (unobs_inds_ravelled, obs_inds_ravelled)."""
    img_w, img_h = image_shape[:2]
    nblocks = int(img_w * img_h / rate ** 2)
    if random:
        shifts = jr.randint(key, (nblocks, 2), 0, rate)
    else:
        shifts = np.ones((nblocks, 2), dtype=int) * (rate // 2)
    inds_w, inds_h = [i for i in range(0, img_w, rate)], [i for i in range(0, img_h, rate)]
    block_inds = np.asarray(list(itertools.product(inds_w, inds_h)))
    block = np.ravel_multi_index([block_inds[:, 0] + shifts[:, 0], block_inds[:, 1] + shifts[:, 1]], (img_w, img_h),
                                 mode='clip')
    unobs = np.setdiff1d(np.arange(img_w * img_h), block, assume_unique=True)
    return unobs.astype(np.int32), block.astype(np.int32)


def gen_inpaint_mask(key, image_shape, width, height):
    """images.py:281-300 -> (shift, unobs_inds_ravelled, obs_inds_ravelled)."""
    img_w, img_h = image_shape[:2]
    width, height = min(width, img_w), min(height, img_h)
    rect_inds = np.asarray(list(itertools.product(range(width), range(height))))
    max_shift = min(img_w, img_h) - max(width, height)
    shift = int(jr.randint(key, (), 0, max_shift))
    rect = np.ravel_multi_index([rect_inds[:, 0] + shift, rect_inds[:, 1] + shift], (img_w, img_h), mode='clip')
    obs = np.setdiff1d(np.arange(img_w * img_h), rect, assume_unique=True)
    return shift, rect.astype(np.int32), obs.astype(np.int32)


def unpack(xy, image_shape, unobs, obs):
    img_w, img_h, img_c = image_shape
    flat = np.reshape(xy, (*xy.shape[:-3], img_w * img_h, img_c))
    return flat[..., unobs, :], flat[..., obs, :]


def concat(x, y, image_shape, unobs, obs):
    img_w, img_h, img_c = image_shape
    img = np.zeros((*x.shape[:-2], img_w * img_h, img_c), dtype=x.dtype)
    img[..., unobs, :] = x
    img[..., obs, :] = y
    return img.reshape(*img.shape[:-2], img_w, img_h, img_c)
