"""Image-restoration masks and the (unobserved, observed) split of an image: the data format on either side of the
score-network CSMC path (reference: fbs/data/images.py:212-363 -- ``InpaintingMask`` / ``SRMask`` / ``ImageRestore``).

Only what the samplers consume is here -- mask generation from a PRNG key (same ``jax.random.randint`` stream, drawn by
the CUDA kernel), ``unpack`` and ``concat`` -- not the dataset readers (MNIST / CelebA files are not part of the path).
Index conventions follow the reference: an image is ``(w, h, c)``, pixels are ravelled row-major over ``(w, h)``, the
inpainting rectangle is the ``width x height`` block shifted by the same random offset along both axes, and the
super-resolution mask observes one pixel of every ``rate x rate`` block.
"""
from typing import NamedTuple, Tuple
import numpy as np
import torch
from .. import random as fr


class InpaintingMask(NamedTuple):          # images.py:212-219
    width: int
    height: int
    shift: int
    unobs_inds_ravelled: np.ndarray
    obs_inds_ravelled: np.ndarray


class SRMask(NamedTuple):                  # images.py:222-225
    rate: int
    unobs_inds_ravelled: np.ndarray
    obs_inds_ravelled: np.ndarray


def _host_key(key):
    return key.detach().cpu().numpy() if isinstance(key, torch.Tensor) else np.asarray(key, dtype=np.uint32)


class ImageRestore:
    """``ImageRestore(task, image_shape, sr_random)`` with ``task`` = ``'inpaint-<s>'`` or ``'supr-<rate>'``."""

    def __init__(self, task: str, image_shape: Tuple[int, int, int], sr_random: bool = True):
        self.image_shape = tuple(int(s) for s in image_shape)
        self.task = task
        w, h, c = self.image_shape
        s = int(task.split('-')[-1])
        if 'inpaint' in task:
            self.unobs_shape = (s ** 2, c)                                  # images.py:237-238
        elif 'supr' in task:
            self.unobs_shape = (int(w * h * (s ** 2 - 1) / s ** 2), c)      # images.py:239-240
        else:
            raise ValueError(f'Unknown task {task}.')
        self.sr_random = sr_random

    # ------------------------------------------------------------------ masks
    def _gen_supr_mask(self, key, rate: int, random: bool = True) -> SRMask:
        """images.py:254-279: one observed pixel per ``rate x rate`` block, at a random position or in the middle."""
        img_w, img_h = self.image_shape[:2]
        nblocks = int(img_w * img_h / rate ** 2)
        if random:
            shifts = np.asarray(fr.randint(_host_key(key), (nblocks, 2), 0, rate)).astype(np.int64)
        else:
            shifts = np.full((nblocks, 2), rate // 2, dtype=np.int64)
        bw, bh = np.meshgrid(np.arange(0, img_w, rate), np.arange(0, img_h, rate), indexing='ij')   # itertools.product order
        rows = np.clip(bw.ravel() + shifts[:, 0], 0, img_w - 1)             # ravel_multi_index(mode='clip')
        cols = np.clip(bh.ravel() + shifts[:, 1], 0, img_h - 1)
        block = (rows * img_h + cols).astype(np.int32)
        unobs = np.setdiff1d(np.arange(img_w * img_h, dtype=np.int32), block, assume_unique=True)
        return SRMask(rate, unobs_inds_ravelled=unobs, obs_inds_ravelled=block)

    def _gen_inpaint_mask(self, key, width: int, height: int) -> InpaintingMask:
        """images.py:281-300: the ``width x height`` rectangle shifted by ``randint(key, (), 0, max_shift)`` along both axes."""
        img_w, img_h = self.image_shape[:2]
        width, height = min(width, img_w), min(height, img_h)
        max_shift = min(img_w, img_h) - max(width, height)
        shift = int(np.asarray(fr.randint(_host_key(key), (), 0, max_shift)).reshape(-1)[0])
        rw, rh = np.meshgrid(np.arange(width), np.arange(height), indexing='ij')
        rows = np.clip(rw.ravel() + shift, 0, img_w - 1)
        cols = np.clip(rh.ravel() + shift, 0, img_h - 1)
        rect = (rows * img_h + cols).astype(np.int32)
        obs = np.setdiff1d(np.arange(img_w * img_h, dtype=np.int32), rect, assume_unique=True)
        return InpaintingMask(width, height, shift, unobs_inds_ravelled=rect, obs_inds_ravelled=obs)

    def gen_mask(self, key):
        s = int(self.task.split('-')[-1])
        if 'inpaint' in self.task:
            return self._gen_inpaint_mask(key, s, s)                        # images.py:303-305
        return self._gen_supr_mask(key, s, random=self.sr_random)           # images.py:306-308

    # ------------------------------------------------------------------ unpack / concat
    def unpack(self, xy, mask):
        """``(..., w, h, c) -> ((..., p, c), (..., q, c))``: unobserved and observed pixels (images.py:330-350)."""
        img_w, img_h, img_c = self.image_shape
        if isinstance(xy, torch.Tensor):
            flat = xy.reshape(*xy.shape[:-3], img_w * img_h, img_c)
            iu = torch.as_tensor(np.asarray(mask.unobs_inds_ravelled), dtype=torch.long, device=xy.device)
            io = torch.as_tensor(np.asarray(mask.obs_inds_ravelled), dtype=torch.long, device=xy.device)
            return flat.index_select(-2, iu), flat.index_select(-2, io)
        xy = np.asarray(xy)
        flat = xy.reshape(*xy.shape[:-3], img_w * img_h, img_c)
        return flat[..., np.asarray(mask.unobs_inds_ravelled), :], flat[..., np.asarray(mask.obs_inds_ravelled), :]

    def concat(self, x, y, mask):
        """The reverse of ``unpack`` (images.py:352-361)."""
        img_w, img_h, img_c = self.image_shape
        if isinstance(x, torch.Tensor):
            img = torch.zeros((*x.shape[:-2], img_w * img_h, img_c), dtype=x.dtype, device=x.device)
            iu = torch.as_tensor(np.asarray(mask.unobs_inds_ravelled), dtype=torch.long, device=x.device)
            io = torch.as_tensor(np.asarray(mask.obs_inds_ravelled), dtype=torch.long, device=x.device)
            img.index_copy_(-2, iu, x)
            img.index_copy_(-2, io, torch.as_tensor(y, dtype=x.dtype, device=x.device).expand(*x.shape[:-2], -1, -1))
            return img.reshape(*img.shape[:-2], img_w, img_h, img_c)
        x, y = np.asarray(x), np.asarray(y)
        img = np.zeros((*x.shape[:-2], img_w * img_h, img_c), dtype=x.dtype)
        img[..., np.asarray(mask.unobs_inds_ravelled), :] = x
        img[..., np.asarray(mask.obs_inds_ravelled), :] = y
        return img.reshape(*img.shape[:-2], img_w, img_h, img_c)
