"""On-disk result layouts of the reference drivers, so that its tabulators can score this package's output.

* image runs (experiments/imgs/inpainting.py:229-251): ``np.save`` of ``restored_imgs [nsamples, w, h, c]`` under
  ``<head>-<method tag>.npy``;
* toy runs (experiments/toy/gp_gibbs.py:193-195): ``np.savez`` with ``samples [nchains, nsamples, d]``, ``gp_mean``,
  ``gp_cov``.
"""
import numpy as np
import torch


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def method_tag(method: str, explicit_backward: bool = False, explicit_final: bool = False, marg: bool = False) -> str:
    """File-name tag the reference builds inline: ``gibbs[-eb][-ef][-marg]`` / ``filter[-marg]`` / ``pmcmc-<delta>``."""
    if method == 'gibbs':
        return 'gibbs' + ('-eb' if explicit_backward else '') + ('-ef' if explicit_final else '') + ('-marg' if marg else '')
    if method == 'filter':
        return 'filter' + ('-marg' if marg else '')
    return method


def save_restored_images(path_head: str, restored_imgs, method: str = 'gibbs', **tag_kwargs) -> str:
    """``np.save(path_head + '-' + tag, restored_imgs)`` with ``restored_imgs [nsamples, w, h, c]`` (inpainting.py:249-251)."""
    imgs = _np(restored_imgs)
    if imgs.ndim != 4:
        raise ValueError(f'restored_imgs must be [nsamples, w, h, c], got {imgs.shape}')
    path = f'{path_head}-{method_tag(method, **tag_kwargs)}'
    np.save(path, imgs)
    return path + '.npy'


def save_chain_samples(path: str, samples, gp_mean, gp_cov) -> str:
    """``np.savez(path, samples=[nchains, nsamples, d], gp_mean=, gp_cov=)`` (gp_gibbs.py:193-195)."""
    s = _np(samples)
    if s.ndim != 3:
        raise ValueError(f'samples must be [nchains, nsamples, d], got {s.shape}')
    np.savez(path, samples=s, gp_mean=_np(gp_mean), gp_cov=_np(gp_cov))
    return path if path.endswith('.npz') else path + '.npz'
