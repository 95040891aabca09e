"""Generate tests/golden/oracle_golden.npz: small input / output vectors of the hot path produced by the ORACLE
(oracle/, the NumPy restatement of the reference).

These are regression pins, NOT reference-generated vectors: the reference (zgbkdlm/fbs) needs jax / jaxlib / flax, which
cannot be imported in this image, so the only vectors that come from outside this repository are the published known
answers checked in tests/test_oracle_random.py (Random123 threefry KATs, the JAX documentation's PRNGKey(0) values).  The
fixtures freeze the oracle's behaviour at the commit that produced them, so that (a) an accidental change to the oracle
is caught on CPU (tests/test_golden.py::test_oracle_reproduces_golden) and (b) the CUDA kernels are checked against files
on disk rather than against code that might drift with them (tests/test_golden.py, -m gpu).

usage: python tests/golden/make_golden.py      (rewrites oracle_golden.npz next to this file)
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import jax_random as jr                      # noqa: E402
from oracle import cond_resampling as ocr                # noqa: E402
from oracle import resampling as orx                     # noqa: E402
from oracle import images as oi                          # noqa: E402


def weights(seed, n):
    rng = np.random.default_rng(seed)
    w = rng.random(n).astype(np.float32) ** 3
    return (w / w.sum(dtype=np.float32)).astype(np.float32)


def build():
    out = {}
    key = jr.PRNGKey(2024)
    out['key'] = key
    out['split_5'] = jr.split(key, 5)
    out['bits_101'] = jr.random_bits(key, 101)
    out['uniform_64'] = jr.uniform(key, (64,))
    out['normal_7x9'] = jr.normal(key, (7, 9))
    out['randint_50'] = jr.randint(key, (50,), 3, 17)
    for n in (10, 100, 257):
        w = weights(n, n)
        k = jr.PRNGKey(1000 + n)
        out[f'w_{n}'] = w
        out[f'cond_killing_{n}'] = ocr.killing(k, w, 3 % n, 7 % n, True)
        out[f'cond_multinomial_{n}'] = ocr.multinomial(k, w, 3 % n, 7 % n, True)
        out[f'cond_systematic_{n}'] = ocr.systematic(k, w, conditional=False)
        out[f'stratified_{n}'] = orx.stratified(w, k)
        out[f'systematic_{n}'] = orx.systematic(w, k)
        out[f'killing_{n}'] = orx.killing(w, k)
    shift, rect, obs = oi.gen_inpaint_mask(jr.PRNGKey(5), (28, 28, 1), 15, 15)
    out['inpaint_shift'] = np.int32(shift)
    out['inpaint_rect'] = rect
    unobs, sr_obs = oi.gen_supr_mask(jr.PRNGKey(6), (28, 28, 1), 4, True)
    out['supr_obs'] = sr_obs
    return out


if __name__ == '__main__':
    np.savez_compressed(os.path.join(HERE, 'oracle_golden.npz'), **build())
    print('wrote', os.path.join(HERE, 'oracle_golden.npz'))
