"""Multi-GPU partitioning: independent chains / conditioning targets are block-partitioned over ranks
with NO data-path collective (the reference's only parallelism is ``vmap`` over chains plus OS-level
fan-out, experiments/toy/gp_gibbs.py:172-173, experiments/bashes/toy_gibbs.sh:22-30).

Per-chain keys are slices of ONE global ``split(key, total_chains)``, so a chain's stream -- and hence
its result -- does not depend on how many GPUs the job runs on.
"""
import numpy as np


def chain_slice(total: int, rank: int, world: int):
    """Contiguous block ``[lo, hi)`` of chains owned by ``rank`` (first ``total % world`` ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError('rank out of range')
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def threefry_split_host(key, num: int) -> np.ndarray:
    """``jax.random.split`` on the host (numpy) -- used only to derive per-rank key slices before any GPU work."""
    key = np.asarray(key, dtype=np.uint32)
    n = 2 * int(num)
    h = n // 2
    x0 = np.arange(h, dtype=np.uint32)
    x1 = np.arange(h, n, dtype=np.uint32)
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    with np.errstate(over='ignore'):
        ks = (key[0], key[1], key[0] ^ key[1] ^ np.uint32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in rot[g % 2]:
                x0 = x0 + x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + np.uint32(g + 1)
    return np.concatenate([x0, x1]).reshape(int(num), 2)


def chain_keys(key, total: int, rank: int, world: int) -> np.ndarray:
    """Keys of the chains owned by ``rank``: ``split(key, total)[lo:hi]``."""
    lo, hi = chain_slice(total, rank, world)
    return np.ascontiguousarray(threefry_split_host(key, total)[lo:hi])
