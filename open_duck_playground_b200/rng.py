"""Host-side restatement of the jax.random key plumbing the runner needs (Threefry-2x32).

``jax.random.PRNGKey(seed)`` / ``jax.random.split`` produce the per-env keys the reference hands to
``env.reset`` and ``randomization_fn`` (Brax ``ppo.train``; common/runner.py:104-118).  JAX is not
installed here, so the same arithmetic is restated in numpy (``jax_threefry_partitionable=True``,
the JAX >= 0.5 default the reference pins at pyproject.toml:8).  The per-step draws inside the env
run on the GPU (csrc/oduck_env.cuh); this module only makes keys.
"""
from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(k0, k1, x0, x1):
    """Vectorised Threefry-2x32, 20 rounds.  All arguments broadcastable uint32 arrays."""
    k0, k1, x0, x1 = (np.asarray(a, dtype=np.uint32) for a in (k0, k1, x0, x1))
    ks = (k0, k1, k0 ^ k1 ^ np.uint32(0x1BD11BDA))
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for blk in range(5):
            for r in _ROT[blk & 1]:
                x0 = x0 + x1
                x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(blk + 1) % 3]
            x1 = x1 + ks[(blk + 2) % 3] + np.uint32(blk + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split(key, num) -> uint32 [num, 2] (also accepts a batch of keys [..., 2] -> [..., num, 2])."""
    key = np.asarray(key, dtype=np.uint32)
    idx = np.arange(num, dtype=np.uint32)
    a, b = threefry2x32(key[..., 0:1], key[..., 1:2], np.uint32(0), idx)
    return np.stack([a, b], axis=-1)


def bits(key, n: int) -> np.ndarray:
    key = np.asarray(key, dtype=np.uint32)
    a, b = threefry2x32(key[..., 0:1], key[..., 1:2], np.uint32(0), np.arange(n, dtype=np.uint32))
    return a ^ b


def uniform(key, n: int, minval=0.0, maxval=1.0) -> np.ndarray:
    u = (((bits(key, n) >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0))
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, u * (hi - lo) + lo)
