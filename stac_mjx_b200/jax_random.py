"""``jax.random.permutation(jax.random.PRNGKey(seed), x, independent=True)`` without JAX (numpy, host side).

The m-phase frame sample of the reference (``stac_mjx/compute_stac.py:136-140``) is
``jax.random.permutation(PRNGKey(0), arange(F), independent=True)[:n_sample]``.  JAX's default PRNG is
Threefry-2x32 (20 rounds, Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11); ``permutation``
of a 1-D array is ``_shuffle``: ``ceil(3 ln n / ln(2^32 - 1))`` rounds of "split the key, draw one uint32 per
element, stable-sort the array by those keys".  Everything is integer arithmetic, hence platform independent and
exactly reproducible here.  JAX itself is not installed in the authoring image, so the implementation is pinned
by the Random123 known-answer vectors for threefry2x32_20 and by the widely published value of
``jax.random.split(PRNGKey(0))`` for the original (non-partitionable) key derivation (tests/test_host_cpu.py).

``partitionable`` selects the counter layout: ``True`` is the ``jax_threefry_partitionable`` default since JAX 0.5.0
(the reference requires ``jax>=0.7.2``, pyproject.toml:36), ``False`` the layout of older releases.
"""

from __future__ import annotations

import math

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key: tuple[int, int], x0, x1) -> tuple[np.ndarray, np.ndarray]:
    """Threefry-2x32, 20 rounds: the block function behind ``jax.random`` (``jax._src.prng._threefry2x32_lowering``)."""
    x0 = np.array(x0, dtype=_U32, copy=True).reshape(-1)
    x1 = np.array(x1, dtype=_U32, copy=True).reshape(-1)
    k0, k1 = _U32(key[0]), _U32(key[1])
    ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x0 ^ x1
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def prng_key(seed: int) -> tuple[int, int]:
    """``jax.random.PRNGKey(seed)`` (``threefry_seed``): the high and low 32 bits of the seed."""
    seed = int(seed)
    return ((seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF)


def _threefry_2x32_flat(key, count: np.ndarray) -> np.ndarray:
    """``prng.threefry_2x32``: a flat counter array is cut in two halves that form the two input words."""
    count = np.asarray(count, dtype=_U32).reshape(-1)
    odd = count.size % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, _U32)])
    h = count.size // 2
    a, b = threefry2x32(key, count[:h], count[h:])
    out = np.concatenate([a, b])
    return out[:-1] if odd else out


def split(key, num: int = 2, partitionable: bool = True) -> list[tuple[int, int]]:
    """``jax.random.split(key, num)``."""
    if partitionable:  # _threefry_split_foldlike: 64-bit iota as (hi, lo) counter words
        idx = np.arange(num, dtype=np.uint64)
        b1, b2 = threefry2x32(key, (idx >> np.uint64(32)).astype(_U32), (idx & np.uint64(0xFFFFFFFF)).astype(_U32))
        return [(int(b1[i]), int(b2[i])) for i in range(num)]
    out = _threefry_2x32_flat(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)  # _threefry_split_original
    return [(int(out[i, 0]), int(out[i, 1])) for i in range(num)]


def random_bits32(key, n: int, partitionable: bool = True) -> np.ndarray:
    """``prng.threefry_random_bits(key, 32, (n,))``."""
    if partitionable:
        idx = np.arange(n, dtype=np.uint64)
        b1, b2 = threefry2x32(key, (idx >> np.uint64(32)).astype(_U32), (idx & np.uint64(0xFFFFFFFF)).astype(_U32))
        return b1 ^ b2
    return _threefry_2x32_flat(key, np.arange(n, dtype=_U32))


def shuffle(key, x: np.ndarray, partitionable: bool = True) -> np.ndarray:
    """``jax._src.random._shuffle`` for a 1-D array: repeated stable sorts by fresh 32-bit keys."""
    x = np.asarray(x)
    n = x.size
    rounds = int(math.ceil(3 * math.log(max(1, n)) / math.log(0xFFFFFFFF)))
    for _ in range(rounds):
        key, sub = split(key, 2, partitionable)
        sort_keys = random_bits32(sub, n, partitionable)
        x = x[np.argsort(sort_keys, kind="stable")]
    return x


def permutation(seed: int, n: int, partitionable: bool = True) -> np.ndarray:
    """``jax.random.permutation(jax.random.PRNGKey(seed), jnp.arange(n), independent=True)``."""
    return shuffle(prng_key(seed), np.arange(n, dtype=np.int64), partitionable)
