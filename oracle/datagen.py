"""numpy twin of cudf_b200/csrc/datagen.cu: x_i = splitmix64(seed + first + i)."""
from __future__ import annotations

import numpy as np

SEED_KEYS, SEED_RIGHT, SEED_PAYLOAD, SEED_VALID, SEED_SELECT = 0x5EED0001, 0x5EED0002, 0x5EED0003, 0x5EED0004, 0x5EED0005


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _counter(n: int, seed: int, first: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        return np.arange(n, dtype=np.uint64) + np.uint64((seed + first) & 0xFFFFFFFFFFFFFFFF)


def fill(n: int, seed: int, first: int = 0, kind: int = 0, modulus: int = 0) -> np.ndarray:
    """kind 0: int64 bits; 1: float64 in [0,1); 2: x % modulus as int64; 3: low 32 bits as int32; 4: validity bits (bool[n])."""
    x = splitmix64(_counter(n, seed, first))
    if kind == 0:
        return x.view(np.int64)
    if kind == 1:
        return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    if kind == 2:
        return (x % np.uint64(modulus)).astype(np.int64)
    if kind == 3:
        return (x & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32)
    if kind == 4:
        return (x >> np.uint64(63)).astype(bool)
    raise ValueError(kind)
