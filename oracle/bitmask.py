"""Arrow validity bitmaps (cpp/include/cudf/utilities/bit.hpp; cpp/src/bitmask/null_mask.cu)."""
from __future__ import annotations

import numpy as np


def allocation_size_bytes(bits: int) -> int:
    """null_mask.hpp:55 — words of 32 bits, padded to 64 B."""
    return ((bits + 31) // 32 * 4 + 63) // 64 * 64


def pack(valid: np.ndarray) -> np.ndarray:
    """bool[n] -> uint32 words, LSB first (bit.hpp word_index / intra_word_index)."""
    b = np.packbits(np.asarray(valid, dtype=bool), bitorder="little")
    out = np.zeros((len(valid) + 31) // 32 * 4, dtype=np.uint8)
    out[: len(b)] = b
    return out.view(np.uint32)


def unpack(words: np.ndarray, n: int, offset: int = 0) -> np.ndarray:
    bits = np.unpackbits(np.asarray(words).view(np.uint8), bitorder="little")
    return bits[offset: offset + n].astype(bool)


def count_set_bits(words: np.ndarray, start: int, stop: int) -> int:
    return int(unpack(words, stop - start, start).sum())


def bitmask_and(valids: list, n: int):
    """null_mask.cu:608-735 — AND of the nullable columns' masks -> (valid | None, null_count)."""
    vs = [v for v in valids if v is not None]
    if not vs:
        return None, 0
    out = np.ones(n, dtype=bool)
    for v in vs:
        out &= v
    return out, int((~out).sum())


def set_null_mask(valid: np.ndarray, begin: int, end: int, value: bool) -> np.ndarray:
    out = valid.copy()
    out[begin:end] = value
    return out
