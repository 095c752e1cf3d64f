"""CPU restatement of cudf::hash_partition's row -> partition mapping — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows cpp/src/partitioning/partitioning.cu:875-945 (hash_partition: row hash, then `hash % num_partitions`, or the
equivalent bit mask for powers of two :54-93), cpp/include/cudf/detail/row_operator/hashing.cuh:40-140 (element hash,
null -> UINT32_MAX, columns combined left to right with hash_combine, first column's hash is the initial value),
cpp/include/cudf/hashing/detail/hashing.hpp:83-86 (hash_combine), murmurhash3_x86_32.cuh:21-67 (floats are normalised:
-0 -> +0, NaN -> the canonical quiet NaN; bool hashes as one byte) and hash_functions.cuh:15-37.

The arithmetic itself is cuco::murmurhash3_32 = the public MurmurHash3_x86_32 (Austin Appleby, SMHasher): cuCollections is
not in /root/reference (fetched by cpp/cmake/thirdparty/get_cucollections.cmake), so this restates the published
algorithm; tests/test_oracle_golden.py pins it on SMHasher's known answers and on scikit-learn's C implementation
(sklearn.utils.murmurhash3_32). The reference's own tests hold no known-answer vectors for fixed-width columns
(cpp/tests/hashing/murmurhash3_x86_32_test.cpp:401-465 checks equalities only), so the per-row VALUES are pinned on the
published algorithm, not on a libcudf output: "parity unpinned" against libcudf itself for this function."""
from __future__ import annotations

import numpy as np

U32 = np.uint32
C1, C2 = U32(0xCC9E2D51), U32(0x1B873593)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def _fmix(h):
    h = h ^ (h >> U32(16))
    h = h * U32(0x85EBCA6B)
    h = h ^ (h >> U32(13))
    h = h * U32(0xC2B2AE35)
    return h ^ (h >> U32(16))


def murmur3_32_fixed(values: np.ndarray, seed: int = 0) -> np.ndarray:
    """MurmurHash3_x86_32 of each element's little-endian bytes (element size 1, 2, 4 or 8)."""
    v = np.ascontiguousarray(values)
    size = v.dtype.itemsize
    raw = v.view(np.uint8).reshape(len(v), size)
    with np.errstate(over="ignore"):
        h = np.full(len(v), seed, dtype=U32)
        for b in range(size // 4):
            k = raw[:, 4 * b:4 * b + 4].copy().view(U32).reshape(-1)
            k = _rotl(k * C1, 15) * C2
            h = _rotl(h ^ k, 13) * U32(5) + U32(0xE6546B64)
        tail = size & 3
        if tail:
            k = np.zeros(len(v), dtype=U32)
            for j in range(tail):
                k |= raw[:, (size // 4) * 4 + j].astype(U32) << U32(8 * j)
            h = h ^ (_rotl(k * C1, 15) * C2)
        return _fmix(h ^ U32(size))


def _normalise(values: np.ndarray) -> np.ndarray:
    v = np.asarray(values)
    if v.dtype.kind == "f":
        v = v.copy()
        v[v == 0] = 0.0                      # -0.0 -> +0.0
        v[np.isnan(v)] = np.nan              # canonical quiet NaN (0x7fc00000 / 0x7ff8000000000000)
        return v
    if v.dtype == np.bool_:
        return v.astype(np.uint8)
    return v


def row_hash(cols, seed: int = 0, identity: bool = False) -> np.ndarray:
    """cols: list of (values, valid-or-None). uint32 row hashes."""
    out = None
    for values, valid in cols:
        v = _normalise(values)
        if identity:
            with np.errstate(over="ignore", invalid="ignore"):
                h = v.astype(np.int64).astype(U32) if v.dtype.kind != "f" else v.astype(np.int64).astype(U32)
        else:
            h = murmur3_32_fixed(v, seed)
        if valid is not None:
            h = np.where(np.asarray(valid, bool), h, U32(0xFFFFFFFF))
        if out is None:
            out = h
        else:
            with np.errstate(over="ignore"):
                out = out ^ (h + U32(0x9E3779B9) + (out << U32(6)) + (out >> U32(2)))
    return out


def hash_partition_ids(cols, num_partitions: int, seed: int = 0, identity: bool = False) -> np.ndarray:
    return (row_hash(cols, seed, identity) % U32(num_partitions)).astype(np.int64)


def hash_partition(table_cols, key_cols, num_partitions: int, seed: int = 0, identity: bool = False):
    """-> (list of partitioned (values, valid) columns, offsets[num_partitions + 1]); rows keep their input order inside a
    partition (the reference leaves that order unspecified: compare per-partition multisets)."""
    n = len(table_cols[0][0]) if table_cols else 0
    if num_partitions <= 0 or n == 0 or not key_cols:
        return [(v[:0], None if m is None else m[:0]) for v, m in table_cols], [0] * (max(num_partitions, 0) + 1)
    ids = hash_partition_ids(key_cols, num_partitions, seed, identity)
    order = np.argsort(ids, kind="stable")
    counts = np.bincount(ids, minlength=num_partitions)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(int).tolist()
    return [(v[order], None if m is None else np.asarray(m)[order]) for v, m in table_cols], offs


def mix64(k: np.ndarray) -> np.ndarray:
    """murmur3's 64-bit finalizer (fmix64, public domain, Appleby): the framework's own table / shuffle hash — not a libcudf
    function, so the oracle restates cudf_b200/csrc/device_utils.cuh::mix64 and is pinned on fmix64's fixed point 0 -> 0 and
    its bijectivity in tests/test_oracle_golden.py."""
    k = np.asarray(k, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xFF51AFD7ED558CCD)
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xC4CEB9FE1A85EC53)
        k ^= k >> np.uint64(33)
    return k


def shuffle_bucket(keys: np.ndarray, num_partitions: int) -> np.ndarray:
    """Destination rank of a row in the sharded join's fused shuffle (b2_range_partition_* with splitters == NULL): the keys in
    radix order (signed: sign bit flipped), plus the golden-ratio constant, through mix64; bucket = high 32 bits * P >> 32.
    The reference shuffles with its row hash modulo P (cpp/src/partitioning/partitioning.cu:hash_partition); any function both
    sides agree on gives the same join result, so this one is framework-defined."""
    k = np.asarray(keys)
    u = k.view(np.uint64).copy() if k.dtype.kind in "iu" and k.dtype.itemsize == 8 else k.astype(np.uint64)
    if k.dtype.kind == "i":
        u ^= np.uint64(1 << 63)
    with np.errstate(over="ignore"):
        h = mix64(u + np.uint64(0x9E3779B97F4A7C15))
    return (((h >> np.uint64(32)) * np.uint64(num_partitions)) >> np.uint64(32)).astype(np.int64)
