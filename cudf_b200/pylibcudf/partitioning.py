"""pylibcudf.partitioning twin (python/pylibcudf/pylibcudf/partitioning.pyx; cpp/include/cudf/partitioning.hpp:58-145)
over b2_partition / b2_hash_partition (cudf_b200/csrc/partition.cu: stable P-way partition, P <= 256).

`partition` follows cudf::partition (rows go to the partition their map entry names; here they also keep their input order
inside a partition, which the reference leaves unspecified). `hash_partition` uses libcudf's row hash — MurmurHash3_x86_32
per key column with the seed, nulls = UINT32_MAX, columns folded with hash_combine, partition = hash % num_partitions
(cpp/src/partitioning/partitioning.cu:54-93,875-945) — so a row lands in the same partition as under libcudf.
Both return `num_partitions + 1` offsets (cpp/tests/partitioning/partition_test.cpp:137, hash_partition_test.cpp:88)."""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from .. import _lib
from .._lib import check, lib
from .column import Column, Table


class HashId(enum.IntEnum):  # cudf::hash_id (partitioning.hpp:32-35)
    HASH_IDENTITY = 0
    HASH_MURMUR3 = 1


DEFAULT_HASH_SEED = 0


def _partition(table: Table, key: Column, mode: int, splitters: Column | None, num_partitions: int, stream):
    out = C.c_void_p()
    offs = (C.c_int32 * (num_partitions + 1))()
    tv, kv = table._view(), key._view()
    sp = C.c_void_p(splitters._data) if splitters is not None and splitters.size() else None
    check(lib.b2_partition(C.byref(tv), C.byref(kv), mode, sp, int(num_partitions), _lib.stream_arg(stream), C.byref(out), offs))
    return Table._from_handle(out.value), list(offs)


def partition(t: Table, partition_map: Column, num_partitions: int, stream=None, mr=None):
    """cudf::partition (partitioning.hpp:58-101): -> (partitioned table, num_partitions + 1 offsets)."""
    if partition_map.has_nulls():
        raise RuntimeError("Unexpected null values in partition_map.")  # cudf::logic_error
    if num_partitions < 0:
        raise ValueError("num_partitions must not be negative")
    if partition_map.size() != (t.num_rows() if t.num_columns() else 0):
        raise RuntimeError("Size mismatch between table and partition map.")
    if num_partitions == 0 or partition_map.size() == 0:
        return _empty_like(t), [0] * (num_partitions + 1)
    if num_partitions > 256 or partition_map.type().numpy_dtype().itemsize == 1 and num_partitions > 127:
        out = C.c_void_p()
        offs = (C.c_int32 * (num_partitions + 1))()
        tv, mv = t._view(), partition_map._view()
        check(lib.b2_partition_by_map(C.byref(tv), C.byref(mv), int(num_partitions), _lib.stream_arg(stream), C.byref(out), offs))
        return Table._from_handle(out.value), list(offs)
    # bucket(row) = number of splitters <= map[row] with splitters 1 .. P-1 is the map entry itself
    dt = partition_map.type().numpy_dtype()
    splitters = Column.from_numpy(np.arange(1, num_partitions, dtype=dt)) if num_partitions > 1 else None
    return _partition(t, partition_map, 0, splitters, num_partitions, stream)


def _empty_like(t: Table) -> Table:
    return Table([Column.from_numpy(np.empty(0, dtype=c.type().numpy_dtype())) for c in t.columns()])


def hash_partition(input: Table, keys, num_partitions: int, hash_function: HashId = HashId.HASH_MURMUR3,  # noqa: A002
                   seed: int = DEFAULT_HASH_SEED, stream=None, mr=None):
    """cudf::hash_partition (partitioning.hpp:103-175): `keys` is a Table of key columns or a list of column indices of
    `input`. -> (partitioned table, num_partitions + 1 offsets)."""
    if isinstance(keys, Table):
        ktab = keys
    else:
        cols = input.columns()
        for i in keys:
            if not 0 <= int(i) < len(cols):
                raise IndexError("columns_to_hash: invalid column index")  # std::out_of_range
        ktab = Table([cols[int(i)] for i in keys])
    if ktab.num_columns() and ktab.num_rows() != input.num_rows():
        raise ValueError("Input table and key table must have same number of rows, or key table should have no columns.")
    out = C.c_void_p()
    offs = (C.c_int32 * (max(num_partitions, 0) + 1))()
    tv, kv = input._view(), ktab._view()
    check(lib.b2_hash_partition(C.byref(tv), C.byref(kv), int(num_partitions), int(hash_function), int(seed) & 0xFFFFFFFF,
                                _lib.stream_arg(stream), C.byref(out), offs))
    return Table._from_handle(out.value), list(offs)
