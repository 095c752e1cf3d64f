"""pylibcudf.partitioning twin (python/pylibcudf/pylibcudf/partitioning.pyx; cpp/include/cudf/partitioning.hpp:58-145)
over b2_partition (cudf_b200/csrc/partition.cu: stable P-way partition, P <= 256).

`partition` follows cudf::partition exactly (rows go to the partition their map entry names; here they also keep their
input order inside a partition, which the reference leaves unspecified). `hash_partition` has the reference's contract
(equal keys land in the same partition, offsets returned) but uses this library's 64-bit mixer on ONE fixed-width key
column, not cudf's murmur3 row hash: partitions are consistent within this library, not with dask_cudf workers running
libcudf."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import check, lib
from .column import Column, Table


def _partition(table: Table, key: Column, mode: int, splitters: Column | None, num_partitions: int, stream):
    out = C.c_void_p()
    offs = (C.c_int32 * (num_partitions + 1))()
    tv, kv = table._view(), key._view()
    sp = C.c_void_p(splitters._data) if splitters is not None and splitters.size() else None
    check(lib.b2_partition(C.byref(tv), C.byref(kv), mode, sp, int(num_partitions), _lib.stream_arg(stream), C.byref(out), offs))
    return Table._from_handle(out.value), list(offs)[:num_partitions]


def partition(t: Table, partition_map: Column, num_partitions: int, stream=None, mr=None):
    """cudf::partition (partitioning.hpp:58-101): -> (partitioned table, offsets of the partitions, length num_partitions)."""
    if partition_map.has_nulls():
        raise ValueError("partition_map contains nulls")
    if num_partitions < 1:
        raise ValueError("num_partitions must be positive")
    if partition_map.size() != (t.num_rows() if t.num_columns() else 0):
        raise RuntimeError("partition_map and the table differ in size")
    # bucket(row) = number of splitters <= map[row] with splitters 1 .. P-1 is the map entry itself
    dt = partition_map.type().numpy_dtype()
    splitters = Column.from_numpy(np.arange(1, num_partitions, dtype=dt)) if num_partitions > 1 else None
    return _partition(t, partition_map, 0, splitters, num_partitions, stream)


def hash_partition(input: Table, columns_to_hash: list, num_partitions: int, stream=None, mr=None):  # noqa: A002
    """cudf::hash_partition (partitioning.hpp:103-145) on one fixed-width key column (see the module docstring)."""
    if len(columns_to_hash) != 1:
        raise ValueError("hash_partition: exactly one key column is supported on this path")
    key = input.columns()[int(columns_to_hash[0])]
    return _partition(input, key, 1, None, num_partitions, stream)
