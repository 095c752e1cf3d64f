"""pylibcudf.contiguous_split twin: pack / unpack / packed_size / pack_metadata of tables of fixed-width columns in
libcudf's wire format (python/pylibcudf/pylibcudf/contiguous_split.pyx; cpp/include/cudf/contiguous_split.hpp:233-317;
layout in cudf_b200/csrc/pack.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import ColumnView, check, lib
from .column import Column, Table
from .types import DataType, TypeId


class _Buffer:
    def __init__(self, handle: int):
        self.handle = handle

    def __del__(self):
        if self.handle:
            lib.b2_buffer_free(C.c_void_p(self.handle))
            self.handle = None


class PackedColumns:
    """cudf::packed_columns: `metadata` (host bytes) + `gpu_data` (one device buffer)."""

    def __init__(self, metadata: bytes, gpu_data_ptr: int, gpu_data_size: int, owner):
        self.metadata = bytes(metadata)
        self._ptr, self._size, self._owner = int(gpu_data_ptr or 0), int(gpu_data_size), owner

    def release(self):
        """-> (metadata as a memoryview, gpu_data as an object exposing __cuda_array_interface__), like pylibcudf."""
        from .column import DeviceSpan

        return memoryview(self.metadata), DeviceSpan(self._ptr, self._size, np.uint8, self._owner)

    @property
    def gpu_data_ptr(self) -> int:
        return self._ptr

    @property
    def gpu_data_size(self) -> int:
        return self._size


def packed_size(input: Table, stream=None) -> int:  # noqa: A002
    out = C.c_size_t(0)
    tv = input._view()
    check(lib.b2_packed_size(C.byref(tv), C.byref(out)))
    return out.value


def pack(input: Table, stream=None, mr=None) -> PackedColumns:  # noqa: A002
    ncols = input.num_columns()
    cap = 16 + 40 * ncols
    md = (C.c_uint8 * cap)()
    mdsz = C.c_size_t(0)
    buf = C.c_void_p()
    tv = input._view()
    check(lib.b2_pack(C.byref(tv), _lib.stream_arg(stream), md, cap, C.byref(mdsz), C.byref(buf)))
    owner = _Buffer(buf.value)
    return PackedColumns(bytes(md[: mdsz.value]), lib.b2_buffer_data(buf), lib.b2_buffer_size(buf), owner)


def pack_metadata(table: Table, contiguous_buffer_ptr: int, buffer_size: int) -> bytes:
    cap = 16 + 40 * table.num_columns()
    md = (C.c_uint8 * cap)()
    mdsz = C.c_size_t(0)
    tv = table._view()
    check(lib.b2_pack_metadata(C.byref(tv), C.c_void_p(contiguous_buffer_ptr), buffer_size, md, cap, C.byref(mdsz)))
    return bytes(md[: mdsz.value])


def unpack_from_memoryviews(metadata, gpu_data_ptr: int, owner=None) -> Table:
    """cudf::unpack(metadata, gpu_data): the columns of the result point into gpu_data (kept alive through `owner`)."""
    md = bytes(metadata)
    ncap = max(0, (len(md) - 16) // 40) + 1
    views = (ColumnView * ncap)()
    ncols, nrows = C.c_int32(0), C.c_int32(0)
    mdbuf = (C.c_uint8 * max(len(md), 1)).from_buffer_copy(md if md else b"\0")
    check(lib.b2_unpack(mdbuf, len(md), C.c_void_p(gpu_data_ptr or 0), views, ncap, C.byref(ncols), C.byref(nrows)))
    cols = []
    for i in range(ncols.value):
        v = views[i]
        cols.append(Column(DataType(TypeId(v.type_id)), v.size, v.data or 0, v.null_mask or 0, v.null_count, 0, [owner]))
    return Table(cols)


def unpack(input: PackedColumns) -> Table:  # noqa: A002
    return unpack_from_memoryviews(input.metadata, input.gpu_data_ptr, input)
