"""pylibcudf.reduce twin (python/pylibcudf/pylibcudf/reduce.pyx:48-157) + segmented_reduce."""
from __future__ import annotations

import ctypes as C
import enum

from .. import _lib
from .._lib import check, lib
from .aggregation import Aggregation
from .column import Column, Scalar
from .types import DataType, NullPolicy


class ScanType(enum.IntEnum):
    INCLUSIVE = 0
    EXCLUSIVE = 1


def reduce(col: Column, agg: Aggregation, data_type: DataType, init: Scalar | None = None, stream=None, mr=None) -> Scalar:
    out = C.c_void_p()
    cv = col._view()
    check(lib.b2_reduce(C.byref(cv), int(agg.kind()), int(data_type.id()), C.c_void_p(init._handle) if init else None,
                        _lib.stream_arg(stream), C.byref(out)))
    return Scalar(out.value)


def scan(col: Column, agg: Aggregation, inclusive: ScanType, null_handling: NullPolicy = NullPolicy.EXCLUDE, stream=None,
         mr=None) -> Column:
    out = C.c_void_p()
    cv = col._view()
    check(lib.b2_scan(C.byref(cv), int(agg.kind()), int(inclusive), int(null_handling), _lib.stream_arg(stream), C.byref(out)))
    return Column._from_handle(out.value)


def segmented_reduce(segmented_values: Column, offsets: Column, agg: Aggregation, data_type: DataType,
                     null_handling: NullPolicy = NullPolicy.EXCLUDE, init: Scalar | None = None, stream=None, mr=None) -> Column:
    """cudf::segmented_reduce (cpp/include/cudf/reduction.hpp); `offsets` is an INT32 device column."""
    out = C.c_void_p()
    cv = segmented_values._view()
    optr = offsets._data + offsets._offset * 4
    check(lib.b2_segmented_reduce(C.byref(cv), C.c_void_p(optr or None), offsets.size(), int(agg.kind()), int(data_type.id()),
                                  int(null_handling), C.c_void_p(init._handle) if init else None, _lib.stream_arg(stream),
                                  C.byref(out)))
    return Column._from_handle(out.value)
