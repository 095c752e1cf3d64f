"""pylibcudf.sorting twin (python/pylibcudf/pylibcudf/sorting.pyx:37-79,333-520) over the C ABI."""
from __future__ import annotations

import ctypes as C

from .. import _lib
from .._lib import check, lib
from .column import Column, Table


def _u8(seq):
    seq = [int(x) for x in (seq or [])]
    arr = (C.c_uint8 * max(len(seq), 1))(*seq)
    return arr, len(seq)


def _sorted_order(source_table, column_order, null_precedence, stable, stream):
    o, no = _u8(column_order)
    p, np_ = _u8(null_precedence)
    out = C.c_void_p()
    tv = source_table._view()
    check(lib.b2_sorted_order(C.byref(tv), o, no, p, np_, stable, _lib.stream_arg(stream), C.byref(out)))
    return Column._from_handle(out.value)


def sorted_order(source_table: Table, column_order: list, null_precedence: list, stream=None, mr=None) -> Column:
    return _sorted_order(source_table, column_order, null_precedence, 0, stream)


def stable_sorted_order(source_table: Table, column_order: list, null_precedence: list, stream=None, mr=None) -> Column:
    return _sorted_order(source_table, column_order, null_precedence, 1, stream)


def _sort(source_table, column_order, null_precedence, stable, stream):
    o, no = _u8(column_order)
    p, np_ = _u8(null_precedence)
    out = C.c_void_p()
    tv = source_table._view()
    check(lib.b2_sort(C.byref(tv), o, no, p, np_, stable, _lib.stream_arg(stream), C.byref(out)))
    return Table._from_handle(out.value)


def sort(source_table: Table, column_order: list, null_precedence: list, stream=None, mr=None) -> Table:
    return _sort(source_table, column_order, null_precedence, 0, stream)


def stable_sort(source_table: Table, column_order: list, null_precedence: list, stream=None, mr=None) -> Table:
    return _sort(source_table, column_order, null_precedence, 1, stream)


def _sort_by_key(values, keys, column_order, null_precedence, stable, stream):
    o, no = _u8(column_order)
    p, np_ = _u8(null_precedence)
    out = C.c_void_p()
    vv, kv = values._view(), keys._view()
    check(lib.b2_sort_by_key(C.byref(vv), C.byref(kv), o, no, p, np_, stable, _lib.stream_arg(stream), C.byref(out)))
    return Table._from_handle(out.value)


def sort_by_key(values: Table, keys: Table, column_order: list, null_precedence: list, stream=None, mr=None) -> Table:
    return _sort_by_key(values, keys, column_order, null_precedence, 0, stream)


def stable_sort_by_key(values: Table, keys: Table, column_order: list, null_precedence: list, stream=None, mr=None) -> Table:
    return _sort_by_key(values, keys, column_order, null_precedence, 1, stream)


# ---- segmented sort / top-k (python/pylibcudf/pylibcudf/sorting.pyx; cpp/include/cudf/sorting.hpp:232-416) ------------
def _segmented_sorted_order(keys, segment_offsets, column_order, null_precedence, stable, stream):
    o, no = _u8(column_order)
    p, np_ = _u8(null_precedence)
    out = C.c_void_p()
    kv, sv = keys._view(), segment_offsets._view()
    check(lib.b2_segmented_sorted_order(C.byref(kv), C.byref(sv), o, no, p, np_, stable, _lib.stream_arg(stream), C.byref(out)))
    return Column._from_handle(out.value)


def segmented_sorted_order(keys: Table, segment_offsets: Column, column_order: list, null_precedence: list, stream=None, mr=None) -> Column:
    return _segmented_sorted_order(keys, segment_offsets, column_order, null_precedence, 0, stream)


def stable_segmented_sorted_order(keys: Table, segment_offsets: Column, column_order: list, null_precedence: list, stream=None,
                                  mr=None) -> Column:
    return _segmented_sorted_order(keys, segment_offsets, column_order, null_precedence, 1, stream)


def _segmented_sort_by_key(values, keys, segment_offsets, column_order, null_precedence, stable, stream):
    o, no = _u8(column_order)
    p, np_ = _u8(null_precedence)
    out = C.c_void_p()
    vv, kv, sv = values._view(), keys._view(), segment_offsets._view()
    check(lib.b2_segmented_sort_by_key(C.byref(vv), C.byref(kv), C.byref(sv), o, no, p, np_, stable, _lib.stream_arg(stream), C.byref(out)))
    return Table._from_handle(out.value)


def segmented_sort_by_key(values: Table, keys: Table, segment_offsets: Column, column_order: list, null_precedence: list, stream=None,
                          mr=None) -> Table:
    return _segmented_sort_by_key(values, keys, segment_offsets, column_order, null_precedence, 0, stream)


def stable_segmented_sort_by_key(values: Table, keys: Table, segment_offsets: Column, column_order: list, null_precedence: list,
                                 stream=None, mr=None) -> Table:
    return _segmented_sort_by_key(values, keys, segment_offsets, column_order, null_precedence, 1, stream)


def top_k(col: Column, k: int, sort_order=1, stream=None, mr=None) -> Column:
    """cudf::top_k (sorting.hpp:370-391); sort_order defaults to DESCENDING (high to low)."""
    out = C.c_void_p()
    cv = col._view()
    check(lib.b2_top_k(C.byref(cv), int(k), int(sort_order), _lib.stream_arg(stream), C.byref(out)))
    return Column._from_handle(out.value)


def top_k_order(col: Column, k: int, sort_order=1, stream=None, mr=None) -> Column:
    out = C.c_void_p()
    cv = col._view()
    check(lib.b2_top_k_order(C.byref(cv), int(k), int(sort_order), _lib.stream_arg(stream), C.byref(out)))
    return Column._from_handle(out.value)


def rank(input_view: Column, method: int, column_order: int, null_handling: int, null_precedence: int, percentage: bool, stream=None,
         mr=None) -> Column:
    """cudf::rank (python/pylibcudf/pylibcudf/sorting.pyx rank; cpp/include/cudf/sorting.hpp:165-230).
    method: RankMethod (0 FIRST, 1 AVERAGE, 2 MIN, 3 MAX, 4 DENSE)."""
    out = C.c_void_p()
    cv = input_view._view()
    check(lib.b2_rank(C.byref(cv), int(method), int(column_order), int(null_handling), int(null_precedence), 1 if percentage else 0,
                      _lib.stream_arg(stream), C.byref(out)))
    return Column._from_handle(out.value)
