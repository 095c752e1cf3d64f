"""pylibcudf.join twin (python/pylibcudf/pylibcudf/join.pyx:63-205) + cudf::hash_join object."""
from __future__ import annotations

import ctypes as C

from .. import _lib
from .._lib import check, lib
from .column import Column, Table
from .types import NullEquality


def _free_join(name, left_keys, right_keys, nulls_equal, stream):
    lo, ro = C.c_void_p(), C.c_void_p()
    lv, rv = left_keys._view(), right_keys._view()
    check(getattr(lib, name)(C.byref(lv), C.byref(rv), int(nulls_equal), _lib.stream_arg(stream), C.byref(lo), C.byref(ro)))
    return Column._from_handle(lo.value), Column._from_handle(ro.value)


def inner_join(left_keys: Table, right_keys: Table, nulls_equal: NullEquality, stream=None, mr=None):
    return _free_join("b2_inner_join", left_keys, right_keys, nulls_equal, stream)


def left_join(left_keys: Table, right_keys: Table, nulls_equal: NullEquality, stream=None, mr=None):
    return _free_join("b2_left_join", left_keys, right_keys, nulls_equal, stream)


def full_join(left_keys: Table, right_keys: Table, nulls_equal: NullEquality, stream=None, mr=None):
    return _free_join("b2_full_join", left_keys, right_keys, nulls_equal, stream)


class HashJoin:
    """cudf::hash_join (cpp/include/cudf/join/hash_join.hpp): build once, probe many."""

    def __init__(self, build: Table, compare_nulls: NullEquality = NullEquality.EQUAL, has_nulls: bool | None = None,
                 load_factor: float = 0.5, stream=None):
        self._build = build
        out = C.c_void_p()
        bv = build._view()
        hn = -1 if has_nulls is None else int(bool(has_nulls))
        check(lib.b2_hash_join_create(C.byref(bv), hn, int(compare_nulls), float(load_factor), _lib.stream_arg(stream), C.byref(out)))
        self._handle = out.value

    def _probe(self, name, probe, output_size, stream):
        lo, ro = C.c_void_p(), C.c_void_p()
        pv = probe._view()
        check(getattr(lib, name)(C.c_void_p(self._handle), C.byref(pv), 0 if output_size is None else 1, int(output_size or 0),
                                 _lib.stream_arg(stream), C.byref(lo), C.byref(ro)))
        return Column._from_handle(lo.value), Column._from_handle(ro.value)

    def _size(self, name, probe, stream):
        out = C.c_size_t(0)
        pv = probe._view()
        check(getattr(lib, name)(C.c_void_p(self._handle), C.byref(pv), _lib.stream_arg(stream), C.byref(out)))
        return out.value

    def inner_join(self, probe: Table, output_size: int | None = None, stream=None):
        return self._probe("b2_hash_join_inner_join", probe, output_size, stream)

    def left_join(self, probe: Table, output_size: int | None = None, stream=None):
        return self._probe("b2_hash_join_left_join", probe, output_size, stream)

    def full_join(self, probe: Table, output_size: int | None = None, stream=None):
        return self._probe("b2_hash_join_full_join", probe, output_size, stream)

    def inner_join_size(self, probe: Table, stream=None) -> int:
        return self._size("b2_hash_join_inner_join_size", probe, stream)

    def left_join_size(self, probe: Table, stream=None) -> int:
        return self._size("b2_hash_join_left_join_size", probe, stream)

    def full_join_size(self, probe: Table, stream=None) -> int:
        return self._size("b2_hash_join_full_join_size", probe, stream)

    def __del__(self):
        if getattr(self, "_handle", 0):
            lib.b2_hash_join_destroy(C.c_void_p(self._handle))
            self._handle = 0
