"""pylibcudf.groupby twin (python/pylibcudf/pylibcudf/groupby.pyx:36-243)."""
from __future__ import annotations

import ctypes as C

from .. import _lib
from .._lib import AggRequest, check, lib
from .column import Column, Table
from .types import NullPolicy, Sorted


class GroupByRequest:
    def __init__(self, values: Column, aggregations: list):
        self._values = values
        self._aggregations = list(aggregations)


class GroupBy:
    def __init__(self, keys: Table, null_handling: NullPolicy = NullPolicy.EXCLUDE, keys_are_sorted: Sorted = Sorted.NO,
                 column_order: list | None = None, null_precedence: list | None = None):
        self._keys = keys  # keep the key buffers alive (groupby.pyx:136-138)
        o = [int(x) for x in (column_order or [])]
        p = [int(x) for x in (null_precedence or [])]
        oa = (C.c_uint8 * max(len(o), 1))(*o)
        pa = (C.c_uint8 * max(len(p), 1))(*p)
        out = C.c_void_p()
        kv = keys._view()
        self._kv = kv
        check(lib.b2_groupby_create(C.byref(kv), int(null_handling), int(keys_are_sorted), oa, len(o), pa, len(p), C.byref(out)))
        self._handle = out.value

    def _run(self, fn, requests, stream):
        n = len(requests)
        arr = (AggRequest * max(n, 1))()
        keep = []
        for i, r in enumerate(requests):
            kinds = (C.c_int32 * max(len(r._aggregations), 1))(*[a.abi_kind() for a in r._aggregations])
            keep.append(kinds)
            arr[i] = AggRequest(r._values._view(), kinds, len(r._aggregations))
        ko, ro = C.c_void_p(), C.c_void_p()
        check(fn(C.c_void_p(self._handle), arr, n, _lib.stream_arg(stream), C.byref(ko), C.byref(ro)))
        keys = Table._from_handle(ko.value)
        flat = Table._from_handle(ro.value).columns()
        results, k = [], 0
        for r in requests:
            m = len(r._aggregations)
            results.append(Table(flat[k: k + m]))
            k += m
        return keys, results

    def aggregate(self, requests: list, stream=None, mr=None):
        """-> (Table group_keys, [Table results per request])  (groupby.pyx:165-201)"""
        return self._run(lib.b2_groupby_aggregate, requests, stream)

    def scan(self, requests: list, stream=None, mr=None):
        return self._run(lib.b2_groupby_scan, requests, stream)

    def __del__(self):
        if getattr(self, "_handle", 0):
            lib.b2_groupby_destroy(C.c_void_p(self._handle))
            self._handle = 0
