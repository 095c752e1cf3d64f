"""pylibcudf.copying.gather twin (python/pylibcudf/pylibcudf/copying.pyx:64-113)."""
from __future__ import annotations

import ctypes as C

from .. import _lib
from .._lib import check, lib
from .column import Column, Table
from .types import OutOfBoundsPolicy


def gather(source_table: Table, gather_map: Column, bounds_policy: OutOfBoundsPolicy, stream=None, mr=None) -> Table:
    out = C.c_void_p()
    tv = source_table._view()
    mv = gather_map._view()
    check(lib.b2_gather(C.byref(tv), C.byref(mv), int(bounds_policy), _lib.stream_arg(stream), C.byref(out)))
    return Table._from_handle(out.value)
