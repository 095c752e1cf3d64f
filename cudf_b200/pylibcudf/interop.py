"""pylibcudf.interop twin for the Arrow C Data / C Device Data interface (python/pylibcudf/pylibcudf/interop.pyx; cpp/include/
cudf/interop.hpp): fixed-width columns <-> ArrowSchema / ArrowArray / ArrowDeviceArray structs (cudf_b200/csrc/arrow_interop.cu).

  to_arrow(column)            -> pyarrow.Array   (host copy through b2_to_arrow_host + pyarrow's C import)
  from_arrow(pyarrow.Array)   -> Column          (pyarrow's C export + b2_from_arrow_host)
  to_arrow_device(column)     -> ArrowDeviceArrayHolder (zero copy; .schema / .device_array are the C structs, __arrow_c_device_array__)
  from_arrow_device(holder or any object with __arrow_c_device_array__) -> Column viewing the producer's device buffers
"""
from __future__ import annotations

import ctypes as C

from .. import _lib
from .._lib import ColumnView, check, lib
from .column import Column
from .types import DataType, TypeId


class ArrowSchema(C.Structure):
    pass


ArrowSchema._fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64), ("n_children", C.c_int64),
                        ("children", C.POINTER(C.POINTER(ArrowSchema))), ("dictionary", C.POINTER(ArrowSchema)),
                        ("release", C.CFUNCTYPE(None, C.POINTER(ArrowSchema))), ("private_data", C.c_void_p)]


class ArrowArray(C.Structure):
    pass


ArrowArray._fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64), ("n_children", C.c_int64),
                       ("buffers", C.POINTER(C.c_void_p)), ("children", C.POINTER(C.POINTER(ArrowArray))), ("dictionary", C.POINTER(ArrowArray)),
                       ("release", C.CFUNCTYPE(None, C.POINTER(ArrowArray))), ("private_data", C.c_void_p)]


class ArrowDeviceArray(C.Structure):
    _fields_ = [("array", ArrowArray), ("device_id", C.c_int64), ("device_type", C.c_int32), ("sync_event", C.c_void_p), ("reserved", C.c_int64 * 3)]


ARROW_DEVICE_CUDA = 2

for _name, _args in (("b2_to_arrow_schema", [C.POINTER(ColumnView), C.c_char_p, C.POINTER(ArrowSchema)]),
                     ("b2_to_arrow_device", [C.POINTER(ColumnView), C.c_void_p, C.POINTER(ArrowDeviceArray)]),
                     ("b2_to_arrow_host", [C.POINTER(ColumnView), C.c_void_p, C.POINTER(ArrowArray)]),
                     ("b2_from_arrow_device", [C.POINTER(ArrowSchema), C.POINTER(ArrowDeviceArray), C.c_void_p, C.POINTER(ColumnView), C.POINTER(C.c_void_p)]),
                     ("b2_from_arrow_host", [C.POINTER(ArrowSchema), C.POINTER(ArrowArray), C.c_void_p, C.POINTER(C.c_void_p)])):
    _fn = getattr(lib, _name)
    _fn.argtypes, _fn.restype = _args, C.c_int
lib.b2_arrow_schema_release.argtypes, lib.b2_arrow_schema_release.restype = [C.POINTER(ArrowSchema)], None
lib.b2_arrow_array_release.argtypes, lib.b2_arrow_array_release.restype = [C.POINTER(ArrowArray)], None


class ArrowDeviceArrayHolder:
    """Owns the exported structs (released on deletion) and keeps the exporting column alive."""

    def __init__(self, column: Column, name: str = "", stream=None):
        self._column = column
        self.schema, self.device_array = ArrowSchema(), ArrowDeviceArray()
        v = column._view()
        check(lib.b2_to_arrow_schema(C.byref(v), name.encode(), C.byref(self.schema)))
        check(lib.b2_to_arrow_device(C.byref(v), _lib.stream_arg(stream), C.byref(self.device_array)))

    def __del__(self):
        lib.b2_arrow_array_release(C.byref(self.device_array.array))
        lib.b2_arrow_schema_release(C.byref(self.schema))


def to_arrow_device(column: Column, name: str = "", stream=None) -> ArrowDeviceArrayHolder:
    return ArrowDeviceArrayHolder(column, name, stream)


def from_arrow_device(holder, stream=None) -> Column:
    """A Column over the producer's device buffers (zero copy; BOOL8 is unpacked into a copy)."""
    view, owner = ColumnView(), C.c_void_p()
    check(lib.b2_from_arrow_device(C.byref(holder.schema), C.byref(holder.device_array), _lib.stream_arg(stream), C.byref(view), C.byref(owner)))
    if owner.value:
        return Column._from_handle(owner.value)
    return Column(DataType(TypeId(view.type_id)), view.size, view.data or 0, view.null_mask or 0, view.null_count, view.offset, [holder])


def to_arrow(column: Column, name: str = "", stream=None):
    """pyarrow.Array with the column's values (host copy) through the C Data interface."""
    import pyarrow as pa

    schema, array = ArrowSchema(), ArrowArray()
    v = column._view()
    check(lib.b2_to_arrow_schema(C.byref(v), name.encode(), C.byref(schema)))
    check(lib.b2_to_arrow_host(C.byref(v), _lib.stream_arg(stream), C.byref(array)))
    # pyarrow moves the structs (and calls their release callbacks when it is done with the buffers)
    return pa.Array._import_from_c(C.addressof(array), C.addressof(schema))


def from_arrow(arr, stream=None) -> Column:
    """Column (device copy) of a pyarrow.Array of a fixed-width type, through the C Data interface."""
    import pyarrow as pa

    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks()
    schema, array = ArrowSchema(), ArrowArray()
    arr._export_to_c(C.addressof(array), C.addressof(schema))
    out = C.c_void_p()
    try:
        check(lib.b2_from_arrow_host(C.byref(schema), C.byref(array), _lib.stream_arg(stream), C.byref(out)))
    finally:
        if array.release:
            array.release(C.byref(array))
        if schema.release:
            schema.release(C.byref(schema))
    return Column._from_handle(out.value)
