"""pylibcudf.null_mask twin (python/pylibcudf/pylibcudf/null_mask.pyx)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import check, lib
from .column import Column, DeviceSpan, Table
from .types import MaskState


class DeviceBuffer:
    """Owning rmm::device_buffer stand-in."""

    def __init__(self, handle: int):
        self._handle = handle

    @property
    def ptr(self) -> int:
        return int(lib.b2_buffer_data(C.c_void_p(self._handle)) or 0)

    @property
    def size(self) -> int:
        return int(lib.b2_buffer_size(C.c_void_p(self._handle)))

    def data_ptr(self) -> int:
        return self.ptr

    def words(self) -> DeviceSpan:
        return DeviceSpan(self.ptr, self.size // 4, np.uint32, self)

    def to_numpy_bits(self, nbits: int):
        import torch

        torch.cuda.current_stream().synchronize()
        if self.size == 0:
            return np.zeros(0, dtype=bool)
        w = torch.as_tensor(DeviceSpan(self.ptr, self.size, np.uint8, self), device="cuda").cpu().numpy()
        return np.unpackbits(w, bitorder="little")[:nbits].astype(bool)

    def __del__(self):
        if getattr(self, "_handle", 0):
            lib.b2_buffer_free(C.c_void_p(self._handle))
            self._handle = 0


def bitmask_allocation_size_bytes(number_of_bits: int) -> int:
    return int(lib.b2_bitmask_allocation_size_bytes(number_of_bits))


def create_null_mask(size: int, state: MaskState = MaskState.UNINITIALIZED, stream=None, mr=None) -> DeviceBuffer:
    out = C.c_void_p()
    check(lib.b2_create_null_mask(size, int(state), _lib.stream_arg(stream), C.byref(out)))
    return DeviceBuffer(out.value)


def copy_bitmask(col: Column, stream=None, mr=None) -> DeviceBuffer:
    out = C.c_void_p()
    check(lib.b2_copy_bitmask(C.c_void_p(col._mask or None), col.offset(), col.offset() + col.size(), _lib.stream_arg(stream),
                              C.byref(out)))
    return DeviceBuffer(out.value)


def bitmask_and(columns, stream=None, mr=None):
    tbl = columns if isinstance(columns, Table) else Table(columns)
    tv = tbl._view()
    out = C.c_void_p()
    nc = C.c_int32(0)
    check(lib.b2_bitmask_and(C.byref(tv), _lib.stream_arg(stream), C.byref(out), C.byref(nc)))
    return DeviceBuffer(out.value), nc.value


def null_count(bitmask_ptr: int, start: int, stop: int, stream=None) -> int:
    out = C.c_int32(0)
    check(lib.b2_null_count(C.c_void_p(bitmask_ptr or None), start, stop, _lib.stream_arg(stream), C.byref(out)))
    return out.value


def count_set_bits(bitmask_ptr: int, start: int, stop: int, stream=None) -> int:
    out = C.c_int32(0)
    check(lib.b2_count_set_bits(C.c_void_p(bitmask_ptr or None), start, stop, _lib.stream_arg(stream), C.byref(out)))
    return out.value


def set_null_mask(bitmask_ptr: int, begin_bit: int, end_bit: int, valid: bool, stream=None) -> None:
    check(lib.b2_set_null_mask(C.c_void_p(bitmask_ptr or None), begin_bit, end_bit, int(valid), _lib.stream_arg(stream)))
