"""Column / Table / Scalar over device memory (python/pylibcudf/pylibcudf/{column,table,scalar}.pyx).

Device memory for inputs comes from torch tensors (plumbing only); outputs are library-owned handles
released when the Python object dies. Both expose ``__cuda_array_interface__`` spans.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .._lib import ColumnView, TableView, check, lib
from .types import DataType, TypeId


class DeviceSpan:
    """A typed window on device memory exposing __cuda_array_interface__; keeps `owner` alive."""

    def __init__(self, ptr: int, nelems: int, dtype, owner):
        self.ptr = int(ptr or 0)
        self.nelems = int(nelems)
        self.dtype = np.dtype(dtype)
        self.owner = owner

    @property
    def __cuda_array_interface__(self):
        typestr = self.dtype.str if self.dtype != np.bool_ else "|b1"
        return {"shape": (self.nelems,), "typestr": typestr, "data": (self.ptr, False), "version": 3, "strides": None}


def _torch():
    import torch

    return torch


class _ColumnHandle:
    def __init__(self, handle: int):
        self.handle = handle

    def __del__(self):
        if self.handle:
            lib.b2_column_free(C.c_void_p(self.handle))
            self.handle = 0


class Column:
    """Non-owning view fields + whatever object keeps the memory alive."""

    def __init__(self, data_type: DataType, size: int, data_ptr: int, mask_ptr: int, null_count: int, offset: int, owners):
        self._type = data_type
        self._size = int(size)
        self._data = int(data_ptr or 0)
        self._mask = int(mask_ptr or 0)
        self._null_count = int(null_count)
        self._offset = int(offset)
        self._owners = owners

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def _from_handle(cls, handle: int) -> "Column":
        h = _ColumnHandle(handle)
        v = ColumnView()
        check(lib.b2_column_view_of(C.c_void_p(handle), C.byref(v)))
        return cls(DataType(TypeId(v.type_id)), v.size, v.data, v.null_mask, v.null_count, v.offset, [h])

    @classmethod
    def from_torch(cls, data, mask=None, null_count: int | None = None, dtype: DataType | None = None, offset: int = 0,
                   size: int | None = None) -> "Column":
        """Zero-copy from a contiguous CUDA tensor; `mask` is a CUDA int32/uint8 tensor holding Arrow validity words."""
        torch = _torch()
        assert data.is_cuda and data.is_contiguous()
        if dtype is None:
            npdt = np.bool_ if data.dtype == torch.bool else np.dtype(str(data.dtype).replace("torch.", ""))
            dtype = DataType.from_numpy(npdt)
        n = data.numel() - offset if size is None else size
        mask_ptr = 0
        if mask is not None:
            assert mask.is_cuda and mask.is_contiguous()
            mask_ptr = mask.data_ptr()
            if null_count is None:
                out = C.c_int32(0)
                check(lib.b2_null_count(C.c_void_p(mask_ptr), offset, offset + n, _lib.stream_arg(None), C.byref(out)))
                null_count = out.value
        return cls(dtype, n, data.data_ptr(), mask_ptr, null_count or 0, offset, [data, mask])

    @classmethod
    def from_numpy(cls, values, valid=None, dtype: DataType | None = None, device="cuda") -> "Column":
        """Host -> device copy. `valid` is an optional boolean array (True = valid)."""
        torch = _torch()
        values = np.ascontiguousarray(values)
        if dtype is None:
            dtype = DataType.from_numpy(values.dtype)
        raw = values.view(np.uint8) if values.dtype != np.bool_ else values.astype(np.uint8)
        t = torch.from_numpy(raw.copy()).to(device)
        mask_t = None
        nulls = 0
        if valid is not None:
            valid = np.asarray(valid, dtype=bool)
            nulls = int((~valid).sum())
            bits = np.packbits(valid, bitorder="little")
            padded = np.zeros(lib.b2_bitmask_allocation_size_bytes(len(valid)) or 64, dtype=np.uint8)
            padded[: len(bits)] = bits
            mask_t = torch.from_numpy(padded).to(device)
        return cls(dtype, len(values), t.data_ptr(), mask_t.data_ptr() if mask_t is not None else 0, nulls, 0, [t, mask_t])

    @classmethod
    def from_arrow(cls, arr, device="cuda") -> "Column":
        import pyarrow as pa

        if isinstance(arr, pa.ChunkedArray):
            arr = arr.combine_chunks()
        np_vals = arr.fill_null(0).to_numpy(zero_copy_only=False) if arr.null_count else arr.to_numpy(zero_copy_only=False)
        valid = None
        if arr.null_count:
            valid = ~np.asarray(arr.is_null().to_numpy(zero_copy_only=False), dtype=bool)
        return cls.from_numpy(np_vals, valid, device=device)

    # ---- accessors (names follow pylibcudf.Column) ----------------------------------------
    def type(self) -> DataType:
        return self._type

    def size(self) -> int:
        return self._size

    def null_count(self) -> int:
        return self._null_count

    def offset(self) -> int:
        return self._offset

    def data(self) -> DeviceSpan:
        dt = self._type.numpy_dtype()
        return DeviceSpan(self._data + self._offset * dt.itemsize, self._size, dt, self)

    def null_mask(self) -> DeviceSpan | None:
        if not self._mask:
            return None
        nwords = (self._offset + self._size + 31) // 32
        return DeviceSpan(self._mask, nwords, np.uint32, self)

    def nullable(self) -> bool:
        return bool(self._mask)

    def has_nulls(self) -> bool:
        return self._null_count > 0

    def with_mask(self, mask, null_count: int) -> "Column":
        return Column(self._type, self._size, self._data, mask.data_ptr() if mask is not None else 0, null_count, self._offset,
                      [self._owners, mask])

    def slice(self, begin: int, end: int) -> "Column":
        """cudf::slice of one range: shares memory, moves `offset` (cpp/include/cudf/copying.hpp slice)."""
        assert 0 <= begin <= end <= self._size
        nulls = 0
        if self._mask and self._null_count:
            out = C.c_int32(0)
            check(lib.b2_null_count(C.c_void_p(self._mask), self._offset + begin, self._offset + end, _lib.stream_arg(None),
                                    C.byref(out)))
            nulls = out.value
        return Column(self._type, end - begin, self._data, self._mask, nulls, self._offset + begin, [self])

    def _view(self) -> ColumnView:
        return ColumnView(int(self._type.id()), self._size, self._data or None, self._mask or None, self._null_count, self._offset)

    # ---- export ---------------------------------------------------------------------------
    def to_torch(self):
        torch = _torch()
        if self._size == 0:
            tdt = getattr(torch, str(self._type.numpy_dtype()) if self._type.id() != TypeId.BOOL8 else "bool")
            return torch.empty(0, dtype=tdt, device="cuda")
        span = self.data()
        if self._type.id() == TypeId.BOOL8:
            span = DeviceSpan(span.ptr, span.nelems, np.uint8, self)
            return torch.as_tensor(span, device="cuda").bool()
        return torch.as_tensor(span, device="cuda")

    # ---- DLPack (python/pylibcudf/pylibcudf/interop.pyx to_dlpack / from_dlpack; data only, like the reference) -----
    def __dlpack__(self, stream=None):
        if self._null_count:
            raise ValueError("DLPack cannot carry a null mask (cudf::to_dlpack rejects columns with nulls)")
        return self.to_torch().__dlpack__(stream=stream) if stream is not None else self.to_torch().__dlpack__()

    def __dlpack_device__(self):
        return self.to_torch().__dlpack_device__()

    @classmethod
    def from_dlpack(cls, obj) -> "Column":
        """Zero-copy import of a 1-D contiguous CUDA DLPack producer (anything with __dlpack__)."""
        t = _torch().from_dlpack(obj)
        if t.dim() != 1 or not t.is_contiguous():
            raise ValueError("from_dlpack: a contiguous 1-D tensor is required")
        return cls.from_torch(t)

    def to_numpy(self):
        """(values, valid) on the host; valid is None when the column has no mask."""
        torch = _torch()
        torch.cuda.current_stream().synchronize()
        dt = self._type.numpy_dtype()
        if self._size == 0:
            vals = np.empty(0, dtype=dt)
        else:
            raw = DeviceSpan(self._data + self._offset * dt.itemsize, self._size * dt.itemsize, np.uint8, self)
            vals = torch.as_tensor(raw, device="cuda").cpu().numpy().view(np.uint8 if dt == np.bool_ else dt)
            if dt == np.bool_:
                vals = vals != 0
        valid = None
        if self._mask:
            span = self.null_mask()
            words = torch.as_tensor(DeviceSpan(span.ptr, span.nelems, np.int32, self), device="cuda").cpu().numpy().view(np.uint32)
            bits = np.unpackbits(words.view(np.uint8), bitorder="little")
            valid = bits[self._offset: self._offset + self._size].astype(bool)
        return vals, valid

    def to_arrow(self):
        import pyarrow as pa

        vals, valid = self.to_numpy()
        return pa.array(vals, mask=None if valid is None else ~valid)

    def __repr__(self):
        return f"Column({self._type!r}, size={self._size}, null_count={self._null_count}, offset={self._offset})"


class Table:
    def __init__(self, columns):
        self._columns = list(columns)
        if self._columns:
            n = self._columns[0].size()
            assert all(c.size() == n for c in self._columns), "Column size mismatch"

    @classmethod
    def _from_handle(cls, handle: int) -> "Table":
        n = lib.b2_table_num_columns(C.c_void_p(handle))
        arr = (C.c_void_p * max(n, 1))()
        check(lib.b2_table_release(C.c_void_p(handle), arr, max(n, 1)))
        lib.b2_table_free(C.c_void_p(handle))
        return cls([Column._from_handle(arr[i]) for i in range(n)])

    def columns(self):
        return list(self._columns)

    def num_columns(self) -> int:
        return len(self._columns)

    def num_rows(self) -> int:
        return self._columns[0].size() if self._columns else 0

    def _view(self):
        n = len(self._columns)
        arr = (ColumnView * max(n, 1))()
        for i, c in enumerate(self._columns):
            arr[i] = c._view()
        tv = TableView(arr, n)
        tv._keepalive = arr
        return tv


class Scalar:
    """numeric_scalar<T> on the device (python/pylibcudf/pylibcudf/scalar.pyx)."""

    def __init__(self, handle: int):
        self._handle = handle

    @classmethod
    def from_py(cls, value, data_type: DataType, valid: bool = True, stream=None) -> "Scalar":
        buf = np.zeros(1, dtype=data_type.numpy_dtype())
        if value is not None:
            buf[0] = value
        raw = np.zeros(8, dtype=np.uint8)
        raw[: buf.itemsize] = buf.view(np.uint8)
        out = C.c_void_p()
        check(lib.b2_scalar_create(int(data_type.id()), raw.ctypes.data_as(C.c_void_p), 1 if (valid and value is not None) else 0,
                                   _lib.stream_arg(stream), C.byref(out)))
        return cls(out.value)

    def type(self) -> DataType:
        return DataType(TypeId(lib.b2_scalar_type(C.c_void_p(self._handle))))

    def _get(self, stream=None):
        raw = np.zeros(8, dtype=np.uint8)
        valid = C.c_int32(0)
        check(lib.b2_scalar_get(C.c_void_p(self._handle), _lib.stream_arg(stream), raw.ctypes.data_as(C.c_void_p), C.byref(valid)))
        dt = self.type().numpy_dtype()
        return raw[: dt.itemsize].view(dt)[0], bool(valid.value)

    def is_valid(self, stream=None) -> bool:
        return self._get(stream)[1]

    def to_py(self, stream=None):
        v, ok = self._get(stream)
        return v.item() if ok else None

    def __del__(self):
        if getattr(self, "_handle", 0):
            lib.b2_scalar_free(C.c_void_p(self._handle))
            self._handle = 0
