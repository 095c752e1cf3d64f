"""Run the pylibcudf-named shim against the CPU emulation of the kernels (tests/emu) — TEST INFRASTRUCTURE ONLY.

`install()` must be called in a dedicated (sub)process: it rebinds every ctypes entry point of `cudf_b200._lib.lib` to
tests/emu/_build/libcudf_b200_emu.so and replaces Column.from_numpy / Column.to_numpy by host-memory versions ("device"
pointers of the emulator are host pointers). The package itself has no switch that does this.
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def install():
    from tests.emu.build_emu import build

    emu = C.CDLL(str(build()))
    import cudf_b200._lib as L
    from cudf_b200.pylibcudf import column as colmod
    from cudf_b200.pylibcudf.types import DataType

    for name in L.DECLARED_SYMBOLS:
        real = getattr(L.lib, name)
        fn = getattr(emu, name)
        fn.argtypes, fn.restype = real.argtypes, real.restype
        setattr(L.lib, name, fn)
    L.current_stream = lambda: 0

    Column = colmod.Column
    emu.emu_alloc.restype, emu.emu_alloc.argtypes = C.c_void_p, [C.c_size_t]
    emu.emu_free.restype, emu.emu_free.argtypes = None, [C.c_void_p]

    class _DevCopy:
        """A host array copied into emulator-owned memory of EXACTLY its size (guard bytes / guard page right behind it)."""

        def __init__(self, arr: np.ndarray):
            self.nbytes = int(arr.nbytes)
            self.ptr = emu.emu_alloc(max(self.nbytes, 1))
            C.memmove(self.ptr, arr.ctypes.data, self.nbytes)

        def __del__(self):
            if self.ptr:
                emu.emu_free(self.ptr)
                self.ptr = None

    def from_numpy(cls, values, valid=None, dtype=None, device="cpu"):
        values = np.ascontiguousarray(values)
        if dtype is None:
            dtype = DataType.from_numpy(values.dtype)
        raw = _DevCopy(np.ascontiguousarray(values.view(np.uint8) if values.dtype != np.bool_ else values.astype(np.uint8)))
        mask_arr, nulls = None, 0
        if valid is not None:
            valid = np.asarray(valid, dtype=bool)
            nulls = int((~valid).sum())
            bits = np.packbits(valid, bitorder="little")
            padded = np.zeros(L.lib.b2_bitmask_allocation_size_bytes(len(valid)) or 64, dtype=np.uint8)
            padded[: len(bits)] = bits
            mask_arr = _DevCopy(padded)
        return cls(dtype, len(values), raw.ptr, mask_arr.ptr if mask_arr is not None else 0, nulls, 0, [raw, mask_arr])

    def _host_bytes(ptr: int, nbytes: int) -> np.ndarray:
        return np.frombuffer(C.string_at(ptr, nbytes), dtype=np.uint8).copy()

    def to_numpy(self):
        dt = self._type.numpy_dtype()
        if self._size == 0:
            vals = np.empty(0, dtype=dt)
        else:
            raw = _host_bytes(self._data + self._offset * dt.itemsize, self._size * dt.itemsize)
            vals = raw.view(np.uint8 if dt == np.bool_ else dt)
            if dt == np.bool_:
                vals = vals != 0
        valid = None
        if self._mask:
            nwords = (self._offset + self._size + 31) // 32
            bits = np.unpackbits(_host_bytes(self._mask, nwords * 4), bitorder="little")
            valid = bits[self._offset: self._offset + self._size].astype(bool)
        return vals, valid

    Column.from_numpy = classmethod(from_numpy)
    Column.to_numpy = to_numpy
    return emu


def install_cy():
    """`install()` + the Cython binding (cudf_b200.pylibcudf_cy) built against the emulator's library: the same _core.pyx, linked
    with tests/emu/_build/libcudf_b200_emu.so and loaded RTLD_DEEPBIND so that its b2_* calls bind to that library although the
    real one sits in the global scope. Returns the package."""
    import importlib.util
    import os

    install()
    from tests.emu.build_emu import LIB, OUT

    # build_cy.py is loaded by path: importing the package would load the product extension first
    bspec = importlib.util.spec_from_file_location("_b2_build_cy", str(ROOT / "cudf_b200" / "pylibcudf_cy" / "build_cy.py"))
    bmod = importlib.util.module_from_spec(bspec)
    bspec.loader.exec_module(bmod)
    build_cy = bmod.build

    so = build_cy(lib=LIB, out_dir=OUT / "cy")
    name = "cudf_b200.pylibcudf_cy._core"
    for m in [m for m in sys.modules if m == "cudf_b200.pylibcudf_cy" or m.startswith("cudf_b200.pylibcudf_cy.")]:
        del sys.modules[m]  # the product extension may have been imported already (e.g. by __graft_entry__.build())
    old = sys.getdlopenflags()
    sys.setdlopenflags(os.RTLD_NOW | os.RTLD_DEEPBIND)
    try:
        spec = importlib.util.spec_from_file_location(name, str(so))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.setdlopenflags(old)
    import cudf_b200.pylibcudf_cy as cy

    assert cy._core is mod
    return cy
