"""cudf::pack / unpack / packed_size / pack_metadata (cpp/include/cudf/contiguous_split.hpp:233-317): round trip, sizes and the
wire bytes against oracle/pack.py (cpp/tests/copying/pack_tests.cpp:20-110: SingleColumnFixedWidth, ...NonNullable,
MultiColumnFixedWidth, EmptyColumns, sliced inputs, corrupted metadata)."""
import ctypes as C
import struct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bytes_at(plc, ptr, n):
    from cudf_b200.pylibcudf.column import DeviceSpan

    c = plc.Column(plc.DataType(plc.TypeId.UINT8), n, ptr, 0, 0, 0, None)
    return c.to_numpy()[0].tobytes() if n else b""


def _roundtrip(plc, cols):
    from oracle import pack as opack
    from tests.helpers import assert_columns_equal

    t = plc.Table([plc.Column.from_numpy(v, m) for v, m in cols])
    cs = plc.contiguous_split
    packed = cs.pack(t)
    assert cs.packed_size(t) == packed.gpu_data_size == opack.packed_size(cols)
    md, data = opack.pack(cols, [int(c.type().id()) for c in t.columns()])
    assert packed.metadata == md
    assert _bytes_at(plc, packed.gpu_data_ptr, packed.gpu_data_size) == data
    out = cs.unpack(packed)
    assert out.num_columns() == len(cols)
    for c, exp in zip(out.columns(), cols):
        assert_columns_equal(c.to_numpy(), exp)
        assert c.null_count() == (0 if exp[1] is None else int((~np.asarray(exp[1], bool)).sum()))
    # metadata of the unpacked (in place) table describes the same buffer
    assert cs.pack_metadata(out, packed.gpu_data_ptr, packed.gpu_data_size) == md
    return packed


def test_pack_gtest_shapes(plc):
    # pack_tests.cpp:69-94: {1..6} with validity {1,1,1,0,1,0}; without mask; three columns of different types
    v = np.array([1, 2, 3, 4, 5, 6], np.int32)
    m = np.array([1, 1, 1, 0, 1, 0], bool)
    _roundtrip(plc, [(v, m)])
    _roundtrip(plc, [(v.astype(np.float32), None)])
    _roundtrip(plc, [(v, m), (np.arange(7, 13, dtype=np.float32), np.array([1, 0, 1, 1, 1, 1], bool)), (np.arange(8, 14, dtype=np.int8), None)])


def test_pack_random_and_sliced(plc):
    rng = np.random.default_rng(4)
    for n in (1, 31, 32, 33, 1000, 70_001):
        cols = []
        for dt in (np.int8, np.uint16, np.int32, np.float32, np.int64, np.float64, np.bool_):
            v = rng.integers(0, 2, n).astype(dt) if dt == np.bool_ else (rng.standard_normal(n) * 100).astype(dt)
            cols.append((v, rng.random(n) < 0.8 if rng.random() < 0.6 else None))
        _roundtrip(plc, cols)
    # sliced input (offset not a multiple of 32): packed as if materialised (pack_tests.cpp NestedSliced / SlicedEmpty)
    from oracle import pack as opack

    n = 5000
    v, m = rng.integers(-1000, 1000, n).astype(np.int64), rng.random(n) < 0.7
    col = plc.Column.from_numpy(v, m).slice(37, 4001)
    packed = plc.contiguous_split.pack(plc.Table([col]))
    md, data = opack.pack([(v[37:4001], m[37:4001])], [int(plc.TypeId.INT64)])
    assert packed.metadata == md and _bytes_at(plc, packed.gpu_data_ptr, packed.gpu_data_size) == data


def test_pack_empty_and_errors(plc):
    cs = plc.contiguous_split
    e = plc.Table([plc.Column.from_numpy(np.empty(0, np.int32)), plc.Column.from_numpy(np.empty(0, np.float64))])
    p = cs.pack(e)                                     # pack_tests.cpp:108-116 EmptyColumns
    assert p.gpu_data_size == 0 and len(p.metadata) == 16 + 2 * 40
    out = cs.unpack(p)
    assert out.num_columns() == 2 and out.num_rows() == 0
    assert struct.unpack_from("<iiii", p.metadata) == (2, 2, 0, 0)
    t = plc.Table([plc.Column.from_numpy(np.arange(5, dtype=np.int32))])
    good = cs.pack(t)
    for bad in (good.metadata[:-3],                    # pack_tests.cpp:639-679 truncated / not a multiple / too long
                good.metadata + b"\0" * 40,
                struct.pack("<iiii", 1, 1, 5, 0) + good.metadata[16:],            # unsupported version
                struct.pack("<iiii", 2, -1, 5, 0) + good.metadata[16:],           # :707-718 negative column count
                struct.pack("<iiii", 2, 1, 4, 0) + good.metadata[16:]):           # row count differs from the columns
        with pytest.raises(RuntimeError):
            cs.unpack_from_memoryviews(bad, good.gpu_data_ptr, good)
