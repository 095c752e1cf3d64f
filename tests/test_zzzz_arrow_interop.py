"""Arrow C Data / C Device Data interface (cudf::to_arrow_schema / to_arrow_host / to_arrow_device / from_arrow /
from_arrow_device_column; cpp/tests/interop/{to_arrow_device_test,from_arrow_device_test,to_arrow_host_test}.cpp check the same
properties): pyarrow on the host is the oracle of the struct layout and of the format strings."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
pa = pytest.importorskip("pyarrow")

TYPES = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64, np.bool_]


def _rand(rng, dt, n):
    if dt == np.bool_:
        return rng.random(n) < 0.5
    if np.dtype(dt).kind == "f":
        return rng.standard_normal(n).astype(dt)
    return rng.integers(0, 100, n).astype(dt)


@pytest.mark.parametrize("dt", TYPES)
def test_host_round_trip_matches_pyarrow(plc, dt):
    rng = np.random.default_rng(3)
    for n in (0, 1, 31, 32, 33, 1000):
        v = _rand(rng, dt, n)
        for valid in (None, rng.random(n) < 0.7):
            col = plc.Column.from_numpy(v, valid)
            arr = plc.interop.to_arrow(col, name="x")
            exp = pa.array(v, mask=None if valid is None else ~valid)
            assert arr.type == exp.type and arr.equals(exp), (dt, n)
            back = plc.interop.from_arrow(exp)
            gv, gm = back.to_numpy()
            em = np.ones(n, bool) if valid is None else valid
            assert back.type().id() == col.type().id() and back.size() == n
            assert back.null_count() == int((~em).sum())
            assert np.array_equal(np.ones(n, bool) if gm is None else gm, em) and np.array_equal(gv[em], v[em])
    # sliced on both sides (bit offsets that are not multiples of 8 / 32)
    n = 500
    v, valid = _rand(rng, dt, n), rng.random(n) < 0.6
    arr = plc.interop.to_arrow(plc.Column.from_numpy(v, valid).slice(37, 401))
    assert arr.equals(pa.array(v[37:401], mask=~valid[37:401]))
    back = plc.interop.from_arrow(pa.array(v, mask=~valid).slice(13, 300))
    gv, gm = back.to_numpy()
    assert np.array_equal(gm, valid[13:313]) and np.array_equal(gv[gm], v[13:313][gm])


def test_temporal_formats(plc):
    t = plc.TypeId
    for tid, patype in [(t.TIMESTAMP_SECONDS, pa.timestamp("s")), (t.TIMESTAMP_MILLISECONDS, pa.timestamp("ms")), (t.TIMESTAMP_MICROSECONDS, pa.timestamp("us")),
                        (t.TIMESTAMP_NANOSECONDS, pa.timestamp("ns")), (t.DURATION_SECONDS, pa.duration("s")), (t.DURATION_NANOSECONDS, pa.duration("ns")),
                        (t.TIMESTAMP_DAYS, pa.date32())]:
        raw = np.arange(5, dtype=np.int32 if tid == t.TIMESTAMP_DAYS else np.int64)
        col = plc.Column.from_numpy(raw, dtype=plc.DataType(tid))
        arr = plc.interop.to_arrow(col)
        assert arr.type == patype and arr.cast(pa.int32() if tid == t.TIMESTAMP_DAYS else pa.int64()).to_pylist() == raw.tolist()
        back = plc.interop.from_arrow(arr)
        assert back.type().id() == tid and np.array_equal(back.to_numpy()[0].view(raw.dtype), raw)


def test_device_array_is_zero_copy(plc):
    rng = np.random.default_rng(5)
    v, valid = rng.integers(0, 1000, 777).astype(np.int64), rng.random(777) < 0.8
    col = plc.Column.from_numpy(v, valid).slice(5, 700)
    h = plc.interop.to_arrow_device(col, name="k")
    a = h.device_array
    assert h.schema.format == b"l" and h.schema.name == b"k"
    assert a.device_type == plc.interop.ARROW_DEVICE_CUDA and a.array.length == 695 and a.array.offset == 5 and a.array.n_buffers == 2
    assert a.array.buffers[1] == col._data and a.array.buffers[0] == col._mask          # the producer's own buffers
    assert a.array.null_count == col.null_count() and a.sync_event
    back = plc.interop.from_arrow_device(h)
    assert back._data == col._data and back.offset() == 5 and back.null_count() == col.null_count()
    gv, gm = back.to_numpy()
    assert np.array_equal(gm, valid[5:700]) and np.array_equal(gv[gm], v[5:700][gm])
    # BOOL8 is bit-packed on the Arrow side: converted both ways
    b, bm = rng.random(100) < 0.5, rng.random(100) < 0.9
    hb = plc.interop.to_arrow_device(plc.Column.from_numpy(b, bm))
    assert hb.schema.format == b"b" and hb.device_array.array.buffers[1] != 0
    gb, gbm = plc.interop.from_arrow_device(hb).to_numpy()
    assert np.array_equal(gbm, bm) and np.array_equal(gb[bm], b[bm])
    with pytest.raises(TypeError):   # cudf::data_type_error: DURATION_DAYS has no Arrow type
        plc.interop.to_arrow(plc.Column.from_numpy(np.arange(3, dtype=np.int32), dtype=plc.DataType(plc.TypeId.DURATION_DAYS)))
