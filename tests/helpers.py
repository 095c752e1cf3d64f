"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np

NUMERIC_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]
ALL_DTYPES = NUMERIC_DTYPES + [np.bool_]


def make_col(values, dtype):
    """list with None -> (values ndarray, valid | None)."""
    valid = np.array([v is not None for v in values], dtype=bool)
    dt = np.dtype(dtype)
    raw = [0 if v is None else v for v in values]
    if dt == np.bool_:
        arr = np.array([bool(v) for v in raw], dtype=bool)
    elif dt.kind == "u":
        arr = np.array(raw, dtype=np.int64).astype(dt)  # negative literals wrap like the C++ wrappers
    else:
        arr = np.array(raw).astype(dt)
    return arr, (None if valid.all() else valid)


def to_plc_column(plc, col):
    values, valid = col
    return plc.Column.from_numpy(values, valid)


def to_plc_table(plc, cols):
    return plc.Table([to_plc_column(plc, c) for c in cols])


def col_to_numpy(column):
    return column.to_numpy()


def assert_columns_equal(got, expected, rtol=0.0, what=""):
    """(values, valid) pairs; values under nulls are ignored (DEVELOPER_GUIDE.md:516-524)."""
    gv, gm = got
    ev, em = expected
    assert len(gv) == len(ev), f"{what}: length {len(gv)} != {len(ev)}"
    gm_ = np.ones(len(gv), bool) if gm is None else np.asarray(gm, bool)
    em_ = np.ones(len(ev), bool) if em is None else np.asarray(em, bool)
    assert np.array_equal(gm_, em_), f"{what}: validity differs\n got {gm_}\n exp {em_}"
    g, e = np.asarray(gv)[em_], np.asarray(ev)[em_]
    if rtol and np.asarray(ev).dtype.kind == "f":
        np.testing.assert_allclose(g, e, rtol=rtol, atol=0, equal_nan=True, err_msg=what)
    else:
        if np.asarray(ev).dtype.kind == "f":
            assert np.array_equal(g, e, equal_nan=True), f"{what}: values differ\n got {g}\n exp {e}"
        else:
            assert np.array_equal(g, e), f"{what}: values differ\n got {g}\n exp {e}"
