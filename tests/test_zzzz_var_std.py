"""groupby SUM_OF_SQUARES / M2 / VARIANCE / STD on the hash path (hash_compound_agg_finalizer.cu:135-186,
cpp/src/groupby/common/m2_var_std.cu). Golden vectors: cpp/tests/groupby/{var,std,sum_of_squares}_tests.cpp.
Runs against the oracle on CPU and the CUDA path on a GPU (first exercised on the emulator, tests/test_emu_kernels.py)."""
import numpy as np
import pytest

from tests.helpers import assert_columns_equal, make_col
from tests.impls import OracleImpl, PlcImpl, sort_groups

F64_RTOL = 1e-6


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def impl(request):
    if request.param == "oracle":
        return OracleImpl()
    return PlcImpl(request.getfixturevalue("plc"))


KEYS = [1, 2, 3, 1, 2, 2, 1, 3, 3, 2]
VALS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
NKEYS = ([1, 2, 3, 1, 2, 2, 1, 3, 3, 2, 4], [1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1])
NVALS = ([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 3], [0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1])


def run(impl, keys, vals, kinds):
    gk, gr = sort_groups(*impl.groupby([keys], [(vals, kinds)]))
    return gk[0], gr[0]


def check(col, values, valid=None):
    v, m = col
    mm = np.ones(len(values), bool) if m is None else np.asarray(m, bool)
    ev = np.ones(len(values), bool) if valid is None else np.asarray(valid, bool)
    assert np.array_equal(mm, ev), (mm, ev)
    np.testing.assert_allclose(np.asarray(v, np.float64)[ev], np.asarray(values, np.float64)[ev], rtol=F64_RTOL)


@pytest.mark.parametrize("vdtype", [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64])
def test_var_std_golden(impl, vdtype):
    keys, vals = make_col(KEYS, np.int32), make_col(VALS, vdtype)
    k, r = run(impl, keys, vals, ["var", "std", "sum_of_squares"])
    assert np.asarray(k[0]).tolist() == [1, 2, 3]
    assert np.asarray(r[0][0]).dtype == np.float64 and np.asarray(r[1][0]).dtype == np.float64
    check(r[0], [9.0, 131.0 / 12, 31.0 / 3])                                   # var_tests.cpp:30-41
    check(r[1], [3.0, np.sqrt(131.0 / 12), np.sqrt(31.0 / 3)])                  # std_tests.cpp:30-41
    assert np.asarray(r[2][0]).dtype == (np.int64 if np.dtype(vdtype).kind == "i" else np.dtype(vdtype))
    check(r[2], [45.0, 123.0, 117.0])                                          # sum_of_squares_tests.cpp:32-38
    # null keys and values; default ddof and ddof = 2 (var_tests.cpp:96-140, std_tests.cpp:92-111)
    keys = (np.array(NKEYS[0], np.int32), np.array(NKEYS[1], bool))
    vals = (np.array(NVALS[0], vdtype), np.array(NVALS[1], bool))
    k, r = run(impl, keys, vals, ["var", "var2", "std"])
    assert np.asarray(k[0]).tolist() == [1, 2, 3, 4]
    check(r[0], [4.5, 49.0 / 3, 18.0, 0.0], [1, 1, 1, 0])
    check(r[1], [0.0, 98.0 / 3, 0.0, 0.0], [0, 1, 0, 0])
    check(r[2], [3 / np.sqrt(2), 7 / np.sqrt(3), 3 * np.sqrt(2), 0.0], [1, 1, 1, 0])
    # zero valid values (var_tests.cpp:79-93): one group, null result
    k, r = run(impl, make_col([1, 1, 1], np.int32), (np.array([3, 4, 5], vdtype), np.zeros(3, bool)), ["var", "std", "sum_of_squares"])
    for c in r:
        assert np.asarray(c[1]).tolist() == [False]


@pytest.mark.parametrize("vdtype", [np.int8, np.int32, np.int64, np.float32, np.float64])
def test_product_golden(impl, vdtype):
    # product_tests.cpp:31-44,90-112
    k, r = run(impl, make_col(KEYS, np.int32), make_col(VALS, vdtype), ["product"])
    assert np.asarray(r[0][0]).dtype == (np.int64 if np.dtype(vdtype).kind == "i" else np.dtype(vdtype))
    check(r[0], [0.0, 180.0, 112.0])
    keys = (np.array(NKEYS[0], np.int32), np.array(NKEYS[1], bool))
    vals = (np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 3], vdtype), np.array([0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0], bool))
    k, r = run(impl, keys, vals, ["product", "sum"])
    check(r[0], [18.0, 36.0, 16.0, 3.0], [1, 1, 1, 0])
    check(r[1], [9.0, 14.0, 10.0, 0.0], [1, 1, 1, 0])


def test_var_std_empty_and_all_null_keys(impl):
    k, r = run(impl, make_col([], np.int32), make_col([], np.float64), ["var", "std", "m2"])
    assert len(k[0]) == 0 and all(len(c[0]) == 0 for c in r)
    k, r = run(impl, (np.array([1, 2, 3], np.int32), np.zeros(3, bool)), make_col([3, 4, 5], np.int64), ["var"])
    assert len(k[0]) == 0 and len(r[0][0]) == 0


@pytest.mark.parametrize("vdtype", [np.int32, np.int64, np.uint16, np.float32, np.float64])
def test_var_std_random(impl, vdtype):
    rng = np.random.default_rng(31)
    o = OracleImpl()
    for n, ng, nf in [(1, 1, 0.0), (2000, 17, 0.0), (60_000, 900, 0.2), (50_000, 40_000, 0.05)]:
        keys = (rng.integers(0, ng, n).astype(np.int64), (rng.random(n) >= nf / 2) if nf else None)
        raw = rng.integers(0, 200, n) if np.dtype(vdtype).kind == "u" else rng.integers(-300, 300, n)
        vals = (raw.astype(vdtype), (rng.random(n) >= nf) if nf else None)
        kinds = ["sum", "sum_of_squares", "m2", "var", "std", "var0", "std2", "mean", "count"]
        if np.dtype(vdtype).kind != "f":
            kinds.append("product")  # integer products wrap identically; float products depend on the multiplication order
        gk, gr = sort_groups(*impl.groupby([keys], [(vals, kinds)]))
        ek, er = sort_groups(*o.groupby([keys], [(vals, kinds)]))
        assert_columns_equal(gk[0], ek[0], what="keys")
        for j, kind in enumerate(kinds):
            g, e = gr[0][j], er[0][j]
            assert np.asarray(g[0]).dtype == np.asarray(e[0]).dtype, kind
            gm = np.ones(len(g[0]), bool) if g[1] is None else np.asarray(g[1], bool)
            em = np.ones(len(e[0]), bool) if e[1] is None else np.asarray(e[1], bool)
            assert np.array_equal(gm, em), kind
            if np.asarray(e[0]).dtype.kind == "f":
                rtol = 2e-4 if np.dtype(vdtype) == np.float32 else F64_RTOL
                # M2 = sumsq - sum^2 / n cancels: compare with a floor scaled by the group's sum of squares
                np.testing.assert_allclose(np.asarray(g[0], np.float64)[em], np.asarray(e[0], np.float64)[em], rtol=rtol, atol=rtol * 9e4)
            else:
                assert np.array_equal(np.asarray(g[0])[em], np.asarray(e[0])[em]), kind


def test_mean_of_wrapping_integer_sums(impl):
    """MEAN = SUM / COUNT with SUM in int64 for every integral source (aggregation.hpp:950-956,
    hash_compound_agg_finalizer.cu:95-131): sums that wrap, and unsigned sums above 2^63, read back as signed."""
    keys = make_col([0, 0, 1, 1, 2], np.int32)
    vals = (np.array([2**64 - 1, 0, 2**63, 2**63, 7], np.uint64), None)
    k, r = run(impl, keys, vals, ["sum", "mean"])
    assert np.asarray(r[0][0]).dtype == np.int64 and np.asarray(r[0][0]).tolist() == [-1, 0, 7]
    np.testing.assert_allclose(np.asarray(r[1][0]), [-0.5, 0.0, 7.0])
    vals = (np.array([2**63 - 1, 1, 5, 6, 7], np.int64), None)
    k, r = run(impl, keys, vals, ["mean"])
    np.testing.assert_allclose(np.asarray(r[0][0]), [-(2.0**63) / 2, 5.5, 7.0])


@pytest.mark.parametrize("vdtype", [np.int8, np.int32, np.int64, np.uint16, np.float32, np.float64])
def test_argmin_argmax(impl, vdtype):
    # argmax_tests.cpp:28-40,82-100 and argmin_tests.cpp:30-35,83-100
    keys, vals = make_col(KEYS, np.int32), make_col([9, 8, 7, 6, 5, 4, 3, 2, 1, 0], vdtype)
    k, r = run(impl, keys, vals, ["argmax", "argmin"])
    assert np.asarray(r[0][0]).dtype == np.int32
    assert np.asarray(r[0][0]).tolist() == [0, 1, 2] and np.asarray(r[1][0]).tolist() == [6, 9, 8]
    keys = (np.array(NKEYS[0], np.int32), np.array([1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1], bool))
    vals = (np.array([9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 4], vdtype), np.array([0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 0], bool))
    k, r = run(impl, keys, vals, ["argmax"])
    check((np.asarray(r[0][0], np.float64), r[0][1]), [3, 4, 7, 0], [1, 1, 1, 0])
    keys = (np.array(NKEYS[0], np.int32), np.array(NKEYS[1], bool))
    vals = (np.array([9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 4], vdtype), np.array([1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 0], bool))
    k, r = run(impl, keys, vals, ["argmin"])
    check((np.asarray(r[0][0], np.float64), r[0][1]), [3, 9, 8, 0], [1, 1, 1, 0])
    # random with ties: the first row among equal extremes (a valid outcome of the reference, which leaves ties open)
    rng = np.random.default_rng(77)
    o = OracleImpl()
    n = 30_000
    keys = (rng.integers(0, 500, n).astype(np.int64), None)
    raw = rng.integers(0, 40, n) if np.dtype(vdtype).kind == "u" else rng.integers(-20, 20, n)
    vals = (raw.astype(vdtype), rng.random(n) < 0.9)
    gk, gr = sort_groups(*impl.groupby([keys], [(vals, ["argmin", "argmax", "min", "max"])]))
    ek, er = sort_groups(*o.groupby([keys], [(vals, ["argmin", "argmax", "min", "max"])]))
    for j in range(4):
        assert_columns_equal(gr[0][j], er[0][j], what=str(j))
    ok = np.asarray(gr[0][0][1], bool) if gr[0][0][1] is not None else np.ones(len(gk[0][0]), bool)
    assert np.array_equal(vals[0][np.asarray(gr[0][0][0])[ok]], np.asarray(gr[0][2][0])[ok])   # values at argmin are the minima
    assert np.array_equal(vals[0][np.asarray(gr[0][1][0])[ok]], np.asarray(gr[0][3][0])[ok])
