"""More known-answer vectors from the reference gtests (scan min/max/count, leading nulls, segmented reduce with
INCLUDE policy / partial offsets, reduce min/max/product) — oracle on CPU, CUDA path on GPU."""
import numpy as np
import pytest

from tests.helpers import assert_columns_equal, make_col
from tests.impls import OracleImpl, PlcImpl

N = None
COL = [5, 4, 6, 0, 1, 6, 5, 3]
COLN = [5, 4, 6, N, 1, 6, 5, 3]
TYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def impl(request):
    return OracleImpl() if request.param == "oracle" else PlcImpl(request.getfixturevalue("plc"))


def ident(dtype, kind):
    dt = np.dtype(dtype)
    if dt.kind == "f":
        return np.inf if kind == "min" else -np.inf
    return int(np.iinfo(dt).max) if kind == "min" else int(np.iinfo(dt).min)


# cpp/tests/reductions/scan_tests.cpp:27-155 (ScanMinTest / ScanMaxTest)
@pytest.mark.parametrize("dtype", TYPES)
def test_scan_min_max_golden(impl, dtype):
    col, coln = make_col(COL, dtype), make_col(COLN, dtype)
    valid = np.array([v is not None for v in COLN])
    cases = [
        ("min", col, True, False, [5, 4, 4, 0, 0, 0, 0, 0], None),
        ("min", col, False, False, [ident(dtype, "min"), 5, 4, 4, 0, 0, 0, 0], None),
        ("min", coln, True, False, [5, 4, 4, 4, 1, 1, 1, 1], valid),
        ("min", coln, True, True, [5, 4, 4, 0, 0, 0, 0, 0], np.arange(8) < 3),
        ("min", coln, False, False, [ident(dtype, "min"), 5, 4, 4, 4, 1, 1, 1], valid),
        ("max", col, True, False, [5, 5, 6, 6, 6, 6, 6, 6], None),
        ("max", col, False, False, [ident(dtype, "max"), 5, 5, 6, 6, 6, 6, 6], None),
        ("max", coln, True, False, [5, 5, 6, 6, 6, 6, 6, 6], valid),
        ("max", coln, True, True, [5, 5, 6, 0, 0, 0, 0, 0], np.arange(8) < 3),
    ]
    for kind, c, inclusive, include, exp, ev in cases:
        got = impl.scan(c, kind, inclusive, include)
        assert_columns_equal(got, (np.array(exp, dtype=dtype), ev), what=f"{kind} incl={inclusive} include={include}")


# scan_tests.cpp:404-437 (ScanLeadingNullsTest)
@pytest.mark.parametrize("dtype", TYPES)
def test_scan_leading_nulls(impl, dtype):
    c = make_col([N, 20, 30], dtype)
    got = impl.scan(c, "min", True, False)
    assert_columns_equal(got, (np.array([0, 20, 20], dtype=dtype), np.array([False, True, True])))
    got = impl.scan(c, "min", True, True)
    assert_columns_equal(got, (np.array([0, 0, 0], dtype=dtype), np.array([False, False, False])))


# scan_tests.cpp:476-545 (ScanCountTest; the strings columns only contribute their validity)
def test_scan_count_golden(impl):
    col, coln = make_col(COL, np.int32), make_col(COLN, np.int32)
    valid = np.array([v is not None for v in COLN])
    assert_columns_equal(impl.scan(col, "count", True, False), (np.arange(1, 9, dtype=np.int32), None))
    assert_columns_equal(impl.scan(col, "count_all", True, False), (np.arange(1, 9, dtype=np.int32), None))
    got = impl.scan(coln, "count", True, False)
    assert np.asarray(got[0]).dtype == np.int32
    assert_columns_equal(got, (np.array([1, 2, 3, 4, 4, 5, 6, 7], dtype=np.int32), valid))
    got = impl.scan(coln, "count_all", True, True)
    assert_columns_equal(got, (np.array([1, 2, 3, 0, 0, 0, 0, 0], dtype=np.int32), np.arange(8) < 3))
    # InclusiveWithOffset: slice [1, 7)
    sl = (coln[0][1:7], coln[1][1:7])
    got = impl.scan(sl, "count", True, False)
    assert_columns_equal(got, (np.array([1, 2, 3, 3, 4, 5], dtype=np.int32), np.array([1, 1, 0, 1, 1, 1], dtype=bool)))


# cpp/tests/reductions/segmented_reduction_tests.cpp:316-365 (SumIncludeNulls) and :649-700 (PartialSegmentReduction)
@pytest.mark.parametrize("dtype", [np.int8, np.int32, np.int64, np.uint32, np.float32, np.float64])
def test_segmented_sum_include_nulls(impl, dtype):
    col = make_col([1, 2, 3, 1, N, 3, 1, N, N, N], dtype)
    offs = [0, 3, 6, 7, 8, 10, 10]
    got = impl.segmented_reduce(col, offs, "sum", dtype, include=True)
    assert_columns_equal(got, (np.array([6, 0, 1, 0, 0, 0], dtype=dtype), np.array([1, 0, 1, 0, 0, 0], dtype=bool)))
    got = impl.segmented_reduce(col, offs, "sum", dtype, include=True, init=(3, True))
    assert_columns_equal(got, (np.array([9, 0, 4, 0, 0, 3], dtype=dtype), np.array([1, 0, 1, 0, 0, 1], dtype=bool)))
    got = impl.segmented_reduce(col, offs, "sum", dtype, include=True, init=(3, False))
    assert np.asarray(got[1]).tolist() == [False] * 6


def test_segmented_partial_offsets(impl):
    vals = np.arange(1, 8, dtype=np.int32)
    col = (vals, np.ones(7, bool))  # nullable column without nulls, as in the reference test
    got = impl.segmented_reduce(col, [1, 3, 4], "sum", np.int32, include=True)
    assert_columns_equal(got, (np.array([5, 4], dtype=np.int32), np.array([True, True])))
    got = impl.segmented_reduce(col, [1, 3, 4], "sum", np.int32, include=True, init=(3, True))
    assert_columns_equal(got, (np.array([8, 7], dtype=np.int32), np.array([True, True])))
    got = impl.segmented_reduce(col, [1, 3, 4], "sum", np.int32, include=True, init=(3, False))
    assert np.asarray(got[1]).tolist() == [False, False]
    # min / max per segment (segmented_reduction_tests.cpp MinExcludeNulls / MaxExcludeNulls shapes)
    c2 = make_col([1, 2, 3, 1, N, 3, 1, N, N, N], np.int32)
    got = impl.segmented_reduce(c2, [0, 3, 6, 7, 8, 10, 10], "min", np.int32)
    assert_columns_equal(got, (np.array([1, 1, 1, 0, 0, 0], dtype=np.int32), np.array([1, 1, 1, 0, 0, 0], dtype=bool)))
    got = impl.segmented_reduce(c2, [0, 3, 6, 7, 8, 10, 10], "max", np.int32)
    assert_columns_equal(got, (np.array([3, 3, 1, 0, 0, 0], dtype=np.int32), np.array([1, 1, 1, 0, 0, 0], dtype=bool)))


# cpp/tests/reductions/reduction_tests.cpp:122-243 (MinMax), :330-367 (sum), product
@pytest.mark.parametrize("dtype", TYPES)
def test_reduce_min_max_product(impl, dtype):
    vals = [5, 0, 9, 8, 3, 6, 1, 2]  # unsigned-safe
    col = make_col(vals, dtype)
    assert impl.reduce(col, "min", dtype) == (np.dtype(dtype).type(0), True)
    assert impl.reduce(col, "max", dtype) == (np.dtype(dtype).type(9), True)
    coln = make_col([5, N, 9, 8, N, 6, 1, 2], dtype)
    assert impl.reduce(coln, "min", dtype) == (np.dtype(dtype).type(1), True)
    assert impl.reduce(coln, "sum", dtype) == (np.dtype(dtype).type(31), True)
    p, ok = impl.reduce(make_col([1, 2, 3, N, 2], dtype), "product", dtype)
    assert ok and p == np.dtype(dtype).type(12)


# cpp/tests/groupby/{min,max,mean,count}_tests.cpp: basic and null_keys_and_values (min :33-40,100-118; max :33-40,104-122;
# mean :100-125; count :104-123).  Groups compared after sorting by key, like test_single_agg.
GKEYS = [1, 2, 3, 1, 2, 2, 1, 3, 3, 2]
GVALS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]
NKEYS = [1, 2, 3, 1, 2, 2, 1, N, 3, 2, 4]
NVALS = [N, 1, 2, 3, 4, N, 6, 7, 8, 9, N]
GB_MORE = [
    ("min", GKEYS, GVALS, [1, 2, 3], [0, 1, 2]),
    ("max", GKEYS, GVALS, [1, 2, 3], [6, 9, 8]),
    ("min", NKEYS, NVALS, [1, 2, 3, 4], [3, 1, 2, N]),
    ("max", NKEYS, NVALS, [1, 2, 3, 4], [6, 9, 8, N]),
    ("mean", NKEYS, NVALS, [1, 2, 3, 4], [4.5, 14.0 / 3, 5.0, N]),
    ("count", NKEYS, NVALS, [1, 2, 3, 4], [2, 3, 2, 0]),
]


@pytest.mark.parametrize("case", GB_MORE, ids=lambda c: f"{c[0]}-{len(c[1])}")
@pytest.mark.parametrize("vdtype", [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64])
def test_groupby_min_max_mean_count_golden(impl, case, vdtype):
    from tests.impls import sort_groups

    kind, keys, vals, ekeys, evals = case
    k, res = sort_groups(*impl.groupby([make_col(keys, np.int32)], [(make_col(vals, vdtype), [kind])]))
    assert np.asarray(k[0][0]).tolist() == ekeys
    got = res[0][0]
    ev = np.array([v is not None for v in evals])
    gm = np.ones(len(evals), bool) if got[1] is None else np.asarray(got[1])
    assert gm.tolist() == ev.tolist()
    exp = np.array([0 if v is None else v for v in evals], dtype=np.float64)
    np.testing.assert_allclose(np.asarray(got[0], dtype=np.float64)[ev], exp[ev], rtol=1e-6)
    want = {"min": np.dtype(vdtype), "max": np.dtype(vdtype), "mean": np.dtype(np.float64), "count": np.dtype(np.int32)}[kind]
    assert np.asarray(got[0]).dtype == want


# cpp/tests/copying/gather_tests.cpp:44-250 (IdentityTest, ReverseIdentityTest, EveryOtherNullOdds/Evens, AllNull,
# MultiColReverseIdentityTest, MultiColNulls) — run against the library only (the oracle's gather is trivial numpy)
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", TYPES)
def test_gather_golden(plc, dtype):
    n = 1000
    data = np.arange(n) % (100 if np.dtype(dtype).itemsize == 1 else n)
    src = plc.Column.from_numpy(data.astype(dtype))
    ident = plc.Column.from_numpy(np.arange(n, dtype=np.int32))
    rev = plc.Column.from_numpy(np.arange(n - 1, -1, -1, dtype=np.int32))
    out = plc.copying.gather(plc.Table([src]), ident, plc.OutOfBoundsPolicy.DONT_CHECK).columns()[0]
    assert np.array_equal(out.to_numpy()[0], data.astype(dtype)) and out.null_count() == 0 and out.null_mask() is None
    out = plc.copying.gather(plc.Table([src, src]), rev, plc.OutOfBoundsPolicy.DONT_CHECK)
    for c in out.columns():
        assert np.array_equal(c.to_numpy()[0], data.astype(dtype)[::-1])
    # every other element null: gathering the null rows gives all nulls, the others none
    valid = (np.arange(n) % 2) != 0
    srcn = plc.Column.from_numpy(data.astype(dtype), valid)
    evens = plc.Column.from_numpy((np.arange(n // 2) * 2).astype(np.int32))
    odds = plc.Column.from_numpy((np.arange(n // 2) * 2 + 1).astype(np.int32))
    o = plc.copying.gather(plc.Table([srcn]), evens, plc.OutOfBoundsPolicy.DONT_CHECK).columns()[0]
    assert o.null_count() == n // 2 and not o.to_numpy()[1].any()
    o = plc.copying.gather(plc.Table([srcn]), odds, plc.OutOfBoundsPolicy.DONT_CHECK).columns()[0]
    v, m = o.to_numpy()
    assert o.null_count() == 0 and m.all() and np.array_equal(v, data.astype(dtype)[1::2])
    alln = plc.Column.from_numpy(data.astype(dtype), np.zeros(n, bool))
    o = plc.copying.gather(plc.Table([alln]), rev, plc.OutOfBoundsPolicy.DONT_CHECK).columns()[0]
    assert o.null_count() == n
    # zero-column table keeps the row count semantics trivially; empty gather map -> empty table
    e = plc.copying.gather(plc.Table([src]), plc.Column.from_numpy(np.empty(0, np.int32)), plc.OutOfBoundsPolicy.DONT_CHECK)
    assert e.num_rows() == 0 and e.columns()[0].type().id() == src.type().id()
