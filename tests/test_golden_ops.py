"""Golden (known-answer) cases of the reference's gtests for join / groupby / scan / segmented reduce,
run against BOTH the oracle (CPU, pins the oracle) and the CUDA path (GPU, parity proper)."""
import numpy as np
import pytest

from tests.golden import misc_cases as G
from tests.helpers import assert_columns_equal, make_col
from tests.impls import OracleImpl, PlcImpl, sort_groups


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def impl(request):
    if request.param == "oracle":
        return OracleImpl()
    return PlcImpl(request.getfixturevalue("plc"))


def icol(vals, dtype=np.int32):
    return make_col(vals, dtype)


def test_join_gold_maps(impl):
    l, r = impl.inner_join([icol(G.JOIN_GOLD_MAPS["probe"])], [icol(G.JOIN_GOLD_MAPS["build"])])
    assert l.tolist() == G.JOIN_GOLD_MAPS["left"] and r.tolist() == G.JOIN_GOLD_MAPS["right"]


@pytest.mark.parametrize("kind", ["inner_join", "left_join", "full_join"])
def test_join_nulls_one_side(impl, kind):
    c = G.JOIN_NULLS_ONE_SIDE
    build = [icol(x) for x in c["build"]]
    probe = [icol(x) for x in c["probe"]]
    l, r = getattr(impl, kind)(probe, build, 0)
    el, er = c[kind]
    # the reference sorts the two index columns independently (join_tests.cpp:2215-2229)
    assert sorted(l.tolist()) == sorted(el) and sorted(r.tolist()) == sorted(er)


def test_join_equal_values_and_empty(impl):
    c = G.JOIN_EQUAL_VALUES
    l, r = impl.inner_join([icol(c["left"])], [icol(c["right"])])
    assert len(l) == c["pairs"] and sorted(zip(l.tolist(), r.tolist())) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    # empty sides: join_tests.cpp:1635-1655,1719-1757,1839-1859
    e = icol([])
    x = icol([1, 2, 3])
    for a, b in ((e, x), (x, e), (e, e)):
        l, r = impl.inner_join([a], [b])
        assert len(l) == 0 and len(r) == 0
    l, r = impl.left_join([x], [e])
    assert l.tolist() == [0, 1, 2] and r.tolist() == [G.NO_MATCH] * 3
    l, r = impl.full_join([e], [x])
    assert sorted(r.tolist()) == [0, 1, 2] and l.tolist() == [G.NO_MATCH] * 3


def test_join_large_output_size(impl):
    n = G.JOIN_LARGE["n"]
    z = (np.zeros(n, np.int32), None)
    assert impl.inner_join_size([z], [z], 1) == n * n  # > INT32_MAX (join_tests.cpp:2299-2314)


def test_join_null_equality(impl):
    l = [icol([1, N := None, 3, None])]
    r = [icol([None, 1, None])]
    a, b = impl.inner_join(l, r, 0)  # EQUAL: null == null
    assert sorted(zip(a.tolist(), b.tolist())) == [(0, 1), (1, 0), (1, 2), (3, 0), (3, 2)]
    a, b = impl.inner_join(l, r, 1)  # UNEQUAL
    assert sorted(zip(a.tolist(), b.tolist())) == [(0, 1)]


def test_join_float_keys(impl):
    nan = float("nan")
    l = [(np.array([0.0, -0.0, nan, 1.5]), None)]
    r = [(np.array([-0.0, nan, -nan, 2.5]), None)]
    a, b = impl.inner_join(l, r)
    assert sorted(zip(a.tolist(), b.tolist())) == [(0, 0), (1, 0), (2, 1), (2, 2)]


@pytest.mark.parametrize("case", G.GROUPBY_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("vdtype", [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64])
def test_groupby_golden(impl, case, vdtype):
    keys = [icol(case["keys"])]
    vals = make_col(case["vals"], vdtype)
    k, res = impl.groupby(keys, [(vals, [case["kind"]])])
    k, res = sort_groups(k, res)
    assert k[0][0].tolist() == case["ekeys"], case["cite"]
    got = res[0][0]
    exp_valid = np.array([v is not None for v in case["evals"]], dtype=bool)
    exp_vals = np.array([0 if v is None else v for v in case["evals"]])
    rdt = {"sum": np.int64 if np.dtype(vdtype).kind in "iu" else vdtype, "count": np.int32, "count_all": np.int32, "mean": np.float64,
           "min": vdtype, "max": vdtype}[case["kind"]]
    assert np.asarray(got[0]).dtype == np.dtype(rdt), f"{case['name']}: result dtype {np.asarray(got[0]).dtype}"
    gm = np.ones(len(exp_vals), bool) if got[1] is None else np.asarray(got[1])
    assert gm.tolist() == exp_valid.tolist()
    np.testing.assert_allclose(np.asarray(got[0], dtype=np.float64)[exp_valid], exp_vals.astype(np.float64)[exp_valid], rtol=1e-6)


@pytest.mark.parametrize("case", G.GROUPBY_SCAN_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("vdtype", [np.int8, np.int32, np.int64, np.float32, np.float64])
def test_groupby_scan_golden(impl, case, vdtype):
    keys = [icol(case["keys"])]
    vals = make_col(case["vals"], vdtype)
    k, res = impl.groupby_scan(keys, [(vals, [case["kind"]])])
    assert np.asarray(k[0][0]).tolist() == case["ekeys"], case["cite"]
    got = res[0][0]
    exp_valid = np.array([v is not None for v in case["evals"]], dtype=bool)
    exp_vals = np.array([0 if v is None else v for v in case["evals"]], dtype=np.float64)
    gm = np.ones(len(exp_vals), bool) if got[1] is None else np.asarray(got[1])
    assert gm.tolist() == exp_valid.tolist()
    np.testing.assert_allclose(np.asarray(got[0], dtype=np.float64)[exp_valid], exp_vals[exp_valid], rtol=1e-6)


def test_groupby_include_null_keys(impl):
    # keys_tests.cpp:25-108: INCLUDE makes the null keys one group
    keys = [icol([1, None, 2, None, 1])]
    vals = icol([1, 2, 3, 4, 5], np.int64)
    k, res = sort_groups(*impl.groupby(keys, [(vals, ["sum", "count_all"])], include_nulls=True))
    assert (k[0][1] is not None) and np.asarray(k[0][1]).tolist() == [False, True, True]
    assert np.asarray(res[0][0][0]).tolist() == [6, 6, 3] and np.asarray(res[0][1][0]).tolist() == [2, 2, 1]


@pytest.mark.parametrize("case", G.SCAN_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint32, np.uint64, np.float32, np.float64])
def test_scan_golden(impl, case, dtype):
    col = make_col(case["vals"], dtype)
    got = impl.scan(col, "sum", case["inclusive"], case["policy"] == "INCLUDE")
    exp = make_col(case["expected"], dtype)
    assert_columns_equal(got, exp, what=case["name"])


@pytest.mark.parametrize("dtype", [np.int8, np.int32, np.int64, np.uint16, np.uint64, np.float32, np.float64])
def test_segmented_sum_golden(impl, dtype):
    c = G.SEGMENTED_SUM
    col = make_col(c["vals"], dtype)
    got = impl.segmented_reduce(col, c["offsets"], "sum", dtype)
    exp = make_col(c["expected"], dtype)
    assert_columns_equal(got, (exp[0], np.array([v is not None for v in c["expected"]])), what="SumExcludeNulls")
    got = impl.segmented_reduce(col, c["offsets"], "sum", dtype, init=(c["init"], True))
    assert_columns_equal(got, (np.array(c["expected_init"], dtype=dtype), np.ones(6, bool)), what="init 3")
    got = impl.segmented_reduce(col, c["offsets"], "sum", dtype, init=(c["init"], False))
    assert_columns_equal(got, (exp[0], np.array([v is not None for v in c["expected"]])), what="null init")


def test_reduce_basics(impl):
    # reduction_tests.cpp:330-367 (sum), :122-243 (min/max), :801-857 (mean), :999-1133 (all-null / dtype cast)
    v = make_col([6, -14, 13, 64, 0, -13, -20, 45], np.int32)
    assert impl.reduce(v, "sum", np.int32) == (81, True)
    assert impl.reduce(v, "sum", np.int64) == (81, True)
    assert impl.reduce(v, "min", np.int32) == (-20, True)
    assert impl.reduce(v, "max", np.int32) == (64, True)
    m, ok = impl.reduce(v, "mean", np.float64)
    assert ok and abs(m - 81 / 8) < 1e-12
    vn = make_col([6, None, 13, None, 0], np.int16)
    assert impl.reduce(vn, "sum", np.int64) == (19, True)
    alln = make_col([None, None], np.float64)
    assert impl.reduce(alln, "sum", np.float64)[1] is False
    assert impl.reduce(make_col([], np.int32), "max", np.int32)[1] is False
    # output dtype != input dtype accumulates in int64 / double then casts (simple.cuh:407-419)
    big = (np.array([2 ** 31 - 1, 2 ** 31 - 1], dtype=np.int32), None)
    assert impl.reduce(big, "sum", np.int64) == (2 ** 32 - 2, True)
    assert impl.reduce(big, "sum", np.int32)[0] == np.int32(-2)
    f = (np.array([0.5, 0.25, 0.125], dtype=np.float32), None)
    assert impl.reduce(f, "sum", np.float64) == (0.875, True)
    assert impl.reduce(v, "sum", np.int32, init=(5, True)) == (86, True)
    assert impl.reduce(v, "sum", np.int32, init=(5, False))[1] is False
