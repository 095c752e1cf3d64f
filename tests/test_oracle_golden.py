"""CPU: the oracle restatement against the reference's own known-answer tests (pins the oracle)."""
import numpy as np
import pytest

import oracle
from oracle import sort as osort
from tests.golden import sort_cases as G
from tests.helpers import ALL_DTYPES, make_col

TYPED = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]


@pytest.mark.parametrize("case", G.SORTED_ORDER_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("dtype", TYPED + [np.bool_])
def test_sorted_order_golden(case, dtype):
    if dtype == np.bool_ and case["skip_bool"]:
        pytest.skip("the reference does not pin bool order for this case")
    cols = [make_col(v, dtype if k == "T" else np.int32) for v, k in case["cols"]]
    got = osort.sorted_order(cols, case["order"], case["nulls"])
    assert got.tolist() == case["expected"], case["cite"]


@pytest.mark.parametrize("dtype", TYPED)
def test_sliced_columns(dtype):
    c1 = make_col(G.SLICED["col1"], np.int32)
    c2 = make_col(G.SLICED["col2"], dtype)
    assert osort.sorted_order([c1, c2], [0, 0]).tolist() == G.SLICED["expected"]
    k = G.SLICED["split"]
    s1 = (c1[0][k:], c1[1][k:])
    assert osort.sorted_order([s1, s1], [0, 0]).tolist() == G.SLICED["expected_sliced"]


@pytest.mark.parametrize("dtype", TYPED)
def test_single_column(dtype):
    for case, valid in ((G.SINGLE_NO_NULL, None), (G.SINGLE_WITH_NULL, G.SINGLE_WITH_NULL["valid"])):
        vals = [v if (valid is None or valid[i]) else None for i, v in enumerate(case["values"])]
        col = make_col(vals, dtype)
        exp = case["expected_unsigned"] if np.dtype(dtype).kind == "u" else case["expected_signed"]
        got = osort.sorted_order([col], [0], [1])
        # the reference compares gathered values (run_stable_sort_test), not indices
        gv = osort.gather([col], got)[0]
        ev = osort.gather([col], exp)[0]
        gm = np.ones(len(got), bool) if gv[1] is None else gv[1]
        em = np.ones(len(got), bool) if ev[1] is None else ev[1]
        assert np.array_equal(gm, em)
        assert np.array_equal(gv[0][gm], ev[0][em])
        if valid is None:
            assert got.tolist() == exp  # no ties among nulls: a stable order is unique


def test_inf_nan():
    col = (np.array(G.INF_NAN["values"], dtype=np.float64), None)
    assert osort.sorted_order([col]).tolist() == G.INF_NAN["expected"]
    # descending (not pinned by the reference): tuple (isnan*(idx+1), f) sorted descending
    got = osort.sorted_order([col], [1]).tolist()
    assert got[:6] == [13, 12, 9, 3, 2, 1]
    assert got[6:] == [4, 10, 6, 8, 7, 0, 14, 5, 11]


def test_size_mismatch_throws():
    # sort_test.cpp:667-698 MismatchInColumnOrderSize / MismatchInNullPrecedenceSize
    c = make_col([1, 2, 3], np.int32)
    with pytest.raises(RuntimeError):
        osort.sorted_order([c, c], [0])
    with pytest.raises(RuntimeError):
        osort.sorted_order([c, c], [0, 0], [1])
    # sort_test.cpp:962-978 sort_by_key size mismatch
    with pytest.raises(RuntimeError):
        osort.sort_by_key([make_col([1, 2], np.int32)], [c])
    # sort_test.cpp:700-717 zero sized
    assert len(osort.sorted_order([(np.empty(0, np.int32), None)])) == 0


def test_join_gold_maps():
    # join_tests.cpp:2316-2337 (gold maps) with the InnerJoinNoNulls column 0 (:1163-1237)
    l, r = oracle.join.inner_join([(np.array([3, 1, 2, 0, 2]), None)], [(np.array([2, 2, 0, 4, 3]), None)])
    assert l.tolist() == [0, 2, 2, 3, 4, 4] and r.tolist() == [4, 0, 1, 2, 0, 1]
