"""cudf::rank (cpp/include/cudf/sorting.hpp:165-230, cpp/src/sort/rank.cu). Golden vectors: the numeric
columns of cpp/tests/sort/rank_test.cpp:62-430 and the header examples. Oracle on CPU, CUDA path on a GPU (first
exercised on the emulator, tests/test_emu_kernels.py)."""
import numpy as np
import pytest

from oracle import sort as osort

FIRST, AVERAGE, MIN, MAX, DENSE = range(5)
# (column_order, null_policy, null_order): rank_test.cpp:47-56 — ASC 0 / DESC 1; EXCLUDE 0 / INCLUDE 1; AFTER 0 / BEFORE 1
ASC_KEEP, ASC_TOP, ASC_BOTTOM = (0, 0, 0), (0, 1, 1), (0, 1, 0)
DESC_KEEP, DESC_TOP, DESC_BOTTOM = (1, 0, 1), (1, 1, 0), (1, 1, 1)
COL = [5, 4, 3, 5, 8, 5]
MASK = [1, 1, 0, 1, 1, 1]
X = None  # masked entry

GOLDEN = [  # (method, args, percentage, expected col1 (no nulls), expected col2 (row 2 null))
    (FIRST, ASC_KEEP, False, [3, 2, 1, 4, 6, 5], [2, 1, X, 3, 5, 4]),
    (DENSE, ASC_TOP, False, [3, 2, 1, 3, 4, 3], [3, 2, 1, 3, 4, 3]),
    (DENSE, ASC_BOTTOM, False, [3, 2, 1, 3, 4, 3], [2, 1, 4, 2, 3, 2]),
    (DENSE, DESC_TOP, False, [2, 3, 4, 2, 1, 2], [3, 4, 1, 3, 2, 3]),
    (DENSE, DESC_BOTTOM, False, [2, 3, 4, 2, 1, 2], [2, 3, 4, 2, 1, 2]),
    (MIN, ASC_TOP, False, [3, 2, 1, 3, 6, 3], [3, 2, 1, 3, 6, 3]),
    (MIN, ASC_BOTTOM, False, [3, 2, 1, 3, 6, 3], [2, 1, 6, 2, 5, 2]),
    (MIN, DESC_TOP, False, [2, 5, 6, 2, 1, 2], [3, 6, 1, 3, 2, 3]),
    (MIN, DESC_BOTTOM, False, [2, 5, 6, 2, 1, 2], [2, 5, 6, 2, 1, 2]),
    (MAX, ASC_TOP, False, [5, 2, 1, 5, 6, 5], [5, 2, 1, 5, 6, 5]),
    (MAX, ASC_BOTTOM, False, [5, 2, 1, 5, 6, 5], [4, 1, 6, 4, 5, 4]),
    (MAX, DESC_TOP, False, [4, 5, 6, 4, 1, 4], [5, 6, 1, 5, 2, 5]),
    (MAX, DESC_BOTTOM, False, [4, 5, 6, 4, 1, 4], [4, 5, 6, 4, 1, 4]),
    (AVERAGE, ASC_KEEP, False, [4, 2, 1, 4, 6, 4], [3, 1, X, 3, 5, 3]),
    (AVERAGE, ASC_TOP, False, [4, 2, 1, 4, 6, 4], [4, 2, 1, 4, 6, 4]),
    (AVERAGE, ASC_BOTTOM, False, [4, 2, 1, 4, 6, 4], [3, 1, 6, 3, 5, 3]),
    (AVERAGE, DESC_KEEP, False, [3, 5, 6, 3, 1, 3], [3, 5, X, 3, 1, 3]),
    (AVERAGE, DESC_TOP, False, [3, 5, 6, 3, 1, 3], [4, 6, 1, 4, 2, 4]),
    (AVERAGE, DESC_BOTTOM, False, [3, 5, 6, 3, 1, 3], [3, 5, 6, 3, 1, 3]),
    (DENSE, ASC_KEEP, True, [0.75, 0.5, 0.25, 0.75, 1.0, 0.75], [2 / 3, 1 / 3, X, 2 / 3, 1.0, 2 / 3]),
    (DENSE, ASC_TOP, True, [0.75, 0.5, 0.25, 0.75, 1.0, 0.75], [0.75, 0.5, 0.25, 0.75, 1.0, 0.75]),
    (DENSE, ASC_BOTTOM, True, [0.75, 0.5, 0.25, 0.75, 1.0, 0.75], [0.5, 0.25, 1.0, 0.5, 0.75, 0.5]),
    (MIN, DESC_KEEP, True, [1 / 3, 5 / 6, 1.0, 1 / 3, 1 / 6, 1 / 3], [0.4, 1.0, X, 0.4, 0.2, 0.4]),
    (MIN, DESC_TOP, True, [1 / 3, 5 / 6, 1.0, 1 / 3, 1 / 6, 1 / 3], [0.5, 1.0, 1 / 6, 0.5, 1 / 3, 0.5]),
    (MIN, DESC_BOTTOM, True, [1 / 3, 5 / 6, 1.0, 1 / 3, 1 / 6, 1 / 3], [1 / 3, 5 / 6, 1.0, 1 / 3, 1 / 6, 1 / 3]),
]


class Oracle:
    def rank(self, col, method, order, policy, nprec, pct):
        return osort.rank(col, method, order, policy, nprec, pct)


class Cuda:
    def __init__(self, plc):
        self.plc = plc

    def rank(self, col, method, order, policy, nprec, pct):
        return self.plc.sorting.rank(self.plc.Column.from_numpy(*col), method, order, policy, nprec, pct).to_numpy()


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def impl(request):
    return Oracle() if request.param == "oracle" else Cuda(request.getfixturevalue("plc"))


def check(got, expected, want_double):
    v, m = got
    assert np.asarray(v).dtype == (np.float64 if want_double else np.int32)
    valid = np.array([e is not X for e in expected])
    gm = np.ones(len(v), bool) if m is None else np.asarray(m, bool)
    assert np.array_equal(gm, valid), (gm, valid)
    exp = np.array([0 if e is X else e for e in expected], np.float64)
    np.testing.assert_allclose(np.asarray(v, np.float64)[valid], exp[valid], rtol=1e-12)


@pytest.mark.parametrize("dtype", [np.int8, np.int32, np.int64, np.uint16, np.float32, np.float64])
def test_rank_golden(impl, dtype):
    c1 = (np.array(COL, dtype), None)
    c2 = (np.array(COL, dtype), np.array(MASK, bool))
    for method, (order, policy, nprec), pct, e1, e2 in GOLDEN:
        dbl = pct or method == AVERAGE
        check(impl.rank(c1, method, order, policy, nprec, pct), e1, dbl)
        check(impl.rank(c2, method, order, policy, nprec, pct), e2, dbl)
    # header examples (sorting.hpp:172-203)
    x = (np.array([3, 4, 5, 4, 1, 2], dtype), None)
    for method, exp in [(FIRST, [3, 4, 6, 5, 1, 2]), (AVERAGE, [3, 4.5, 6, 4.5, 1, 2]), (MIN, [3, 4, 6, 4, 1, 2]), (MAX, [3, 5, 6, 5, 1, 2]),
                        (DENSE, [3, 4, 5, 4, 1, 2])]:
        check(impl.rank(x, method, 0, 1, 0, False), exp, method == AVERAGE)


def test_rank_random(impl):
    rng = np.random.default_rng(41)
    o = Oracle()
    for n in (1, 2, 777, 30_000):
        for dtype in (np.int16, np.float64):
            v = rng.integers(0, max(2, n // 7), n).astype(dtype)
            if dtype == np.float64 and n > 10:
                v[::53] = np.nan
                v[::47] = -0.0
            col = (v, (rng.random(n) < 0.9) if n > 1 else None)
            for method in range(5):
                for order, policy, nprec in (ASC_KEEP, ASC_TOP, DESC_TOP, DESC_KEEP):
                    for pct in (False, True):
                        g, e = impl.rank(col, method, order, policy, nprec, pct), o.rank(col, method, order, policy, nprec, pct)
                        gm = np.ones(n, bool) if g[1] is None else np.asarray(g[1], bool)
                        em = np.ones(n, bool) if e[1] is None else np.asarray(e[1], bool)
                        assert np.asarray(g[0]).dtype == np.asarray(e[0]).dtype and np.array_equal(gm, em)
                        np.testing.assert_allclose(np.asarray(g[0], np.float64)[em], np.asarray(e[0], np.float64)[em], rtol=1e-12)


def test_rank_empty(impl):
    v, m = impl.rank((np.array([], np.int32), None), AVERAGE, 0, 1, 0, False)
    assert len(v) == 0
