"""segmented_sorted_order / segmented_sort_by_key / top_k / top_k_order (cpp/include/cudf/sorting.hpp:232-416).
Golden vectors: cpp/tests/sort/segmented_sort_tests.cpp:69-260 and the sorting.hpp examples. Oracle on CPU, CUDA path on
a GPU (first exercised on the emulator, tests/test_emu_kernels.py)."""
import numpy as np
import pytest

from oracle import sort as osort

SEG = np.array([0, 3, 5, 5, 5, 6, 11, 13, 14, 16], np.int32)


class Oracle:
    name = "oracle"

    def seg_order(self, keys, offs, order=None, prec=None, stable=False):
        return osort.segmented_sorted_order(keys, offs, order, prec)

    def seg_sort_by_key(self, vals, keys, offs, order=None, prec=None, stable=False):
        return osort.segmented_sort_by_key(vals, keys, offs, order, prec)

    def top_k(self, col, k, order=1):
        return osort.top_k(col, k, order)

    def top_k_order(self, col, k, order=1):
        return osort.top_k_order(col, k, order)


class Cuda:
    name = "cuda"

    def __init__(self, plc):
        self.plc = plc

    def _tbl(self, cols):
        return self.plc.Table([self.plc.Column.from_numpy(v, m) for v, m in cols])

    def seg_order(self, keys, offs, order=None, prec=None, stable=False):
        s = self.plc.sorting
        f = s.stable_segmented_sorted_order if stable else s.segmented_sorted_order
        return f(self._tbl(keys), self.plc.Column.from_numpy(np.asarray(offs)), order or [], prec or []).to_numpy()[0]

    def seg_sort_by_key(self, vals, keys, offs, order=None, prec=None, stable=False):
        s = self.plc.sorting
        f = s.stable_segmented_sort_by_key if stable else s.segmented_sort_by_key
        out = f(self._tbl(vals), self._tbl(keys), self.plc.Column.from_numpy(np.asarray(offs)), order or [], prec or [])
        return [c.to_numpy() for c in out.columns()]

    def top_k(self, col, k, order=1):
        return self.plc.sorting.top_k(self.plc.Column.from_numpy(*col), k, order).to_numpy()

    def top_k_order(self, col, k, order=1):
        return self.plc.sorting.top_k_order(self.plc.Column.from_numpy(*col), k, order).to_numpy()[0]


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def impl(request):
    return Oracle() if request.param == "oracle" else Cuda(request.getfixturevalue("plc"))


def col(vals, dtype, valid=None):
    return (np.array(vals, dtype), None if valid is None else np.array(valid, bool))


def eq(got, exp_vals, exp_valid=None):
    v, m = got
    ev = np.ones(len(exp_vals), bool) if exp_valid is None else np.array(exp_valid, bool)
    gm = np.ones(len(v), bool) if m is None else np.asarray(m, bool)
    assert np.array_equal(gm, ev), (gm, ev)
    assert np.array_equal(np.asarray(v)[ev], np.array(exp_vals, np.asarray(v).dtype)[ev]), (v, exp_vals)


def test_header_examples(impl):
    keys = [col([9, 8, 7, 6, 5, 4, 3, 2, 1, 0], np.int32)]
    assert impl.seg_order(keys, np.array([0, 3, 7, 10], np.int32)).tolist() == [2, 1, 0, 6, 5, 4, 3, 9, 8, 7]   # sorting.hpp:238-244
    assert impl.seg_order(keys, np.array([3, 7], np.int32)).tolist() == [0, 1, 2, 6, 5, 4, 3, 7, 8, 9]          # sorting.hpp:252-258
    assert impl.seg_order(keys, np.array([], np.int32)).tolist() == list(range(10))                            # :246-247
    assert impl.seg_order(keys, np.array([4], np.int32)).tolist() == list(range(10))


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.int32, np.int64, np.uint32, np.float32, np.float64])
def test_segmented_sort_golden(impl, dtype):
    # SegmentedSort.NoNull (segmented_sort_tests.cpp:69-104)
    c1 = col([10, 36, 14, 32, 49, 23, 10, 34, 12, 45, 12, 37, 43, 26, 21, 16], dtype)
    c2 = col([10, 63, 41, 23, 94, 32, 10, 43, 21, 54, 22, 73, 34, 62, 12, 61], dtype)
    eq(impl.seg_sort_by_key([c1], [c1], SEG, [0])[0], [10, 14, 36, 32, 49, 23, 10, 12, 12, 34, 45, 37, 43, 26, 16, 21])
    eq(impl.seg_sort_by_key([c1], [c1], SEG, [1])[0], [36, 14, 10, 49, 32, 23, 45, 34, 12, 12, 10, 43, 37, 26, 21, 16])
    r = impl.seg_sort_by_key([c1, c2], [c1, c2], SEG, [])
    eq(r[0], [10, 14, 36, 32, 49, 23, 10, 12, 12, 34, 45, 37, 43, 26, 16, 21])
    eq(r[1], [10, 41, 63, 23, 94, 32, 10, 21, 22, 43, 54, 73, 34, 62, 61, 12])
    r = impl.seg_sort_by_key([c1, c2], [c1, c2], SEG, [0, 1])
    eq(r[1], [10, 41, 63, 23, 94, 32, 10, 22, 21, 43, 54, 73, 34, 62, 61, 12])
    # SegmentedSort.Null (segmented_sort_tests.cpp:106-160)
    n1 = col([1, 3, 2, 4, 5, 23, 6, 8, 7, 9, 7, 37, 43, 26, 21, 16], dtype, [1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1])
    eq(impl.seg_sort_by_key([n1], [n1], SEG, [], [0])[0], [1, 3, 2, 4, 5, 23, 6, 7, 7, 8, 9, 37, 43, 26, 16, 21],
       [1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1])
    eq(impl.seg_sort_by_key([n1], [n1], SEG, [], [1])[0], [2, 1, 3, 4, 5, 23, 9, 6, 7, 7, 8, 37, 43, 26, 16, 21],
       [0, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    eq(impl.seg_sort_by_key([n1], [n1], SEG, [1], [0])[0], [2, 3, 1, 5, 4, 23, 9, 8, 7, 7, 6, 43, 37, 26, 21, 16],
       [0, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1])
    eq(impl.seg_sort_by_key([n1], [n1], SEG, [1], [1])[0], [3, 1, 2, 5, 4, 23, 8, 7, 7, 6, 9, 43, 37, 26, 21, 16],
       [1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1])
    # SegmentedSort.StableNoNulls (segmented_sort_tests.cpp:162-183)
    v = col([10, 36, 14, 32, 49, 23, 10, 34, 12, 45, 11, 37, 43, 26, 21, 16], dtype)
    k = col([10, 63, 10, 23, 94, 32, 10, 43, 22, 43, 22, 34, 34, 62, 62, 61], dtype)
    eq(impl.seg_sort_by_key([v], [k], SEG, [0], stable=True)[0], [10, 14, 36, 32, 49, 23, 10, 12, 11, 34, 45, 37, 43, 26, 16, 21])
    eq(impl.seg_sort_by_key([v], [k], SEG, [1], stable=True)[0], [36, 10, 14, 49, 32, 23, 34, 45, 12, 11, 10, 37, 43, 26, 21, 16])


def test_segmented_sort_random(impl):
    rng = np.random.default_rng(17)
    o = Oracle()
    for n, nseg in [(1, 1), (50, 4), (20_000, 300), (30_000, 3)]:
        cuts = np.sort(rng.integers(0, n + 1, nseg + 1)).astype(np.int32)
        cuts[0] = max(0, cuts[0] - 0)
        k1 = (rng.integers(-50, 50, n).astype(np.int32), rng.random(n) < 0.9)
        k2 = (rng.standard_normal(n), None)
        vals = (np.arange(n, dtype=np.int64), None)
        for order, prec in [([0, 1], [1, 0]), ([1, 0], [0, 1])]:
            # the row index as a last key makes the expected order unique (the unstable variant leaves ties open)
            keys = [k1, k2, (np.arange(n, dtype=np.int32), None)]
            got = impl.seg_order(keys, cuts, order + [0], prec + [1])
            assert np.array_equal(got, o.seg_order(keys, cuts, order + [0], prec + [1])), (n, nseg, order)
            got = impl.seg_sort_by_key([vals], [k1, k2], cuts, order, prec, stable=True)[0][0]
            assert np.array_equal(got, o.seg_sort_by_key([vals], [k1, k2], cuts, order, prec)[0][0])


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.uint8, np.float64])
def test_top_k(impl, dtype):
    rng = np.random.default_rng(18)
    o = Oracle()
    for n, k in [(1, 1), (10, 3), (5000, 17), (5000, 5000), (300, 400), (40, 0)]:
        c = (rng.integers(0, 200, n).astype(dtype), (rng.random(n) < 0.85) if n > 5 else None)
        for order in (0, 1):
            idx = impl.top_k_order(c, k, order)
            exp = o.top_k_order(c, k, order)
            assert idx.dtype == np.int32 and len(idx) == len(exp)
            gv, gm = impl.top_k(c, k, order)
            ev, em = o.top_k(c, k, order)
            # any order is allowed: compare as multisets of (validity, value)
            def canon(v, m):
                m = np.ones(len(v), bool) if m is None else np.asarray(m, bool)
                return sorted((bool(a), float(b) if a else 0.0) for a, b in zip(m, v))
            assert canon(gv, gm) == canon(ev, em), (n, k, order)
            assert canon(c[0][idx], None if c[1] is None else c[1][idx]) == canon(ev, em)


@pytest.mark.gpu
def test_segmented_errors(plc):
    t = plc.Table([plc.Column.from_numpy(np.arange(5, dtype=np.int32))])
    with pytest.raises(RuntimeError):  # cudf::logic_error: offsets are not size_type (sorting.hpp:236)
        plc.sorting.segmented_sorted_order(t, plc.Column.from_numpy(np.array([0, 5], np.int64)), [], [])
    with pytest.raises(RuntimeError):  # values / keys row mismatch (segmented_sort_tests.cpp:47-52)
        plc.sorting.segmented_sort_by_key(plc.Table([plc.Column.from_numpy(np.arange(4, dtype=np.int32))]), t,
                                          plc.Column.from_numpy(np.array([0, 5], np.int32)), [], [])
    with pytest.raises(ValueError):    # k < 0 (top_k.cu:106)
        plc.sorting.top_k(t.columns()[0], -1)
