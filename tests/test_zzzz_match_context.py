"""hash_join match context and partitioned probes (cpp/include/cudf/join/hash_join.hpp:254-440).

Golden vectors are the fixed-width cases of cpp/tests/join/join_tests.cpp:2339-2570 (the string key columns of those
tests are out of scope, the int32 columns carry the same expectations). Runs against the oracle on CPU and against the
CUDA path on a GPU; the CUDA kernels of this file were first exercised on the emulator (tests/test_emu_kernels.py).
"""
import numpy as np
import pytest

from tests.helpers import make_col
from tests.impls import OracleImpl, PlcImpl


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def impl(request):
    if request.param == "oracle":
        return OracleImpl()
    return PlcImpl(request.getfixturevalue("plc"))


def icol(vals, dtype=np.int32):
    return make_col(vals, dtype)


PROBE, BUILD = [3, 1, 2, 0, 2], [2, 2, 0, 4, 3]


@pytest.mark.parametrize("ne", [0, 1])
def test_match_context_golden(impl, ne):
    # HashJoinInnerMatchContext / Left / Full, single int32 column (join_tests.cpp:2363-2377,2424-2437,2476-2483)
    assert impl.match_counts([icol(PROBE)], [icol(BUILD)], ne, "inner").tolist() == [1, 0, 2, 1, 2]
    assert impl.match_counts([icol(PROBE)], [icol(BUILD)], ne, "left").tolist() == [1, 1, 2, 1, 2]
    assert impl.match_counts([icol(PROBE)], [icol(BUILD)], ne, "full").tolist() == [1, 1, 2, 1, 2]
    # HashJoinMatchContextDuplicatesAndEdgeCases (join_tests.cpp:2527-2563)
    assert impl.match_counts([icol([1, 1, 2, 2, 3])], [icol([1, 1, 1, 2, 4])], ne, "inner").tolist() == [3, 3, 1, 1, 0]


def test_match_context_empty_right(impl):
    # HashJoinMatchContextEmptyRight (join_tests.cpp:2494-2525)
    l, r = [icol([3, 1, 2])], [icol([])]
    assert impl.match_counts(l, r, 0, "inner").tolist() == [0, 0, 0]
    assert impl.match_counts(l, r, 0, "left").tolist() == [1, 1, 1]
    assert impl.match_counts(l, r, 0, "full").tolist() == [1, 1, 1]


@pytest.mark.parametrize("kind", ["inner", "left", "full"])
def test_partitioned_join_golden(impl, kind):
    o = OracleImpl()
    l, r = [icol(PROBE)], [icol(BUILD)]
    exp = getattr(o, f"{kind}_join")(l, r, 0)
    for bounds in [(0,), (2,), (1, 3, 4), (0, 1, 2, 3, 4, 5)]:
        got = impl.partitioned_join(l, r, 0, kind, bounds)
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]), (kind, bounds)


@pytest.mark.parametrize("kdtype", [np.int64, np.int32, np.float64, np.int16])
def test_match_context_random(impl, kdtype):
    """Counts sum to *_join_size; the partitioned probes reproduce the one-shot join for any cut of the left table."""
    rng = np.random.default_rng(123)
    o = OracleImpl()
    for nl, nr, hi in [(1, 1, 2), (1000, 700, 300), (30_000, 9_000, 4000)]:
        l = [(rng.integers(0, hi, nl).astype(kdtype), rng.random(nl) < 0.9)]
        r = [(rng.integers(0, hi, nr).astype(kdtype), rng.random(nr) < 0.95)]
        for ne in (0, 1):
            for kind in ("inner", "left", "full"):
                c = impl.match_counts(l, r, ne, kind)
                assert c.dtype == np.int32 and np.array_equal(c, o.match_counts(l, r, ne, kind)), (nl, ne, kind)
                exp = getattr(o, f"{kind}_join")(l, r, ne)
                bounds = sorted(rng.integers(0, nl + 1, 3).tolist())
                got = impl.partitioned_join(l, r, ne, kind, bounds)
                assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]), (nl, ne, kind, bounds)


def test_match_context_wide_keys(impl):
    rng = np.random.default_rng(124)
    o = OracleImpl()
    nl, nr = 8000, 5000
    l = [(rng.integers(0, 40, nl).astype(np.int64), None), (rng.integers(0, 30, nl).astype(np.int64), rng.random(nl) < 0.9)]
    r = [(rng.integers(0, 40, nr).astype(np.int64), None), (rng.integers(0, 30, nr).astype(np.int64), rng.random(nr) < 0.9)]
    for kind in ("inner", "left", "full"):
        assert np.array_equal(impl.match_counts(l, r, 0, kind), o.match_counts(l, r, 0, kind))
        exp = getattr(o, f"{kind}_join")(l, r, 0)
        got = impl.partitioned_join(l, r, 0, kind, (1000, 4321))
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]), kind


@pytest.mark.gpu
def test_partition_context_errors(plc):
    t = plc.Table([plc.Column.from_numpy(np.array(PROBE, np.int32))])
    hj = plc.join.HashJoin(plc.Table([plc.Column.from_numpy(np.array(BUILD, np.int32))]), plc.NullEquality.EQUAL)
    ctx = hj.inner_join_match_context(t)
    for a, b in [(-1, 2), (3, 2), (0, 6)]:  # outside the left table -> std::invalid_argument (hash_join.hpp:343-345)
        with pytest.raises(ValueError):
            hj.partitioned_inner_join(plc.join.JoinPartitionContext(ctx, a, b))
    with pytest.raises(ValueError):
        hj.partitioned_inner_join(plc.join.JoinPartitionContext(None, 0, 1))
    l, r = hj.partitioned_inner_join(plc.join.JoinPartitionContext(ctx, 2, 2))
    assert l.size() == 0 and r.size() == 0
