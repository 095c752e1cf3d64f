"""The compiled (Cython) binding `cudf_b200.pylibcudf_cy` on the GPU: the hot-path operations through typed calls into
libcudf_b200.so, compared with the oracle exactly like the ctypes twin's parity tests (tests/test_parity_gpu.py)."""
import numpy as np
import pytest

from oracle import sort as osort
from tests.helpers import assert_columns_equal
from tests.impls import OracleImpl, PlcImpl, sort_groups

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cy(plc):
    from tests.conftest import EMU_RUN, _has_gpu

    if EMU_RUN and not _has_gpu():
        from tests.emu.harness import install_cy

        return install_cy()
    import cudf_b200.pylibcudf_cy as cy

    return cy


def test_sort_paths(cy):
    rng = np.random.default_rng(101)
    n = 300_007
    keys = rng.integers(-2**62, 2**62, n).astype(np.int64)
    vals = rng.random(n)
    before = cy._core.kernel_launch_count()
    got = cy.sorting.sort_by_key(cy.Table([cy.Column.from_numpy(vals)]), cy.Table([cy.Column.from_numpy(keys)]), [cy.Order.ASCENDING], [])
    assert np.array_equal(got.columns()[0].to_numpy()[0], osort.sort_by_key([(vals, None)], [(keys, None)], [0])[0][0])
    assert cy._core.kernel_launch_count() > before
    valid = rng.random(n) < 0.9
    k32 = rng.integers(-1000, 1000, n).astype(np.int32)
    for order, nulls in ((0, 0), (1, 1)):
        so = cy.sorting.stable_sorted_order(cy.Table([cy.Column.from_numpy(k32, valid)]), [order], [nulls])
        assert np.array_equal(so.to_numpy()[0], osort.sorted_order([(k32, valid)], [order], [nulls]))
    out = cy.sorting.sort(cy.Table([cy.Column.from_numpy(keys)]), [1], [])
    assert np.array_equal(out.columns()[0].to_numpy()[0], np.sort(keys)[::-1])


def test_join_groupby_reduce(cy):
    rng = np.random.default_rng(102)
    cu, o = PlcImpl(cy), OracleImpl()
    l = [(rng.integers(0, 20_000, 60_000), rng.random(60_000) < 0.97)]
    r = [(rng.integers(0, 20_000, 25_000), rng.random(25_000) < 0.97)]
    for ne in (0, 1):
        for kind in ("inner_join", "left_join", "full_join"):
            g, e = getattr(cu, kind)(l, r, ne), getattr(o, kind)(l, r, ne)
            assert np.array_equal(g[0], e[0]) and np.array_equal(g[1], e[1]), (kind, ne)
    assert cu.inner_join_size(l, r) == o.inner_join_size(l, r)
    k = [(rng.integers(0, 5000, 200_000).astype(np.int64), None)]
    v = (rng.integers(-1000, 1000, 200_000).astype(np.int64), rng.random(200_000) < 0.8)
    kinds = ["sum", "min", "max", "count", "count_all", "mean"]
    gk, gr = sort_groups(*cu.groupby(k, [(v, kinds)]))
    ek, er = sort_groups(*o.groupby(k, [(v, kinds)]))
    assert_columns_equal(gk[0], ek[0], what="keys")
    for j, kind in enumerate(kinds):
        assert_columns_equal(gr[0][j], er[0][j], what=kind)
    x = (rng.integers(-100, 100, 400_000).astype(np.int64), rng.random(400_000) < 0.9)
    assert_columns_equal(cu.scan(x, "sum"), o.scan(x, "sum"), what="scan")
    assert cu.reduce(x, "sum", np.int64) == o.reduce(x, "sum", np.int64)
    offs = np.sort(rng.integers(0, 400_000, 300)).astype(np.int32)
    offs[0] = 0
    assert_columns_equal(cu.segmented_reduce(x, offs, "sum", np.int64), o.segmented_reduce(x, offs, "sum", np.int64), what="segmented")


def test_error_classes_and_interop(cy, plc):
    keys = np.arange(10, dtype=np.int64)
    with pytest.raises(RuntimeError):  # cudf::logic_error
        cy.sorting.sort_by_key(cy.Table([cy.Column.from_numpy(keys[:5])]), cy.Table([cy.Column.from_numpy(keys)]), [0], [])
    with pytest.raises(ValueError):  # std::invalid_argument
        cy.join.HashJoin(cy.Table([cy.Column.from_numpy(keys)]), 0, None, 1.5)
    pc = plc.Column.from_numpy(keys)
    assert np.array_equal(cy.Column.from_plc(pc).to_plc().to_numpy()[0], keys)
